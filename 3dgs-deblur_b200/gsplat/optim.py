"""Adam over the trainer's flat parameter buffer ("next" row f-2).

The reference builds one `torch.optim.Adam(eps=1e-15)` per Splatfacto parameter group
(nerfstudio/engine/optimizers.py:158-171, splatfacto.py:1063-1100).  `FlatAdam` keeps the two moment buffers flat, next
to the flat parameter / gradient buffers of `gsplat.dp.FlatGaussians`, and updates any `[begin, end)` slice with one
kernel (`b200_adam_step`): a data-parallel trainer can update slice i while the gradient exchange of slice i+1 is still
in flight, the gradient slice is cleared in the same pass (no separate zero_grad fill), and a step costs one C call
instead of a Python optimizer round trip per slice.  CUDA only."""
import torch

from . import _lib
from ._lib import check, stream


class FlatAdam:
    def __init__(self, flat: torch.Tensor, flat_grad: torch.Tensor, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-15):
        _lib.require_cuda(flat, flat_grad)
        if flat.dtype != torch.float32 or flat_grad.dtype != torch.float32 or flat.shape != flat_grad.shape or flat.dim() != 1:
            raise ValueError("FlatAdam: expected two 1-D float32 buffers of equal length")
        self.flat, self.flat_grad = flat, flat_grad
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.step_count = 0
        self._ptrs = (flat.data_ptr(), flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr())
        self._lib = _lib.load()

    # ---- device-resident step state (no host sync, CUDA-graph friendly) ------------------------------------------
    def use_device_state(self):
        """Switch to the device-resident step count / bias corrections / veto (b200_adam_prepare,
        b200_adam_step_state): `prepare(veto_flag)` once per step, then `update_state(begin, end)` per slice."""
        if getattr(self, "state", None) is None:
            nbytes = self._lib.b200_adam_state_bytes()
            self.state = torch.zeros(nbytes // 4, dtype=torch.int32, device=self.flat.device)
        return self

    def prepare(self, veto_flag=None):
        """Once per step, before the first `update_state`.  veto_flag: device int32[1] (or None); non-zero skips this step's
        update on the device (gradients are still cleared) and is reset to zero."""
        with _lib.on_device(self.flat.device):
            check(self._lib.b200_adam_prepare(self.state.data_ptr(), None if veto_flag is None else veto_flag.data_ptr(),
                                              self.lr, self.betas[0], self.betas[1], stream()))

    def update_state(self, begin: int = 0, end: int = None, grad_scale: float = 1.0, zero_grad: bool = True,
                     background_ctas: int = 0):
        """background_ctas > 0: run as a background kernel with that many resident CTAs per SM
        (b200_adam_step_state_background) -- for an update queued on a second stream beside latency-bound kernels."""
        end = self.flat.numel() if end is None else end
        if not (0 <= begin <= end <= self.flat.numel()) or begin % 4:
            raise ValueError(f"FlatAdam.update_state: bad slice [{begin}, {end}) (must start on a 16-byte boundary)")
        off = 4 * begin
        p, g, m, v = self._ptrs
        with _lib.on_device(self.flat.device):
            if background_ctas > 0:
                check(self._lib.b200_adam_step_state_background(
                    end - begin, p + off, g + off, m + off, v + off, self.state.data_ptr(), self.betas[0], self.betas[1],
                    self.eps, float(grad_scale), 1 if zero_grad else 0, int(background_ctas), stream()))
            else:
                check(self._lib.b200_adam_step_state(end - begin, p + off, g + off, m + off, v + off, self.state.data_ptr(),
                                                     self.betas[0], self.betas[1], self.eps, float(grad_scale),
                                                     1 if zero_grad else 0, stream()))

    def rebind(self, flat: torch.Tensor, flat_grad: torch.Tensor):
        """The parameter buffers were reallocated (densification): fresh zero moments of the new size; the caller fills them
        (gsplat.densify carries the surviving rows over).  Step count and hyper-parameters stay."""
        _lib.require_cuda(flat, flat_grad)
        self.flat, self.flat_grad = flat, flat_grad
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self._ptrs = (flat.data_ptr(), flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr())

    def begin_step(self):
        """Once per optimisation step, before the first `update` of that step."""
        self.step_count += 1

    def update(self, begin: int = 0, end: int = None, grad_scale: float = 1.0, zero_grad: bool = True):
        """Adam update of flat[begin:end] from flat_grad[begin:end] on the current stream."""
        end = self.flat.numel() if end is None else end
        if not (0 <= begin <= end <= self.flat.numel()):
            raise ValueError(f"FlatAdam.update: bad slice [{begin}, {end})")
        off = 4 * begin
        p, g, m, v = self._ptrs
        with _lib.on_device(self.flat.device):
            check(self._lib.b200_adam_step(end - begin, p + off, g + off, m + off, v + off, self.step_count, self.lr,
                                           self.betas[0], self.betas[1], self.eps, float(grad_scale), 1 if zero_grad else 0,
                                           stream()))

    def step(self, zero_grad: bool = True):
        """Whole-buffer update (single-GPU training)."""
        self.begin_step()
        self.update(0, None, 1.0, zero_grad)
