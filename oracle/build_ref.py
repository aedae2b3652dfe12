"""Build recipe for the UNMODIFIED reference gsplat CUDA extension -> oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported by the product
package; only tests/, __graft_entry__ (build + smoke checker) and bench.py's
baseline legs may touch it.

The reference sources are compiled *where they lie* under
/root/reference/gsplat/gsplat/cuda/csrc (forward.cu, backward.cu, bindings.cu,
ext.cpp + vendored glm); nothing is copied into this repository.  Output is a
single pybind module oracle/_ref/gsplat_ref_csrc.so (git-ignored, travels to
the GPU box with gpurun).  Flags follow the reference's own AOT build
(/root/reference/gsplat/setup.py:76  "-O3 --use_fast_math"), which is what
scripts/install.sh:19-23 (`pip install -e .`) produces for users.

Usage:  python oracle/build_ref.py        (needs /root/reference; ~4 min)
"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_CSRC = "/root/reference/gsplat/gsplat/cuda/csrc"
NAME = "gsplat_ref_csrc"


def so_path():
    return os.path.join(OUT, NAME + ".so")


def build(verbose=False):
    if not os.path.isdir(REF_CSRC):
        return None  # GPU box: prebuilt .so only
    os.makedirs(OUT, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(REF_CSRC, "*.cu"))) + [os.path.join(REF_CSRC, "ext.cpp")]
    newest = max(os.path.getmtime(s) for s in srcs)
    if os.path.exists(so_path()) and os.path.getmtime(so_path()) > newest:
        return so_path()
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load

    load(
        name=NAME,
        sources=srcs,
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3", "--use_fast_math", "--expt-relaxed-constexpr"],
        extra_include_paths=[os.path.join(REF_CSRC, "third_party", "glm")],
        build_directory=OUT,
        verbose=verbose,
        is_python_module=False,  # do not dlopen here (no GPU needed to build)
    )
    return so_path()


def load_ref():
    """Import the prebuilt reference extension (GPU box or here)."""
    import importlib.util

    import torch  # noqa: F401  (libtorch symbols must be loaded first)

    p = so_path()
    if not os.path.exists(p):
        raise FileNotFoundError(p)
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
