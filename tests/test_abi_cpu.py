"""CPU: the C-ABI library loads, exports every symbol include/b200splat.h declares, the ctypes table
matches the header, and argument errors are reported without touching a GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200splat.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"B200_API\s+([\w\s\*]+?)\b(b200_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        nargs = 0 if args in ("", "void") else len(args.split(","))
        out[m.group(2)] = nargs
    return out


@pytest.fixture(scope="module")
def lib():
    from gsplat import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from gsplat import _lib

    decl = _declared()
    assert len(decl) >= 18
    for name, nargs in decl.items():
        assert hasattr(lib, name), f"{name} declared in b200splat.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype"
        assert len(_lib.SIGNATURES[name][1]) == nargs, f"{name}: header has {nargs} args, ctypes table {len(_lib.SIGNATURES[name][1])}"
    assert set(_lib.SIGNATURES) == set(decl)
    assert lib.b200_abi_version() == 1
    assert lib.b200_packed_record_bytes() == 64


def test_argument_errors_are_reported_before_any_gpu_work(lib):
    # invalid arguments return B200_ERR_INVALID (-1) with a message, like the reference's TORCH_CHECKs
    rc = lib.b200_rasterize_forward(10, 8, 8, 16, 11, *([None] * 4), 0.0, 0.0, *([None] * 9))
    assert rc == -1 and b"unsupported blur size" in lib.b200_last_error()
    rc = lib.b200_project_gaussians_forward(0, None, None, 1.0, None, None, None, 0.0, 0.0, None, 1.0, 1.0, 0.0, 0.0, 8, 8,
                                            16, 0.01, *([None] * 10))
    assert rc == -1 and b"num_points" in lib.b200_last_error()
    rc = lib.b200_compute_sh_forward(7, 4, 3, 3, None, None, None, None)
    assert rc == -1 and b"Invalid method" in lib.b200_last_error()


def test_no_cpu_fallback_in_the_product_package():
    """The product must not route through oracle/ or any CPU path: no import of it anywhere in the package."""
    pkg = os.path.join(ROOT, "3dgs-deblur_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, os.path.join(dp, f)


def test_operators_raise_on_cpu_tensors():
    import torch

    import gsplat

    with pytest.raises(RuntimeError, match="CUDA"):
        gsplat.spherical_harmonics(0, torch.zeros(4, 3), torch.zeros(4, 1, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        gsplat.compute_cov2d_bounds(torch.ones(4, 3))
    # the loss / optimizer / dispatcher helpers are CUDA-only as well
    from gsplat.data import ImagePrefetcher
    from gsplat.losses import l1_loss, photometric_loss
    from gsplat.optim import FlatAdam

    with pytest.raises(RuntimeError, match="CUDA"):
        l1_loss(torch.zeros(16, 16, 3), torch.zeros(16, 16, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        photometric_loss(torch.zeros(16, 16, 3), torch.zeros(16, 16, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        FlatAdam(torch.zeros(8), torch.zeros(8))
    with pytest.raises(RuntimeError, match="CUDA"):
        ImagePrefetcher([torch.zeros(4, 4, 3, dtype=torch.uint8)], [torch.zeros(21)], "cpu")


def test_missing_library_fails_loudly(tmp_path):
    """No silent fallback: with the shared library absent every entry point raises, naming the build command."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import torch, gsplat\n"
            "from gsplat import _lib\n"
            "try:\n"
            "    _lib.load()\n"
            "except RuntimeError as e:\n"
            "    assert 'not found' in str(e) and 'no CPU fallback' in str(e), e\n"
            "    print('LOUD')\n" % os.path.join(ROOT, "3dgs-deblur_b200"))
    env = dict(os.environ, B200SPLAT_LIB=str(tmp_path / "nope" / "libb200splat.so"))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "LOUD" in r.stdout, r.stdout + r.stderr
