"""CPU tests: the C-ABI library loads and exports every symbol include/plslam_b200.h declares,
and fails loudly (no fallback) when no CUDA device is present."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    txt = (ROOT / "include" / "plslam_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(plf_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(built):
    import plslam_b200 as plf
    lib = plf.load_library()
    syms = declared_symbols()
    assert len(syms) >= 8
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in header but not exported: {missing}"
    assert lib.plf_abi_version() == 1


def test_struct_sizes_match(built):
    """ctypes mirrors of the POD structs have the same size as the C side assumes (defaults round-trip)."""
    import plslam_b200 as plf
    p = plf.default_params()
    assert p.orb_nlevels == 4 and abs(p.min_ratio_12_p - 0.9) < 1e-6 and p.lsd_n_bins == 1024
    assert p.homog_th == 1e-7 and p.max_iters_ref == 10
    l = plf.default_limits()
    assert l.max_batch > 0 and l.max_keypoints >= 2000


def test_no_cpu_fallback(built):
    """Without a GPU plf_create must fail with PLF_ERR_NO_DEVICE, not silently compute on the CPU."""
    import torch
    import plslam_b200 as plf
    if torch.cuda.is_available():
        pytest.skip("GPU present; the failure path is exercised on the CPU box")
    with pytest.raises(plf.PlfError, match="no CUDA device"):
        plf.Frontend(device=0)
