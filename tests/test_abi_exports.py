"""CPU: the C-ABI library loads without a GPU and exports every symbol include/zs_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "zs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zs_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_surface():
    names = _declared()
    for must in ("zs_normal_sample", "zs_normal_logprob_fwd", "zs_normal_logprob_bwd", "zs_bernoulli_logpmf_fwd",
                 "zs_bernoulli_logpmf_bwd", "zs_bernoulli_sample", "zs_categorical_logpmf_fwd", "zs_iw_objective",
                 "zs_iw_bernoulli_fused", "zs_sgld_step", "zs_sghmc_post", "zs_iw_step_host"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from zhusuan import _backend
    if not os.path.exists(_backend.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_backend.LIB_PATH)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    lib.zs_abi_version.restype = ctypes.c_int
    assert lib.zs_abi_version() == _backend.ABI_VERSION == 3
    lib.zs_strerror.restype = ctypes.c_char_p
    assert lib.zs_strerror(0) == b"ok" and b"dtype" in lib.zs_strerror(-2)


def test_binding_signatures_cover_header():
    from zhusuan import _backend
    _backend.load()
    assert sorted(_backend.EXPORTS) == _declared()


def test_no_cpu_fallback():
    import torch
    from zhusuan import _backend
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_backend.BackendError):
        _backend.require_cuda()
