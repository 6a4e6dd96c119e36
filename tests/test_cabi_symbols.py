"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/hevc_b200.h declares."""
import ctypes

from hevc_b200 import _cabi


def test_library_exports_every_declared_symbol():
    _cabi.build()
    names = _cabi.declared_symbols()
    assert len(names) >= 20
    L = ctypes.CDLL(str(_cabi.LIB_PATH))
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.hb_abi_version() == 1


def test_no_device_means_loud_failure():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('device present')
    with pytest.raises(_cabi.BackendUnavailable):
        _cabi.Context(0)
