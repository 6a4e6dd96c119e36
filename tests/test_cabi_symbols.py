"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/hevc_b200.h declares."""
import ctypes

import pytest

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


def test_host_escape_matches_oracle():
    """hb_escape_rbsp is pure host code (usable without a GPU): emulation prevention must equal the oracle's orc_escape on
    random payloads rich in zero runs, on the adversarial cases, and on empty input."""
    import ctypes as C

    import numpy as np

    from hevc_b200 import _cabi
    from oracle import cmodel
    L = _cabi.lib()
    L.hb_escape_rbsp.restype = C.c_size_t
    L.hb_escape_rbsp.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    O = cmodel.lib()
    O.orc_escape.restype = C.c_size_t
    O.orc_escape.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    rng = np.random.default_rng(5)
    cases = [b'', b'\x00', b'\x00\x00', b'\x00\x00\x00', b'\x00\x00\x01', b'\x00\x00\x03', b'\x00\x00\x04', b'\x00\x00\x00\x00\x00\x00',
             b'\x01\x00\x00\x02\x00\x00\x03\x00\x00', bytes(1000)]
    for _ in range(40):
        n = int(rng.integers(1, 5000))
        a = rng.integers(0, 256, n, dtype=np.uint8)
        a[rng.random(n) < 0.6] = 0                              # long zero runs
        a[rng.random(n) < 0.1] = rng.integers(0, 4)
        cases.append(a.tobytes())
    for data in cases:
        src = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data or b'\x00')
        cap = 2 * len(data) + 16
        got, want = (C.c_uint8 * cap)(), (C.c_uint8 * cap)()
        ng = L.hb_escape_rbsp(src, len(data), got, cap)
        nw = O.orc_escape(want, cap, src, len(data))
        assert ng == nw and bytes(got[:ng]) == bytes(want[:nw])
        assert b'\x00\x00\x00' not in bytes(got[:ng]) and b'\x00\x00\x01' not in bytes(got[:ng]) and b'\x00\x00\x02' not in bytes(got[:ng])
    assert L.hb_escape_rbsp(src, len(data), got, 1) == 0       # too small: refused, nothing written past cap


@pytest.mark.parametrize('w,h,depth,hdr', [(1920, 1080, 8, False), (3840, 2160, 10, True), (200, 120, 10, True), (64, 64, 8, False)])
def test_host_parameter_sets_match_oracle(w, h, depth, hdr):
    """hb_param_sets (pure host code, no device): VPS + SPS + PPS must equal the CPU model's headers byte for byte and parse
    back to the requested geometry."""
    import ctypes as C

    from hevc_b200 import _cabi, compliance
    from hevc_b200 import encoder as E
    from hevc_b200.mp4 import split_nals
    from oracle import encoder_model as em
    from tests import enc_common as ec
    p = ec.b200_params(w, h, depth, keyint=60, hdr10=hdr)
    cp = E.to_c_params(p, qp=(24, 26), hash_sei=False)
    L = _cabi.lib()
    L.hb_param_sets.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    buf, n = (C.c_uint8 * 1024)(), C.c_size_t()
    assert L.hb_param_sets(C.byref(cp), buf, 1024, C.byref(n)) == 0
    got = bytes(buf[:n.value])
    model = em.ModelEncoder(ec.model_params(p, 24, 26, False))
    want = model.headers()
    model.close()
    assert got == want
    nals = split_nals(got)
    assert [(x[0] >> 1) & 0x3f for x in nals] == [32, 33, 34]
    sps = compliance.parse_sps(nals[1])
    assert (sps['width'], sps['height'], sps['bit_depth']) == (w, h, depth)
    assert L.hb_param_sets(C.byref(cp), buf, 8, C.byref(n)) != 0          # too small


@pytest.mark.parametrize('seed', range(4))
def test_host_rate_controller_matches_oracle(seed):
    """hb_rc_simulate steps the controller the device kernels run (csrc/enc_dev.cuh) on the host; against oracle/hevc_rc.c on
    random size-estimate sequences (overshoots, starvation, key frames) the chosen QPs must be identical."""
    import ctypes as C

    import numpy as np

    from hevc_b200 import _cabi
    from hevc_b200 import encoder as E
    from oracle import cmodel
    from oracle import encoder_model as em
    from tests import enc_common as ec
    rng = np.random.default_rng(seed)
    p = ec.b200_params(1920, 1080, 8 + 2 * (seed & 1), keyint=30)
    p.vbv_maxrate_kbps, p.vbv_bufsize_kbit = int(rng.integers(500, 30000)), int(rng.integers(600, 40000))
    qp = (int(rng.integers(10, 30)), int(rng.integers(12, 34)))
    cp = E.to_c_params(p, qp=qp, hash_sei=False, rate_control=True)
    n = 400
    is_idr = (np.arange(n) % 30 == 0).astype(np.int32)
    t16 = p.vbv_maxrate_kbps * 1000 * 16 * p.fps_den // p.fps_num
    est = (t16 * rng.lognormal(0.0, 1.2, n) * np.where(is_idr, 6.0, 1.0)).astype(np.int64)
    qps = np.zeros(n, np.int32)
    L = _cabi.lib()
    L.hb_rc_simulate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    assert L.hb_rc_simulate(C.byref(cp), est.ctypes.data, is_idr.ctypes.data, n, qps.ctypes.data) == 0
    O = cmodel.lib()
    mp = ec.model_params(p, qp[0], qp[1], False, rate_control=True)

    class OrcRc(C.Structure):
        _fields_ = [('t16', C.c_longlong), ('b16', C.c_longlong), ('fullness', C.c_longlong), ('have', C.c_int * 2), ('qp_prev', C.c_int * 2),
                    ('est_prev', C.c_longlong * 2), ('poc', C.c_int), ('cascade', C.c_int)]
    rc = OrcRc()
    O.orc_rc_init(C.byref(rc), C.byref(mp))
    O.orc_rc_pick_qp.restype = C.c_int
    O.orc_rc_update.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_longlong]
    want = []
    for i in range(n):
        q = O.orc_rc_pick_qp(C.byref(rc), C.byref(mp), int(is_idr[i]))
        want.append(q)
        O.orc_rc_update(C.byref(rc), int(is_idr[i]), q, int(est[i]))
    assert qps.tolist() == want
    assert max(want) > min(want)                 # the sequences really exercise the controller
