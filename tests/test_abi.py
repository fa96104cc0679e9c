"""The C-ABI library loads and exports every symbol include/synthsr_b200.h declares (no compute without a GPU)."""
import ctypes
import os

from synthsr_b200 import _lib


def test_header_parses():
    protos = _lib.parse_header()
    assert len(protos) >= 30
    for must in ('ssr_deform_labels_nearest', 'ssr_gmm_bias_minmax', 'ssr_conv3d_fwd_tc', 'ssr_conv3d_wgrad_tc',
                 'ssr_bn_stats', 'ssr_head_loss', 'ssr_adam_flat', 'ssr_last_error'):
        assert must in protos, must
    ret, args = protos['ssr_adam_flat']
    assert ret == 'int' and [a[1] for a in args][:5] == ['p', 'g', 'm', 'v', 'n'] and args[-1][0] == 'void*'


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), 'run python -m synthsr_b200.build'
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.parse_header():
        assert hasattr(dll, name), name
    dll.ssr_abi_version.restype = ctypes.c_int
    assert dll.ssr_abi_version() == 1


def test_argument_errors_are_reported_not_swallowed():
    """bad arguments return a negative code with a message (no CUDA call is made before validation)."""
    dll = ctypes.CDLL(_lib.LIB_PATH)
    dll.ssr_last_error.restype = ctypes.c_char_p
    dll.ssr_resize.restype = ctypes.c_int
    r = dll.ssr_resize(None, None, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, None)
    assert r == -1 and b'invalid argument' in dll.ssr_last_error()
