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


def test_packed_weight_sizes_of_every_pack_mode():
    """ssr_conv3d_packed_size is host arithmetic (no CUDA): floats of the K-major packed copy per pack mode.
    0 / 1: 32-channel TF32 chunks of the forward / data-gradient kernel; 2, 3, 4, 6, 8: the k2n layout (9 tiles of 96 rows);
    5: hi / lo TF32 chunks; 7: TF32 hi chunks + bf16 [w_hi ; w_lo] chunks; 9: bf16 w1 chunks + bf16 w2 chunks (64 channels each)."""
    dll = ctypes.CDLL(_lib.LIB_PATH)
    f = dll.ssr_conv3d_packed_size
    f.restype = ctypes.c_longlong
    row = 27 * 32          # floats per output row of one chunk: 27 taps x 128 bytes

    def up16(n):
        return (n + 15) // 16 * 16
    for cin, cout in ((24, 24), (48, 96), (96, 48), (192, 192), (384, 384), (40, 24)):
        n32, n64 = (cin + 31) // 32, (cin + 63) // 64
        assert f(cin, 0, cout, 0) == n32 * row * up16(cout)
        assert f(cin, 0, cout, 1) == ((cout + 31) // 32) * row * up16(cin)
        assert f(cin, cin, cout, 5) == 2 * n32 * row * up16(cout)
        assert f(cin, cin, cout, 7) == (n32 + (2 * cin + 63) // 64) * row * up16(cout)
        assert f(cin, cin, cout, 9) == 2 * n64 * row * up16(cout)
        assert f(cin, cin, cout, 9) <= f(cin, cin, cout, 7)          # bf16x3 never streams more weight bytes than hybrid
    for mode in (2, 3, 4, 6, 8):
        assert f(24, 0, 24, mode) == 9 * 96 * 32
    # a channel part of a concatenated kernel: Cin2 = (first channel << 12) | channels
    assert f(144, (48 << 12) | 96, 48, 9) == 2 * 2 * row * 48
    # bad pack arguments are reported before any CUDA call
    dll.ssr_last_error.restype = ctypes.c_char_p
    dll.ssr_conv3d_pack_weights.restype = ctypes.c_int
    assert dll.ssr_conv3d_pack_weights(None, None, 24, 0, 24, 10, None) == -1 and b'pack args' in dll.ssr_last_error()


def test_split_k_choice_is_host_arithmetic_and_bounded():
    """ssr_conv3d_fwd_tc_ksplit (the tile model of conv3d_tc_kernel, no CUDA): split-K only on the small deep levels of the
    160^3 network (<= 27000 voxels), never more parts than 8 or than K chunks, and off under SSR_NO_SPLIT_K."""
    import subprocess
    import sys
    dll = ctypes.CDLL(_lib.LIB_PATH)
    f = dll.ssr_conv3d_fwd_tc_ksplit
    f.restype = ctypes.c_int
    small = [(192, 384, 10, 5), (384, 384, 10, 5), (384, 384, 10, 0), (384, 192, 10, 0)]
    for c, co, d, comp in small:
        k = f(c, 0, co, 1, d, d, d, comp)
        assert 2 <= k <= 8, (c, co, d, comp, k)                      # 10^3: <= 64 tiles without it
    for c, co, d, comp in [(96, 96, 40, 5), (48, 48, 80, 5), (24, 48, 80, 4), (24, 24, 160, 0)]:
        assert f(c, 0, co, 1, d, d, d, comp) == 1, (c, co, d)        # enough tiles; y traffic would dominate
    assert f(8, 0, 16, 1, 10, 10, 10, 0) == 1                        # one K chunk: nothing to split
    code = ("import ctypes; from synthsr_b200 import _lib; f = ctypes.CDLL(_lib.LIB_PATH).ssr_conv3d_fwd_tc_ksplit; "
            "f.restype = ctypes.c_int; print(f(384, 0, 384, 1, 10, 10, 10, 5))")
    env = dict(os.environ, SSR_NO_SPLIT_K='1')
    out = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(_lib.__file__)))
    assert out.stdout.strip() == '1', (out.stdout, out.stderr)
