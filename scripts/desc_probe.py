"""Which shifted / overlapping shared-memory operand views does tcgen05.mma read consistently with the TMA swizzle?
GPU box only.  Prints max |error| against the 'logical rows' expectation for each descriptor configuration."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthsr_b200._lib import lib, stream_ptr

def tf32(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)

rng = np.random.default_rng(0)
A = tf32(rng.standard_normal((256, 32)).astype(np.float32))
Bm = tf32(rng.standard_normal((256, 32)).astype(np.float32))
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(Bm).cuda()

def run(swz, N, mn, nk, layout, a, b):
    p = np.array([N, mn, nk, layout, *a, *b], dtype=np.int32)
    d = torch.full((128, N), float('nan'), device='cuda')
    lib.ssr_tc_desc_probe(dA, 256, dB, 256, d, swz, p.ctypes.data, stream_ptr())
    torch.cuda.synchronize()
    return d.cpu().numpy()

def expect_k(N, a_off, a_sbo, b_off, b_sbo):
    r = np.arange(128); ra = a_off // 128 + (r // 8) * (a_sbo // 128) + r % 8
    n = np.arange(N); rb = b_off // 128 + (n // 8) * (b_sbo // 128) + n % 8
    return A[ra].astype(np.float64) @ Bm[rb].astype(np.float64).T

def expect_mn(N, a_off, a_lbo, b_off, b_lbo):
    out = np.zeros((128, N))
    for j in range(4):
        ra = a_off // 128 + j * (a_lbo // 128) + np.arange(8)
        for i in range(N // 32):
            rb = b_off // 128 + i * (b_lbo // 128) + np.arange(8)
            out[j * 32:(j + 1) * 32, i * 32:(i + 1) * 32] = A[ra].astype(np.float64).T @ Bm[rb].astype(np.float64)
    return out

print('== K-major SWIZZLE_128B, A = 128 rows in 16 groups of 8 (stride SBO), K = 32 (4 MMAs, +32 B)')
for a_off, a_sbo in ((0, 1024), (128, 1024), (256, 1024), (896, 1024), (0, 1280), (128, 1280), (256, 1280), (0, 2048)):
    res = []
    for bo in range(8):
        d = run(0, 32, 0, 4, 2, (a_off, 16, a_sbo, bo, 32), (0, 16, 1024, 0, 32))
        res.append(np.nanmax(np.abs(d - expect_k(32, a_off, a_sbo, 0, 1024))))
    print('A start +%4d B  SBO %4d : err by base_offset 0..7 = %s' % (a_off, a_sbo, ' '.join('%.1e' % e for e in res)))
print('== K-major, B shifted (N = 32 rows in 4 groups)')
for b_off, b_sbo in ((128, 1024), (0, 1280), (128, 1280)):
    res = []
    for bo in range(8):
        d = run(0, 32, 0, 4, 2, (0, 16, 1024, 0, 32), (b_off, 16, b_sbo, bo, 32))
        res.append(np.nanmax(np.abs(d - expect_k(32, 0, 1024, b_off, b_sbo))))
    print('B start +%4d B  SBO %4d : err by base_offset 0..7 = %s' % (b_off, b_sbo, ' '.join('%.1e' % e for e in res)))
print('== MN-major SWIZZLE_128B_ATOM_32B (UMMA layout 1, SBO 512), K = 8 rows, M = 4 atoms (LBO), N = 96 = 3 atoms (LBO)')
for a_off, a_lbo, b_off, b_lbo in ((0, 1024, 0, 1024), (0, 1280, 0, 1024), (128, 1280, 0, 1024), (0, 1024, 128, 1024),
                                   (0, 1024, 0, 128), (0, 1024, 0, 1280), (0, 1024, 1280, 128), (0, 1024, 256, 128),
                                   (0, 128, 0, 1024)):
    res = []
    for bo in range(8):
        d = run(1, 96, 1, 1, 1, (a_off, a_lbo, 512, bo if a_off % 1024 or a_lbo % 1024 else 0, 0),
                (b_off, b_lbo, 512, bo if b_off % 1024 or b_lbo % 1024 else 0, 0))
        res.append(np.nanmax(np.abs(d - expect_mn(96, a_off, a_lbo, b_off, b_lbo))))
    print('A +%4d LBO %4d | B +%4d LBO %4d : err by base_offset 0..7 = %s' % (a_off, a_lbo, b_off, b_lbo, ' '.join('%.1e' % e for e in res)))
