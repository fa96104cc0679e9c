"""diagnostics of the work-in-progress conv3d_tc_up_k2n_kernel against conv3d_tc_up_kernel<1> (per parity class / plane)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from synthsr_b200._lib import lib, stream_ptr
rng = np.random.default_rng(1)
for (dl, cu) in [([1, 8, 14], 16), ([1, 8, 14], 32), ([1, 8, 14], 48), ([2, 8, 14], 32), ([4, 16, 16], 48), ([16, 16, 16], 48)]:
    cs, co = 24, 24
    df = [2 * v for v in dl]
    nl, nf = int(np.prod(dl)), int(np.prod(df))
    st = stream_ptr()
    w = torch.from_numpy((rng.normal(size=(3, 3, 3, cs + cu, co)) / np.sqrt(27 * (cs + cu))).astype(np.float32)).cuda()
    low = torch.from_numpy(rng.normal(size=(nl, cu)).astype(np.float32)).cuda()
    wskip, weff = torch.empty(27 * cs * co, device='cuda'), torch.empty(8 * 27 * cu * co, device='cuda')
    lib.ssr_conv3d_up_weights(w, cs, cu, co, wskip, weff, st)
    nfw = lib.ssr_conv3d_packed_size(cu, 0, co, 0)
    fwd8 = torch.empty(8 * nfw, device='cuda')
    for par in range(8):
        lib.ssr_conv3d_pack_weights(weff[par * 27 * cu * co:(par + 1) * 27 * cu * co], fwd8[par * nfw:(par + 1) * nfw], cu, 0, co, 0, st)
    y = torch.full((nf, co), float('nan'), device='cuda')
    lib.ssr_conv3d_fwd_tc_up(low, cu, fwd8, y, 1, *dl, co, st)
    wpk = torch.empty(4 * 8 * 96 * 32, device='cuda')
    lib.ssr_conv3d_pack_up_k2n(weff, wpk, cu, st)
    yk = torch.full((nf, co), float('nan'), device='cuda')
    try:
        lib.ssr_conv3d_fwd_tc_up_k2n(low, cu, wpk, yk, 1, *dl, co, st)
        torch.cuda.synchronize()
    except Exception as e:
        print(dl, cu, 'FAILED', e); break
    a, b = y.view(*df, co).cpu().numpy(), yk.view(*df, co).cpu().numpy()
    print(dl, cu, 'nan', int(np.isnan(b).sum()), 'max err / max', float(np.nanmax(np.abs(a - b)) / np.abs(a).max()))
    for par in range(8):
        p0, p1, p2 = (par >> 2) & 1, (par >> 1) & 1, par & 1
        e = np.abs(a[p0::2, p1::2, p2::2] - b[p0::2, p1::2, p2::2])
        print('   class', (p0, p1, p2), 'err %.2e' % np.nanmax(e), ' by low-res plane:', ['%.1e' % np.nanmax(e[z]) for z in range(min(dl[0], 6))],
              ' by channel block:', ['%.1e' % np.nanmax(e[..., c:c + 8]) for c in (0, 8, 16)])
