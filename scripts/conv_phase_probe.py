"""per-CTA phase timing of conv3d_tc_kernel (clock64 stamps of sampled CTAs).  GPU box only."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthsr_b200._lib import lib, stream_ptr
import scripts.profile_conv as pc  # noqa

def probe(name):
    dbg = torch.zeros(128 * 16, dtype=torch.int64, device='cuda')
    lib.ssr_tc_set_debug(dbg)
    pc.run(name, 1)
    torch.cuda.synchronize()
    lib.ssr_tc_set_debug(None)
    d = dbg.cpu().numpy().reshape(128, 16)
    d = d[d[:, 0] > 0]
    if len(d) == 0:
        print(name, 'no samples'); return
    t0 = d[:, 0:1]
    rel = (d[:, :8] - t0)
    names = ['start', 'setup done', 'producer done', 'mma loop start', 'mma loop end', 'acc ready (epi)', 'epi done', 'exit']
    print('== %s: %d sampled CTAs (cycles, median)' % (name, len(d)))
    for i, n in enumerate(names):
        print('   %-18s %8.0f' % (n, np.median(rel[:, i])))
    print('   producer wait(emptyA) %8.0f | mma wait(fullA) %8.0f | mma wait(fullB) %8.0f' % (
        np.median(d[:, 8]), np.median(d[:, 9]), np.median(d[:, 10])))

for n in (sys.argv[1].split(',') if len(sys.argv) > 1 else ['fwd24', 'fwd72', 'dgrad72', 'fwd96', 'wgrad24', 'wgrad72', 'wgrad96']):
    probe(n)
