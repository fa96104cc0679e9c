"""What does a TMA load through a TFLOAT32 tensor map do to fp32 data?  Convolve with a centre-tap identity kernel
(weights exactly 1.0) so that the output equals the operand the tensor core saw, and compare bitwise with candidates."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthsr_b200._lib import lib, stream_ptr
d, c = [8, 16, 8], 32
nv = int(np.prod(d))
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn((nv, c), device='cuda', generator=g) * 3
w = torch.zeros((3, 3, 3, c, c), device='cuda'); w[1, 1, 1] = torch.eye(c, device='cuda')
y = torch.empty((nv, c), device='cuda')
wp = torch.empty(lib.ssr_conv3d_packed_size(c, 0, c, 0), device='cuda')
lib.ssr_conv3d_pack_weights(w, wp, c, 0, c, 0, stream_ptr())
lib.ssr_conv3d_fwd_tc(x, c, None, 0, wp, None, y, 1, *d, c, 0, stream_ptr())
torch.cuda.synchronize()
xi, yi = x.view(torch.int32), y.view(torch.int32)
cands = {'truncate (RZ)': xi & ~0x1FFF, 'rna (ties away)': (xi + 0x1000) & ~0x1FFF,
         'rne (ties even)': (xi + 0xFFF + ((xi >> 13) & 1)) & ~0x1FFF, 'unchanged fp32': xi}
for k, v in cands.items():
    print('%-18s matches %.4f%% of elements' % (k, 100.0 * (v == yi).float().mean().item()))
print('low 13 bits zero in output: %.4f%%' % (100.0 * ((yi & 0x1FFF) == 0).float().mean().item()))
ex = (yi != cands['rna (ties away)']).nonzero()[:5]
for i, j in ex.tolist():
    print('  x=%08x y=%08x rna=%08x' % (xi[i, j].item() & 0xffffffff, yi[i, j].item() & 0xffffffff, cands['rna (ties away)'][i, j].item() & 0xffffffff))
