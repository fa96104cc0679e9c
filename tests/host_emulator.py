"""TEST INFRASTRUCTURE ONLY -- a NumPy stand-in for the generator entry points of libsynthsr_b200.so.

The product generator (synthsr_b200/generator.py) is a host-side orchestration: it decides which kernel runs on which buffer
with which shapes, strides and parameter blocks.  The kernels themselves are validated on the GPU (tests/test_generator_gpu.py);
this module lets the ORCHESTRATION be executed in the CPU suite: every `ssr_*` call the generator makes is carried out here
on the host memory behind the very pointers the generator passes (tensors live on the CPU, staging "device" pointers are
plain host addresses), with the oracle's primitives doing the arithmetic.  The result is compared with the oracle's
end-to-end graph and with the reference's own graph outputs (tests/test_generator_host_orchestration.py).

Nothing under synthsr_b200/ imports this file; a SynthGenerator only accepts a non-CUDA device when its `lib` has been
replaced by an object carrying `host_emulation = True`, which only the tests do (monkeypatch).
Argument conventions follow include/synthsr_b200.h exactly (same names, same order)."""
import ctypes

import numpy as np

from oracle import generator as OG

f32 = np.float32


def view(p, shape, dtype=np.float32):
    """numpy view of `prod(shape)` elements at a tensor's storage or at a raw host address."""
    if p is None:
        return None
    addr = p.data_ptr() if hasattr(p, 'data_ptr') else int(p)
    n = int(np.prod(shape))
    if hasattr(p, 'numel'):                                # a whole tensor was passed: the launch must fit inside it
        assert p.numel() * p.element_size() >= n * np.dtype(dtype).itemsize, \
            'launch touches %d bytes of a %d-byte buffer' % (n * np.dtype(dtype).itemsize, p.numel() * p.element_size())
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def _addr(p):
    return p.data_ptr() if hasattr(p, 'data_ptr') else int(p)


def disjoint(src, n_src, dst, n_dst, what):
    """gather / stencil kernels read neighbours of what other threads write: source and destination must not overlap
    (the emulator works on copies, which would hide such a hazard -- so it is asserted instead).  Sizes in float32."""
    a0, b0 = _addr(src), _addr(dst)
    assert a0 + 4 * n_src <= b0 or b0 + 4 * n_dst <= a0, '%s: source and destination buffers overlap' % what


def _pad3(vol, pads):
    return np.pad(vol, [(int(p), int(p)) for p in pads] + [(0, 0)] * (vol.ndim - 3))


class HostEmulator:
    host_emulation = True

    def __init__(self):
        self.calls = []                                    # (name, args) in launch order, for orchestration asserts

    def __getattr__(self, name):
        raise AttributeError('host emulator has no %s -- the generator called an entry point that is not emulated' % name)

    def _log(self, name, *args):
        self.calls.append((name, args))

    # -- nrn_layers.Resize ---------------------------------------------------------------------------------------------
    def ssr_resize(self, src, dst, B, s0, s1, s2, d0, d1, d2, C, nearest, dst_stride, dst_off, stream):
        self._log('ssr_resize', (s0, s1, s2), (d0, d1, d2), C, nearest, dst_stride, dst_off)
        if dst_stride <= 0:
            dst_stride, dst_off = C, 0
        x = view(src, (B, s0, s1, s2, C)).copy()
        nv = d0 * d1 * d2
        disjoint(src, B * s0 * s1 * s2 * C, dst, B * nv * dst_stride, 'ssr_resize')
        out = view(dst, (B * nv, dst_stride))
        for b in range(B):
            r = OG.resize(x[b], [d0, d1, d2], 'nearest' if nearest else 'linear')
            out[b * nv:(b + 1) * nv, dst_off:dst_off + C] = r.reshape(nv, C)
        return 0

    # -- nrn_layers.VecInt ---------------------------------------------------------------------------------------------
    def ssr_svf_integrate(self, vec, tmp, B, n0, n1, n2, nb_steps, stream):
        self._log('ssr_svf_integrate', (n0, n1, n2), nb_steps)
        v = view(vec, (B, n0, n1, n2, 3))
        for b in range(B):
            v[b] = OG.integrate_vec(v[b].copy(), nb_steps)
        return 0

    # -- pad + field + affine + interpolation + crop + flip (+ swap) -----------------------------------------------------
    def _deform(self, vol, b, aff, fh, n, pads, h, crop_idx, c, flip, method):
        vol = _pad3(vol, pads)
        assert list(vol.shape[:3]) == list(n)
        field = OG.resize(fh[b].copy(), list(n), 'linear') if fh is not None else None
        if aff is not None:
            vol = OG.spatial_transformer(vol, aff[b], field, method)
        elif field is not None:
            vol = OG.transform(vol, field, method)
        if crop_idx is not None:
            vol = OG.random_crop(vol, crop_idx[b], list(c))
        assert list(vol.shape[:3]) == list(c)
        return vol

    def ssr_deform_labels_nearest(self, labels, out, aff, field_half, B, n0, n1, n2, p0, p1, p2, h0, h1, h2, crop_idx,
                                  c0, c1, c2, flip, swap_lut, lut_len, stream):
        self._log('ssr_deform_labels_nearest', (n0, n1, n2), (p0, p1, p2), (h0, h1, h2), (c0, c1, c2))
        lab = view(labels, (B, n0 - 2 * p0, n1 - 2 * p1, n2 - 2 * p2), np.int32)
        A = view(aff, (B, 4, 4))
        fh = view(field_half, (B, h0, h1, h2, 3)) if field_half is not None else None
        ci = view(crop_idx, (B, 3), np.int32)
        fl = view(flip, (B,), np.uint8)
        lut = view(swap_lut, (lut_len,), np.int32) if swap_lut is not None else None
        o = view(out, (B, c0, c1, c2), np.int32)
        for b in range(B):
            v = lab[b].astype(f32)[..., None]
            if A is not None or fh is not None:
                v = self._deform(v, b, A, fh, (n0, n1, n2), (p0, p1, p2), None, ci, (c0, c1, c2), None, 'nearest')
            else:
                v = _pad3(v, (p0, p1, p2))
                if ci is not None:
                    v = OG.random_crop(v, ci[b], [c0, c1, c2])
            v = v[..., 0].astype(np.int32)
            if fl is not None and fl[b]:
                if lut is not None:
                    v = lut[v]
                v = v[::-1]
            o[b] = v
        return 0

    def ssr_warp_linear(self, image, out, aff, field_half, B, n0, n1, n2, p0, p1, p2, h0, h1, h2, crop_idx, c0, c1, c2,
                        flip, stream):
        self._log('ssr_warp_linear', (n0, n1, n2), (p0, p1, p2), (h0, h1, h2), (c0, c1, c2))
        img = view(image, (B, n0 - 2 * p0, n1 - 2 * p1, n2 - 2 * p2)).copy()
        disjoint(image, img.size, out, B * c0 * c1 * c2, 'ssr_warp_linear')
        A = view(aff, (B, 4, 4))
        fh = view(field_half, (B, h0, h1, h2, 3)) if field_half is not None else None
        ci = view(crop_idx, (B, 3), np.int32)
        fl = view(flip, (B,), np.uint8)
        o = view(out, (B, c0, c1, c2))
        for b in range(B):
            v = self._deform(img[b][..., None], b, A, fh, (n0, n1, n2), (p0, p1, p2), None, ci, (c0, c1, c2), None,
                             'linear')[..., 0]
            if fl is not None and fl[b]:
                v = v[::-1]
            o[b] = v
        return 0

    # -- SampleConditionalGMM + BiasFieldCorruption + clip + min/max -----------------------------------------------------
    def ssr_gmm_bias_minmax(self, labels, lut_mean, lut_std, lut_len, noise, seed, stream_id, bias_small, b0, b1, b2,
                            apply_bias, clip_max, out, minmax, B, n0, n1, n2, stream):
        self._log('ssr_gmm_bias_minmax', (n0, n1, n2), (b0, b1, b2), apply_bias)
        lab = view(labels, (B, n0, n1, n2), np.int32)
        lm, ls = view(lut_mean, (B, lut_len)), view(lut_std, (B, lut_len))
        if noise is not None:
            nz = view(noise, (B, n0, n1, n2))
        else:   # throughput mode: the kernel draws Philox normals from (seed, stream_id); any standard normals keyed the same way
            nz = np.random.default_rng([int(seed) & 0xffffffff, int(stream_id) & 0xffffffff]).standard_normal(
                (B, n0, n1, n2)).astype(f32)
        o = view(out, (B, n0, n1, n2))
        mm = view(minmax, (B, 2))                          # the emulator keeps min/max as plain floats in the int32 slots
        small = view(bias_small, (B, b0, b1, b2)) if bias_small is not None else None
        for b in range(B):
            x = ((ls[b][lab[b]] * nz[b]).astype(f32) + lm[b][lab[b]]).astype(f32)
            if small is not None and apply_bias:
                bias = np.exp(OG.resize(small[b][..., None].copy(), [n0, n1, n2], 'linear')[..., 0]).astype(f32)
                x = (bias * x).astype(f32)
            x = np.clip(x, f32(0), f32(clip_max))
            o[b] = x
            mm[b] = [x.min(), x.max()]
        return 0

    def ssr_minmax(self, x, minmax, B, nvox, stream):
        self._log('ssr_minmax', nvox)
        v = view(x, (B, nvox))
        mm = view(minmax, (B, 2))
        for b in range(B):
            mm[b] = [v[b].min(), v[b].max()]
        return 0

    # -- GaussianBlur with optional fused normalisation + gamma ----------------------------------------------------------
    def ssr_blur3d(self, src, dst, kern, k0, k1, k2, minmax, gamma_exp, B, n0, n1, n2, src_stride, src_off, dst_stride,
                   dst_off, stream):
        self._log('ssr_blur3d', (k0, k1, k2), (n0, n1, n2), minmax is not None, gamma_exp is not None, src_stride, dst_stride,
                  dst_off)
        nv = n0 * n1 * n2
        x = view(src, (B * nv, src_stride))[:, src_off].reshape(B, n0, n1, n2).copy()
        disjoint(src, B * nv * src_stride, dst, B * nv * dst_stride, 'ssr_blur3d')
        k = view(kern, (k0, k1, k2)).copy()
        mm = view(minmax, (B, 2))
        ge = view(gamma_exp, (B,))
        o = view(dst, (B * nv, dst_stride))
        for b in range(B):
            v = x[b]
            if mm is not None:
                m, M = mm[b]
                v = np.clip(v, m, M)
                v = ((v - m) / ((M - m).astype(f32) + f32(1e-7)).astype(f32)).astype(f32)
            if ge is not None:
                v = np.power(v, ge[b]).astype(f32)
            v = OG.conv3d_same(v, k) if (k0, k1, k2) != (1, 1, 1) or k[0, 0, 0] != 1 else v
            o[b * nv:(b + 1) * nv, dst_off] = v.reshape(-1)
        return 0

    # -- MimicAcquisition -------------------------------------------------------------------------------------------------
    def ssr_mimic_acquisition(self, src, dst, dist, params, B, n0, n1, n2, o0, o1, o2, dst_stride, dst_off, dist_stride,
                              dist_off, stream):
        self._log('ssr_mimic_acquisition', (n0, n1, n2), (o0, o1, o2), dst_stride, dst_off, dist is not None)
        x = view(src, (B, n0, n1, n2, 1)).copy()
        P = view(params, (B, 9))
        nv = o0 * o1 * o2
        disjoint(src, x.size, dst, B * nv * dst_stride, 'ssr_mimic_acquisition')
        if dist is not None and _addr(dist) != _addr(dst):
            disjoint(src, x.size, dist, B * nv * dist_stride, 'ssr_mimic_acquisition (dist)')
        o = view(dst, (B * nv, dst_stride))
        dd = view(dist, (B * nv, dist_stride)) if dist is not None else None
        inshape = [n0, n1, n2]
        for b in range(B):
            down_zoom, up_zoom, res = P[b, 0:3], P[b, 3:6], P[b, 6:9]
            grid = OG._grid(inshape)
            down_loc = [np.clip((grid[d] / down_zoom[d]).astype(f32), f32(0), f32(inshape[d])) for d in range(3)]
            low = OG.interpn_nearest(x[b], down_loc)
            ugrid = OG._grid([o0, o1, o2])
            up_loc = [(ugrid[d] / up_zoom[d]).astype(f32) for d in range(3)]
            out = OG.interpn_linear(low, up_loc)
            o[b * nv:(b + 1) * nv, dst_off] = out.reshape(-1)
            if dd is not None:
                dsq = None
                for d in range(3):
                    fd = (up_loc[d] - np.floor(up_loc[d])).astype(f32)
                    cd = (np.ceil(up_loc[d]) - up_loc[d]).astype(f32)
                    sq = np.square((np.minimum(fd, cd) * res[d]).astype(f32)).astype(f32)
                    dsq = sq if dsq is None else (dsq + sq).astype(f32)
                dd[b * nv:(b + 1) * nv, dist_off] = np.sqrt(dsq).astype(f32).reshape(-1)
        return 0

    def ssr_copy_strided(self, src, dst, n, src_stride, src_off, dst_stride, dst_off, stream):
        self._log('ssr_copy_strided', n, src_stride, src_off, dst_stride, dst_off)
        s = view(src, (n, src_stride))
        d = view(dst, (n, dst_stride))
        d[:, dst_off] = s[:, src_off].copy()
        return 0

    def ssr_fill_outer3(self, dst, f0, f1, f2, B, n0, n1, n2, dst_stride, dst_off, stream):
        self._log('ssr_fill_outer3', (n0, n1, n2), dst_stride, dst_off, f0 is not None)
        fs = [view(f, (n,), np.float64) if f is not None else np.ones(n) for f, n in zip((f0, f1, f2), (n0, n1, n2))]
        val = (fs[0][:, None, None] * fs[1][None, :, None] * fs[2][None, None, :]).astype(f32).reshape(-1)
        nv = n0 * n1 * n2
        d = view(dst, (B * nv, dst_stride))
        for b in range(B):
            d[b * nv:(b + 1) * nv, dst_off] = val
        return 0
