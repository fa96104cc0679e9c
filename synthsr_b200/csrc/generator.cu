// On-the-fly synthetic-scan generator kernels (the labels_to_image_model hot path), sm_100a.
//
// Replaces the TF graph built by SynthSR/labels_to_image_model.py:32-266 and the Keras layers of
// ext/lab2im/layers.py / ext/neuron/{layers,utils}.py it instantiates.  Every kernel is HBM/L2-bound
// gather / stencil / elementwise work; the design goal is one pass over each full-size volume.
//
// This translation unit is compiled with -fmad=false: the coordinate arithmetic that decides which label a
// voxel receives must round exactly like the float32 op-by-op evaluation of the reference graph
// (ext/neuron/utils.py:67-122, 147-154, 267-286, 316-317), so no multiply-add contraction is allowed.
// Volumes are [B][d0][d1][d2](,C) with d2 (the reference's last spatial axis) contiguous.
#include "common.cuh"
#include <cuda.h>
#include <math_constants.h>
#include <mutex>

namespace {

// ---------------------------------------------------------------------------------------------------------
// trilinear sampling with the reference's clamping / weight conventions (ext/neuron/utils.py:67-110)
// ---------------------------------------------------------------------------------------------------------
struct Axis {
  int i0, i1;
  float w0, w1;  // w0 multiplies corner i0 (= diff_loc1), w1 multiplies corner i1 (= 1 - diff_loc1)
};

__device__ __forceinline__ Axis lin_axis(float loc, int n) {
  const float mx = (float)(n - 1);
  const float l0 = floorf(loc);
  const float cl = fminf(fmaxf(loc, 0.f), mx);
  const float l0c = fminf(fmaxf(l0, 0.f), mx);
  const float l1 = fminf(fmaxf(__fadd_rn(l0c, 1.f), 0.f), mx);
  Axis a;
  a.w0 = __fsub_rn(l1, cl);
  a.w1 = __fsub_rn(1.f, a.w0);
  a.i0 = (int)l0c;
  a.i1 = (int)l1;
  return a;
}

// resize() sampling location for output index j: loc = g + (g / zoom - g)   (ext/neuron/utils.py:150,317)
__device__ __forceinline__ float resize_loc(int j, float zoom) {
  const float g = (float)j;
  return __fadd_rn(g, __fsub_rn(__fdiv_rn(g, zoom), g));
}

// sample C (<=3) channels of vol [n0][n1][n2][C] at (l0,l1,l2); corner order = itertools.product([0,1],repeat=3)
template <int C>
__device__ __forceinline__ void sample_linear(const float* __restrict__ vol, int n0, int n1, int n2, float l0,
                                              float l1, float l2, float* out) {
  const Axis a0 = lin_axis(l0, n0), a1 = lin_axis(l1, n1), a2 = lin_axis(l2, n2);
  const int i0[2] = {a0.i0, a0.i1}, i1[2] = {a1.i0, a1.i1}, i2[2] = {a2.i0, a2.i1};
  const float w0[2] = {a0.w0, a0.w1}, w1[2] = {a1.w0, a1.w1}, w2[2] = {a2.w0, a2.w1};
  bool first = true;
#pragma unroll
  for (int c0 = 0; c0 < 2; ++c0)
#pragma unroll
    for (int c1 = 0; c1 < 2; ++c1)
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        const float wt = __fmul_rn(__fmul_rn(w0[c0], w1[c1]), w2[c2]);
        const long long idx = (((long long)i0[c0] * n1 + i1[c1]) * n2 + i2[c2]) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float t = __fmul_rn(wt, __ldg(vol + idx + c));
          out[c] = first ? t : __fadd_rn(out[c], t);
        }
        first = false;
      }
}

// ---------------------------------------------------------------------------------------------------------
// K1: generic resize (ext/neuron/layers.py:361-394 -> utils.resize :127-154), linear or nearest, C channels
// ---------------------------------------------------------------------------------------------------------
template <int C, bool NEAREST>
__global__ void resize_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int s0, int s1, int s2,
                              int d0, int d1, int d2, float z0, float z1, float z2, int dst_stride, int dst_off) {
  const long long nvox = (long long)d0 * d1 * d2;
  const long long total = nvox * B;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(t / nvox);
    long long v = t - (long long)b * nvox;
    const int k = (int)(v % d2);
    v /= d2;
    const int j = (int)(v % d1);
    const int i = (int)(v / d1);
    const float* sv = src + (long long)b * s0 * s1 * s2 * C;
    const float l0 = resize_loc(i, z0), l1 = resize_loc(j, z1), l2 = resize_loc(k, z2);
    float out[C];
    if (NEAREST) {
      int r0 = (int)rintf(l0), r1 = (int)rintf(l1), r2 = (int)rintf(l2);
      r0 = min(max(r0, 0), s0 - 1);
      r1 = min(max(r1, 0), s1 - 1);
      r2 = min(max(r2, 0), s2 - 1);
      const long long idx = (((long long)r0 * s1 + r1) * s2 + r2) * C;
#pragma unroll
      for (int c = 0; c < C; ++c) out[c] = sv[idx + c];
    } else {
      sample_linear<C>(sv, s0, s1, s2, l0, l1, l2, out);
    }
    float* o = dst + ((long long)b * nvox + (t - (long long)b * nvox)) * dst_stride + dst_off;
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = out[c];
  }
}

// ---------------------------------------------------------------------------------------------------------
// MimicAcquisition (ext/lab2im/layers.py:921-987), both resamplings fused: the low-resolution acquisition is never
// materialised.  The reference keeps it in a full-size tensor V[i] = vol[nearest(clip(i / down_zoom, 0, n))] and then
// samples V linearly at j / up_zoom; here each of the 8 corners of the linear interpolation evaluates V on the fly.
// params [B][9] = down_zoom[3] | up_zoom[3] | acquisition resolution[3] (float32, computed on the host exactly as the
// reference does, :935-938).  dist = distance (mm) to the nearest acquired voxel (:973-986), optional.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int mimic_low_index(int i, float down_zoom, int n) {
  float l = __fdiv_rn((float)i, down_zoom);
  l = fminf(fmaxf(l, 0.f), (float)n);                       // K.clip(down_loc, 0, inshape)
  int r = (int)rintf(l);                                    // tf.round: half to even
  return min(max(r, 0), n - 1);
}

__global__ void mimic_acquisition_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                         float* __restrict__ dist, const float* __restrict__ params, int B, int n0,
                                         int n1, int n2, int o0, int o1, int o2, int dst_stride, int dst_off,
                                         int dist_stride, int dist_off) {
  const long long nout = (long long)o0 * o1 * o2, total = nout * B;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(t / nout);
    long long v = t - (long long)b * nout;
    const int k = (int)(v % o2);
    v /= o2;
    const int j = (int)(v % o1);
    const int i = (int)(v / o1);
    const float* P = params + b * 9;
    const float* sv = src + (long long)b * n0 * n1 * n2;
    const float u0 = __fdiv_rn((float)i, P[3]), u1 = __fdiv_rn((float)j, P[4]), u2 = __fdiv_rn((float)k, P[5]);
    const Axis a0 = lin_axis(u0, n0), a1 = lin_axis(u1, n1), a2 = lin_axis(u2, n2);
    const int i0[2] = {mimic_low_index(a0.i0, P[0], n0), mimic_low_index(a0.i1, P[0], n0)};
    const int i1[2] = {mimic_low_index(a1.i0, P[1], n1), mimic_low_index(a1.i1, P[1], n1)};
    const int i2[2] = {mimic_low_index(a2.i0, P[2], n2), mimic_low_index(a2.i1, P[2], n2)};
    const float w0[2] = {a0.w0, a0.w1}, w1[2] = {a1.w0, a1.w1}, w2[2] = {a2.w0, a2.w1};
    float out = 0.f;
    bool first = true;
#pragma unroll
    for (int c0 = 0; c0 < 2; ++c0)
#pragma unroll
      for (int c1 = 0; c1 < 2; ++c1)
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const float wt = __fmul_rn(__fmul_rn(w0[c0], w1[c1]), w2[c2]);
          const float tv = __fmul_rn(wt, __ldg(sv + ((long long)i0[c0] * n1 + i1[c1]) * n2 + i2[c2]));
          out = first ? tv : __fadd_rn(out, tv);
          first = false;
        }
    const long long ov = (long long)b * nout + (t - (long long)b * nout);
    dst[ov * dst_stride + dst_off] = out;
    if (dist) {
      const float uu[3] = {u0, u1, u2};
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float f = __fsub_rn(uu[d], floorf(uu[d])), c = __fsub_rn(ceilf(uu[d]), uu[d]);
        const float m = __fmul_rn(fminf(f, c), P[6 + d]);
        const float sq = __fmul_rn(m, m);
        acc = d == 0 ? sq : __fadd_rn(acc, sq);
      }
      dist[ov * dist_stride + dist_off] = __fsqrt_rn(acc);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// K2: one scaling-and-squaring step  v_out = v + interp_linear(v, idx + v)   (ext/neuron/utils.py:366-369)
// ---------------------------------------------------------------------------------------------------------
__global__ void svf_step_kernel(const float* __restrict__ vin, float* __restrict__ vout, int B, int n0, int n1, int n2,
                                float prescale) {
  const long long nvox = (long long)n0 * n1 * n2;
  const long long total = nvox * B;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(t / nvox);
    long long v = t - (long long)b * nvox;
    const int k = (int)(v % n2);
    v /= n2;
    const int j = (int)(v % n1);
    const int i = (int)(v / n1);
    const float* base = vin + (long long)b * nvox * 3;
    const float* p = vin + t * 3;
    if (prescale != 0.f) {  // only the "vec / 2**nb_steps" pre-pass (exact: power of two)
      vout[t * 3 + 0] = __fmul_rn(p[0], prescale);
      vout[t * 3 + 1] = __fmul_rn(p[1], prescale);
      vout[t * 3 + 2] = __fmul_rn(p[2], prescale);
      continue;
    }
    const float u0 = p[0], u1 = p[1], u2 = p[2];
    float s[3];
    sample_linear<3>(base, n0, n1, n2, __fadd_rn((float)i, u0), __fadd_rn((float)j, u1), __fadd_rn((float)k, u2), s);
    vout[t * 3 + 0] = __fadd_rn(u0, s[0]);
    vout[t * 3 + 1] = __fadd_rn(u1, s[1]);
    vout[t * 3 + 2] = __fadd_rn(u2, s[2]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// K3: fused spatial deformation:  (half-res integrated SVF --linear resize--> full res)  o  affine  o  crop
//     o flip  o  L/R label swap, nearest (labels, bit exact) or linear (real image) sampling.
//     ext/lab2im/layers.py:196-211 (Resize + SpatialTransformer), :252-270 (RandomCrop), :391-427 (RandomFlip),
//     ext/neuron/utils.py:222-286 (affine o field), :112-122 (nearest), ext/lab2im/layers.py:1754 (pad).
// ---------------------------------------------------------------------------------------------------------
struct DeformParams {
  int B;
  int n0, n1, n2;     // (padded) grid on which the deformation is defined
  int p0, p1, p2;     // padding margin (source volume is [n0-2p0][n1-2p1][n2-2p2])
  int h0, h1, h2;     // half-res field grid (0 => no elastic field)
  int c0, c1, c2;     // crop (= output) shape
  float z0, z1, z2;   // float32(full / half) zoom factors
  float m0, m1, m2;   // float32((n-1)/2) centres
  int has_aff;
  int lut_len;
};

template <bool NEAREST, typename T>
__global__ void deform_kernel(const T* __restrict__ src, T* __restrict__ dst, const float* __restrict__ aff,
                              const float* __restrict__ field, const int* __restrict__ crop_idx,
                              const unsigned char* __restrict__ flip, const int* __restrict__ swap_lut,
                              DeformParams P) {
  const long long nout = (long long)P.c0 * P.c1 * P.c2;
  const long long total = nout * P.B;
  const int s0 = P.n0 - 2 * P.p0, s1 = P.n1 - 2 * P.p1, s2 = P.n2 - 2 * P.p2;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(t / nout);
    long long v = t - (long long)b * nout;
    const int ok = (int)(v % P.c2);
    v /= P.c2;
    const int oj = (int)(v % P.c1);
    int oi = (int)(v / P.c1);
    const bool fl = flip != nullptr && flip[b] != 0;
    if (fl) oi = P.c0 - 1 - oi;                       // tf.reverse along axis 0 after the crop
    const int i = oi + (crop_idx ? crop_idx[b * 3 + 0] : 0);
    const int j = oj + (crop_idx ? crop_idx[b * 3 + 1] : 0);
    const int k = ok + (crop_idx ? crop_idx[b * 3 + 2] : 0);
    const float g0 = (float)i, g1 = (float)j, g2 = (float)k;
    float u[3] = {0.f, 0.f, 0.f};
    if (P.h0 > 0) {                                    // Resize(linear) of the integrated half-res field
      const float* fb = field + (long long)b * P.h0 * P.h1 * P.h2 * 3;
      sample_linear<3>(fb, P.h0, P.h1, P.h2, resize_loc(i, P.z0), resize_loc(j, P.z1), resize_loc(k, P.z2), u);
    }
    float l0, l1, l2;
    if (P.has_aff) {
      const float* A = aff + b * 16;
      const float mc0 = __fsub_rn(g0, P.m0), mc1 = __fsub_rn(g1, P.m1), mc2 = __fsub_rn(g2, P.m2);
      float q0 = mc0, q1 = mc1, q2 = mc2;
      if (P.h0 > 0) { q0 = __fadd_rn(mc0, u[0]); q1 = __fadd_rn(mc1, u[1]); q2 = __fadd_rn(mc2, u[2]); }
      float r[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        float a = __fmul_rn(A[d * 4 + 0], q0);
        a = __fadd_rn(a, __fmul_rn(A[d * 4 + 1], q1));
        a = __fadd_rn(a, __fmul_rn(A[d * 4 + 2], q2));
        r[d] = __fadd_rn(a, A[d * 4 + 3]);
      }
      l0 = __fadd_rn(g0, __fsub_rn(r[0], mc0));
      l1 = __fadd_rn(g1, __fsub_rn(r[1], mc1));
      l2 = __fadd_rn(g2, __fsub_rn(r[2], mc2));
    } else {
      l0 = __fadd_rn(g0, u[0]); l1 = __fadd_rn(g1, u[1]); l2 = __fadd_rn(g2, u[2]);
    }
    const T* sb = src + (long long)b * s0 * s1 * s2;
    T outv;
    if (NEAREST) {
      int r0 = (int)rintf(l0), r1 = (int)rintf(l1), r2 = (int)rintf(l2);   // tf.round: half to even
      r0 = min(max(r0, 0), P.n0 - 1) - P.p0;
      r1 = min(max(r1, 0), P.n1 - 1) - P.p1;
      r2 = min(max(r2, 0), P.n2 - 1) - P.p2;
      T val = (T)0;
      if (r0 >= 0 && r0 < s0 && r1 >= 0 && r1 < s1 && r2 >= 0 && r2 < s2)
        val = sb[((long long)r0 * s1 + r1) * s2 + r2];
      outv = val;
    } else {
      const Axis a0 = lin_axis(l0, P.n0), a1 = lin_axis(l1, P.n1), a2 = lin_axis(l2, P.n2);
      const int i0[2] = {a0.i0 - P.p0, a0.i1 - P.p0}, i1[2] = {a1.i0 - P.p1, a1.i1 - P.p1},
                i2[2] = {a2.i0 - P.p2, a2.i1 - P.p2};
      const float w0[2] = {a0.w0, a0.w1}, w1[2] = {a1.w0, a1.w1}, w2[2] = {a2.w0, a2.w1};
      float acc = 0.f;
      bool first = true;
#pragma unroll
      for (int c0 = 0; c0 < 2; ++c0)
#pragma unroll
        for (int c1 = 0; c1 < 2; ++c1)
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {
            const float wt = __fmul_rn(__fmul_rn(w0[c0], w1[c1]), w2[c2]);
            float val = 0.f;
            if (i0[c0] >= 0 && i0[c0] < s0 && i1[c1] >= 0 && i1[c1] < s1 && i2[c2] >= 0 && i2[c2] < s2)
              val = (float)sb[((long long)i0[c0] * s1 + i1[c1]) * s2 + i2[c2]];
            const float tt = __fmul_rn(wt, val);
            acc = first ? tt : __fadd_rn(acc, tt);
            first = false;
          }
      outv = (T)acc;
    }
    if (NEAREST && fl && swap_lut != nullptr) {
      const int li = (int)outv;
      if (li >= 0 && li < P.lut_len) outv = (T)swap_lut[li];
    }
    dst[t] = outv;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller (throughput mode: GMM noise generated on the fly instead of injected)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// 4 standard normals for counter `ctr`
__device__ __forceinline__ void philox_normal4(unsigned long long seed, unsigned long long stream,
                                               unsigned long long ctr, float* z) {
  uint32_t r[4];
  philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)stream, (uint32_t)(stream >> 32), (uint32_t)seed,
                (uint32_t)(seed >> 32), r);
  const float r0 = sqrtf(-2.f * __logf(u01(r[0]))), r1 = sqrtf(-2.f * __logf(u01(r[2])));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u01(r[1]), &s0, &c0);
  __sincosf(6.283185307179586f * u01(r[3]), &s1, &c1);
  z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
}

__global__ void philox_normal_kernel(float* __restrict__ out, long long n, unsigned long long seed,
                                     unsigned long long stream) {
  const long long ng = (n + 3) / 4;
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < ng; g += (long long)gridDim.x * blockDim.x) {
    float z[4];
    philox_normal4(seed, stream, (unsigned long long)g, z);
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (g * 4 + e < n) out[g * 4 + e] = z[e];
  }
}

// ordered-uint encoding so that atomicMin/atomicMax on uint32 order like floats
__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  const uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}

__global__ void minmax_init_kernel(uint32_t* mm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mm[i] = (i & 1) ? 0u : 0xffffffffu;   // [min, max] pairs
}

// ---------------------------------------------------------------------------------------------------------
// K4: SampleConditionalGMM + BiasFieldCorruption + clip + global min/max   (one synthetic channel)
//     ext/lab2im/layers.py:480-498, 1067-1097, 1214-1215, 1230-1231
// ---------------------------------------------------------------------------------------------------------
struct GmmParams {
  int B;
  int n0, n1, n2;
  int lut_len;
  int b0, b1, b2;     // small bias grid (0 => no bias field)
  float z0, z1, z2;   // float32(full / small)
  int apply_bias;
  float clip_max;     // <= 0: no clipping
  unsigned long long seed, stream;
};

__global__ void gmm_bias_kernel(const int* __restrict__ labels, const float* __restrict__ lut_mean,
                                const float* __restrict__ lut_std, const float* __restrict__ noise,
                                const float* __restrict__ bias_small, float* __restrict__ out,
                                uint32_t* __restrict__ minmax, GmmParams P) {
  const long long nvox = (long long)P.n0 * P.n1 * P.n2;
  const long long ngrp = (nvox + 3) / 4;
  const long long total = ngrp * P.B;
  float lmin = CUDART_INF_F, lmax = -CUDART_INF_F;
  int cur_b = -1;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(t / ngrp);
    const long long g = t - (long long)b * ngrp;
    if (b != cur_b && cur_b >= 0) {   // flush per-batch-item extrema (grid-stride may cross items)
      atomicMin(minmax + 2 * cur_b, f2ord(lmin));
      atomicMax(minmax + 2 * cur_b + 1, f2ord(lmax));
      lmin = CUDART_INF_F; lmax = -CUDART_INF_F;
    }
    cur_b = b;
    float z[4];
    if (noise == nullptr) philox_normal4(P.seed, P.stream + (unsigned long long)b, (unsigned long long)g, z);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long v = g * 4 + e;
      if (v >= nvox) break;
      const long long gi = (long long)b * nvox + v;
      const int lab = labels[gi];
      float mu = 0.f, sd = 0.f;
      if (lab >= 0 && lab < P.lut_len) { mu = lut_mean[b * P.lut_len + lab]; sd = lut_std[b * P.lut_len + lab]; }
      const float nz = noise ? noise[gi] : z[e];
      float x = __fadd_rn(__fmul_rn(sd, nz), mu);
      if (P.b0 > 0) {
        long long vv = v;
        const int k = (int)(vv % P.n2); vv /= P.n2;
        const int j = (int)(vv % P.n1);
        const int i = (int)(vv / P.n1);
        float bf;
        sample_linear<1>(bias_small + (long long)b * P.b0 * P.b1 * P.b2, P.b0, P.b1, P.b2, resize_loc(i, P.z0),
                         resize_loc(j, P.z1), resize_loc(k, P.z2), &bf);
        if (P.apply_bias) x = __fmul_rn(expf(bf), x);
      }
      if (P.clip_max > 0.f) x = fminf(fmaxf(x, 0.f), P.clip_max);
      out[gi] = x;
      lmin = fminf(lmin, x); lmax = fmaxf(lmax, x);
    }
  }
  // warp-shuffle reduction, one atomic per warp (every lane reaches this point: no early exits above)
  {
    const unsigned full = 0xffffffffu;
    const int bref = __shfl_sync(full, cur_b, 0);
    const bool uniform = __all_sync(full, cur_b == bref);
    if (uniform) {
      if (bref >= 0) {
        for (int o = 16; o > 0; o >>= 1) {
          lmin = fminf(lmin, __shfl_xor_sync(full, lmin, o));
          lmax = fmaxf(lmax, __shfl_xor_sync(full, lmax, o));
        }
        if ((threadIdx.x & 31) == 0) {
          atomicMin(minmax + 2 * bref, f2ord(lmin));
          atomicMax(minmax + 2 * bref + 1, f2ord(lmax));
        }
      }
    } else if (cur_b >= 0) {
      atomicMin(minmax + 2 * cur_b, f2ord(lmin));
      atomicMax(minmax + 2 * cur_b + 1, f2ord(lmax));
    }
  }
}

// plain min/max of a float volume per batch item (real-image target: IntensityAugmentation(normalise=True))
__global__ void minmax_kernel(const float* __restrict__ x, long long nvox, uint32_t* __restrict__ minmax) {
  const int b = blockIdx.y;
  float lmin = CUDART_INF_F, lmax = -CUDART_INF_F;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvox; v += (long long)gridDim.x * blockDim.x) {
    const float f = x[(long long)b * nvox + v];
    lmin = fminf(lmin, f); lmax = fmaxf(lmax, f);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
    lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(minmax + 2 * b, f2ord(lmin));
    atomicMax(minmax + 2 * b + 1, f2ord(lmax));
  }
}

// ---------------------------------------------------------------------------------------------------------
// K5/K6: 3-D stencil with zero padding ('SAME'), optional fused min-max normalisation + gamma on the loads
//     ext/lab2im/layers.py:1235-1242 (normalise, gamma) + :748-759 (tf.nn.conv3d 'SAME' per channel)
//     shared-memory halo tile: block = 8 x 8 x 32 outputs.
// ---------------------------------------------------------------------------------------------------------
constexpr int BT0 = 8, BT1 = 8, BT2 = 32;

struct BlurParams {
  int B, n0, n1, n2;
  int k0, k1, k2;        // kernel window (odd)
  int src_stride, src_off, dst_stride, dst_off;
  int normalise;         // 1: (x-m)/(M-m+1e-7) from minmax
  int use_gamma;         // 1: pow(x, gamma_exp[b])
};

__global__ void blur3d_kernel(const float* __restrict__ src, float* __restrict__ dst, const float* __restrict__ kern,
                              const uint32_t* __restrict__ minmax, const float* __restrict__ gamma_exp, BlurParams P) {
  extern __shared__ float tile[];
  const int r0 = P.k0 / 2, r1 = P.k1 / 2, r2 = P.k2 / 2;
  const int t0 = BT0 + 2 * r0, t1 = BT1 + 2 * r1, t2 = BT2 + 2 * r2;
  const int nb2 = (P.n2 + BT2 - 1) / BT2, nb1 = (P.n1 + BT1 - 1) / BT1, nb0 = (P.n0 + BT0 - 1) / BT0;
  long long blk = blockIdx.x;
  const int bz = (int)(blk % nb2); blk /= nb2;
  const int by = (int)(blk % nb1); blk /= nb1;
  const int bx = (int)(blk % nb0);
  const int b = (int)(blk / nb0);
  const int o0 = bx * BT0 - r0, o1 = by * BT1 - r1, o2 = bz * BT2 - r2;
  const long long nvox = (long long)P.n0 * P.n1 * P.n2;
  float m = 0.f, inv_den = 1.f, ge = 1.f;
  if (P.normalise) {
    m = ord2f(minmax[2 * b]);
    const float M = ord2f(minmax[2 * b + 1]);
    inv_den = (M - m) + 1e-7f;
    if (P.use_gamma) ge = gamma_exp[b];
  }
  const int tsz = t0 * t1 * t2;
  for (int e = threadIdx.x; e < tsz; e += blockDim.x) {
    const int c = e % t2, bb = (e / t2) % t1, a = e / (t2 * t1);
    const int i = o0 + a, j = o1 + bb, k = o2 + c;
    float val = 0.f;
    if (i >= 0 && i < P.n0 && j >= 0 && j < P.n1 && k >= 0 && k < P.n2) {
      val = src[((long long)b * nvox + ((long long)i * P.n1 + j) * P.n2 + k) * P.src_stride + P.src_off];
      if (P.normalise) {
        val = (val - m) / inv_den;
        if (P.use_gamma) val = powf(val, ge);
      }
    }
    tile[e] = val;
  }
  __syncthreads();
  // 256 threads: thread -> (a in 0..7, c in 0..31), loops over the 8 values of the middle axis
  const int c = threadIdx.x % BT2, a = threadIdx.x / BT2;
  const int i = bx * BT0 + a, k = bz * BT2 + c;
  if (i >= P.n0 || k >= P.n2) return;
  for (int bb = 0; bb < BT1; ++bb) {
    const int j = by * BT1 + bb;
    if (j >= P.n1) break;
    float acc = 0.f;
    for (int x = 0; x < P.k0; ++x)
      for (int y = 0; y < P.k1; ++y) {
        const float* row = tile + ((a + x) * t1 + (bb + y)) * t2 + c;
        const float* kr = kern + (x * P.k1 + y) * P.k2;
        for (int z = 0; z < P.k2; ++z) acc += kr[z] * row[z];
      }
    dst[((long long)b * nvox + ((long long)i * P.n1 + j) * P.n2 + k) * P.dst_stride + P.dst_off] = acc;
  }
}

// 3x3x3 window (sigma 0.5 blur of the target and the acquisition blur at the label resolution: both blurs of the default
// training configuration): same tile, same per-output accumulation order (k0, k1, k2) as blur3d_kernel, but the 27 taps
// live in registers, a thread keeps its 3 x 10 x 3 input window in registers for its 8 outputs (90 shared-memory loads
// instead of 216 + 216 global tap loads), and the halo load walks rows without per-element divisions.
__global__ void __launch_bounds__(256, 2)
blur3d_333_kernel(const float* __restrict__ src, float* __restrict__ dst, const float* __restrict__ kern,
                  const uint32_t* __restrict__ minmax, const float* __restrict__ gamma_exp, BlurParams P) {
  constexpr int T0 = BT0 + 2, T1 = BT1 + 2, T2 = BT2 + 2;
  __shared__ float tile[T0 * T1 * T2];
  const int nb2 = (P.n2 + BT2 - 1) / BT2, nb1 = (P.n1 + BT1 - 1) / BT1, nb0 = (P.n0 + BT0 - 1) / BT0;
  long long blk = blockIdx.x;
  const int bz = (int)(blk % nb2); blk /= nb2;
  const int by = (int)(blk % nb1); blk /= nb1;
  const int bx = (int)(blk % nb0);
  const int b = (int)(blk / nb0);
  const int o0 = bx * BT0 - 1, o1 = by * BT1 - 1, o2 = bz * BT2 - 1;
  const long long nvox = (long long)P.n0 * P.n1 * P.n2;
  float m = 0.f, inv_den = 1.f, ge = 1.f;
  if (P.normalise) {
    m = ord2f(minmax[2 * b]);
    const float M = ord2f(minmax[2 * b + 1]);
    inv_den = (M - m) + 1e-7f;
    if (P.use_gamma) ge = gamma_exp[b];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = warp; row < T0 * T1; row += 8) {
    const int a = row / T1, bb = row - a * T1;
    const int i = o0 + a, j = o1 + bb;
    const bool row_ok = i >= 0 && i < P.n0 && j >= 0 && j < P.n1;
    const long long base = (long long)b * nvox + ((long long)i * P.n1 + j) * P.n2;
    for (int c = lane; c < T2; c += 32) {
      const int k = o2 + c;
      float val = 0.f;
      if (row_ok && k >= 0 && k < P.n2) {
        val = src[(base + k) * P.src_stride + P.src_off];
        if (P.normalise) {
          val = (val - m) / inv_den;
          if (P.use_gamma) val = powf(val, ge);
        }
      }
      tile[row * T2 + c] = val;
    }
  }
  float kr[27];
#pragma unroll
  for (int t = 0; t < 27; ++t) kr[t] = kern[t];
  __syncthreads();
  const int c = lane, a = warp;
  const int i = bx * BT0 + a, k = bz * BT2 + c;
  if (i >= P.n0 || k >= P.n2) return;
  float v[3][T1][3];
#pragma unroll
  for (int x = 0; x < 3; ++x)
#pragma unroll
    for (int r = 0; r < T1; ++r)
#pragma unroll
      for (int z = 0; z < 3; ++z) v[x][r][z] = tile[((a + x) * T1 + r) * T2 + c + z];
#pragma unroll
  for (int bb = 0; bb < BT1; ++bb) {
    const int j = by * BT1 + bb;
    if (j < P.n1) {
      float acc = 0.f;
#pragma unroll
      for (int x = 0; x < 3; ++x)
#pragma unroll
        for (int y = 0; y < 3; ++y)
#pragma unroll
          for (int z = 0; z < 3; ++z) acc += kr[(x * 3 + y) * 3 + z] * v[x][bb + y][z];
      dst[((long long)b * nvox + ((long long)i * P.n1 + j) * P.n2 + k) * P.dst_stride + P.dst_off] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// TMA-staged, persistent version of blur3d_333_kernel (the two 3x3x3 blurs of the default training configuration).
// The halo tile of a block of 8 x 8 x 32 outputs is ONE 4-D TMA box {40, 10, 10, 1} of the dense source volume
// (cp.async.bulk.tensor, completion on an mbarrier): out-of-volume elements arrive as zeros = the 'SAME' zero padding of
// tf.nn.conv3d (ext/lab2im/layers.py:748,758), no per-element bounds logic or index arithmetic on the load path.  CTAs are
// persistent (grid = a multiple of the SM count) and double-buffered: thread 0 issues the box of tile t + 1 before the CTA
// works on tile t, so the global -> shared latency that bounded the one-tile-per-CTA kernel (2 CTAs / SM, load -> sync ->
// compute -> exit; profiles/r02_generator_ncu_baseline.txt: 305 - 523 GB/s) overlaps the normalise / gamma / stencil work.
// Same arithmetic and accumulation order as blur3d_333_kernel (bit-identical results).
// ---------------------------------------------------------------------------------------------------------
constexpr int TBX = 4;                                        // the box starts TBX floats left of the tile: the innermost start of a
                                                              // TMA box must be 16-byte aligned (a start at x0 - 1 is an illegal
                                                              // instruction at run time: profiles/r02_tma_blur_sanitizer.txt)
constexpr int TB2 = 40;                                       // box width: TBX + BT2 + 1, rounded up to a multiple of 4 floats
constexpr int TMA_TILE_FLOATS = (BT0 + 2) * (BT1 + 2) * TB2;  // 4000 floats = 16000 B per box
constexpr int TMA_STAGE_FLOATS = (TMA_TILE_FLOATS + 31) / 32 * 32;   // stage stride: TMA destinations are 128-byte aligned

__device__ __forceinline__ uint32_t gen_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gen_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gen_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gen_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gen_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gen_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = gen_smem_u32(bar);
  uint32_t done = 0;
  for (int spin = 0;; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity), "r"(0x989680u) : "memory");
    if (done) break;
    if (spin > 2000) { printf("generator: mbarrier timeout (block %d)\n", blockIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void gen_tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                                int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(gen_smem_u32(dst)), "l"((uint64_t)map), "r"(gen_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

__device__ __forceinline__ void gen_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gen_smem_u32(bar)) : "memory");
}

// 288 threads: warps 0-7 compute (one d0 plane of the tile each), warp 8 is the TMA producer (same warp-specialised
// full / empty mbarrier ring as the convolution kernels)
__global__ void __launch_bounds__(288, 2)
blur3d_333_tma_kernel(const __grid_constant__ CUtensorMap map_src, float* __restrict__ dst, const float* __restrict__ kern,
                      const uint32_t* __restrict__ minmax, const float* __restrict__ gamma_exp, BlurParams P) {
  constexpr int T1 = BT1 + 2, T2 = TB2;
  __shared__ __align__(128) float tiles[2][TMA_STAGE_FLOATS];
  __shared__ __align__(8) uint64_t full[2], empty[2];
  const int nb2 = (P.n2 + BT2 - 1) / BT2, nb1 = (P.n1 + BT1 - 1) / BT1, nb0 = (P.n0 + BT0 - 1) / BT0;
  const long long ntiles = (long long)P.B * nb0 * nb1 * nb2;
  const long long nvox = (long long)P.n0 * P.n1 * P.n2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    gen_mbar_init(full + 0, 1);
    gen_mbar_init(full + 1, 1);
    gen_mbar_init(empty + 0, 8);
    gen_mbar_init(empty + 1, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (warp == 8) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&map_src) : "memory");
      int it = 0;
      for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int stage = it & 1;
        long long blk = tile;
        const int bz = (int)(blk % nb2); blk /= nb2;
        const int by = (int)(blk % nb1); blk /= nb1;
        const int bx = (int)(blk % nb0);
        const int b = (int)(blk / nb0);
        gen_mbar_wait(empty + stage, (uint32_t)(((it >> 1) & 1) ^ 1));      // the consumers are done with this stage
        gen_mbar_expect_tx(full + stage, TMA_TILE_FLOATS * 4);
        gen_tma_load_4d(&map_src, full + stage, tiles[stage], bz * BT2 - TBX, by * BT1 - 1, bx * BT0 - 1, b);
      }
    }
    return;
  }
  // ================================ consumers (256 threads) ================================
  float kr[27];
#pragma unroll
  for (int t = 0; t < 27; ++t) kr[t] = kern[t];
  int it = 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int stage = it & 1;
    long long blk = tile;
    const int bz = (int)(blk % nb2); blk /= nb2;
    const int by = (int)(blk % nb1); blk /= nb1;
    const int bx = (int)(blk % nb0);
    const int b = (int)(blk / nb0);
    gen_mbar_wait(full + stage, (uint32_t)((it >> 1) & 1));
    float* tile_s = tiles[stage];
    if (P.normalise) {                                       // IntensityAugmentation on the staged tile (layers.py:1235-1242)
      const float m = ord2f(minmax[2 * b]);
      const float inv_den = (ord2f(minmax[2 * b + 1]) - m) + 1e-7f;
      const float ge = P.use_gamma ? gamma_exp[b] : 1.f;
      const int o0 = bx * BT0 - 1, o1 = by * BT1 - 1, o2 = bz * BT2 - 1;
      for (int row = warp; row < (BT0 + 2) * T1; row += 8) {
        const int a = row / T1, bb = row - a * T1;
        const int i = o0 + a, j = o1 + bb;
        const bool row_ok = i >= 0 && i < P.n0 && j >= 0 && j < P.n1;
        for (int c = lane; c < BT2 + 2; c += 32) {
          const int k = o2 + c;
          float val = 0.f;                                   // padding stays exactly zero (it is not normalised)
          if (row_ok && k >= 0 && k < P.n2) {
            val = (tile_s[row * T2 + (TBX - 1) + c] - m) / inv_den;
            // x^g on [0, 1] as ex2(g * lg2(x)) (two SFU instructions; ~1e-6 relative, the bar on images is 1e-3): the
            // accurate powf made this kernel issue-bound (73 M warp instructions for 4.1 M voxels, ncu r02f)
            if (P.use_gamma) val = __powf(val, ge);
          }
          tile_s[row * T2 + (TBX - 1) + c] = val;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");         // consumers only: the producer warp is not part of it
    }
    const int c = lane, a = warp;
    const int i = bx * BT0 + a, k = bz * BT2 + c;
    if (i < P.n0 && k < P.n2) {
      float v[3][T1][3];
#pragma unroll
      for (int x = 0; x < 3; ++x)
#pragma unroll
        for (int r = 0; r < T1; ++r)
#pragma unroll
          for (int z = 0; z < 3; ++z) v[x][r][z] = tile_s[((a + x) * T1 + r) * T2 + (TBX - 1) + c + z];
#pragma unroll
      for (int bb = 0; bb < BT1; ++bb) {
        const int j = by * BT1 + bb;
        if (j < P.n1) {
          float acc = 0.f;
#pragma unroll
          for (int x = 0; x < 3; ++x)
#pragma unroll
            for (int y = 0; y < 3; ++y)
#pragma unroll
              for (int z = 0; z < 3; ++z) acc += kr[(x * 3 + y) * 3 + z] * v[x][bb + y][z];
          dst[((long long)b * nvox + ((long long)i * P.n1 + j) * P.n2 + k) * P.dst_stride + P.dst_off] = acc;
        }
      }
    }
    __syncwarp();
    if (lane == 0) gen_mbar_arrive(empty + stage);           // this warp's reads of tiles[stage] are complete
  }
}

// elementwise normalise(+gamma) without blur (window 1x1x1 special case is handled by blur3d too; this is for
// the real-image target where no blur follows)
__global__ void copy_strided_kernel(const float* __restrict__ src, float* __restrict__ dst, long long n, int ss, int so,
                                    int ds, int dofs) {
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n; v += (long long)gridDim.x * blockDim.x)
    dst[v * ds + dofs] = src[v * ss + so];
}

// reliability map = outer product of per-axis factors (ext/lab2im/edit_tensors.py:313-329), or constant 1 (:333)
__global__ void outer3_kernel(float* __restrict__ dst, const double* __restrict__ f0, const double* __restrict__ f1,
                              const double* __restrict__ f2, int B, int n0, int n1, int n2, int ds, int dofs) {
  const long long nvox = (long long)n0 * n1 * n2;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nvox * B;
       t += (long long)gridDim.x * blockDim.x) {
    long long v = t % nvox;
    const int k = (int)(v % n2); v /= n2;
    const int j = (int)(v % n1);
    const int i = (int)(v / n1);
    const float val = f0 ? (float)(__dmul_rn(__dmul_rn(f0[i], f1[j]), f2[k])) : 1.f;
    dst[t * ds + dofs] = val;
  }
}

int grid_for(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  const long long cap = 148LL * 16;   // grid-stride kernels: 16 resident CTAs of 256 threads per SM
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

// =========================================================================================================
// C ABI (declared in include/synthsr_b200.h)
// =========================================================================================================
extern "C" {

int ssr_resize(const float* src, float* dst, int B, int s0, int s1, int s2, int d0, int d1, int d2, int C,
               int nearest, int dst_stride, int dst_off, void* stream) {
  SSR_CHECK_ARG(src && dst && B > 0 && C >= 1 && C <= 3, "src/dst/B/C");
  SSR_CHECK_ARG(s0 > 0 && s1 > 0 && s2 > 0 && d0 > 0 && d1 > 0 && d2 > 0, "shapes");
  if (dst_stride <= 0) { dst_stride = C; dst_off = 0; }
  const float z0 = (float)((double)d0 / (double)s0), z1 = (float)((double)d1 / (double)s1),
              z2 = (float)((double)d2 / (double)s2);
  const long long total = (long long)B * d0 * d1 * d2;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_for(total);
#define LAUNCH(CC, NN)                                                                                             \
  resize_kernel<CC, NN><<<g, 256, 0, st>>>(src, dst, B, s0, s1, s2, d0, d1, d2, z0, z1, z2, dst_stride, dst_off)
  if (C == 1) { if (nearest) LAUNCH(1, true); else LAUNCH(1, false); }
  else if (C == 2) { if (nearest) LAUNCH(2, true); else LAUNCH(2, false); }
  else { if (nearest) LAUNCH(3, true); else LAUNCH(3, false); }
#undef LAUNCH
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

int ssr_svf_integrate(float* vec, float* tmp, int B, int n0, int n1, int n2, int nb_steps, void* stream) {
  SSR_CHECK_ARG(vec && tmp && B > 0 && n0 > 0 && n1 > 0 && n2 > 0 && nb_steps >= 0 && nb_steps < 30, "args");
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)B * n0 * n1 * n2;
  const int g = grid_for(total);
  // vec / 2**nb_steps (exact), then nb_steps ping-pong squaring passes; result ends in `vec`
  float* a = vec;
  float* b = tmp;
  svf_step_kernel<<<g, 256, 0, st>>>(a, b, B, n0, n1, n2, 1.0f / (float)(1 << nb_steps));
  SSR_COUNT_LAUNCH();
  { float* t = a; a = b; b = t; }
  for (int s = 0; s < nb_steps; ++s) {
    svf_step_kernel<<<g, 256, 0, st>>>(a, b, B, n0, n1, n2, 0.f);
    SSR_COUNT_LAUNCH();
    float* t = a; a = b; b = t;
  }
  SSR_CHECK_LAUNCH();
  if (a != vec) SSR_CHECK_CUDA(cudaMemcpyAsync(vec, a, (size_t)total * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return SSR_OK;
}

static int deform_common(const void* src, void* dst, int is_label, const float* aff, const float* field, int B, int n0,
                         int n1, int n2, int p0, int p1, int p2, int h0, int h1, int h2, const int* crop_idx, int c0,
                         int c1, int c2, const unsigned char* flip, const int* swap_lut, int lut_len, void* stream) {
  SSR_CHECK_ARG(src && dst && B > 0, "src/dst");
  SSR_CHECK_ARG(n0 > 2 * p0 && n1 > 2 * p1 && n2 > 2 * p2 && p0 >= 0 && p1 >= 0 && p2 >= 0, "grid/padding");
  SSR_CHECK_ARG(c0 > 0 && c1 > 0 && c2 > 0 && c0 <= n0 && c1 <= n1 && c2 <= n2, "crop shape");
  SSR_CHECK_ARG((field == nullptr) == (h0 == 0), "field/half shape");
  DeformParams P;
  P.B = B; P.n0 = n0; P.n1 = n1; P.n2 = n2; P.p0 = p0; P.p1 = p1; P.p2 = p2;
  P.h0 = field ? h0 : 0; P.h1 = h1; P.h2 = h2; P.c0 = c0; P.c1 = c1; P.c2 = c2;
  P.z0 = field ? (float)((double)n0 / (double)h0) : 1.f;
  P.z1 = field ? (float)((double)n1 / (double)h1) : 1.f;
  P.z2 = field ? (float)((double)n2 / (double)h2) : 1.f;
  P.m0 = (float)((n0 - 1) / 2.0); P.m1 = (float)((n1 - 1) / 2.0); P.m2 = (float)((n2 - 1) / 2.0);
  P.has_aff = aff != nullptr; P.lut_len = lut_len;
  const long long total = (long long)B * c0 * c1 * c2;
  cudaStream_t st = (cudaStream_t)stream;
  if (is_label)
    deform_kernel<true, int><<<grid_for(total), 256, 0, st>>>((const int*)src, (int*)dst, aff, field, crop_idx, flip,
                                                              swap_lut, P);
  else
    deform_kernel<false, float><<<grid_for(total), 256, 0, st>>>((const float*)src, (float*)dst, aff, field, crop_idx,
                                                                 flip, nullptr, P);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

int ssr_deform_labels_nearest(const int* labels, int* out, const float* aff, const float* field_half, int B, int n0,
                              int n1, int n2, int p0, int p1, int p2, int h0, int h1, int h2, const int* crop_idx,
                              int c0, int c1, int c2, const unsigned char* flip, const int* swap_lut, int lut_len,
                              void* stream) {
  return deform_common(labels, out, 1, aff, field_half, B, n0, n1, n2, p0, p1, p2, h0, h1, h2, crop_idx, c0, c1, c2,
                       flip, swap_lut, lut_len, stream);
}

int ssr_warp_linear(const float* image, float* out, const float* aff, const float* field_half, int B, int n0, int n1,
                    int n2, int p0, int p1, int p2, int h0, int h1, int h2, const int* crop_idx, int c0, int c1, int c2,
                    const unsigned char* flip, void* stream) {
  return deform_common(image, out, 0, aff, field_half, B, n0, n1, n2, p0, p1, p2, h0, h1, h2, crop_idx, c0, c1, c2,
                       flip, nullptr, 0, stream);
}

int ssr_philox_normal(float* out, long long n, unsigned long long seed, unsigned long long stream_id, void* stream) {
  SSR_CHECK_ARG(out && n > 0, "out/n");
  philox_normal_kernel<<<grid_for((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(out, n, seed, stream_id);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

int ssr_gmm_bias_minmax(const int* labels, const float* lut_mean, const float* lut_std, int lut_len,
                        const float* noise, unsigned long long seed, unsigned long long stream_id,
                        const float* bias_small, int b0, int b1, int b2, int apply_bias, float clip_max, float* out,
                        unsigned int* minmax, int B, int n0, int n1, int n2, void* stream) {
  SSR_CHECK_ARG(labels && lut_mean && lut_std && out && minmax && B > 0 && lut_len > 0, "pointers");
  SSR_CHECK_ARG((bias_small == nullptr) == (b0 == 0), "bias grid");
  GmmParams P;
  P.B = B; P.n0 = n0; P.n1 = n1; P.n2 = n2; P.lut_len = lut_len;
  P.b0 = bias_small ? b0 : 0; P.b1 = b1; P.b2 = b2;
  P.z0 = bias_small ? (float)((double)n0 / (double)b0) : 1.f;
  P.z1 = bias_small ? (float)((double)n1 / (double)b1) : 1.f;
  P.z2 = bias_small ? (float)((double)n2 / (double)b2) : 1.f;
  P.apply_bias = apply_bias; P.clip_max = clip_max; P.seed = seed; P.stream = stream_id;
  cudaStream_t st = (cudaStream_t)stream;
  minmax_init_kernel<<<1, 2 * B < 32 ? 32 : ((2 * B + 31) / 32) * 32, 0, st>>>(minmax, 2 * B);
  SSR_COUNT_LAUNCH();
  const long long total = (long long)B * (((long long)n0 * n1 * n2 + 3) / 4);
  gmm_bias_kernel<<<grid_for(total), 256, 0, st>>>(labels, lut_mean, lut_std, noise, bias_small, out, minmax, P);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

int ssr_minmax(const float* x, unsigned int* minmax, int B, long long nvox, void* stream) {
  SSR_CHECK_ARG(x && minmax && B > 0 && nvox > 0, "args");
  cudaStream_t st = (cudaStream_t)stream;
  minmax_init_kernel<<<1, 2 * B < 32 ? 32 : ((2 * B + 31) / 32) * 32, 0, st>>>(minmax, 2 * B);
  SSR_COUNT_LAUNCH();
  dim3 g(grid_for(nvox), B);
  minmax_kernel<<<g, 256, 0, st>>>(x, nvox, minmax);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

typedef CUresult (*GenEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static GenEncodeTiledFn gen_get_encode() {
  static GenEncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (GenEncodeTiledFn)p;
  });
  return fn;
}

// dense single-channel volume [B][n0][n1][n2] as a 4-D tensor map with the blur's halo box; false when TMA's alignment
// rules (16-byte base and strides) do not hold for this volume -- the caller then uses the plain shared-memory kernel
static bool gen_blur_map(CUtensorMap* m, const float* src, int B, int n0, int n1, int n2) {
  GenEncodeTiledFn enc = gen_get_encode();
  if (!enc || ((uintptr_t)src & 15) != 0 || (n2 & 3) != 0) return false;
  cuuint64_t dims[4] = {(cuuint64_t)n2, (cuuint64_t)n1, (cuuint64_t)n0, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)n2 * 4, (cuuint64_t)n1 * n2 * 4, (cuuint64_t)n0 * n1 * n2 * 4};
  cuuint32_t box[4] = {TB2, BT1 + 2, BT0 + 2, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int ssr_blur3d(const float* src, float* dst, const float* kern, int k0, int k1, int k2, const unsigned int* minmax,
               const float* gamma_exp, int B, int n0, int n1, int n2, int src_stride, int src_off, int dst_stride,
               int dst_off, void* stream) {
  SSR_CHECK_ARG(src && dst && kern && B > 0, "pointers");
  SSR_CHECK_ARG((k0 & 1) && (k1 & 1) && (k2 & 1) && k0 > 0 && k1 > 0 && k2 > 0, "odd window");
  BlurParams P;
  P.B = B; P.n0 = n0; P.n1 = n1; P.n2 = n2; P.k0 = k0; P.k1 = k1; P.k2 = k2;
  P.src_stride = src_stride > 0 ? src_stride : 1; P.src_off = src_off;
  P.dst_stride = dst_stride > 0 ? dst_stride : 1; P.dst_off = dst_off;
  P.normalise = minmax != nullptr; P.use_gamma = gamma_exp != nullptr;
  const size_t smem = (size_t)(BT0 + k0 - 1) * (BT1 + k1 - 1) * (BT2 + k2 - 1) * sizeof(float);
  SSR_CHECK_ARG(smem <= 200 * 1024, "blur window too large for the shared-memory tile");
  if (smem > 48 * 1024)
    SSR_CHECK_CUDA(cudaFuncSetAttribute(blur3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nblk = (long long)B * ssr_div_up(n0, BT0) * ssr_div_up(n1, BT1) * ssr_div_up(n2, BT2);
  CUtensorMap map;
  if (k0 == 3 && k1 == 3 && k2 == 3 && P.src_stride == 1 && P.src_off == 0 && !getenv("SSR_NO_TMA_BLUR") &&
      gen_blur_map(&map, src, B, n0, n1, n2)) {
    static int num_sms = 0;
    if (!num_sms) {
      int dev = 0;
      SSR_CHECK_CUDA(cudaGetDevice(&dev));
      SSR_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long grid = nblk < 2LL * num_sms ? nblk : 2LL * num_sms;        // persistent: 2 CTAs per SM
    blur3d_333_tma_kernel<<<(unsigned)grid, 288, 0, (cudaStream_t)stream>>>(map, dst, kern, minmax, gamma_exp, P);
  } else if (k0 == 3 && k1 == 3 && k2 == 3)
    blur3d_333_kernel<<<(unsigned)nblk, 256, 0, (cudaStream_t)stream>>>(src, dst, kern, minmax, gamma_exp, P);
  else
    blur3d_kernel<<<(unsigned)nblk, 256, smem, (cudaStream_t)stream>>>(src, dst, kern, minmax, gamma_exp, P);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

int ssr_mimic_acquisition(const float* src, float* dst, float* dist, const float* params, int B, int n0, int n1,
                          int n2, int o0, int o1, int o2, int dst_stride, int dst_off, int dist_stride, int dist_off,
                          void* stream) {
  SSR_CHECK_ARG(src && dst && params && B > 0 && n0 > 0 && n1 > 0 && n2 > 0 && o0 > 0 && o1 > 0 && o2 > 0, "args");
  mimic_acquisition_kernel<<<grid_for((long long)B * o0 * o1 * o2), 256, 0, (cudaStream_t)stream>>>(
      src, dst, dist, params, B, n0, n1, n2, o0, o1, o2, dst_stride > 0 ? dst_stride : 1, dst_off,
      dist_stride > 0 ? dist_stride : 1, dist_off);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

int ssr_copy_strided(const float* src, float* dst, long long n, int src_stride, int src_off, int dst_stride,
                     int dst_off, void* stream) {
  SSR_CHECK_ARG(src && dst && n > 0 && src_stride > 0 && dst_stride > 0, "args");
  copy_strided_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(src, dst, n, src_stride, src_off, dst_stride,
                                                                    dst_off);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

int ssr_fill_outer3(float* dst, const double* f0, const double* f1, const double* f2, int B, int n0, int n1, int n2,
                    int dst_stride, int dst_off, void* stream) {
  SSR_CHECK_ARG(dst && B > 0 && dst_stride > 0, "args");
  SSR_CHECK_ARG((f0 == nullptr) == (f1 == nullptr) && (f1 == nullptr) == (f2 == nullptr), "factors");
  outer3_kernel<<<grid_for((long long)B * n0 * n1 * n2), 256, 0, (cudaStream_t)stream>>>(dst, f0, f1, f2, B, n0, n1, n2,
                                                                                        dst_stride, dst_off);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

}  // extern "C"
