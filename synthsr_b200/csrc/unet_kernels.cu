// U-Net training-step kernels (fp32 CUDA-core reference convolutions + all non-GEMM layers), sm_100a.
//
// Replaces the Keras layers instantiated by ext/neuron/models.py:301-360 (conv_enc) and :420-498 (conv_dec):
// Conv3D(+bias+ELU), BatchNormalization(axis=-1), MaxPooling3D(2), UpSampling3D(2)+concatenate, the 1x1x1
// likelihood head, the L1/L2 loss of SynthSR/metrics_model.py:53-104 and Keras' Adam (SynthSR/training.py:444).
// The direct convolutions here are the exact-fp32 "parity mode" and the cross-check for the tcgen05 path in
// conv_tc.cu; everything else is shared by both modes.
//
// Tensors are [B][d0][d1][d2][C] float32, channels contiguous; kernels (k,k,k,Cin,Cout) exactly as Keras stores
// them, so checkpoints interchange without reshuffling.
#include "common.cuh"
#include <math_constants.h>

namespace {

__device__ __forceinline__ float elu_f(float a) { return a > 0.f ? a : (expf(a) - 1.f); }
// branch-free ELU on the SFU exponential (same form as the tensor-core epilogues): the divergent expf of elu_f costs as
// many instructions as the 27-tap first-layer convolution itself
__device__ __forceinline__ float elu_fast(float a) { const float neg = __expf(fminf(a, 0.f)) - 1.f; return a > 0.f ? a : neg; }
// derivative of ELU expressed through its output h = elu(a): 1 if a > 0 (h > 0) else exp(a) = h + 1
__device__ __forceinline__ float elu_grad_from_out(float h) { return h > 0.f ? 1.f : (h + 1.f); }

struct ConvGeom {
  int B, d0, d1, d2;
  int C1, C2;     // input channels of source 1 / source 2 (logical concat [x1, x2]; C2 = 0 if single source)
  int Cout;
  int k;          // cubic kernel size (odd)
  int act;        // 1: ELU
};

// ---------------------------------------------------------------------------------------------------------
// direct convolution, 'same' zero padding, cross-correlation (Keras Conv3D).  One thread: 1 voxel x 8 couts.
// ---------------------------------------------------------------------------------------------------------
constexpr int CO_T = 8;

__global__ void __launch_bounds__(256)
conv3d_direct_kernel(const float* __restrict__ x1, const float* __restrict__ x2, const float* __restrict__ w,
                     const float* __restrict__ bias, float* __restrict__ y, ConvGeom G) {
  const long long nvox = (long long)G.B * G.d0 * G.d1 * G.d2;
  const long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int co0 = blockIdx.y * CO_T;
  if (v >= nvox) return;
  long long r = v;
  const int i2 = (int)(r % G.d2); r /= G.d2;
  const int i1 = (int)(r % G.d1); r /= G.d1;
  const int i0 = (int)(r % G.d0);
  const int b = (int)(r / G.d0);
  const int Cin = G.C1 + G.C2;
  const int rad = G.k / 2;
  float acc[CO_T];
#pragma unroll
  for (int e = 0; e < CO_T; ++e) acc[e] = 0.f;
  const bool full = (co0 + CO_T <= G.Cout) && ((G.Cout & 3) == 0);
  for (int a = 0; a < G.k; ++a) {
    const int j0 = i0 + a - rad;
    if (j0 < 0 || j0 >= G.d0) continue;
    for (int bb = 0; bb < G.k; ++bb) {
      const int j1 = i1 + bb - rad;
      if (j1 < 0 || j1 >= G.d1) continue;
      for (int c = 0; c < G.k; ++c) {
        const int j2 = i2 + c - rad;
        if (j2 < 0 || j2 >= G.d2) continue;
        const long long nv = (((long long)b * G.d0 + j0) * G.d1 + j1) * G.d2 + j2;
        const int tap = (a * G.k + bb) * G.k + c;
        const float* wt = w + (long long)tap * Cin * G.Cout + co0;
        const float* p1 = x1 + nv * G.C1;
        for (int ci = 0; ci < G.C1; ++ci) {
          const float xv = p1[ci];
          const float* wr = wt + (long long)ci * G.Cout;
          if (full) {
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(wr));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(wr) + 1);
            acc[0] += xv * w0.x; acc[1] += xv * w0.y; acc[2] += xv * w0.z; acc[3] += xv * w0.w;
            acc[4] += xv * w1.x; acc[5] += xv * w1.y; acc[6] += xv * w1.z; acc[7] += xv * w1.w;
          } else {
#pragma unroll
            for (int e = 0; e < CO_T; ++e)
              if (co0 + e < G.Cout) acc[e] += xv * __ldg(wr + e);
          }
        }
        if (G.C2 > 0) {
          const float* p2 = x2 + nv * G.C2;
          for (int ci = 0; ci < G.C2; ++ci) {
            const float xv = p2[ci];
            const float* wr = wt + (long long)(G.C1 + ci) * G.Cout;
            if (full) {
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(wr));
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(wr) + 1);
              acc[0] += xv * w0.x; acc[1] += xv * w0.y; acc[2] += xv * w0.z; acc[3] += xv * w0.w;
              acc[4] += xv * w1.x; acc[5] += xv * w1.y; acc[6] += xv * w1.z; acc[7] += xv * w1.w;
            } else {
#pragma unroll
              for (int e = 0; e < CO_T; ++e)
                if (co0 + e < G.Cout) acc[e] += xv * __ldg(wr + e);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < CO_T; ++e) {
    if (co0 + e < G.Cout) {
      float o = acc[e] + (bias ? bias[co0 + e] : 0.f);
      if (G.act) o = elu_f(o);
      y[v * G.Cout + co0 + e] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// first U-Net layer (Cin <= 2 -> Cout <= 32, 3x3x3): K = 27*Cin is far too small for the tensor cores, and the
// generic direct kernel re-reads every input 3x.  Tile 4 x 8 x 32 outputs per block, input halo + weights in
// shared memory, every thread produces all Cout channels of 4 voxels (96-byte contiguous rows out).
// ---------------------------------------------------------------------------------------------------------
constexpr int FT0 = 4, FT1 = 8, FT2 = 32;

template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
conv3d_first_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ y, int B, int d0, int d1, int d2, int act, uint16_t* __restrict__ y2) {
  // y2 (optional): [bf16(y_lo) | bf16(y_hi)], 2 COUT bf16 per voxel = the operand of the next layer's hybrid compensated
  // forward (conv_tc.cu: tf32_split_bf16_kernel), written from the same staged rows instead of a separate pass over y
  __shared__ float sx[(FT0 + 2) * (FT1 + 2) * (FT2 + 2) * CIN];
  __shared__ __align__(16) float sw[27 * CIN * COUT];
  __shared__ float sb[COUT];
  __shared__ __align__(16) float sst[FT1 * FT2 * COUT];   // output staging for one plane
  const int nb2 = (d2 + FT2 - 1) / FT2, nb1 = (d1 + FT1 - 1) / FT1, nb0 = (d0 + FT0 - 1) / FT0;
  long long blk = blockIdx.x;
  const int b2 = (int)(blk % nb2); blk /= nb2;
  const int b1 = (int)(blk % nb1); blk /= nb1;
  const int b0 = (int)(blk % nb0);
  const int b = (int)(blk / nb0);
  for (int e = threadIdx.x; e < 27 * CIN * COUT; e += 256) sw[e] = w[e];
  if (threadIdx.x < COUT) sb[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  const int o0 = b0 * FT0 - 1, o1 = b1 * FT1 - 1, o2 = b2 * FT2 - 1;
  constexpr int T1 = FT1 + 2, T2 = FT2 + 2;
  for (int e = threadIdx.x; e < (FT0 + 2) * T1 * T2; e += 256) {
    const int c = e % T2, bb = (e / T2) % T1, a = e / (T2 * T1);
    const int i = o0 + a, j = o1 + bb, k = o2 + c;
    const bool ok = i >= 0 && i < d0 && j >= 0 && j < d1 && k >= 0 && k < d2;
    const long long src = ((((long long)b * d0 + i) * d1 + j) * d2 + k) * CIN;
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) sx[e * CIN + ci] = ok ? x[src + ci] : 0.f;
  }
  __syncthreads();
  const int t2 = threadIdx.x % FT2, t1 = threadIdx.x / FT2;
  // two output planes per pass: every weight vector read from shared memory feeds both (the kernel is bound by its
  // shared-memory loads: 6 x LDS.128 of weights + the input value per 24 FMAs when done plane by plane)
  for (int a0 = 0; a0 < FT0; a0 += 2) {
    const int i0 = b0 * FT0 + a0;
    if (i0 >= d0) break;                                   // block-uniform
    float acc[2][COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) { acc[0][co] = sb[co]; acc[1][co] = sb[co]; }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int bb = 0; bb < 3; ++bb)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float* px0 = sx + (((a0 + a) * T1 + (t1 + bb)) * T2 + (t2 + c)) * CIN;
          const float* px1 = px0 + T1 * T2 * CIN;          // plane a0 + 1 (inside the halo tile: FT0 is even)
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) {
            const float xv0 = px0[ci], xv1 = px1[ci];
            const float4* pw = reinterpret_cast<const float4*>(sw + (((a * 3 + bb) * 3 + c) * CIN + ci) * COUT);
#pragma unroll
            for (int q = 0; q < COUT / 4; ++q) {
              const float4 wv = pw[q];
              acc[0][q * 4 + 0] += xv0 * wv.x; acc[0][q * 4 + 1] += xv0 * wv.y;
              acc[0][q * 4 + 2] += xv0 * wv.z; acc[0][q * 4 + 3] += xv0 * wv.w;
              acc[1][q * 4 + 0] += xv1 * wv.x; acc[1][q * 4 + 1] += xv1 * wv.y;
              acc[1][q * 4 + 2] += xv1 * wv.z; acc[1][q * 4 + 3] += xv1 * wv.w;
            }
          }
        }
    // stage each plane's 256 x COUT outputs in shared memory and write them as fully coalesced float4 rows (a
    // thread-per-voxel store pattern touches every 32-byte sector in six separate instructions)
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
      if (i0 + pl >= d0) break;                            // block-uniform
      __syncthreads();                                     // previous plane's copy-out has finished reading sst
#pragma unroll
      for (int q = 0; q < COUT / 4; ++q) {
        float4 v = make_float4(acc[pl][q * 4], acc[pl][q * 4 + 1], acc[pl][q * 4 + 2], acc[pl][q * 4 + 3]);
        if (act) { v.x = elu_fast(v.x); v.y = elu_fast(v.y); v.z = elu_fast(v.z); v.w = elu_fast(v.w); }
        reinterpret_cast<float4*>(sst + threadIdx.x * COUT)[q] = v;
      }
      __syncthreads();
      for (int e = threadIdx.x; e < FT1 * FT2 * (COUT / 4); e += 256) {
        const int vox = e / (COUT / 4), part = e % (COUT / 4);
        const int j1 = b1 * FT1 + (vox >> 5), j2 = b2 * FT2 + (vox & 31);
        if (j1 < d1 && j2 < d2) {
          const long long vidx = (((long long)b * d0 + i0 + pl) * d1 + j1) * d2 + j2;
          const float4 v = reinterpret_cast<const float4*>(sst)[e];
          reinterpret_cast<float4*>(y + vidx * COUT)[part] = v;
          if (y2 != nullptr) {
            const float in[4] = {v.x, v.y, v.z, v.w};
            uint32_t lo[4], hi[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t u = __float_as_uint(in[q]);
              u = (u + 0xFFFu + ((u >> 13) & 1u)) & ~0x1FFFu;          // rne_tf32, as the TMA load rounds
              const float h = __uint_as_float(u);
              const uint32_t uh = __float_as_uint(h), ul = __float_as_uint(in[q] - h);
              hi[q] = (uh + 0x7FFFu + ((uh >> 16) & 1u)) >> 16;        // rne bf16
              lo[q] = (ul + 0x7FFFu + ((ul >> 16) & 1u)) >> 16;
            }
            uint16_t* r2 = y2 + vidx * (2 * COUT) + part * 4;
            *reinterpret_cast<uint2*>(r2) = make_uint2(lo[0] | (lo[1] << 16), lo[2] | (lo[3] << 16));
            *reinterpret_cast<uint2*>(r2 + COUT) = make_uint2(hi[0] | (hi[1] << 16), hi[2] | (hi[3] << 16));
          }
        }
      }
    }
  }
}


// weights for the data-gradient expressed as a forward convolution of dy:
//   wd[tap'][co][ci] = w[K-1-tap'][ci][co]   (flip all three axes, swap channel roles)
__global__ void flip_transpose_kernel(const float* __restrict__ w, float* __restrict__ wd, int ntap, int Cin, int Cout) {
  const long long n = (long long)ntap * Cin * Cout;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(t % Cin);
    const int co = (int)((t / Cin) % Cout);
    const int tp = (int)(t / ((long long)Cin * Cout));
    wd[t] = w[((long long)(ntap - 1 - tp) * Cin + ci) * Cout + co];
  }
}

// ---------------------------------------------------------------------------------------------------------
// weight gradient: dW[tap][ci][co] = sum_v x[v + tap][ci] * dy[v][co].  block = (tap, ci, voxel split),
// threads run over co (coalesced dy rows, broadcast x).  fp32 accumulate per thread, atomicAdd across splits.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
wgrad_direct_kernel(const float* __restrict__ x1, const float* __restrict__ x2, const float* __restrict__ dy,
                    float* __restrict__ dw, ConvGeom G, int nsplit) {
  const int Cin = G.C1 + G.C2;
  const int tap = blockIdx.x;
  const int ci = blockIdx.y;
  const int split = blockIdx.z;
  const int rad = G.k / 2;
  const int c = tap % G.k, bb = (tap / G.k) % G.k, a = tap / (G.k * G.k);
  const float* xs = ci < G.C1 ? x1 : x2;
  const int xc = ci < G.C1 ? ci : ci - G.C1;
  const int xC = ci < G.C1 ? G.C1 : G.C2;
  const long long nvox = (long long)G.B * G.d0 * G.d1 * G.d2;
  const long long per = (nvox + nsplit - 1) / nsplit;
  const long long v0 = split * per, v1 = min(nvox, v0 + per);
  for (int co = threadIdx.x; co < G.Cout; co += blockDim.x) {
    float acc = 0.f;
    for (long long v = v0; v < v1; ++v) {
      long long r = v;
      const int i2 = (int)(r % G.d2); r /= G.d2;
      const int i1 = (int)(r % G.d1); r /= G.d1;
      const int i0 = (int)(r % G.d0);
      const int b = (int)(r / G.d0);
      const int j0 = i0 + a - rad, j1 = i1 + bb - rad, j2 = i2 + c - rad;
      if (j0 < 0 || j0 >= G.d0 || j1 < 0 || j1 >= G.d1 || j2 < 0 || j2 >= G.d2) continue;
      const long long nv = (((long long)b * G.d0 + j0) * G.d1 + j1) * G.d2 + j2;
      acc += xs[nv * xC + xc] * dy[v * G.Cout + co];
    }
    float* o = dw + ((long long)tap * Cin + ci) * G.Cout + co;
    if (nsplit == 1) *o += acc; else atomicAdd(o, acc);
  }
}

// ---------------------------------------------------------------------------------------------------------
// weight gradient for very small Cin (the first layer: Cin = 1 or 2): dW[(tap,ci)][co] = sum_v X[v+tap][ci] dY[v][co]
// as a skinny GEMM [27*Cin x V] . [V x Cout].  Persistent blocks stage 128 voxels (their k^3*Cin input neighbours and
// their Cout output gradients) in shared memory, every thread owns up to WS_MAXOUT (row, col) outputs in registers,
// one atomicAdd per output per block at the end.
// ---------------------------------------------------------------------------------------------------------
constexpr int WS_VOX = 128;

// thread (r, cg): row r of the [k^3*Cin] input-patch matrix, column group cg of 4 output channels
__global__ void __launch_bounds__(384)
wgrad_small_cin_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, ConvGeom G) {
  extern __shared__ __align__(16) float wsm[];
  const int Cin = G.C1, K = G.k * G.k * G.k * Cin, ncg = G.Cout / 4;
  float* sdy = wsm;                            // [WS_VOX][Cout]   (16-byte aligned rows: Cout % 4 == 0)
  float* sx = wsm + WS_VOX * G.Cout;           // [WS_VOX][K + 1]
  const int rad = G.k / 2;
  const long long nvox = (long long)G.B * G.d0 * G.d1 * G.d2;
  const int r = threadIdx.x / ncg, cg = threadIdx.x % ncg;
  const bool active = r < K;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long base = (long long)blockIdx.x * WS_VOX; base < nvox; base += (long long)gridDim.x * WS_VOX) {
    if (threadIdx.x < WS_VOX) {
      const int vi = threadIdx.x;
      const long long v = base + vi;
      long long q = v;
      const int i2 = (int)(q % G.d2); q /= G.d2;
      const int i1 = (int)(q % G.d1); q /= G.d1;
      const int i0 = (int)(q % G.d0);
      const int b = (int)(q / G.d0);
      float* row = sx + vi * (K + 1);
      for (int a = 0; a < G.k; ++a)
        for (int bb = 0; bb < G.k; ++bb)
          for (int c = 0; c < G.k; ++c) {
            const int j0 = i0 + a - rad, j1 = i1 + bb - rad, j2 = i2 + c - rad;
            const bool ok = v < nvox && j0 >= 0 && j0 < G.d0 && j1 >= 0 && j1 < G.d1 && j2 >= 0 && j2 < G.d2;
            const long long nv = (((long long)b * G.d0 + j0) * G.d1 + j1) * G.d2 + j2;
            for (int ci = 0; ci < Cin; ++ci) row[((a * G.k + bb) * G.k + c) * Cin + ci] = ok ? x[nv * Cin + ci] : 0.f;
          }
    }
    for (int e = threadIdx.x; e < WS_VOX * ncg; e += blockDim.x) {
      const long long v = base + e / ncg;
      reinterpret_cast<float4*>(sdy)[e] =
          v < nvox ? reinterpret_cast<const float4*>(dy)[v * ncg + e % ncg] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int vi = 0; vi < WS_VOX; ++vi) {
        const float xv = sx[vi * (K + 1) + r];
        const float4 d = reinterpret_cast<const float4*>(sdy)[vi * ncg + cg];
        acc.x += xv * d.x; acc.y += xv * d.y; acc.z += xv * d.z; acc.w += xv * d.w;
      }
    }
    __syncthreads();
  }
  if (active) {
    float* o = dw + (long long)r * G.Cout + cg * 4;     // dw layout (tap, ci, co): row r = tap*Cin + ci
    atomicAdd(o + 0, acc.x); atomicAdd(o + 1, acc.y); atomicAdd(o + 2, acc.z); atomicAdd(o + 3, acc.w);
  }
}

// First-layer weight gradient (Cin <= 2, Cout = 24): register-tiled.  A thread owns the (k0, k1) tap pair, an
// 8-channel group of dy and all three k2 taps (24 * CIN accumulators); nine such thread sets stride over the voxels of a
// plane held in shared memory (x halo tile + dy plane), so every dy value is loaded from shared memory once per
// 27 threads and feeds 24 FMAs.  Persistent blocks: one atomic flush per block.
// The dy planes (the only large operand: 96 B per voxel) are double-buffered with cp.async (zero-fill outside the
// volume): plane p + 1 streams in while plane p is being accumulated, tiles are WT0 = 8 planes deep so the x halo tile is
// re-loaded once per 8 planes (the first version loaded every plane synchronously and spent ~2/3 of its time waiting).
constexpr int WT0 = 8;

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int src_size = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(src_size) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int CIN>
__global__ void __launch_bounds__(256)
wgrad_first_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, int B, int d0,
                   int d1, int d2) {
  constexpr int COUT = 24, T1 = FT1 + 2, T2 = FT2 + 2;
  extern __shared__ __align__(16) float wf_smem[];
  float* sdy = wf_smem;                                        // [2][FT1 * FT2 * COUT]
  float* sx = sdy + 2 * FT1 * FT2 * COUT;                      // [(WT0 + 2) * T1 * T2 * CIN]
  float* sacc = sx + (WT0 + 2) * T1 * T2 * CIN;                // [27 * CIN * COUT]
  const int tid = threadIdx.x;
  for (int e = tid; e < 27 * CIN * COUT; e += 256) sacc[e] = 0.f;
  const int owner = tid % 27, stream = tid / 27;          // stream 9 (threads 243..255) only helps with the loads
  const int k01 = owner / 3, cg = owner % 3, k0 = k01 / 3, k1 = k01 % 3;
  float acc[CIN][3][8];
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[ci][c][j] = 0.f;
  const int nb2 = (d2 + FT2 - 1) / FT2, nb1 = (d1 + FT1 - 1) / FT1, nb0 = (d0 + WT0 - 1) / WT0;
  const long long ntiles = (long long)B * nb0 * nb1 * nb2;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    long long blk = tile;
    const int b2 = (int)(blk % nb2); blk /= nb2;
    const int b1 = (int)(blk % nb1); blk /= nb1;
    const int b0 = (int)(blk % nb0);
    const int b = (int)(blk / nb0);
    const int o0 = b0 * WT0 - 1, o1 = b1 * FT1 - 1, o2 = b2 * FT2 - 1;
    const int np = min(WT0, d0 - b0 * WT0);                 // planes of this tile
    auto prefetch_plane = [&](int a0) {                     // dy plane b0 * WT0 + a0 -> buffer a0 & 1
      float* dst = sdy + (a0 & 1) * (FT1 * FT2 * COUT);
      const int i0 = b0 * WT0 + a0;
      for (int e = tid; e < FT1 * FT2 * (COUT / 4); e += 256) {
        const int vox = e / (COUT / 4), part = e % (COUT / 4);
        const int i1 = b1 * FT1 + (vox >> 5), i2 = b2 * FT2 + (vox & 31);
        const bool ok = i1 < d1 && i2 < d2;
        const float* src = ok ? dy + ((((long long)b * d0 + i0) * d1 + i1) * d2 + i2) * COUT + part * 4 : dy;
        cp_async16_zfill(dst + e * 4, src, ok);
      }
      cp_async_commit();
    };
    __syncthreads();                                      // previous tile's readers are done with sx / sdy
    prefetch_plane(0);
    for (int e = tid; e < (np + 2) * T1 * T2; e += 256) {
      const int c = e % T2, bb = (e / T2) % T1, a = e / (T2 * T1);
      const int i = o0 + a, j = o1 + bb, k = o2 + c;
      const bool ok = i >= 0 && i < d0 && j >= 0 && j < d1 && k >= 0 && k < d2;
      const long long src = ((((long long)b * d0 + i) * d1 + j) * d2 + k) * CIN;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) sx[e * CIN + ci] = ok ? x[src + ci] : 0.f;
    }
    for (int a0 = 0; a0 < np; ++a0) {
      cp_async_wait<0>();                                 // this thread's part of plane a0 has landed ...
      __syncthreads();                                    // ... everyone's has; sx visible; plane a0 - 1 fully consumed
      if (a0 + 1 < np) prefetch_plane(a0 + 1);            // streams in behind the accumulation of plane a0
      if (stream < 9) {
        const float* pdy = sdy + (a0 & 1) * (FT1 * FT2 * COUT);
        for (int vi = stream; vi < FT1 * FT2; vi += 9) {
          const int t1 = vi >> 5, t2 = vi & 31;
          const float4 da = reinterpret_cast<const float4*>(pdy + vi * COUT + cg * 8)[0];
          const float4 db = reinterpret_cast<const float4*>(pdy + vi * COUT + cg * 8)[1];
          const float* px = sx + (((a0 + k0) * T1 + (t1 + k1)) * T2 + t2) * CIN;
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float xv = px[c * CIN + ci];
              acc[ci][c][0] += xv * da.x; acc[ci][c][1] += xv * da.y; acc[ci][c][2] += xv * da.z; acc[ci][c][3] += xv * da.w;
              acc[ci][c][4] += xv * db.x; acc[ci][c][5] += xv * db.y; acc[ci][c][6] += xv * db.z; acc[ci][c][7] += xv * db.w;
            }
        }
      }
    }
  }
  __syncthreads();
  if (stream < 9) {
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j)     // dw layout (tap, ci, co), tap = (k0 * 3 + k1) * 3 + k2
          atomicAdd(&sacc[(((k0 * 3 + k1) * 3 + c) * CIN + ci) * COUT + cg * 8 + j], acc[ci][c][j]);
  }
  __syncthreads();
  for (int e = tid; e < 27 * CIN * COUT; e += 256) atomicAdd(dw + e, sacc[e]);
}

// per-channel sum over voxels of t[v][C] (bias gradients): out[c] += sum_v t[v][c]
__global__ void channel_sum_kernel(const float* __restrict__ t, long long nvox, int C, float* __restrict__ out) {
  extern __shared__ double sh[];
  const int cx = threadIdx.x;          // channel lane
  const int ry = threadIdx.y;          // voxel sub-lane
  const int ny = blockDim.y;
  for (int c0 = 0; c0 < C; c0 += blockDim.x) {
    const int c = c0 + cx;
    double acc = 0.0;
    if (c < C)
      for (long long v = blockIdx.x * (long long)ny + ry; v < nvox; v += (long long)gridDim.x * ny) acc += t[v * C + c];
    sh[ry * blockDim.x + cx] = acc;
    __syncthreads();
    if (ry == 0 && c < C) {
      double s = 0.0;
      for (int y = 0; y < ny; ++y) s += sh[y * blockDim.x + cx];
      atomicAdd(out + c, (float)s);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// BatchNormalization (training mode): statistics, finalize, apply (+pool / +upsample), backward
// ---------------------------------------------------------------------------------------------------------
// sums[0..C) = sum x, sums[C..2C) = sum x^2   (double)
__global__ void bn_stats_kernel(const float* __restrict__ x, long long nvox, int C, double* __restrict__ sums) {
  extern __shared__ double sh[];
  const int cx = threadIdx.x, ry = threadIdx.y, ny = blockDim.y, nx = blockDim.x;
  for (int c0 = 0; c0 < C; c0 += nx) {
    const int c = c0 + cx;
    double s = 0.0, q = 0.0;
    if (c < C)
      for (long long v = blockIdx.x * (long long)ny + ry; v < nvox; v += (long long)gridDim.x * ny) {
        const double f = x[v * C + c];
        s += f; q += f * f;
      }
    sh[(ry * nx + cx) * 2] = s;
    sh[(ry * nx + cx) * 2 + 1] = q;
    __syncthreads();
    if (ry == 0 && c < C) {
      double ts = 0.0, tq = 0.0;
      for (int y = 0; y < ny; ++y) { ts += sh[(y * nx + cx) * 2]; tq += sh[(y * nx + cx) * 2 + 1]; }
      atomicAdd(sums + c, ts);
      atomicAdd(sums + C + c, tq);
    }
    __syncthreads();
  }
}

// Vectorised per-channel double sums over voxels for C % 4 == 0 (C <= 1024): thread (r, q) owns the float4 of channels
// 4q..4q+3 of every (gridDim * R)-th voxel; loads are 16-byte and coalesced across the threads of a row, four
// independent loads are in flight per thread, fp32 partials of four voxels are folded into double accumulators.
//   MODE 0: sums = [sum a | sum a^2]                       (BatchNorm statistics)
//   MODE 1: sums = [sum a | sum a * (b - mean) * invstd]    (BatchNorm backward: a = dy, b = x)
template <int MODE>
__global__ void __launch_bounds__(256)
colsum2_vec_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ stats,
                   long long nvox, int C, double* __restrict__ sums) {
  __shared__ double sh[256 * 8];
  const int nq = C >> 2, R = blockDim.x / nq;
  const int q = threadIdx.x % nq, r = threadIdx.x / nq;
  double s[4] = {0., 0., 0., 0.}, t[4] = {0., 0., 0., 0.};
  if (r < R) {
    float4 mean = make_float4(0.f, 0.f, 0.f, 0.f), inv = mean;
    if (MODE == 1) {
      mean = *reinterpret_cast<const float4*>(stats + 4 * q);
      inv = *reinterpret_cast<const float4*>(stats + C + 4 * q);
    }
    const long long stride = (long long)gridDim.x * R;
    constexpr int U = MODE == 0 ? 8 : 4;       // independent 16-byte loads in flight per thread (x U, + U for b)
    for (long long v0 = (long long)blockIdx.x * R + r; v0 < nvox; v0 += U * stride) {
      float4 x[U], y[MODE == 1 ? U : 1];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long v = v0 + u * stride;
        x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == 1) y[u] = mean;
        if (v < nvox) {
          x[u] = *reinterpret_cast<const float4*>(a + v * C + 4 * q);
          if (MODE == 1) y[u] = *reinterpret_cast<const float4*>(b + v * C + 4 * q);
        }
      }
      float4 fs = make_float4(0.f, 0.f, 0.f, 0.f), ft = fs;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        fs.x += x[u].x; fs.y += x[u].y; fs.z += x[u].z; fs.w += x[u].w;
        if (MODE == 0) {
          ft.x += x[u].x * x[u].x; ft.y += x[u].y * x[u].y; ft.z += x[u].z * x[u].z; ft.w += x[u].w * x[u].w;
        } else {
          const float4 yy = y[MODE == 1 ? u : 0];
          ft.x += x[u].x * ((yy.x - mean.x) * inv.x); ft.y += x[u].y * ((yy.y - mean.y) * inv.y);
          ft.z += x[u].z * ((yy.z - mean.z) * inv.z); ft.w += x[u].w * ((yy.w - mean.w) * inv.w);
        }
      }
      s[0] += fs.x; s[1] += fs.y; s[2] += fs.z; s[3] += fs.w;
      t[0] += ft.x; t[1] += ft.y; t[2] += ft.z; t[3] += ft.w;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { sh[threadIdx.x * 8 + j] = s[j]; sh[threadIdx.x * 8 + 4 + j] = t[j]; }
  __syncthreads();
  if (threadIdx.x < nq) {
    double ts[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
    for (int rr = 0; rr < R; ++rr)
#pragma unroll
      for (int j = 0; j < 8; ++j) ts[j] += sh[(rr * nq + q) * 8 + j];
#pragma unroll
    for (int j = 0; j < 4; ++j) { atomicAdd(sums + 4 * q + j, ts[j]); atomicAdd(sums + C + 4 * q + j, ts[4 + j]); }
  }
}

template <int MODE>
static void launch_colsum2(const float* a, const float* b, const float* stats, long long nvox, int C, double* sums,
                           cudaStream_t st) {
  const int nq = C / 4, R = 256 / nq;
  const int U = MODE == 0 ? 8 : 4;
  long long nb = (nvox + (long long)R * U - 1) / ((long long)R * U);
  if (nb > 148 * 8) nb = 148 * 8;
  colsum2_vec_kernel<MODE><<<(unsigned)nb, nq * R, 0, st>>>(a, b, stats, nvox, C, sums);
}
static bool colsum_vec_ok(int C, const void* a, const void* b) {
  return C % 4 == 0 && C <= 1024 && ((uintptr_t)a & 15) == 0 && (!b || ((uintptr_t)b & 15) == 0);
}

// stats layout (float, 4*C): mean | invstd | scale (= gamma*invstd) | shift (= beta - mean*scale)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, long long n, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ moving_mean,
                                   float* __restrict__ moving_var, float eps, float momentum, float* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = sums[c] / (double)n;
  double var = sums[C + c] / (double)n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float scale = gamma[c] * invstd;
  stats[c] = (float)mean;
  stats[C + c] = invstd;
  stats[2 * C + c] = scale;
  stats[3 * C + c] = beta[c] - (float)mean * scale;
  if (moving_mean) {   // Keras 2.3.1: moving = moving*m + batch*(1-m); variance with n/(n-(1+eps)) correction
    moving_mean[c] = moving_mean[c] * momentum + (float)mean * (1.f - momentum);
    const double corr = (double)n / ((double)n - (1.0 + (double)eps));
    moving_var[c] = moving_var[c] * momentum + (float)(var * corr) * (1.f - momentum);
  }
}

// inference-mode stats from the moving averages
__global__ void bn_stats_from_moving_kernel(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                                            const float* __restrict__ mm, const float* __restrict__ mv, float eps,
                                            float* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = 1.f / sqrtf(mv[c] + eps);
  const float scale = gamma[c] * invstd;
  stats[c] = mm[c]; stats[C + c] = invstd; stats[2 * C + c] = scale; stats[3 * C + c] = beta[c] - mm[c] * scale;
}

// mode 0: y = BN(x);  mode 1: y = maxpool2('same')(BN(x));  mode 2: y = upsample2(BN(x)) into a channel slice
// VEC = 4: float4 over channels (C, dst_stride, dst_off multiples of 4), VEC = 1: scalar fallback
template <int VEC>
__device__ __forceinline__ void ldv(const float* p, float* v) {
  if (VEC == 4) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  else v[0] = p[0];
}

template <int VEC>
__global__ void bn_apply_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ stats,
                                int B, int d0, int d1, int d2, int C, int mode, int dst_stride, int dst_off) {
  const float* scale = stats + 2 * C;
  const float* shift = stats + 3 * C;
  const int CV = C / VEC;
  int o0 = d0, o1 = d1, o2 = d2;
  if (mode == 1) { o0 = (d0 + 1) / 2; o1 = (d1 + 1) / 2; o2 = (d2 + 1) / 2; }
  if (mode == 2) { o0 = d0 * 2; o1 = d1 * 2; o2 = d2 * 2; }
  const long long n = (long long)B * o0 * o1 * o2 * CV;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(t % CV);
    const long long vo = t / CV;
    float sc[VEC], sh[VEC], out[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) { sc[e] = scale[cv * VEC + e]; sh[e] = shift[cv * VEC + e]; }
    if (mode == 0) {
      float v[VEC];
      ldv<VEC>(x + vo * C + cv * VEC, v);
#pragma unroll
      for (int e = 0; e < VEC; ++e) out[e] = v[e] * sc[e] + sh[e];
    } else {
      long long r = vo;
      const int k = (int)(r % o2); r /= o2;
      const int j = (int)(r % o1); r /= o1;
      const int i = (int)(r % o0);
      const int b = (int)(r / o0);
      if (mode == 2) {
        float v[VEC];
        ldv<VEC>(x + ((((long long)b * d0 + i / 2) * d1 + j / 2) * d2 + k / 2) * C + cv * VEC, v);
#pragma unroll
        for (int e = 0; e < VEC; ++e) out[e] = v[e] * sc[e] + sh[e];
      } else {
#pragma unroll
        for (int e = 0; e < VEC; ++e) out[e] = -CUDART_INF_F;
        for (int a = 0; a < 2; ++a)
          for (int bb = 0; bb < 2; ++bb)
            for (int cc = 0; cc < 2; ++cc) {
              const int s0 = 2 * i + a, s1 = 2 * j + bb, s2 = 2 * k + cc;
              if (s0 < d0 && s1 < d1 && s2 < d2) {
                float v[VEC];
                ldv<VEC>(x + ((((long long)b * d0 + s0) * d1 + s1) * d2 + s2) * C + cv * VEC, v);
#pragma unroll
                for (int e = 0; e < VEC; ++e) out[e] = fmaxf(out[e], v[e] * sc[e] + sh[e]);
              }
            }
      }
    }
    float* q = y + vo * dst_stride + dst_off + cv * VEC;
    if (VEC == 4) *reinterpret_cast<float4*>(q) = make_float4(out[0], out[1], out[2], out[3]);
    else q[0] = out[0];
  }
}

// sums2[0..C) = sum dy, sums2[C..2C) = sum dy * xhat     (xhat = (x - mean) * invstd)
__global__ void bn_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                     const float* __restrict__ stats, long long nvox, int C, double* __restrict__ sums2) {
  extern __shared__ double sh[];
  const int cx = threadIdx.x, ry = threadIdx.y, ny = blockDim.y, nx = blockDim.x;
  for (int c0 = 0; c0 < C; c0 += nx) {
    const int c = c0 + cx;
    double s = 0.0, q = 0.0;
    if (c < C) {
      const float mean = stats[c], invstd = stats[C + c];
      for (long long v = blockIdx.x * (long long)ny + ry; v < nvox; v += (long long)gridDim.x * ny) {
        const double g = dy[v * C + c];
        s += g; q += g * (double)((x[v * C + c] - mean) * invstd);
      }
    }
    sh[(ry * nx + cx) * 2] = s;
    sh[(ry * nx + cx) * 2 + 1] = q;
    __syncthreads();
    if (ry == 0 && c < C) {
      double ts = 0.0, tq = 0.0;
      for (int y = 0; y < ny; ++y) { ts += sh[(y * nx + cx) * 2]; tq += sh[(y * nx + cx) * 2 + 1]; }
      atomicAdd(sums2 + c, ts);
      atomicAdd(sums2 + C + c, tq);
    }
    __syncthreads();
  }
}

// dgamma += sum dy*xhat ; dbeta += sum dy
__global__ void bn_param_grad_kernel(const double* __restrict__ sums2, int C, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dgamma[c] += (float)sums2[C + c];
  dbeta[c] += (float)sums2[c];
}

// Block-level reduction of per-thread float4 partial sums that belong to channel group `cv` (threads of a block with
// the same cv), then one atomicAdd per channel per block: the conv bias gradient db[c] = sum_v da[v][c] comes for free
// out of the kernel that writes da (saves a full extra pass over the tensor).
constexpr int EW_THREADS = 192;      // multiple of every C/4 used by the U-Net (6, 12, 24, 48, 96) and of 2, 4, 8
__device__ __forceinline__ void block_reduce_dbias(float4 part, int cv, int CV, float* __restrict__ dbias) {
  __shared__ float4 sred[EW_THREADS];
  sred[threadIdx.x] = part;
  __syncthreads();
  if ((int)threadIdx.x < CV) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = threadIdx.x; i < EW_THREADS; i += CV) { const float4 p = sred[i]; s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w; }
    atomicAdd(dbias + cv * 4 + 0, s.x); atomicAdd(dbias + cv * 4 + 1, s.y);
    atomicAdd(dbias + cv * 4 + 2, s.z); atomicAdd(dbias + cv * 4 + 3, s.w);
  }
}

// dx = scale * (dy - mean(dy) - xhat * mean(dy*xhat))  [+ add]  [* elu'(x)]      (x = BN input = ELU output)
// float4 over channels; every thread keeps the same channel group for its whole grid-stride loop (EW_THREADS % CV == 0)
__global__ void __launch_bounds__(EW_THREADS)
bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ stats,
                    const double* __restrict__ sums2, long long nvox, int C, const float* __restrict__ add,
                    int add_stride, int add_off, int elu, float* __restrict__ dx, float* __restrict__ dbias) {
  const int CV = C >> 2;
  const long long n = nvox * CV;
  const long long tid = blockIdx.x * (long long)EW_THREADS + threadIdx.x;
  const int cv = (int)(tid % CV);
  const double inv_n = 1.0 / (double)nvox;
  float mean[4], invstd[4], scale[4], m1[4], m2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = cv * 4 + e;
    mean[e] = stats[c]; invstd[e] = stats[C + c]; scale[e] = stats[2 * C + c];
    m1[e] = (float)(sums2[c] * inv_n); m2[e] = (float)(sums2[C + c] * inv_n);
  }
  float4 part = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long t = tid; t < n; t += (long long)gridDim.x * EW_THREADS) {
    const long long v = t / CV;
    const float4 g4 = reinterpret_cast<const float4*>(dy)[t];
    const float4 x4 = reinterpret_cast<const float4*>(x)[t];
    const float gi[4] = {g4.x, g4.y, g4.z, g4.w}, xi[4] = {x4.x, x4.y, x4.z, x4.w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float xhat = (xi[e] - mean[e]) * invstd[e];
      o[e] = scale[e] * (gi[e] - m1[e] - xhat * m2[e]);
    }
    if (add) {
      const float4 a4 = *reinterpret_cast<const float4*>(add + v * add_stride + add_off + cv * 4);
      o[0] += a4.x; o[1] += a4.y; o[2] += a4.z; o[3] += a4.w;
    }
    if (elu) {
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] *= elu_grad_from_out(xi[e]);
    }
    reinterpret_cast<float4*>(dx)[t] = make_float4(o[0], o[1], o[2], o[3]);
    part.x += o[0]; part.y += o[1]; part.z += o[2]; part.w += o[3];
  }
  if (dbias) block_reduce_dbias(part, cv, CV, dbias);
}

// scalar fallback (C not a multiple of 4 or EW_THREADS not a multiple of C/4)
__global__ void bn_bwd_apply_scalar_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                           const float* __restrict__ stats, const double* __restrict__ sums2,
                                           long long nvox, int C, const float* __restrict__ add, int add_stride,
                                           int add_off, int elu, float* __restrict__ dx) {
  const long long n = nvox * C;
  const double inv_n = 1.0 / (double)nvox;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const float xv = x[t];
    const float xhat = (xv - stats[c]) * stats[C + c];
    float g = stats[2 * C + c] * (dy[t] - (float)(sums2[c] * inv_n) - xhat * (float)(sums2[C + c] * inv_n));
    if (add) g += add[(t / C) * add_stride + add_off + c];
    if (elu) g *= elu_grad_from_out(xv);
    dx[t] = g;
  }
}

// gradient of maxpool2('same')(BN(x)) w.r.t. BN(x): route dp to the first maximum of each 2x2x2 window
__global__ void maxpool_bwd_kernel(const float* __restrict__ dp, const float* __restrict__ x,
                                   const float* __restrict__ stats, int B, int d0, int d1, int d2, int C,
                                   float* __restrict__ dyf) {
  const float* scale = stats + 2 * C;
  const float* shift = stats + 3 * C;
  const int o0 = (d0 + 1) / 2, o1 = (d1 + 1) / 2, o2 = (d2 + 1) / 2;
  const long long n = (long long)B * o0 * o1 * o2 * C;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    long long r = t / C;
    const int k = (int)(r % o2); r /= o2;
    const int j = (int)(r % o1); r /= o1;
    const int i = (int)(r % o0);
    const int b = (int)(r / o0);
    float m = -CUDART_INF_F;
    long long arg = -1;
    for (int a = 0; a < 2; ++a)
      for (int bb = 0; bb < 2; ++bb)
        for (int cc = 0; cc < 2; ++cc) {
          const int s0 = 2 * i + a, s1 = 2 * j + bb, s2 = 2 * k + cc;
          if (s0 < d0 && s1 < d1 && s2 < d2) {
            const long long idx = ((((long long)b * d0 + s0) * d1 + s1) * d2 + s2) * C + c;
            const float v = x[idx] * scale[c] + shift[c];
            dyf[idx] = 0.f;
            if (v > m) { m = v; arg = idx; }
          }
        }
    if (arg >= 0) dyf[arg] = dp[t];
  }
}

// float4 version (C % 4 == 0): a thread owns 4 channels of one pooled voxel; eight 16-byte loads in flight, eight
// 16-byte stores
__global__ void __launch_bounds__(256)
maxpool_bwd_vec_kernel(const float* __restrict__ dp, const float* __restrict__ x, const float* __restrict__ stats, int B,
                       int d0, int d1, int d2, int C, float* __restrict__ dyf) {
  const int CV = C >> 2;
  const int o0 = (d0 + 1) / 2, o1 = (d1 + 1) / 2, o2 = (d2 + 1) / 2;
  const long long n = (long long)B * o0 * o1 * o2 * CV;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(t % CV);
    long long r = t / CV;
    const int k = (int)(r % o2); r /= o2;
    const int j = (int)(r % o1); r /= o1;
    const int i = (int)(r % o0);
    const int b = (int)(r / o0);
    const float4 sc = *reinterpret_cast<const float4*>(stats + 2 * C + 4 * cv);
    const float4 sf = *reinterpret_cast<const float4*>(stats + 3 * C + 4 * cv);
    const float4 g = *reinterpret_cast<const float4*>(dp + t * 4);
    float4 xv[8];
    long long idx[8];
    bool ok[8];
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int s0 = 2 * i + (w >> 2), s1 = 2 * j + ((w >> 1) & 1), s2 = 2 * k + (w & 1);
      ok[w] = s0 < d0 && s1 < d1 && s2 < d2;
      idx[w] = ((((long long)b * d0 + s0) * d1 + s1) * d2 + s2) * C + 4 * cv;
      xv[w] = ok[w] ? *reinterpret_cast<const float4*>(x + idx[w]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 m = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    int ax = -1, ay = -1, az = -1, aw = -1;                 // first maximum in window order, like the scalar kernel
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (!ok[w]) continue;
      const float vx = xv[w].x * sc.x + sf.x, vy = xv[w].y * sc.y + sf.y, vz = xv[w].z * sc.z + sf.z, vw = xv[w].w * sc.w + sf.w;
      if (vx > m.x) { m.x = vx; ax = w; }
      if (vy > m.y) { m.y = vy; ay = w; }
      if (vz > m.z) { m.z = vz; az = w; }
      if (vw > m.w) { m.w = vw; aw = w; }
    }
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (!ok[w]) continue;
      *reinterpret_cast<float4*>(dyf + idx[w]) =
          make_float4(ax == w ? g.x : 0.f, ay == w ? g.y : 0.f, az == w ? g.z : 0.f, aw == w ? g.w : 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Encoder levels: backward of  maxpool2('same')(BN(x))  +  skip gradient, in two passes over x instead of four over x
// and a full-resolution dy (maxpool_bwd_vec_kernel + colsum2_vec_kernel<1> + bn_bwd_apply_kernel): the pooled
// gradient dp is routed to the first maximum of each 2x2x2 window on the fly, dy is never materialised.
// A thread owns 4 channels of one pooled voxel (EW_THREADS % (C/4) == 0 keeps the channel group fixed per thread).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pool_window_load(const float* __restrict__ x, int b, int i, int j, int k, int d0, int d1,
                                                 int d2, int C, int cv, float4* xv, long long* idx, bool* ok) {
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const int s0 = 2 * i + (w >> 2), s1 = 2 * j + ((w >> 1) & 1), s2 = 2 * k + (w & 1);
    ok[w] = s0 < d0 && s1 < d1 && s2 < d2;
    idx[w] = ((((long long)b * d0 + s0) * d1 + s1) * d2 + s2) * C + 4 * cv;
    xv[w] = ok[w] ? *reinterpret_cast<const float4*>(x + idx[w]) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
// first maximum of BN(x) in window order (same rule as maxpool_bwd_vec_kernel)
__device__ __forceinline__ void pool_window_argmax(const float4* xv, const bool* ok, const float4 sc, const float4 sf,
                                                   int* arg) {
  float4 m = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
  arg[0] = arg[1] = arg[2] = arg[3] = -1;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    if (!ok[w]) continue;
    const float vx = xv[w].x * sc.x + sf.x, vy = xv[w].y * sc.y + sf.y, vz = xv[w].z * sc.z + sf.z, vw = xv[w].w * sc.w + sf.w;
    if (vx > m.x) { m.x = vx; arg[0] = w; }
    if (vy > m.y) { m.y = vy; arg[1] = w; }
    if (vz > m.z) { m.z = vz; arg[2] = w; }
    if (vw > m.w) { m.w = vw; arg[3] = w; }
  }
}

// sums2[0..C) = sum_v dy, sums2[C..2C) = sum_v dy * xhat  with dy = unpool(dp)
__global__ void __launch_bounds__(EW_THREADS)
pool_bn_bwd_reduce_kernel(const float* __restrict__ dp, const float* __restrict__ x, const float* __restrict__ stats, int B,
                          int d0, int d1, int d2, int C, double* __restrict__ sums2) {
  __shared__ double sh[EW_THREADS * 8];
  const int CV = C >> 2;
  const int o0 = (d0 + 1) / 2, o1 = (d1 + 1) / 2, o2 = (d2 + 1) / 2;
  const long long n = (long long)B * o0 * o1 * o2 * CV;
  const long long tid = blockIdx.x * (long long)EW_THREADS + threadIdx.x;
  const int cv = (int)(tid % CV);
  const float4 mean = *reinterpret_cast<const float4*>(stats + 4 * cv);
  const float4 inv = *reinterpret_cast<const float4*>(stats + C + 4 * cv);
  const float4 sc = *reinterpret_cast<const float4*>(stats + 2 * C + 4 * cv);
  const float4 sf = *reinterpret_cast<const float4*>(stats + 3 * C + 4 * cv);
  double s[4] = {0., 0., 0., 0.}, q[4] = {0., 0., 0., 0.};
  for (long long t = tid; t < n; t += (long long)gridDim.x * EW_THREADS) {
    long long r = t / CV;
    const int k = (int)(r % o2); r /= o2;
    const int j = (int)(r % o1); r /= o1;
    const int i = (int)(r % o0);
    const int b = (int)(r / o0);
    const float4 g = *reinterpret_cast<const float4*>(dp + t * 4);
    float4 xv[8]; long long idx[8]; bool ok[8]; int arg[4];
    pool_window_load(x, b, i, j, k, d0, d1, d2, C, cv, xv, idx, ok);
    pool_window_argmax(xv, ok, sc, sf, arg);
    float xa[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (arg[0] == w) xa[0] = xv[w].x;
      if (arg[1] == w) xa[1] = xv[w].y;
      if (arg[2] == w) xa[2] = xv[w].z;
      if (arg[3] == w) xa[3] = xv[w].w;
    }
    s[0] += g.x; s[1] += g.y; s[2] += g.z; s[3] += g.w;
    q[0] += g.x * ((xa[0] - mean.x) * inv.x); q[1] += g.y * ((xa[1] - mean.y) * inv.y);
    q[2] += g.z * ((xa[2] - mean.z) * inv.z); q[3] += g.w * ((xa[3] - mean.w) * inv.w);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) { sh[threadIdx.x * 8 + e] = s[e]; sh[threadIdx.x * 8 + 4 + e] = q[e]; }
  __syncthreads();
  if ((int)threadIdx.x < CV) {
    double ts[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
    for (int i = threadIdx.x; i < EW_THREADS; i += CV)
#pragma unroll
      for (int e = 0; e < 8; ++e) ts[e] += sh[i * 8 + e];
#pragma unroll
    for (int e = 0; e < 4; ++e) { atomicAdd(sums2 + 4 * cv + e, ts[e]); atomicAdd(sums2 + C + 4 * cv + e, ts[4 + e]); }
  }
}

// dx = scale * (unpool(dp) - mean(dy) - xhat * mean(dy*xhat)) [+ add] [* elu'(x)] ; dbias += column sums of dx
__global__ void __launch_bounds__(EW_THREADS)
pool_bn_bwd_apply_kernel(const float* __restrict__ dp, const float* __restrict__ x, const float* __restrict__ stats,
                         const double* __restrict__ sums2, int B, int d0, int d1, int d2, int C,
                         const float* __restrict__ add, int add_stride, int add_off, int elu, float* __restrict__ dx,
                         float* __restrict__ dbias) {
  const int CV = C >> 2;
  const int o0 = (d0 + 1) / 2, o1 = (d1 + 1) / 2, o2 = (d2 + 1) / 2;
  const long long n = (long long)B * o0 * o1 * o2 * CV;
  const long long tid = blockIdx.x * (long long)EW_THREADS + threadIdx.x;
  const int cv = (int)(tid % CV);
  const double inv_n = 1.0 / ((double)B * d0 * d1 * d2);
  const float4 sc = *reinterpret_cast<const float4*>(stats + 2 * C + 4 * cv);
  const float4 sf = *reinterpret_cast<const float4*>(stats + 3 * C + 4 * cv);
  float mean[4], invstd[4], scale[4], m1[4], m2[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = cv * 4 + e;
    mean[e] = stats[c]; invstd[e] = stats[C + c]; scale[e] = stats[2 * C + c];
    m1[e] = (float)(sums2[c] * inv_n); m2[e] = (float)(sums2[C + c] * inv_n);
  }
  float4 part = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long t = tid; t < n; t += (long long)gridDim.x * EW_THREADS) {
    long long r = t / CV;
    const int k = (int)(r % o2); r /= o2;
    const int j = (int)(r % o1); r /= o1;
    const int i = (int)(r % o0);
    const int b = (int)(r / o0);
    const float4 g = *reinterpret_cast<const float4*>(dp + t * 4);
    const float gi[4] = {g.x, g.y, g.z, g.w};
    float4 xv[8]; long long idx[8]; bool ok[8]; int arg[4];
    pool_window_load(x, b, i, j, k, d0, d1, d2, C, cv, xv, idx, ok);
    float4 av[8];
    if (add) {
#pragma unroll
      for (int w = 0; w < 8; ++w)
        av[w] = ok[w] ? *reinterpret_cast<const float4*>(add + (idx[w] / C) * add_stride + add_off + 4 * cv)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    pool_window_argmax(xv, ok, sc, sf, arg);
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (!ok[w]) continue;
      const float xi[4] = {xv[w].x, xv[w].y, xv[w].z, xv[w].w};
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xhat = (xi[e] - mean[e]) * invstd[e];
        o[e] = scale[e] * ((arg[e] == w ? gi[e] : 0.f) - m1[e] - xhat * m2[e]);
      }
      if (add) { o[0] += av[w].x; o[1] += av[w].y; o[2] += av[w].z; o[3] += av[w].w; }
      if (elu) {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] *= elu_grad_from_out(xi[e]);
      }
      *reinterpret_cast<float4*>(dx + idx[w]) = make_float4(o[0], o[1], o[2], o[3]);
      part.x += o[0]; part.y += o[1]; part.z += o[2]; part.w += o[3];
    }
  }
  if (dbias) block_reduce_dbias(part, cv, CV, dbias);
}

__global__ void __launch_bounds__(256)
upsample_bwd_vec_kernel(const float* __restrict__ du, int du_stride, int du_off, int B, int d0, int d1, int d2, int C,
                        float* __restrict__ dlow) {
  const int CV = C >> 2;
  const long long n = (long long)B * d0 * d1 * d2 * CV;
  const int f0 = 2 * d0, f1 = 2 * d1, f2 = 2 * d2;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(t % CV);
    long long r = t / CV;
    const int k = (int)(r % d2); r /= d2;
    const int j = (int)(r % d1); r /= d1;
    const int i = (int)(r % d0);
    const int b = (int)(r / d0);
    float4 v[8];
#pragma unroll
    for (int w = 0; w < 8; ++w)
      v[w] = *reinterpret_cast<const float4*>(
          du + ((((long long)b * f0 + 2 * i + (w >> 2)) * f1 + 2 * j + ((w >> 1) & 1)) * f2 + 2 * k + (w & 1)) * du_stride +
          du_off + 4 * cv);
    float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < 8; ++w) { sacc.x += v[w].x; sacc.y += v[w].y; sacc.z += v[w].z; sacc.w += v[w].w; }   // same order as scalar
    *reinterpret_cast<float4*>(dlow + t * 4) = sacc;
  }
}

// gradient of upsample2: dlow[v][c] = sum of the 8 children of du (du may be a channel slice of a wider tensor)
__global__ void upsample_bwd_kernel(const float* __restrict__ du, int du_stride, int du_off, int B, int d0, int d1,
                                    int d2, int C, float* __restrict__ dlow) {
  const long long n = (long long)B * d0 * d1 * d2 * C;   // low-res element count
  const int f0 = 2 * d0, f1 = 2 * d1, f2 = 2 * d2;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    long long r = t / C;
    const int k = (int)(r % d2); r /= d2;
    const int j = (int)(r % d1); r /= d1;
    const int i = (int)(r % d0);
    const int b = (int)(r / d0);
    float s = 0.f;
    for (int a = 0; a < 2; ++a)
      for (int bb = 0; bb < 2; ++bb)
        for (int cc = 0; cc < 2; ++cc)
          s += du[((((long long)b * f0 + 2 * i + a) * f1 + 2 * j + bb) * f2 + 2 * k + cc) * du_stride + du_off + c];
    dlow[t] = s;
  }
}

// da = (dh [+ add]) * elu'(h)      (dh may be a channel slice of a wider tensor); float4 over channels + fused db
__global__ void __launch_bounds__(EW_THREADS)
elu_bwd_kernel(const float* __restrict__ dh, int dh_stride, int dh_off, const float* __restrict__ h,
               const float* __restrict__ add, long long nvox, int C, float* __restrict__ da, float* __restrict__ dbias) {
  const int CV = C >> 2;
  const long long n = nvox * CV;
  const long long tid = blockIdx.x * (long long)EW_THREADS + threadIdx.x;
  const int cv = (int)(tid % CV);
  float4 part = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long t = tid; t < n; t += (long long)gridDim.x * EW_THREADS) {
    const long long v = t / CV;
    float4 g = *reinterpret_cast<const float4*>(dh + v * dh_stride + dh_off + cv * 4);
    if (add) { const float4 a4 = reinterpret_cast<const float4*>(add)[t]; g.x += a4.x; g.y += a4.y; g.z += a4.z; g.w += a4.w; }
    const float4 h4 = reinterpret_cast<const float4*>(h)[t];
    g.x *= elu_grad_from_out(h4.x); g.y *= elu_grad_from_out(h4.y);
    g.z *= elu_grad_from_out(h4.z); g.w *= elu_grad_from_out(h4.w);
    reinterpret_cast<float4*>(da)[t] = g;
    part.x += g.x; part.y += g.y; part.z += g.z; part.w += g.w;
  }
  if (dbias) block_reduce_dbias(part, cv, CV, dbias);
}

__global__ void elu_bwd_scalar_kernel(const float* __restrict__ dh, int dh_stride, int dh_off, const float* __restrict__ h,
                                      const float* __restrict__ add, long long nvox, int C, float* __restrict__ da) {
  const long long n = nvox * C;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    float g = dh[(t / C) * dh_stride + dh_off + c];
    if (add) g += add[t];
    da[t] = g * elu_grad_from_out(h[t]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// 1x1x1 likelihood head (models.py:480-481) + residual + crop + L1/L2 loss (metrics_model.py:53-104), forward and
// backward in one pass: writes pred, dL/dfeat, accumulates dW_head, db_head and the loss.
// ---------------------------------------------------------------------------------------------------------
struct HeadParams {
  int B, d0, d1, d2;
  int C;            // input features
  int L;            // nb_labels (outputs), <= 4
  int metric;       // 1: l1, 2: l2
  int res_stride;   // residual: image tensor channel stride (0: none)
  int res_idx[4];   // image channel added to output l
  int tgt_stride;   // target channels (== L)
  int c0, c1, c2, cb0, cb1, cb2;   // loss crop window (size / begin); c0 == 0 -> no crop
  int train;        // 1: also write dfeat / accumulate parameter gradients
  double inv_count; // 1 / (number of elements entering the mean)
};

__global__ void __launch_bounds__(256)
head_loss_kernel(const float* __restrict__ feat, const float* __restrict__ w, const float* __restrict__ bias,
                 const float* __restrict__ image, const float* __restrict__ target, float* __restrict__ pred,
                 float* __restrict__ dfeat, float* __restrict__ dw, float* __restrict__ db, double* __restrict__ loss,
                 HeadParams P) {
  __shared__ float sw[4 * 512];
  __shared__ double sloss[8];
  const int C = P.C, L = P.L;
  for (int e = threadIdx.x; e < C * L; e += blockDim.x) sw[e] = w[e];   // w[c][l]
  __syncthreads();
  const long long nvox = (long long)P.B * P.d0 * P.d1 * P.d2;
  double lloss = 0.0;
  float ldb[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvox; v += (long long)gridDim.x * blockDim.x) {
    const float* f = feat + v * C;
    float o[4];
    for (int l = 0; l < L; ++l) o[l] = bias[l];
    for (int c = 0; c < C; ++c) {
      const float fv = f[c];
      for (int l = 0; l < L; ++l) o[l] += fv * sw[c * L + l];
    }
    bool inside = true;
    if (P.c0 > 0) {
      long long r = v;
      const int i2 = (int)(r % P.d2); r /= P.d2;
      const int i1 = (int)(r % P.d1); r /= P.d1;
      const int i0 = (int)(r % P.d0);
      inside = i0 >= P.cb0 && i0 < P.cb0 + P.c0 && i1 >= P.cb1 && i1 < P.cb1 + P.c1 && i2 >= P.cb2 && i2 < P.cb2 + P.c2;
    }
    float g[4];
    for (int l = 0; l < L; ++l) {
      if (pred) pred[v * L + l] = o[l];
      float p = o[l];
      if (P.res_stride > 0) p += image[v * P.res_stride + P.res_idx[l]];
      const float err = p - target[v * P.tgt_stride + l];
      g[l] = 0.f;
      if (inside) {
        if (P.metric == 1) {
          lloss += fabsf(err);
          g[l] = (err > 0.f ? 1.f : (err < 0.f ? -1.f : 0.f)) * (float)P.inv_count;
        } else {
          lloss += (double)err * err;
          g[l] = 2.f * err * (float)P.inv_count;
        }
      }
      ldb[l] += g[l];
    }
    if (P.train) {
      for (int c = 0; c < C; ++c) {
        float d = 0.f;
        for (int l = 0; l < L; ++l) d += g[l] * sw[c * L + l];
        dfeat[v * C + c] = d;
      }
    }
  }
  // loss + db: block reduction
  for (int o = 16; o > 0; o >>= 1) {
    lloss += __shfl_xor_sync(0xffffffffu, lloss, o);
    for (int l = 0; l < L; ++l) ldb[l] += __shfl_xor_sync(0xffffffffu, ldb[l], o);
  }
  if ((threadIdx.x & 31) == 0) {
    sloss[threadIdx.x >> 5] = lloss;
    if (P.train)
      for (int l = 0; l < L; ++l) atomicAdd(db + l, ldb[l]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sloss[i];
    atomicAdd(loss, s * P.inv_count);
  }
}

// Fused, coalesced version of head_loss_kernel + head_gout_kernel + head_wgrad_kernel for C % 4 == 0, C <= 128:
// a group of GS lanes (GS = power of two >= C/4) owns one voxel, lane q holds the float4 of channels 4q..4q+3, the
// 1x1x1 convolution is a shuffle reduction inside the group, dfeat is written as float4 and the head weight gradient
// feat^T * g is accumulated in registers (one block-level reduction at the end).  feat is read once, dfeat written once.
template <int GS, int LT>        // LT: compile-time bound of L (1 or 4) so the per-output arrays stay in few registers
__global__ void __launch_bounds__(256)
head_loss_vec_kernel(const float* __restrict__ feat, const float* __restrict__ feat_stats, const float* __restrict__ w,
                     const float* __restrict__ bias, const float* __restrict__ image, const float* __restrict__ target,
                     float* __restrict__ pred, float* __restrict__ dfeat, float* __restrict__ dw, float* __restrict__ db,
                     double* __restrict__ loss, float* __restrict__ xdot, HeadParams P) {
  // xdot (optional, with feat_stats): xdot[c][l] += sum_v g[v][l] * xhat[v][c] -- with db it gives the two reductions of
  // the BatchNorm backward of the folded layer (sum dy, sum dy * xhat for dy = g w^T) without another pass over feat
  __shared__ float sw[4 * 128];
  __shared__ float sdw[4 * 128];
  __shared__ float sdx[4 * 128];
  __shared__ double sloss[8];
  const int C = P.C, L = P.L, nq = C >> 2;
  for (int e = threadIdx.x; e < C * L; e += blockDim.x) { sw[e] = w[e]; sdw[e] = 0.f; sdx[e] = 0.f; }   // w[c][l]
  __syncthreads();
  const int q = threadIdx.x % GS;
  const bool active = q < nq;
  float wq[4][LT];                                // this lane's 4 channels x L outputs
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int l = 0; l < LT; ++l) wq[j][l] = (active && l < L) ? sw[(4 * q + j) * L + l] : 0.f;
  float bl[LT];
#pragma unroll
  for (int l = 0; l < LT; ++l) bl[l] = l < L ? bias[l] : 0.f;
  // optional BatchNorm of the feature map folded into the load (feat_stats = the layer's 4*C stats: scale at 2C, shift
  // at 3C): the normalised tensor is never written to memory
  float4 fsc = make_float4(1.f, 1.f, 1.f, 1.f), fsh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (feat_stats && active) {
    fsc = *reinterpret_cast<const float4*>(feat_stats + 2 * C + 4 * q);
    fsh = *reinterpret_cast<const float4*>(feat_stats + 3 * C + 4 * q);
  }
  const bool want_x = xdot != nullptr && feat_stats != nullptr;
  float4 fmean = make_float4(0.f, 0.f, 0.f, 0.f), finv = fmean;
  if (want_x && active) {
    fmean = *reinterpret_cast<const float4*>(feat_stats + 4 * q);
    finv = *reinterpret_cast<const float4*>(feat_stats + C + 4 * q);
  }
  float wacc[4][LT], xacc[4][LT];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int l = 0; l < LT; ++l) { wacc[j][l] = 0.f; xacc[j][l] = 0.f; }
  const long long nvox = (long long)P.B * P.d0 * P.d1 * P.d2;
  const int gpb = blockDim.x / GS;
  double lloss = 0.0;
  float ldb[LT];
#pragma unroll
  for (int l = 0; l < LT; ++l) ldb[l] = 0.f;
  // all lanes of a warp run the same number of iterations (shuffles below are full-warp); two voxels per iteration
  // keep two independent 16-byte loads (plus the image / target loads) in flight per lane
  const long long vstep = (long long)gridDim.x * gpb;
  const long long niter = (nvox + 2 * vstep - 1) / (2 * vstep);
  long long vbase = (long long)blockIdx.x * gpb + threadIdx.x / GS;
  for (long long it = 0; it < niter; ++it, vbase += 2 * vstep) {
    float4 f[2], xh[2];
    float o[2][LT], tg[2][LT], im[2][LT];
    bool vok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long v = vbase + u * vstep;
      vok[u] = v < nvox;
      f[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      xh[u] = f[u];
      if (vok[u] && active) {
        f[u] = *reinterpret_cast<const float4*>(feat + v * C + 4 * q);
        xh[u] = make_float4((f[u].x - fmean.x) * finv.x, (f[u].y - fmean.y) * finv.y, (f[u].z - fmean.z) * finv.z,
                            (f[u].w - fmean.w) * finv.w);
        f[u].x = f[u].x * fsc.x + fsh.x; f[u].y = f[u].y * fsc.y + fsh.y;
        f[u].z = f[u].z * fsc.z + fsh.z; f[u].w = f[u].w * fsc.w + fsh.w;
      }
#pragma unroll
      for (int l = 0; l < LT; ++l) {
        tg[u][l] = (vok[u] && l < L) ? target[v * P.tgt_stride + l] : 0.f;
        im[u][l] = (vok[u] && l < L && P.res_stride > 0) ? image[v * P.res_stride + P.res_idx[l]] : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int l = 0; l < LT; ++l) o[u][l] = f[u].x * wq[0][l] + f[u].y * wq[1][l] + f[u].z * wq[2][l] + f[u].w * wq[3][l];
#pragma unroll
    for (int off = GS / 2; off > 0; off >>= 1)
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int l = 0; l < LT; ++l)
          if (l < L) o[u][l] += __shfl_xor_sync(0xffffffffu, o[u][l], off);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!vok[u]) continue;
      const long long v = vbase + u * vstep;
      bool inside = true;
      if (P.c0 > 0) {
        long long r = v;
        const int i2 = (int)(r % P.d2); r /= P.d2;
        const int i1 = (int)(r % P.d1); r /= P.d1;
        const int i0 = (int)(r % P.d0);
        inside = i0 >= P.cb0 && i0 < P.cb0 + P.c0 && i1 >= P.cb1 && i1 < P.cb1 + P.c1 && i2 >= P.cb2 && i2 < P.cb2 + P.c2;
      }
      float g[LT];
#pragma unroll
      for (int l = 0; l < LT; ++l) g[l] = 0.f;
#pragma unroll
      for (int l = 0; l < LT; ++l) {
        if (l >= L) break;
        const float ol = o[u][l] + bl[l];
        if (pred && q == 0) pred[v * L + l] = ol;
        const float err = ol + im[u][l] - tg[u][l];
        if (inside) {
          if (P.metric == 1) {
            if (q == 0) lloss += fabsf(err);
            g[l] = (err > 0.f ? 1.f : (err < 0.f ? -1.f : 0.f)) * (float)P.inv_count;
          } else {
            if (q == 0) lloss += (double)err * err;
            g[l] = 2.f * err * (float)P.inv_count;
          }
        }
        if (q == 0) ldb[l] += g[l];
      }
      if (P.train && active) {
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int l = 0; l < LT; ++l) {
          d.x += g[l] * wq[0][l]; d.y += g[l] * wq[1][l]; d.z += g[l] * wq[2][l]; d.w += g[l] * wq[3][l];
        }
        *reinterpret_cast<float4*>(dfeat + v * C + 4 * q) = d;
#pragma unroll
        for (int l = 0; l < LT; ++l) {
          wacc[0][l] += f[u].x * g[l]; wacc[1][l] += f[u].y * g[l]; wacc[2][l] += f[u].z * g[l]; wacc[3][l] += f[u].w * g[l];
          xacc[0][l] += xh[u].x * g[l]; xacc[1][l] += xh[u].y * g[l]; xacc[2][l] += xh[u].z * g[l]; xacc[3][l] += xh[u].w * g[l];
        }
      }
    }
  }
  __syncwarp();
  // head weight gradient: reduce over the groups of the warp, then over the warps through shared memory
  if (P.train) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int l = 0; l < LT; ++l) {
        if (l >= L) break;
        float a = wacc[j][l];
        for (int off = GS; off < 32; off <<= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        if ((threadIdx.x & 31) < GS && active) atomicAdd(&sdw[(4 * q + j) * L + l], a);
        if (want_x) {                                  // block-uniform
          float x2 = xacc[j][l];
          for (int off = GS; off < 32; off <<= 1) x2 += __shfl_xor_sync(0xffffffffu, x2, off);
          if ((threadIdx.x & 31) < GS && active) atomicAdd(&sdx[(4 * q + j) * L + l], x2);
        }
      }
  }
  for (int o2 = 16; o2 > 0; o2 >>= 1) {
    lloss += __shfl_xor_sync(0xffffffffu, lloss, o2);
#pragma unroll
    for (int l = 0; l < LT; ++l) ldb[l] += __shfl_xor_sync(0xffffffffu, ldb[l], o2);
  }
  if ((threadIdx.x & 31) == 0) {
    sloss[threadIdx.x >> 5] = lloss;
    if (P.train)
      for (int l = 0; l < L; ++l) atomicAdd(db + l, ldb[l]);
  }
  __syncthreads();
  if (P.train)
    for (int e = threadIdx.x; e < C * L; e += blockDim.x) {
      atomicAdd(dw + e, sdw[e]);
      if (want_x) atomicAdd(xdot + e, sdx[e]);
    }
  if (threadIdx.x == 0) {
    double s2 = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s2 += sloss[i];
    atomicAdd(loss, s2 * P.inv_count);
  }
}

// dW_head[c][l] += sum_v feat[v][c] * g[v][l] where g is recomputed from dfeat is not possible -> separate pass
// using the stored per-voxel output gradient gout[v][l].
__global__ void head_wgrad_kernel(const float* __restrict__ feat, const float* __restrict__ gout, long long nvox, int C,
                                  int L, float* __restrict__ dw) {
  extern __shared__ double sh[];
  const int cx = threadIdx.x, ry = threadIdx.y, ny = blockDim.y, nx = blockDim.x;
  for (int l = 0; l < L; ++l)
    for (int c0 = 0; c0 < C; c0 += nx) {
      const int c = c0 + cx;
      double s = 0.0;
      if (c < C)
        for (long long v = blockIdx.x * (long long)ny + ry; v < nvox; v += (long long)gridDim.x * ny)
          s += (double)feat[v * C + c] * (double)gout[v * L + l];
      sh[ry * nx + cx] = s;
      __syncthreads();
      if (ry == 0 && c < C) {
        double ts = 0.0;
        for (int y = 0; y < ny; ++y) ts += sh[y * nx + cx];
        atomicAdd(dw + c * L + l, (float)ts);
      }
      __syncthreads();
    }
}

// per-voxel output gradient g[v][l] of the loss (same formula as in head_loss_kernel), for head_wgrad_kernel
__global__ void head_gout_kernel(const float* __restrict__ pred, const float* __restrict__ image,
                                 const float* __restrict__ target, float* __restrict__ gout, HeadParams P) {
  const long long nvox = (long long)P.B * P.d0 * P.d1 * P.d2;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvox; v += (long long)gridDim.x * blockDim.x) {
    bool inside = true;
    if (P.c0 > 0) {
      long long r = v;
      const int i2 = (int)(r % P.d2); r /= P.d2;
      const int i1 = (int)(r % P.d1); r /= P.d1;
      const int i0 = (int)(r % P.d0);
      inside = i0 >= P.cb0 && i0 < P.cb0 + P.c0 && i1 >= P.cb1 && i1 < P.cb1 + P.c1 && i2 >= P.cb2 && i2 < P.cb2 + P.c2;
    }
    for (int l = 0; l < P.L; ++l) {
      float p = pred[v * P.L + l];
      if (P.res_stride > 0) p += image[v * P.res_stride + P.res_idx[l]];
      const float err = p - target[v * P.tgt_stride + l];
      float g = 0.f;
      if (inside) g = P.metric == 1 ? (err > 0.f ? 1.f : (err < 0.f ? -1.f : 0.f)) * (float)P.inv_count
                                    : 2.f * err * (float)P.inv_count;
      gout[v * P.L + l] = g;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Keras-2.3.1 Adam on one flat buffer:  m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr_t m/(sqrt(v)+eps)
// ---------------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr_t, float b1, float b2, float eps,
                            float gscale) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const float gg = g[t] * gscale;
    const float mm = b1 * m[t] + (1.f - b1) * gg;
    const float vv = b2 * v[t] + (1.f - b2) * gg * gg;
    m[t] = mm; v[t] = vv;
    p[t] = p[t] - lr_t * mm / (sqrtf(vv) + eps);
  }
}

int grid_for(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

extern "C" {

static int conv3d_first_launch(const float* x1, int C1, const float* w, const float* bias, float* y, uint16_t* y2, int B,
                               int d0, int d1, int d2, int Cout, int act, void* stream) {
  const long long nblk = (long long)B * ((d0 + FT0 - 1) / FT0) * ((d1 + FT1 - 1) / FT1) * ((d2 + FT2 - 1) / FT2);
  cudaStream_t st = (cudaStream_t)stream;
  if (C1 == 1 && Cout == 24) conv3d_first_kernel<1, 24><<<(unsigned)nblk, 256, 0, st>>>(x1, w, bias, y, B, d0, d1, d2, act, y2);
  else if (C1 == 2 && Cout == 24) conv3d_first_kernel<2, 24><<<(unsigned)nblk, 256, 0, st>>>(x1, w, bias, y, B, d0, d1, d2, act, y2);
  else if (C1 == 1 && Cout == 8) conv3d_first_kernel<1, 8><<<(unsigned)nblk, 256, 0, st>>>(x1, w, bias, y, B, d0, d1, d2, act, y2);
  else conv3d_first_kernel<2, 8><<<(unsigned)nblk, 256, 0, st>>>(x1, w, bias, y, B, d0, d1, d2, act, y2);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// first layer (Cin <= 2, Cout 8 / 24, k = 3) that also writes y2 = [bf16(y_lo) | bf16(y_hi)] -- what
// ssr_tf32_split_bf16(y) would produce -- for the hybrid compensated forward of the next layer
int ssr_conv3d_first_fwd_split(const float* x, int C1, const float* w, const float* bias, float* y, void* y2, int B, int d0,
                               int d1, int d2, int Cout, int act, void* stream) {
  SSR_CHECK_ARG(x && w && y && y2 && (C1 == 1 || C1 == 2) && (Cout == 24 || Cout == 8) && (((uintptr_t)y | (uintptr_t)y2) & 15) == 0,
                "first-layer split forward: Cin 1 / 2, Cout 8 / 24, 16-byte aligned outputs");
  return conv3d_first_launch(x, C1, w, bias, y, reinterpret_cast<uint16_t*>(y2), B, d0, d1, d2, Cout, act, stream);
}

int ssr_conv3d_fwd_ref(const float* x1, int C1, const float* x2, int C2, const float* w, const float* bias, float* y,
                       int B, int d0, int d1, int d2, int Cout, int k, int act, void* stream) {
  SSR_CHECK_ARG(x1 && w && y && C1 > 0 && C2 >= 0 && (C2 == 0 || x2) && Cout > 0 && (k & 1), "conv args");
  ConvGeom G{B, d0, d1, d2, C1, C2, Cout, k, act};
  const long long nvox = (long long)B * d0 * d1 * d2;
  if (k == 3 && C2 == 0 && C1 <= 2 && (Cout == 24 || Cout == 8) && (((uintptr_t)y) & 15) == 0)     // first-layer kernel
    return conv3d_first_launch(x1, C1, w, bias, y, nullptr, B, d0, d1, d2, Cout, act, stream);
  dim3 grid((unsigned)((nvox + 255) / 256), (unsigned)((Cout + CO_T - 1) / CO_T));
  conv3d_direct_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x1, x2, w, bias, y, G);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// dx[v][Cin] = sum_tap sum_co dy[v - tap][co] w[tap][ci][co];  wd_scratch: k^3*Cin*Cout floats
int ssr_conv3d_dgrad_ref(const float* dy, const float* w, float* wd_scratch, float* dx, int B, int d0, int d1, int d2,
                         int Cin, int Cout, int k, void* stream) {
  SSR_CHECK_ARG(dy && w && wd_scratch && dx && Cin > 0 && Cout > 0 && (k & 1), "dgrad args");
  cudaStream_t st = (cudaStream_t)stream;
  const int ntap = k * k * k;
  flip_transpose_kernel<<<grid_for((long long)ntap * Cin * Cout), 256, 0, st>>>(w, wd_scratch, ntap, Cin, Cout);
  SSR_COUNT_LAUNCH();
  ConvGeom G{B, d0, d1, d2, Cout, 0, Cin, k, 0};
  const long long nvox = (long long)B * d0 * d1 * d2;
  dim3 grid((unsigned)((nvox + 255) / 256), (unsigned)((Cin + CO_T - 1) / CO_T));
  conv3d_direct_kernel<<<grid, 256, 0, st>>>(dy, nullptr, wd_scratch, nullptr, dx, G);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// dw (k,k,k,C1+C2,Cout) += x (*) dy ;  db[Cout] += sum dy   (db may be NULL)
int ssr_conv3d_wgrad_ref(const float* x1, int C1, const float* x2, int C2, const float* dy, float* dw, float* db, int B,
                         int d0, int d1, int d2, int Cout, int k, void* stream) {
  SSR_CHECK_ARG(x1 && dy && dw && C1 > 0 && C2 >= 0 && (C2 == 0 || x2) && Cout > 0 && (k & 1), "wgrad args");
  cudaStream_t st = (cudaStream_t)stream;
  ConvGeom G{B, d0, d1, d2, C1, C2, Cout, k, 0};
  const long long nvox = (long long)B * d0 * d1 * d2;
  const int Ksmall = k * k * k * C1;
  if (k == 3 && C2 == 0 && C1 <= 2 && Cout == 24 && ((uintptr_t)dy & 15) == 0) {       // first layer of the U-Net
    const long long ntiles = (long long)B * ((d0 + WT0 - 1) / WT0) * ((d1 + FT1 - 1) / FT1) * ((d2 + FT2 - 1) / FT2);
    const unsigned nb = (unsigned)(ntiles < 148 * 3 ? ntiles : 148 * 3);
    const size_t smem = sizeof(float) * (2 * FT1 * FT2 * 24 + (size_t)(WT0 + 2) * (FT1 + 2) * (FT2 + 2) * C1 + 27 * C1 * 24);
    static bool attr_set = false;
    if (!attr_set) {
      SSR_CHECK_CUDA(cudaFuncSetAttribute(wgrad_first_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      SSR_CHECK_CUDA(cudaFuncSetAttribute(wgrad_first_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr_set = true;
    }
    if (C1 == 1) wgrad_first_kernel<1><<<nb, 256, smem, st>>>(x1, dy, dw, B, d0, d1, d2);
    else wgrad_first_kernel<2><<<nb, 256, smem, st>>>(x1, dy, dw, B, d0, d1, d2);
  } else
  if (C2 == 0 && C1 <= 4 && Cout % 4 == 0 && Ksmall * (Cout / 4) <= 384 && nvox >= 4096 &&
      ((uintptr_t)dy & 15) == 0) {
    const size_t smem = (size_t)WS_VOX * (Ksmall + 1 + Cout) * sizeof(float);
    if (smem > 48 * 1024)
      SSR_CHECK_CUDA(cudaFuncSetAttribute(wgrad_small_cin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long nb = (nvox + WS_VOX - 1) / WS_VOX;
    if (nb > 148 * 4) nb = 148 * 4;
    int nthr = (Ksmall * (Cout / 4) + 31) / 32 * 32;
    if (nthr < WS_VOX) nthr = WS_VOX;
    wgrad_small_cin_kernel<<<(unsigned)nb, nthr, smem, st>>>(x1, dy, dw, G);
  } else {
    int nsplit = (int)((nvox + 32767) / 32768);
    if (nsplit > 64) nsplit = 64;
    dim3 grid(k * k * k, C1 + C2, nsplit);
    wgrad_direct_kernel<<<grid, 128, 0, st>>>(x1, x2, dy, dw, G, nsplit);
  }
  SSR_COUNT_LAUNCH();
  if (db) {
    dim3 blk(32, 8);
    int g = (int)((nvox + 7) / 8); if (g > 148 * 8) g = 148 * 8;
    channel_sum_kernel<<<g, blk, 32 * 8 * sizeof(double), st>>>(dy, nvox, Cout, db);
    SSR_COUNT_LAUNCH();
  }
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

int ssr_channel_sum(const float* t, long long nvox, int C, float* out, void* stream) {
  SSR_CHECK_ARG(t && out && nvox > 0 && C > 0, "args");
  dim3 blk(32, 8);
  int g = (int)((nvox + 7) / 8); if (g > 148 * 8) g = 148 * 8;
  channel_sum_kernel<<<g, blk, 32 * 8 * sizeof(double), (cudaStream_t)stream>>>(t, nvox, C, out);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// training-mode BN statistics: sums (2*C doubles, zeroed here) -> stats (4*C floats) ; updates moving stats if given
int ssr_bn_stats(const float* x, long long nvox, int C, const float* gamma, const float* beta, float* moving_mean,
                 float* moving_var, float eps, float momentum, double* sums_scratch, float* stats, void* stream) {
  SSR_CHECK_ARG(x && gamma && beta && sums_scratch && stats && nvox > 0 && C > 0, "args");
  cudaStream_t st = (cudaStream_t)stream;
  SSR_CHECK_CUDA(cudaMemsetAsync(sums_scratch, 0, 2 * C * sizeof(double), st));
  dim3 blk(32, 8);
  int g = (int)((nvox + 7) / 8); if (g > 148 * 8) g = 148 * 8;
  if (colsum_vec_ok(C, x, nullptr)) launch_colsum2<0>(x, nullptr, nullptr, nvox, C, sums_scratch, st);
  else bn_stats_kernel<<<g, blk, 32 * 8 * 2 * sizeof(double), st>>>(x, nvox, C, sums_scratch);
  SSR_COUNT_LAUNCH();
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums_scratch, nvox, C, gamma, beta, moving_mean, moving_var, eps,
                                                      momentum, stats);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// second half of ssr_bn_stats for sums produced elsewhere (the fused epilogue of ssr_conv3d_fwd_tc_k2n_stats)
int ssr_bn_finalize(const double* sums, long long nvox, int C, const float* gamma, const float* beta, float* moving_mean,
                    float* moving_var, float eps, float momentum, float* stats, void* stream) {
  SSR_CHECK_ARG(sums && gamma && beta && stats && nvox > 0 && C > 0, "args");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, nvox, C, gamma, beta, moving_mean, moving_var,
                                                                         eps, momentum, stats);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

int ssr_bn_stats_inference(int C, const float* gamma, const float* beta, const float* moving_mean,
                           const float* moving_var, float eps, float* stats, void* stream) {
  SSR_CHECK_ARG(gamma && beta && moving_mean && moving_var && stats && C > 0, "args");
  bn_stats_from_moving_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(C, gamma, beta, moving_mean,
                                                                                moving_var, eps, stats);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// mode 0 plain, 1 +maxpool2('same'), 2 +upsample2 ; (d0,d1,d2) is the INPUT spatial shape
int ssr_bn_apply(const float* x, float* y, const float* stats, int B, int d0, int d1, int d2, int C, int mode,
                 int dst_stride, int dst_off, void* stream) {
  SSR_CHECK_ARG(x && y && stats && mode >= 0 && mode <= 2, "args");
  if (dst_stride <= 0) { dst_stride = C; dst_off = 0; }
  long long n = (long long)B * d0 * d1 * d2 * C;
  if (mode == 1) n = (long long)B * ((d0 + 1) / 2) * ((d1 + 1) / 2) * ((d2 + 1) / 2) * C;
  if (mode == 2) n *= 8;
  const bool vec = (C % 4 == 0) && (dst_stride % 4 == 0) && (dst_off % 4 == 0) && (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
  if (vec)
    bn_apply_kernel<4><<<grid_for(n / 4), 256, 0, (cudaStream_t)stream>>>(x, y, stats, B, d0, d1, d2, C, mode,
                                                                         dst_stride, dst_off);
  else
    bn_apply_kernel<1><<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, y, stats, B, d0, d1, d2, C, mode, dst_stride,
                                                                      dst_off);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// BN backward: dx = BN'(dy) [+ add] [* elu'(x)] ; dgamma/dbeta accumulated.  sums_scratch: 2*C doubles.
static bool ew_vec_ok(int C, const void* a, const void* b, const void* c, int stride, int off) {
  return C % 4 == 0 && EW_THREADS % (C / 4) == 0 && stride % 4 == 0 && off % 4 == 0 &&
         (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}

static int ew_grid(long long n) {       // grid * EW_THREADS must stay a multiple of C/4: any grid works (EW_THREADS is)
  long long g = (n + EW_THREADS - 1) / EW_THREADS;
  const long long cap = 148LL * 10;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// dbias (optional): += sum_v dx[v][c]  (the bias gradient of the convolution that produced x)
static int bn_bwd_impl(const float* dy, const float* x, const float* stats, long long nvox, int C, const float* add,
               int add_stride, int add_off, int elu, float* dx, float* dgamma, float* dbeta, float* dbias,
               double* sums_scratch, void* stream, int have_sums) {
  SSR_CHECK_ARG(dy && x && stats && dx && sums_scratch && nvox > 0 && C > 0, "args");
  cudaStream_t st = (cudaStream_t)stream;
  if (!have_sums) {
    SSR_CHECK_CUDA(cudaMemsetAsync(sums_scratch, 0, 2 * C * sizeof(double), st));
    dim3 blk(32, 8);
    int g = (int)((nvox + 7) / 8); if (g > 148 * 8) g = 148 * 8;
    if (colsum_vec_ok(C, dy, x) && ((uintptr_t)stats & 15) == 0) launch_colsum2<1>(dy, x, stats, nvox, C, sums_scratch, st);
    else bn_bwd_reduce_kernel<<<g, blk, 32 * 8 * 2 * sizeof(double), st>>>(dy, x, stats, nvox, C, sums_scratch);
    SSR_COUNT_LAUNCH();
  }
  if (dgamma && dbeta) {
    bn_param_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums_scratch, C, dgamma, dbeta);
    SSR_COUNT_LAUNCH();
  }
  if (add && add_stride <= 0) { add_stride = C; add_off = 0; }
  if (ew_vec_ok(C, dy, x, dx, add ? add_stride : 4, add ? add_off : 0) && (!add || ((uintptr_t)add & 15) == 0)) {
    bn_bwd_apply_kernel<<<ew_grid(nvox * (C / 4)), EW_THREADS, 0, st>>>(dy, x, stats, sums_scratch, nvox, C, add,
                                                                        add_stride, add_off, elu, dx, dbias);
    SSR_COUNT_LAUNCH();
  } else {
    bn_bwd_apply_scalar_kernel<<<grid_for(nvox * C), 256, 0, st>>>(dy, x, stats, sums_scratch, nvox, C, add, add_stride,
                                                                   add_off, elu, dx);
    SSR_COUNT_LAUNCH();
    if (dbias) { int rc = ssr_channel_sum(dx, nvox, C, dbias, stream); if (rc) return rc; }
  }
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// Encoder level backward: gradient dp w.r.t. the POOLED BatchNorm output -> dx w.r.t. the BatchNorm input x
// (= ssr_maxpool_bwd + ssr_bn_bwd without the full-resolution intermediate).  (d0,d1,d2): shape of x.
int ssr_pool_bn_bwd(const float* dp, const float* x, const float* stats, int B, int d0, int d1, int d2, int C,
                    const float* add, int add_stride, int add_off, int elu, float* dx, float* dgamma, float* dbeta,
                    float* dbias, double* sums_scratch, void* stream) {
  SSR_CHECK_ARG(dp && x && stats && dx && sums_scratch && B > 0 && d0 > 0 && d1 > 0 && d2 > 0 && C > 0, "args");
  if (add && add_stride <= 0) { add_stride = C; add_off = 0; }
  SSR_CHECK_ARG(C % 4 == 0 && EW_THREADS % (C / 4) == 0 && (!add || (add_stride % 4 == 0 && add_off % 4 == 0)) &&
                (((uintptr_t)dp | (uintptr_t)x | (uintptr_t)stats | (uintptr_t)dx | (uintptr_t)add) & 15) == 0,
                "ssr_pool_bn_bwd needs C % 4 == 0, 192 % (C/4) == 0 and 16-byte aligned tensors");
  cudaStream_t st = (cudaStream_t)stream;
  SSR_CHECK_CUDA(cudaMemsetAsync(sums_scratch, 0, 2 * C * sizeof(double), st));
  const long long n = (long long)B * ((d0 + 1) / 2) * ((d1 + 1) / 2) * ((d2 + 1) / 2) * (C / 4);
  pool_bn_bwd_reduce_kernel<<<ew_grid(n), EW_THREADS, 0, st>>>(dp, x, stats, B, d0, d1, d2, C, sums_scratch);
  SSR_COUNT_LAUNCH();
  if (dgamma && dbeta) {
    bn_param_grad_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums_scratch, C, dgamma, dbeta);
    SSR_COUNT_LAUNCH();
  }
  pool_bn_bwd_apply_kernel<<<ew_grid(n), EW_THREADS, 0, st>>>(dp, x, stats, sums_scratch, B, d0, d1, d2, C, add, add_stride,
                                                              add_off, elu, dx, dbias);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

int ssr_bn_bwd(const float* dy, const float* x, const float* stats, long long nvox, int C, const float* add,
               int add_stride, int add_off, int elu, float* dx, float* dgamma, float* dbeta, float* dbias,
               double* sums_scratch, void* stream) {
  return bn_bwd_impl(dy, x, stats, nvox, C, add, add_stride, add_off, elu, dx, dgamma, dbeta, dbias, sums_scratch, stream, 0);
}
// same with the two reductions already in sums2 (= [sum dy | sum dy * xhat], e.g. from ssr_head_loss_bnsums)
int ssr_bn_bwd_sums(const float* dy, const float* x, const float* stats, long long nvox, int C, const float* add,
                    int add_stride, int add_off, int elu, float* dx, float* dgamma, float* dbeta, float* dbias,
                    double* sums2, void* stream) {
  return bn_bwd_impl(dy, x, stats, nvox, C, add, add_stride, add_off, elu, dx, dgamma, dbeta, dbias, sums2, stream, 1);
}

int ssr_maxpool_bwd(const float* dp, const float* x, const float* stats, int B, int d0, int d1, int d2, int C,
                    float* dy_full, void* stream) {
  SSR_CHECK_ARG(dp && x && stats && dy_full, "args");
  const long long n = (long long)B * ((d0 + 1) / 2) * ((d1 + 1) / 2) * ((d2 + 1) / 2) * C;
  if (C % 4 == 0 && (((uintptr_t)dp | (uintptr_t)x | (uintptr_t)stats | (uintptr_t)dy_full) & 15) == 0)
    maxpool_bwd_vec_kernel<<<grid_for(n / 4), 256, 0, (cudaStream_t)stream>>>(dp, x, stats, B, d0, d1, d2, C, dy_full);
  else
    maxpool_bwd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(dp, x, stats, B, d0, d1, d2, C, dy_full);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// (d0,d1,d2) is the LOW-res shape; du is the full-res gradient (channel slice du_off..du_off+C of du_stride)
int ssr_upsample_bwd(const float* du, int du_stride, int du_off, int B, int d0, int d1, int d2, int C, float* dlow,
                     void* stream) {
  SSR_CHECK_ARG(du && dlow && du_stride >= C, "args");
  if (C % 4 == 0 && du_stride % 4 == 0 && du_off % 4 == 0 && (((uintptr_t)du | (uintptr_t)dlow) & 15) == 0)
    upsample_bwd_vec_kernel<<<grid_for((long long)B * d0 * d1 * d2 * (C / 4)), 256, 0, (cudaStream_t)stream>>>(
        du, du_stride, du_off, B, d0, d1, d2, C, dlow);
  else
    upsample_bwd_kernel<<<grid_for((long long)B * d0 * d1 * d2 * C), 256, 0, (cudaStream_t)stream>>>(
        du, du_stride, du_off, B, d0, d1, d2, C, dlow);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

int ssr_elu_bwd(const float* dh, int dh_stride, int dh_off, const float* h, const float* add, long long nvox, int C,
                float* da, float* dbias, void* stream) {
  SSR_CHECK_ARG(dh && h && da && nvox > 0 && C > 0, "args");
  if (dh_stride <= 0) { dh_stride = C; dh_off = 0; }
  if (ew_vec_ok(C, dh, h, da, dh_stride, dh_off) && (!add || ((uintptr_t)add & 15) == 0)) {
    elu_bwd_kernel<<<ew_grid(nvox * (C / 4)), EW_THREADS, 0, (cudaStream_t)stream>>>(dh, dh_stride, dh_off, h, add, nvox,
                                                                                     C, da, dbias);
    SSR_COUNT_LAUNCH();
  } else {
    elu_bwd_scalar_kernel<<<grid_for(nvox * C), 256, 0, (cudaStream_t)stream>>>(dh, dh_stride, dh_off, h, add, nvox, C, da);
    SSR_COUNT_LAUNCH();
    if (dbias) { int rc = ssr_channel_sum(da, nvox, C, dbias, stream); if (rc) return rc; }
  }
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// head + loss forward/backward.  loss: 1 double (zeroed here).  gout_scratch: nvox*L floats (train only).
// sums2[c] = sum_l w[c][l] db[l] (= sum_v dy[v][c]), sums2[C + c] = sum_l w[c][l] xdot[c][l] (= sum_v dy * xhat) for
// dy = g w^T: the reductions of the BatchNorm backward of the layer folded into the head
__global__ void head_bn_sums_kernel(const float* __restrict__ w, const float* __restrict__ db, const float* __restrict__ xdot,
                                    int C, int L, double* __restrict__ sums2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int l = 0; l < L; ++l) { a += (double)w[c * L + l] * (double)db[l]; b += (double)w[c * L + l] * (double)xdot[c * L + l]; }
  sums2[c] = a;
  sums2[C + c] = b;
}

static int head_loss_impl(const float* feat, const float* feat_stats, const float* w, const float* bias, const float* image,
                  int image_channels,
                  const int* res_idx, const float* target, float* pred, float* dfeat, float* dw, float* db,
                  double* loss, float* gout_scratch, int B, int d0, int d1, int d2, int C, int L, int metric,
                  const int* crop_size, const int* crop_begin, int train, void* stream, float* xdot) {
  SSR_CHECK_ARG(feat && w && bias && target && loss && pred, "pointers");
  SSR_CHECK_ARG(L >= 1 && L <= 4 && C * L <= 2048 && (metric == 1 || metric == 2), "head shape/metric");
  SSR_CHECK_ARG(!train || (dfeat && dw && db && gout_scratch), "train buffers");
  HeadParams P;
  P.B = B; P.d0 = d0; P.d1 = d1; P.d2 = d2; P.C = C; P.L = L; P.metric = metric;
  P.res_stride = (image && res_idx) ? image_channels : 0;
  for (int l = 0; l < 4; ++l) P.res_idx[l] = (res_idx && l < L) ? res_idx[l] : 0;
  P.tgt_stride = L;
  long long count = (long long)B * d0 * d1 * d2 * L;
  if (crop_size) {
    P.c0 = crop_size[0]; P.c1 = crop_size[1]; P.c2 = crop_size[2];
    P.cb0 = crop_begin[0]; P.cb1 = crop_begin[1]; P.cb2 = crop_begin[2];
    count = (long long)B * P.c0 * P.c1 * P.c2 * L;
  } else { P.c0 = P.c1 = P.c2 = 0; P.cb0 = P.cb1 = P.cb2 = 0; }
  P.train = train;
  P.inv_count = 1.0 / (double)count;
  cudaStream_t st = (cudaStream_t)stream;
  SSR_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(double), st));
  const long long nvox = (long long)B * d0 * d1 * d2;
  const bool head_vec = C % 4 == 0 && C <= 128 && ((uintptr_t)feat & 15) == 0 && (!train || ((uintptr_t)dfeat & 15) == 0) &&
                        (!feat_stats || ((uintptr_t)feat_stats & 15) == 0);
  SSR_CHECK_ARG(!feat_stats || head_vec, "folded BatchNorm needs the vectorised head path (C % 4 == 0, C <= 128, aligned)");
  if (head_vec) {
    // fused + coalesced path: loss, prediction, dfeat and the head weight gradient in one pass over feat
    const int nq = C / 4;
    const int gs = nq <= 1 ? 1 : nq <= 2 ? 2 : nq <= 4 ? 4 : nq <= 8 ? 8 : nq <= 16 ? 16 : 32;
    long long nb = (nvox + (256 / gs) - 1) / (256 / gs);
    if (nb > 148 * 8) nb = 148 * 8;
#define SSR_HEAD_LAUNCH(GS_)                                                                                              \
  do {                                                                                                                  \
    if (L == 1) head_loss_vec_kernel<GS_, 1><<<(unsigned)nb, 256, 0, st>>>(feat, feat_stats, w, bias, image, target, pred, dfeat, dw, db, loss, xdot, P); \
    else head_loss_vec_kernel<GS_, 4><<<(unsigned)nb, 256, 0, st>>>(feat, feat_stats, w, bias, image, target, pred, dfeat, dw, db, loss, xdot, P);        \
  } while (0)
    switch (gs) {
      case 1: SSR_HEAD_LAUNCH(1); break;
      case 2: SSR_HEAD_LAUNCH(2); break;
      case 4: SSR_HEAD_LAUNCH(4); break;
      case 8: SSR_HEAD_LAUNCH(8); break;
      case 16: SSR_HEAD_LAUNCH(16); break;
      default: SSR_HEAD_LAUNCH(32); break;
    }
#undef SSR_HEAD_LAUNCH
    SSR_COUNT_LAUNCH();
    SSR_CHECK_LAUNCH();
    return SSR_OK;
  }
  SSR_CHECK_ARG(!xdot, "the BatchNorm-sums output needs the vectorised head path");
  head_loss_kernel<<<grid_for(nvox), 256, 0, st>>>(feat, w, bias, image, target, pred, dfeat, dw, db, loss, P);
  SSR_COUNT_LAUNCH();
  if (train) {
    head_gout_kernel<<<grid_for(nvox), 256, 0, st>>>(pred, image, target, gout_scratch, P);
    SSR_COUNT_LAUNCH();
    dim3 blk(32, 8);
    int g = (int)((nvox + 7) / 8); if (g > 148 * 8) g = 148 * 8;
    head_wgrad_kernel<<<g, blk, 32 * 8 * sizeof(double), st>>>(feat, gout_scratch, nvox, C, L, dw);
    SSR_COUNT_LAUNCH();
  }
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

int ssr_head_loss(const float* feat, const float* feat_stats, const float* w, const float* bias, const float* image,
                  int image_channels,
                  const int* res_idx, const float* target, float* pred, float* dfeat, float* dw, float* db,
                  double* loss, float* gout_scratch, int B, int d0, int d1, int d2, int C, int L, int metric,
                  const int* crop_size, const int* crop_begin, int train, void* stream) {
  return head_loss_impl(feat, feat_stats, w, bias, image, image_channels, res_idx, target, pred, dfeat, dw, db, loss,
                        gout_scratch, B, d0, d1, d2, C, L, metric, crop_size, crop_begin, train, stream, nullptr);
}
// same with feat_stats given (BatchNorm folded into the head) and train = 1; additionally writes the two reductions of
// that BatchNorm's backward, sums2 = [sum_v dy | sum_v dy * xhat] (2*C doubles) for dy = dfeat, so that ssr_bn_bwd_sums
// needs no reduction pass.  db_head must be ZERO on entry (it is read back as sum_v g); xdot_scratch: C*L floats.
int ssr_head_loss_bnsums(const float* feat, const float* feat_stats, const float* w, const float* bias, const float* image,
                         int image_channels, const int* res_idx, const float* target, float* pred, float* dfeat,
                         float* dw, float* db, double* loss, float* gout_scratch, int B, int d0, int d1, int d2, int C,
                         int L, int metric, const int* crop_size, const int* crop_begin, float* xdot_scratch,
                         double* sums2, void* stream) {
  SSR_CHECK_ARG(feat_stats && xdot_scratch && sums2 && db, "bn sums buffers");
  cudaStream_t st = (cudaStream_t)stream;
  SSR_CHECK_CUDA(cudaMemsetAsync(xdot_scratch, 0, (size_t)C * L * sizeof(float), st));
  int rc = head_loss_impl(feat, feat_stats, w, bias, image, image_channels, res_idx, target, pred, dfeat, dw, db, loss,
                          gout_scratch, B, d0, d1, d2, C, L, metric, crop_size, crop_begin, 1, stream, xdot_scratch);
  if (rc) return rc;
  head_bn_sums_kernel<<<(C + 127) / 128, 128, 0, st>>>(w, db, xdot_scratch, C, L, sums2);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

int ssr_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr_t, float beta1, float beta2,
                  float eps, float grad_scale, void* stream) {
  SSR_CHECK_ARG(p && g && m && v && n > 0, "args");
  adam_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr_t, beta1, beta2, eps, grad_scale);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

}  // extern "C"
