// 3x3x3 'same' convolution on the 5th-generation tensor cores (tcgen05, TF32 operands, FP32 accumulate in TMEM).
//
// Replaces cuDNN-via-TF for KL.Conv3D (ext/neuron/models.py:316,444): forward and, with flipped/transposed packed
// weights, the data gradient.  Implicit GEMM, nothing is materialised:
//
//   GEMM   D[M x N] += A[M x K] * B[N x K]^T      M = 128 output voxels (16 along d1 x 8 along d2 at one d0)
//                                                 N = output channels (NT <= 192 per CTA)
//                                                 K = (tap, input-channel chunk of 32)
//   A      activation "slab": one TMA box (32 ch, 8 d2, 18 d1, 1 d0) of the NDHWC tensor -> 144 rows x 128 B in
//          shared memory, SWIZZLE_128B.  The three d1 taps are three UMMA descriptors into the SAME slab
//          (+k1 * 1024 B, swizzle phase preserved), the d2 tap is a shifted TMA box, the d0 tap selects which of
//          the CTA's TZ accumulators the slab feeds.  Out-of-bounds parts of a box are zero filled by TMA =
//          'same' padding, and so are channels beyond C when C is not a multiple of 32.
//   B      packed weights [chunk][k2][k0][k1][Npad][32] (K-major, zero padded), TMA box (32, NT) per tap.
//   D      TZ accumulators of NT fp32 columns in TMEM (TZ*NT <= 512); epilogue = tcgen05.ld -> +bias -> ELU -> store.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue.  mbarrier rings: A slabs (SA stages), B tap groups (2 stages), one "accumulators ready".
#include "common.cuh"
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <string>
#include <cstdlib>

extern "C" int ssr_channel_sum(const float* t, long long nvox, int C, float* out, void* stream);

namespace {

// optional phase timestamps (clock64) of sampled CTAs, for profiling only: ssr_tc_set_debug(buffer)
__device__ long long* g_dbg = nullptr;
#define DBG_STAMP(slot) do { if (dbg) dbg[slot] = clock64(); } while (0)

// ---------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU.  The suspend-time hint lets the
// hardware put the waiting warp to sleep, so warps parked on a barrier do not steal issue slots from the MMA warp.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  for (int spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(0x989680u)
        : "memory");
    if (done) break;
    if (spin > 2000) { printf("conv_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

// one elected lane of a fully converged warp (keeps descriptor operands in uniform registers: no per-MMA waterfall)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when all previously issued tcgen05 ops of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);        // start address
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major) = 16 B
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset = 1024 B
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
// Descriptors as (lo, hi) halves so the issue loop only adds byte offsets (>> 4) to `lo`:
//   lo = start_address>>4 | LBO>>4 << 16 ;  hi = SBO>>4 | version(1) << 14 | layout_type << 29
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
constexpr uint32_t DESC_HI_K_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);        // K-major, SWIZZLE_128B, SBO 1024
constexpr uint32_t DESC_HI_MN_SW128_32B = (512u >> 4) | (1u << 14) | (1u << 29);    // MN-major tf32: SWIZZLE_128B_BASE32B,
                                                                                    // K groups of 4 rows (512 B)
__device__ __forceinline__ void umma_tf32_lh(uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_tf32_lh2(uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc,
                                              uint32_t accumulate) {
  umma_tf32_lh(d_tmem, alo, blo, hi, idesc, accumulate);
}

// NK back-to-back MMAs into one accumulator, both descriptors advancing by `STEP` (in 16-byte units) per MMA, emitted
// as ONE asm block: the issuing warp is alone on its instruction stream, so every extra (dependent, uniform-datapath)
// instruction between two UTCHMMAs costs ~10 cycles; a tight chain sustains the tensor-core rate
// (profiles/r01_mma_issue_microbench.txt: 40 cycles per M128xN32xK8 MMA).
#define SSR_MMA1 "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t"
#define SSR_MMAN(STEP) "add.s64 da, da, " #STEP ";\n\tadd.s64 db, db, " #STEP ";\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, q;\n\t"
#define SSR_MMA_HEAD "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t" \
                     "setp.ne.b32 p, %5, 0;\n\tsetp.ne.b32 q, %6, 0;\n\t"
#define SSR_MMA_ARGS ::"r"(d_tmem), "r"(alo), "r"(blo), "r"(hi), "r"(idesc), "r"(accumulate), "r"(1u) : "memory"

template <int NK>
__device__ __forceinline__ void umma_chain_k(uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc,
                                             uint32_t accumulate) {   // K-major operands: +32 B per K-step
  if (NK == 1) asm volatile(SSR_MMA_HEAD SSR_MMA1 "}" SSR_MMA_ARGS);
  if (NK == 2) asm volatile(SSR_MMA_HEAD SSR_MMA1 SSR_MMAN(2) "}" SSR_MMA_ARGS);
  if (NK == 3) asm volatile(SSR_MMA_HEAD SSR_MMA1 SSR_MMAN(2) SSR_MMAN(2) "}" SSR_MMA_ARGS);
  if (NK == 4) asm volatile(SSR_MMA_HEAD SSR_MMA1 SSR_MMAN(2) SSR_MMAN(2) SSR_MMAN(2) "}" SSR_MMA_ARGS);
}
// same chains with kind::f16 (bf16 operands, 16 elements = 32 B per K-step): the correction term of the compensated forward
#define SSR_MMA1_H "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
#define SSR_MMAN_H(STEP) "add.s64 da, da, " #STEP ";\n\tadd.s64 db, db, " #STEP ";\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, q;\n\t"
template <int NK>
__device__ __forceinline__ void umma_chain_k16(uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc,
                                               uint32_t accumulate) {
  if (NK == 1) asm volatile(SSR_MMA_HEAD SSR_MMA1_H "}" SSR_MMA_ARGS);
  if (NK == 2) asm volatile(SSR_MMA_HEAD SSR_MMA1_H SSR_MMAN_H(2) "}" SSR_MMA_ARGS);
  if (NK == 3) asm volatile(SSR_MMA_HEAD SSR_MMA1_H SSR_MMAN_H(2) SSR_MMAN_H(2) "}" SSR_MMA_ARGS);
  if (NK == 4) asm volatile(SSR_MMA_HEAD SSR_MMA1_H SSR_MMAN_H(2) SSR_MMAN_H(2) SSR_MMAN_H(2) "}" SSR_MMA_ARGS);
}
// nks K-steps of a K-major chunk, TF32 (f16 == false) or bf16 operands
__device__ __forceinline__ void umma_chain_any(bool f16, int nks, uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t idesc,
                                               uint32_t accumulate) {
  if (!f16) {
    if (nks == 4) umma_chain_k<4>(d_tmem, alo, blo, DESC_HI_K_SW128, idesc, accumulate);
    else if (nks == 3) umma_chain_k<3>(d_tmem, alo, blo, DESC_HI_K_SW128, idesc, accumulate);
    else if (nks == 2) umma_chain_k<2>(d_tmem, alo, blo, DESC_HI_K_SW128, idesc, accumulate);
    else umma_chain_k<1>(d_tmem, alo, blo, DESC_HI_K_SW128, idesc, accumulate);
  } else {
    if (nks == 4) umma_chain_k16<4>(d_tmem, alo, blo, DESC_HI_K_SW128, idesc, accumulate);
    else if (nks == 3) umma_chain_k16<3>(d_tmem, alo, blo, DESC_HI_K_SW128, idesc, accumulate);
    else if (nks == 2) umma_chain_k16<2>(d_tmem, alo, blo, DESC_HI_K_SW128, idesc, accumulate);
    else umma_chain_k16<1>(d_tmem, alo, blo, DESC_HI_K_SW128, idesc, accumulate);
  }
}

// 16 K-steps of the weight-gradient tile (MN-major operands: +1024 B per K-step)
__device__ __forceinline__ void umma_chain_mn16(uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(SSR_MMA_HEAD SSR_MMA1 SSR_MMAN(64) SSR_MMAN(64) SSR_MMAN(64) SSR_MMAN(64) SSR_MMAN(64) SSR_MMAN(64)
               SSR_MMAN(64) SSR_MMAN(64) SSR_MMAN(64) SSR_MMAN(64) SSR_MMAN(64) SSR_MMAN(64) SSR_MMAN(64) SSR_MMAN(64)
               SSR_MMAN(64) "}" SSR_MMA_ARGS);
}

// weight-gradient chain with the dY tile stored as rows of 10 voxels (+1280 B per K-step) and X as rows of 8 (+1024 B)
#define SSR_MMAN2(SA_, SB_) "add.s64 da, da, " #SA_ ";\n\tadd.s64 db, db, " #SB_ ";\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, q;\n\t"
__device__ __forceinline__ void umma_chain_mn16_ab(uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t hi, uint32_t idesc,
                                                   uint32_t accumulate) {
  asm volatile(SSR_MMA_HEAD SSR_MMA1 SSR_MMAN2(64, 80) SSR_MMAN2(64, 80) SSR_MMAN2(64, 80) SSR_MMAN2(64, 80)
               SSR_MMAN2(64, 80) SSR_MMAN2(64, 80) SSR_MMAN2(64, 80) SSR_MMAN2(64, 80) SSR_MMAN2(64, 80)
               SSR_MMAN2(64, 80) SSR_MMAN2(64, 80) SSR_MMAN2(64, 80) SSR_MMAN2(64, 80) SSR_MMAN2(64, 80)
               SSR_MMAN2(64, 80) "}" SSR_MMA_ARGS);
}
// same, `n` (< 16) K-steps: tiles that stick out of the volume along d1 skip their all-zero rows
__device__ __forceinline__ void umma_chain_mn_ab(int n, uint32_t d_tmem, uint32_t alo, uint32_t blo, uint32_t hi,
                                                 uint32_t idesc, uint32_t accumulate) {
  for (int i = 0; i < n; ++i) {
    umma_tf32_lh2(d_tmem, alo, blo, hi, idesc, accumulate);
    accumulate = 1u; alo += 64u; blo += 80u;
  }
}

__device__ __forceinline__ uint32_t bf16_bits(float v) {       // round to nearest even bf16 (operands are finite)
  const uint32_t u = __float_as_uint(v);
  return (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
}
// the two bf16 operands of the hybrid forward's correction term for one fp32 activation: hi = bf16(rne_tf32(v)),
// lo = bf16(v - rne_tf32(v))  (identical to tf32_split_bf16_kernel, so producers may emit them in their epilogue)
__device__ __forceinline__ void split_bf16(float v, uint32_t& lo, uint32_t& hi) {
  uint32_t u = __float_as_uint(v);
  u = (u + 0xFFFu + ((u >> 13) & 1u)) & ~0x1FFFu;
  const float h = __uint_as_float(u);
  hi = bf16_bits(h);
  lo = bf16_bits(v - h);
}

// instruction descriptor: D=F32, A=B=TF32, both K-major, M=128, N
__device__ __forceinline__ uint32_t make_idesc_tf32(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// same with A = B = BF16 (kind::f16; cute::UMMA::InstrDescriptor: a_format / b_format 1 = BF16)
__device__ __forceinline__ uint32_t make_idesc_bf16(int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------------------
constexpr bool G_FENCE_IN_LOOP = false;
constexpr int TM1 = 16, TM2 = 8;                  // output tile: 16 (d1) x 8 (d2) voxels = 128 GEMM rows
constexpr int SLAB_ROWS = (TM1 + 2) * TM2;         // 144 rows x 128 B
constexpr int SLAB_BYTES = SLAB_ROWS * 128;        // 18432 (multiple of 1024)
constexpr int SB = 2;                              // B-group stages
constexpr int TC_TAIL_BYTES = 3072;              // conv3d_tc_kernel: 52 barrier words + 640 floats of bias behind the stages
constexpr int MAX_CHUNKS = 40;                   // 12 chunks (384 channels) x 3 terms of a compensated convolution

struct TcGeom {
  int B, D0, D1, D2;
  int Cout;          // real output channels
  int Npad;          // Cout rounded up to 16 (rows per tap in the packed weights)
  int NT;            // output channels per CTA (multiple of 16, <= 192)
  int TZ;            // d0 planes per CTA tile (one MMA-issuing warp each, <= 8)
  int nslot;         // ring of plane accumulators in TMEM (NT columns each): plane zo of the CTA's it-th tile lives in slot
                     // (it * TZ + zo) % nslot, so the epilogue drains plane by plane while the next tile's planes start in
                     // the slots already freed (nslot = 2 TZ is the old "two accumulator sets")
  int KG;            // d0 taps per B group: 3 (all 9 taps resident) or 1
  int SA;            // A stages
  int nchunks;
  int n1tiles, n2tiles, n0tiles, nNtiles;
  int act;
  int accumulate;    // epilogue adds the partial result already in y (before bias / activation)
  int ksplit;        // split-K: the chunk list is cut into ksplit contiguous parts, one CTA tile each; the parts add their raw
                     // partial sums into y with red.global.add (y zeroed or holding earlier partial sums; bias / activation
                     // in a follow-up pass).  For the deep levels: few voxels, long K -- without it <= 120 CTAs walk serial
                     // MMA chains of 1300 - 5000 instructions at N = 32
  int pl_pitch;      // PL mode: d2 pitch of the zero-padded plane (D2 + 2); M windows = n2tiles (n1tiles = 1)
  int pl_slab;       // PL mode: bytes of a slab stage (padded plane + over-read slack, multiple of 1024)
  int pl_tx;         // PL mode: bytes one TMA load of a padded plane writes
  int tmem_cols;
  unsigned char chunk_src[MAX_CHUNKS];    // 0: x1, 1: x2
  unsigned char chunk_ks[MAX_CHUNKS];     // K-steps of 8 channels actually present in the chunk (1..4)
  short chunk_c0[MAX_CHUNKS];             // first channel of the chunk inside its source
  unsigned char chunk_w[MAX_CHUNKS];      // chunk of the PACKED WEIGHTS this chunk multiplies (== index, except in the
                                          // compensated forward, where [x | x_lo | x] meet [w_hi | w_hi | w_lo])
  unsigned char chunk_f16[MAX_CHUNKS];    // 1: the chunk holds 64 bf16 channels (kind::f16 MMAs) instead of 32 fp32 / TF32
};

__device__ __forceinline__ bool slab_needed(const TcGeom& G, int z0, int k0g, int zin) {
  const int din = z0 + k0g + zin - 1;
  if (din < 0 || din >= G.D0) return false;
  for (int kk = 0; kk < G.KG; ++kk) {
    const int zo = zin - kk;
    if (zo >= 0 && zo < G.TZ && z0 + zo < G.D0) return true;
  }
  return false;
}
__device__ __forceinline__ bool group_needed(const TcGeom& G, int z0, int k0g) {
  for (int zin = 0; zin < G.TZ + G.KG - 1; ++zin)
    if (slab_needed(G, z0, k0g, zin)) return true;
  return false;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Column sums over the 32 lanes of a warp of 16 values per lane in 16 shuffles ("transposed" butterfly: every exchange
// halves the number of values a lane keeps): returns the sum of v[ch] over all lanes with
// ch = 8 * bit4(lane) + 4 * bit3 + 2 * bit2 + bit1 (both lanes of a pair hold it).
__device__ __forceinline__ float colsum16_transpose(const float (&v)[16], int lane) {
  float a[8], b[4], c[2];
  bool hi = (lane & 16) != 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float recv = __shfl_xor_sync(0xffffffffu, hi ? v[i] : v[i + 8], 16);
    a[i] = (hi ? v[i + 8] : v[i]) + recv;
  }
  hi = (lane & 8) != 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float recv = __shfl_xor_sync(0xffffffffu, hi ? a[i] : a[i + 4], 8);
    b[i] = (hi ? a[i + 4] : a[i]) + recv;
  }
  hi = (lane & 4) != 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float recv = __shfl_xor_sync(0xffffffffu, hi ? b[i] : b[i + 2], 4);
    c[i] = (hi ? b[i + 2] : b[i]) + recv;
  }
  hi = (lane & 2) != 0;
  const float recv = __shfl_xor_sync(0xffffffffu, hi ? c[0] : c[1], 2);
  float d = (hi ? c[1] : c[0]) + recv;
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}

// Persistent: one CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The TMEM holds a RING of plane
// accumulators (G.nslot slots of NT columns, each with its own full / empty barrier), so the epilogue of a plane (TMEM ->
// registers -> bias/ELU -> global) overlaps the MMAs of the following planes and of the next tile, and the TMA producer
// runs ahead across tile boundaries (no pipeline refill, no per-tile setup).  Deep tiles (TZ up to 8 planes) are what cuts
// the L2 -> shared-memory traffic these kernels are bound by (profiles/r02_conv_schemes_ncu.txt: 6.3 TB/s on every
// variant): the weights of a chunk are streamed once per tile and the d0 halo is 2 planes per TZ.
// EPI (epilogue fusions for the levels below full resolution, the counterpart of the k2n kernel's):
//   1 (data gradient): out *= elu'(elu_h) and dbias[c] += sum_v out[v][c]      -- replaces elu_bwd_kernel
//   2 (forward):       sums[c] += sum_v out, sums[Cout + c] += sum_v out^2    -- replaces colsum2_vec_kernel<0>
// Column sums: transposed warp butterfly per 16-channel block -> shared-memory floats per CTA -> one atomic per channel.
// PL ("plane-linearised", the small deep levels): a tile is a window of 128 consecutive rows of the whole zero-padded
// (D1 + 2) x (D2 + 2) plane, loaded as ONE slab; output voxel (i1, i2) <-> row r = i1 * pitch + i2 and every (k1, k2) tap is
// the row shift k1 * pitch + k2 of the same slab (row-shifted K-major views read consistently with the TMA swizzle,
// profiles/r01_desc_probe_unaligned_views.txt).  With 16 x 8 tiles a 10^3 / 20^3 plane fills 39 % / 52 % of the GEMM rows
// it pays for, a linearised plane 78 %; rows that fall into the halo columns are computed and dropped.
template <int EPI, bool PL = false>
__global__ void __launch_bounds__(416, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_x2,
                 const __grid_constant__ CUtensorMap map_w, const float* __restrict__ bias, float* __restrict__ y,
                 const TcGeom G, const float* __restrict__ elu_h, float* __restrict__ dbias, double* __restrict__ sums) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A stages][B stages][barriers][bias]
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  const int bgroup_bytes = G.KG * 3 * G.NT * 128;
  const int slab_bytes = PL ? G.pl_slab : SLAB_BYTES;
  uint8_t* sB = sA + (size_t)G.SA * slab_bytes;
  uint64_t* bars = (uint64_t*)(sB + (size_t)SB * bgroup_bytes);
  uint64_t* fullA = bars;                          // [SA <= 8]
  uint64_t* emptyA = bars + 8;                     // [8]
  uint64_t* fullB = bars + 16;                     // [SB]
  uint64_t* emptyB = bars + 18;                    // [SB]
  uint64_t* accFull = bars + 20;                   // [nslot <= 10]
  uint64_t* accEmpty = bars + 30;                  // [10]
  uint64_t* kindBar = bars + 40;                   // [TZ <= 8] one per MMA warp: TF32 -> bf16 switch of the hybrid forward
  uint32_t* tmem_slot = (uint32_t*)(bars + 48);
  float* sbias = (float*)(bars + 52);              // Npad floats (<= 576), 16-byte aligned, zero padded
  float* sred = sbias + 640;                       // EPI: 2 x Npad per-CTA channel sums (host reserves the space)
  if (EPI != 0)
    for (int i = threadIdx.x; i < 2 * G.Npad; i += blockDim.x) sred[i] = 0.f;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* dbg = (g_dbg && (blockIdx.x % 2) == 0 && blockIdx.x / 2 < 74) ? g_dbg + (blockIdx.x / 2) * 16 : nullptr;
  if (threadIdx.x == 0) DBG_STAMP(0);

  if (threadIdx.x == 0) {
    // NW = G.TZ MMA warps: each of them releases every ring stage and signals every accumulator set
    for (int i = 0; i < G.SA; ++i) { mbar_init(fullA + i, 1); mbar_init(emptyA + i, G.TZ); }
    for (int i = 0; i < SB; ++i) { mbar_init(fullB + i, 1); mbar_init(emptyB + i, G.TZ); }
    for (int i = 0; i < G.nslot; ++i) { mbar_init(accFull + i, 1); mbar_init(accEmpty + i, 4); }   // owner warp / 4 epilogue warps
    for (int i = 0; i < 8; ++i) mbar_init(kindBar + i, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  for (int i = threadIdx.x; i < G.Npad; i += blockDim.x) sbias[i] = (bias && i < G.Cout) ? bias[i] : 0.f;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x1);
    tma_prefetch_desc(&map_x2);
    tma_prefetch_desc(&map_w);
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)G.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) DBG_STAMP(1);
  const int ntiles = G.B * G.n0tiles * G.n1tiles * G.n2tiles * G.nNtiles * G.ksplit;

#define DECODE_TILE(tile)                                                       \
  int t_ = (tile);                                                               \
  const int kpart = t_ % G.ksplit; t_ /= G.ksplit;                               \
  const int ch_lo = kpart * G.nchunks / G.ksplit, ch_hi = (kpart + 1) * G.nchunks / G.ksplit; \
  const int nt = t_ % G.nNtiles; t_ /= G.nNtiles;                                \
  const int t2 = t_ % G.n2tiles; t_ /= G.n2tiles;                                \
  const int t1 = t_ % G.n1tiles; t_ /= G.n1tiles;                                \
  const int t0 = t_ % G.n0tiles;                                                 \
  const int b = t_ / G.n0tiles;                                                  \
  const int x0 = t2 * TM2, y0 = t1 * TM1, z0 = t0 * G.TZ, n0 = nt * G.NT;        \
  const int nz = min(G.TZ, G.D0 - z0);                                           \
  (void)x0; (void)y0; (void)n0; (void)b; (void)nz; (void)ch_lo; (void)ch_hi;

  // Closed-form slab ranges (identical in the producer and the MMA warp; no per-slab predicate evaluation in the issue
  // loop).  nz = valid output planes of the tile.  For the d0-tap group starting at k0g, slab index zin (input plane
  // z0 + k0g + zin - 1) contributes to accumulators zo = zin - kk, kk in [kk_lo, kk_hi]:
  //   zin in [zin_lo, zin_hi),  zin_lo = 1 iff the first plane would be -1,  zin_hi clipped at the volume end.
  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int sa = 0, pa = 0, sb = 0, pb = 0;
      long long wait_empty = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        DECODE_TILE(tile)
        for (int ch = ch_lo; ch < ch_hi; ++ch) {
          const CUtensorMap* mx = G.chunk_src[ch] ? &map_x2 : &map_x1;
          const int c0 = G.chunk_c0[ch];
          for (int k2 = 0; k2 < 3; ++k2) {
            for (int k0g = 0; k0g < 3; k0g += G.KG) {
              const int zin_lo = (z0 + k0g == 0) ? 1 : 0;
              const int zin_hi = min(nz + G.KG - 1, G.D0 - (z0 + k0g - 1));
              if (zin_lo >= zin_hi) continue;
              mbar_wait(emptyB + sb, pb ^ 1);
              mbar_expect_tx(fullB + sb, (uint32_t)bgroup_bytes);
              for (int kk = 0; kk < G.KG; ++kk)
                for (int k1 = 0; k1 < 3; ++k1) {
                  const int row = (((G.chunk_w[ch] * 3 + k2) * 3 + (k0g + kk)) * 3 + k1) * G.Npad + n0;
                  tma_load_2d(&map_w, fullB + sb, sB + (size_t)sb * bgroup_bytes + (size_t)(kk * 3 + k1) * G.NT * 128, 0, row);
                }
              if (++sb == SB) { sb = 0; pb ^= 1; }
              for (int zin = zin_lo; zin < zin_hi; ++zin) {
                { const long long w0 = dbg ? clock64() : 0; mbar_wait(emptyA + sa, pa ^ 1); if (dbg) wait_empty += clock64() - w0; }
                if (PL) {      // the whole zero-padded plane (the d2 tap is a row shift of the operand view, not of the box)
                  mbar_expect_tx(fullA + sa, (uint32_t)G.pl_tx);
                  tma_load_5d(mx, fullA + sa, sA + (size_t)sa * slab_bytes, c0, -1, -1, z0 + k0g + zin - 1, b);
                } else {
                  mbar_expect_tx(fullA + sa, SLAB_BYTES);
                  tma_load_5d(mx, fullA + sa, sA + (size_t)sa * SLAB_BYTES, c0, x0 + k2 - 1, y0 - 1, z0 + k0g + zin - 1, b);
                }
                if (++sa == G.SA) { sa = 0; pa ^= 1; }
              }
            }
          }
        }
      }
      DBG_STAMP(2);
      if (dbg) dbg[8] = wait_empty;
    }
  } else if (warp <= G.TZ) {
    // ================================ MMA issuers: warp w owns accumulator (output plane) zo = w - 1 ================
    // The thread that issues tcgen05.mma stalls while the tensor pipe is busy (shallow MMA queue), so with a single
    // issuer every scalar instruction between two MMAs is exposed (measured 68 cycles per N=32 MMA against the 40-cycle
    // pipe rate).  One issuing warp per accumulator: the chains are independent, so one warp's waits / descriptor
    // arithmetic hide behind the other warps' MMAs.  Every warp walks all slabs (so that the ring barriers see NW
    // arrivals per stage) and issues only the d0 tap kk = zin - zo that lands in its own accumulator.
    {
      const int zo = warp - 1;
      const uint32_t idesc32 = make_idesc_tf32(G.NT), idesc16 = make_idesc_bf16(G.NT);
      const int KG = G.KG, SA = G.SA, nchunks = G.nchunks, D0 = G.D0;
      const uint32_t NT = (uint32_t)G.NT;
      const uint32_t btile16 = (NT * 128u) >> 4;
      const uint32_t a_base = desc_lo(smem_u32(sA), 16), b_base = desc_lo(smem_u32(sB), 16);
      int sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
      uint32_t kind_phase = 0;
      long long wait_a = 0, wait_b = 0;
      if (warp == 1 && lane == 0) DBG_STAMP(3);
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        DECODE_TILE(tile)
        const int gpl = it * G.TZ + zo, slot = gpl % G.nslot;     // this plane's accumulator slot and how often it was used
        const uint32_t use = (uint32_t)(gpl / G.nslot);
        const uint32_t dcol = tmem_base + (uint32_t)slot * NT;
        const bool active = zo < nz;
        // epilogue has drained this slot.  (Deferring this wait to the warp's first MMA of the tile -- so that warps without
        // a tap in the first slabs keep walking the ring -- measured 5 % SLOWER on every generic layer: forward 5.04 -> 5.28
        // ms, data gradient 3.00 -> 3.16 ms per step, scripts/gpu/r02_q.sh against r02_s.sh.)
        mbar_wait(accEmpty + slot, (use & 1u) ^ 1u);
        tc_fence_after();
        uint32_t acc = 0u;                         // first MMA of the tile into this accumulator overwrites
        bool prev_f16 = false;
        for (int ch = ch_lo; ch < ch_hi; ++ch) {
          const int nks = G.chunk_ks[ch];
          const bool f16 = G.chunk_f16[ch] != 0;
          const uint32_t idesc = f16 ? idesc16 : idesc32;
          if (f16 != prev_f16) {
            // tcgen05.mma instructions of DIFFERENT kinds are not ordered against each other: the TF32 MMAs into this
            // accumulator must have completed before the first bf16 MMA reads it (one bubble per tile and warp; the
            // other MMA warps keep the tensor pipe busy meanwhile).  Nothing to wait for when this warp has not issued
            // into the accumulator yet (a split-K part that starts with the bf16 chunks).
            if (acc != 0u) {
              if (elect_one()) umma_commit(kindBar + zo);
              __syncwarp();
              mbar_wait(kindBar + zo, kind_phase);
              kind_phase ^= 1u;
              tc_fence_after();
            }
            prev_f16 = f16;
          }
          for (int k2 = 0; k2 < 3; ++k2) {
            for (int k0g = 0; k0g < 3; k0g += KG) {
              const int zin_lo = (z0 + k0g == 0) ? 1 : 0;
              const int zin_hi = min(nz + KG - 1, D0 - (z0 + k0g - 1));
              if (zin_lo >= zin_hi) continue;
              { const long long w0 = dbg ? clock64() : 0; mbar_wait(fullB + sb, pb); if (dbg) wait_b += clock64() - w0; }
              const uint32_t blo0 = b_base + (uint32_t)sb * ((uint32_t)bgroup_bytes >> 4);
              for (int zin = zin_lo; zin < zin_hi; ++zin) {
                { const long long w0 = dbg ? clock64() : 0; mbar_wait(fullA + sa, pa); if (dbg) wait_a += clock64() - w0; }
                const int kk = zin - zo;
                if (active && kk >= 0 && kk < KG) {            // warp-uniform
                  if (elect_one()) {
                    uint32_t alo = a_base + (uint32_t)sa * ((uint32_t)slab_bytes >> 4);
                    if (PL) alo += (uint32_t)(t2 * 128 + k2) * 8u;      // window start + d2 tap, in 128-byte rows
                    const uint32_t k1_step = PL ? (uint32_t)G.pl_pitch * 8u : (uint32_t)(TM2 * 128 >> 4);
                    uint32_t blo = blo0 + (uint32_t)(kk * 3) * btile16;
#pragma unroll
                    for (int k1 = 0; k1 < 3; ++k1) {
                      umma_chain_any(f16, nks, dcol, alo, blo, idesc, acc);
                      acc = 1u;
                      alo += k1_step;
                      blo += btile16;
                    }
                    umma_commit(emptyA + sa);        // this warp's reads of the slab are done once these MMAs complete
                  }
                  acc = 1u;
                } else if (lane == 0) {
                  mbar_arrive(emptyA + sa);          // not my tap: release immediately
                }
                __syncwarp();
                if (++sa == SA) { sa = 0; pa ^= 1; }
              }
              if (elect_one()) umma_commit(emptyB + sb);   // arrives when this warp's MMAs on the group are complete
              __syncwarp();
              if (++sb == SB) { sb = 0; pb ^= 1; }
            }
          }
        }
        if (elect_one()) umma_commit(accFull + slot);            // (an inactive plane of a clipped tile arrives at once)
        __syncwarp();
        if (warp == 1 && lane == 0 && it == 0) { DBG_STAMP(4); if (dbg) { dbg[9] = wait_a; dbg[10] = wait_b; } }
      }
      if (warp == 1 && lane == 0) DBG_STAMP(11);
    }
  } else {
    // ================================ epilogue (last four warps) ================================
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                    // GEMM row = voxel inside the tile
    const bool vec_ok = (G.Cout & 3) == 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      DECODE_TILE(tile)
      int i1 = y0 + (r >> 3), i2 = x0 + (r & 7);
      if (PL) { const int rr = t2 * 128 + r; i1 = rr / G.pl_pitch; i2 = rr - i1 * G.pl_pitch; }   // halo columns: i2 >= D2
      const bool vox_ok = i1 < G.D1 && i2 < G.D2;
      for (int zo = 0; zo < G.TZ; ++zo) {
        const int gpl = it * G.TZ + zo, slot = gpl % G.nslot;
        mbar_wait(accFull + slot, (uint32_t)(gpl / G.nslot) & 1u);
        tc_fence_after();
        if (warp == G.TZ + 1 && lane == 0 && it == 0 && zo == 0) DBG_STAMP(5);
        const uint32_t acc_base = tmem_base + (uint32_t)(slot * G.NT);
        const int i0 = z0 + zo;
        float* orow = y + ((((long long)b * G.D0 + i0) * G.D1 + i1) * G.D2 + i2) * G.Cout + n0;
        for (int cb = 0; cb < (zo < nz ? G.NT : 0); cb += 16) {       // planes past the volume end: nothing to drain
          uint32_t v[16];
          tmem_ld16(acc_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cb, v);
          tmem_ld_wait();
          float o[16];
          const float4* bs = reinterpret_cast<const float4*>(sbias + n0 + cb);      // zero padded, 16-byte aligned
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 bb = bs[e4];
            o[e4 * 4 + 0] = __uint_as_float(v[e4 * 4 + 0]) + bb.x;
            o[e4 * 4 + 1] = __uint_as_float(v[e4 * 4 + 1]) + bb.y;
            o[e4 * 4 + 2] = __uint_as_float(v[e4 * 4 + 2]) + bb.z;
            o[e4 * 4 + 3] = __uint_as_float(v[e4 * 4 + 3]) + bb.w;
          }
          if (G.accumulate && vox_ok) {        // partial sums written earlier (the upsampled part of a decoder convolution)
            const int nv = G.Cout - (n0 + cb);
            if (nv >= 16 && vec_ok) {
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                const float4 pv = *reinterpret_cast<const float4*>(orow + cb + e);
                o[e] += pv.x; o[e + 1] += pv.y; o[e + 2] += pv.z; o[e + 3] += pv.w;
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (e < nv) o[e] += orow[cb + e];
            }
          }
          if (G.act) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {             // ELU without a branch: exp(min(x,0)) - 1 selected where x <= 0
              const float neg = __expf(fminf(o[e], 0.f)) - 1.f;
              o[e] = o[e] > 0.f ? o[e] : neg;
            }
          }
          if constexpr (EPI != 0) {
            const int nv = G.Cout - (n0 + cb);
            if constexpr (EPI == 1) {                  // elu'(pre) from the ELU output h: 1 where h > 0, h + 1 elsewhere
              const float* hrow = elu_h + (orow - y);
              if (vox_ok) {
                if (nv >= 16 && vec_ok) {
#pragma unroll
                  for (int e = 0; e < 16; e += 4) {
                    const float4 hv = __ldg(reinterpret_cast<const float4*>(hrow + cb + e));
                    o[e] *= hv.x > 0.f ? 1.f : hv.x + 1.f; o[e + 1] *= hv.y > 0.f ? 1.f : hv.y + 1.f;
                    o[e + 2] *= hv.z > 0.f ? 1.f : hv.z + 1.f; o[e + 3] *= hv.w > 0.f ? 1.f : hv.w + 1.f;
                  }
                } else {
#pragma unroll
                  for (int e = 0; e < 16; ++e)
                    if (e < nv) { const float hv = __ldg(hrow + cb + e); o[e] *= hv > 0.f ? 1.f : hv + 1.f; }
                }
              }
            }
            float sv[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) sv[e] = (vox_ok && e < nv) ? o[e] : 0.f;
            const int chn = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            const float cs = colsum16_transpose(sv, lane);
            if (!(lane & 1)) atomicAdd(sred + n0 + cb + chn, cs);
            if constexpr (EPI == 2) {
#pragma unroll
              for (int e = 0; e < 16; ++e) sv[e] *= sv[e];
              const float cq = colsum16_transpose(sv, lane);
              if (!(lane & 1)) atomicAdd(sred + G.Npad + n0 + cb + chn, cq);
            }
          }
          if (!vox_ok) continue;
          const int nvalid = G.Cout - (n0 + cb);       // channels of this 16-block that exist
          if (G.ksplit > 1) {                          // split-K part: add the raw partial sums (host: no bias / act / EPI)
            if (nvalid >= 16 && vec_ok) {
#pragma unroll
              for (int e = 0; e < 16; e += 4) red_add_v4(orow + cb + e, o[e], o[e + 1], o[e + 2], o[e + 3]);
            } else {
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (e < nvalid) atomicAdd(orow + cb + e, o[e]);
            }
            continue;
          }
          if (nvalid >= 16 && vec_ok) {
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              *reinterpret_cast<float4*>(orow + cb + e) = make_float4(o[e], o[e + 1], o[e + 2], o[e + 3]);
          } else if (nvalid >= 8 && vec_ok) {
            *reinterpret_cast<float4*>(orow + cb) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(orow + cb + 4) = make_float4(o[4], o[5], o[6], o[7]);
#pragma unroll
            for (int e = 8; e < 16; ++e)
              if (e < nvalid) orow[cb + e] = o[e];
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (e < nvalid) orow[cb + e] = o[e];
          }
        }
        // slot drained: hand it back to the MMA warp that uses it next
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(accEmpty + slot);
      }
      if (warp == G.TZ + 1 && lane == 0 && it == 0) DBG_STAMP(6);
    }
  }
#undef DECODE_TILE
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)G.tmem_cols);
  if (threadIdx.x == 0) DBG_STAMP(7);
  if constexpr (EPI == 1) {
    for (int c = threadIdx.x; c < G.Cout; c += blockDim.x) atomicAdd(dbias + c, sred[c]);
  }
  if constexpr (EPI == 2) {
    for (int c = threadIdx.x; c < G.Cout; c += blockDim.x) {
      atomicAdd(sums + c, (double)sred[c]);
      atomicAdd(sums + G.Cout + c, (double)sred[G.Npad + c]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Convolution over a nearest-neighbour 2x upsampled tensor WITHOUT the upsampled tensor (decoder levels:
// UpSampling3D -> concatenate -> Conv3D, ext/neuron/models.py:425-446).  For an output voxel o = 2 i + p (parity p per
// axis) the three taps of an axis read up[o - 1], up[o], up[o + 1] = low[i - 1 + p], low[i], low[i + p]: two low-res
// voxels, so per parity class the 3x3x3 kernel collapses to an effective 2x2x2 one (sums of the original taps, formed in
// fp32 by up_weights_kernel): 8 taps instead of 27 on the upsampled part of K, i.e. 3.4x fewer MMAs, and the 8x larger
// upsampled tensor is neither written nor read.
//   MODE 1 (forward):   y[2 i + p][co] = sum_{k in taps(p), ci} low[i + k - 1][ci] * Weff_p[k][ci][co]
//                       tile index carries the parity; taps(p) per axis = {p, p + 1}; raw partial sums (no bias /
//                       activation) go to the parity sub-lattice of the full-resolution output; the skip part of the
//                       concatenation is added by a normal convolution with `accumulate`.
//   MODE 2 (gradient):  dlow[j][ci] = sum_p sum_{k'} dy_p[j + k' - 1][co] * Weff_p[2 - k'][ci][co],  dy_p[i] = dy[2 i + p]
//                       the 8 parity classes are extra K chunks: each has its own strided TMA view of dy (doubled global
//                       strides, base offset p), its own taps {1 - p, 2 - p} per axis and its own packed weights.  This
//                       is the gradient w.r.t. the LOW-resolution tensor: the 2x2x2 sum of UpSampling3D's backward is
//                       part of the GEMM.
// Same tiling / pipeline / warp roles as conv3d_tc_kernel (its protocol is kept line by line).
// ---------------------------------------------------------------------------------------------------------
constexpr int UP_MAX_CHUNKS = 24;                 // 8 chunks (256 channels) x 3 terms of a compensated convolution

struct UpGeom {
  int B, D0, D1, D2;     // LOW-resolution grid = GEMM rows
  int Cout, Npad, NT, TZ, KG, SA;
  int nchunks;           // 32-channel chunks of the input (per parity class in MODE 2)
  int n1tiles, n2tiles, n0tiles, nNtiles;
  int tmem_cols;
  int par_rows;          // rows of the packed weights per parity class (= nchunks * 27 * Npad)
  unsigned char chunk_ks[UP_MAX_CHUNKS];
  short chunk_c0[UP_MAX_CHUNKS];
  unsigned char chunk_src[UP_MAX_CHUNKS];   // MODE 1: tensor map of the chunk (0: x, 1: its TF32 residual x_lo)
  unsigned char chunk_w[UP_MAX_CHUNKS];     // chunk of the packed weights (see TcGeom::chunk_w)
  unsigned char chunk_f16[UP_MAX_CHUNKS];   // 1: 64 bf16 channels, kind::f16 MMAs (see TcGeom::chunk_f16)
};
struct UpMaps { CUtensorMap x[8]; };

template <int MODE>
__global__ void __launch_bounds__(288, 1)
conv3d_tc_up_kernel(const __grid_constant__ UpMaps maps, const __grid_constant__ CUtensorMap map_w,
                    float* __restrict__ y, const UpGeom G) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  const int bgroup_bytes = G.KG * 3 * G.NT * 128;
  uint8_t* sB = sA + (size_t)G.SA * SLAB_BYTES;
  uint64_t* bars = (uint64_t*)(sB + (size_t)SB * bgroup_bytes);
  uint64_t* fullA = bars;
  uint64_t* emptyA = bars + G.SA;
  uint64_t* fullB = bars + 2 * G.SA;
  uint64_t* emptyB = fullB + SB;
  uint64_t* accFull = emptyB + SB;                 // [2]
  uint64_t* accEmpty = accFull + 2;                // [2]
  uint32_t* tmem_slot = (uint32_t*)(accEmpty + 2);
  uint64_t* kindBar = bars + 26;                   // [TZ <= 4]: TF32 -> bf16 switch of the hybrid forward (see conv3d_tc_kernel)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < G.SA; ++i) { mbar_init(fullA + i, 1); mbar_init(emptyA + i, G.TZ); }
    for (int i = 0; i < SB; ++i) { mbar_init(fullB + i, 1); mbar_init(emptyB + i, G.TZ); }
    for (int i = 0; i < 2; ++i) { mbar_init(accFull + i, G.TZ); mbar_init(accEmpty + i, 4); }
    for (int i = 0; i < 4; ++i) mbar_init(kindBar + i, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < (MODE == 2 ? 8 : 2); ++i) tma_prefetch_desc(&maps.x[i]);
    tma_prefetch_desc(&map_w);
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)G.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr int NPT = MODE == 1 ? 8 : 1;           // parity classes folded into the tile index
  constexpr int NPK = MODE == 2 ? 8 : 1;           // parity classes folded into K
  const int ntiles = G.B * G.n0tiles * G.n1tiles * G.n2tiles * G.nNtiles * NPT;
  const uint32_t set_cols = (uint32_t)(G.TZ * G.NT);

#define UP_DECODE_TILE(tile)                                                    \
  int t_ = (tile);                                                               \
  const int tpar = MODE == 1 ? (t_ & 7) : 0; if (MODE == 1) t_ >>= 3;            \
  const int nt = t_ % G.nNtiles; t_ /= G.nNtiles;                                \
  const int t2 = t_ % G.n2tiles; t_ /= G.n2tiles;                                \
  const int t1 = t_ % G.n1tiles; t_ /= G.n1tiles;                                \
  const int t0 = t_ % G.n0tiles;                                                 \
  const int b = t_ / G.n0tiles;                                                  \
  const int x0 = t2 * TM2, y0 = t1 * TM1, z0 = t0 * G.TZ, n0 = nt * G.NT;        \
  const int nz = min(G.TZ, G.D0 - z0);                                           \
  (void)x0; (void)y0; (void)n0; (void)b; (void)nz; (void)tpar;
  // taps of parity class `par` per axis (bit 2: d0, bit 1: d1, bit 0: d2): MODE 1 {p, p + 1}, MODE 2 {1 - p, 2 - p}
#define UP_TAPS(par)                                                             \
  const int p0_ = ((par) >> 2) & 1, p1_ = ((par) >> 1) & 1, p2_ = (par) & 1;     \
  const int k0lo = MODE == 1 ? p0_ : 1 - p0_, k1lo = MODE == 1 ? p1_ : 1 - p1_, k2lo = MODE == 1 ? p2_ : 1 - p2_;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int sa = 0, pa = 0, sb = 0, pb = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        UP_DECODE_TILE(tile)
        for (int pk = 0; pk < NPK; ++pk) {
          const int par = MODE == 1 ? tpar : pk;
          UP_TAPS(par)
          for (int ch = 0; ch < G.nchunks; ++ch) {
            const CUtensorMap* mx = &maps.x[MODE == 2 ? pk : G.chunk_src[ch]];
            const int c0 = G.chunk_c0[ch];
            const int brow = par * G.par_rows + G.chunk_w[ch] * 27 * G.Npad + n0;
            for (int k2 = k2lo; k2 <= k2lo + 1; ++k2) {
              for (int k0g = 0; k0g < 3; k0g += G.KG) {
                const int kk_lo = max(k0lo - k0g, 0), kk_hi = min(k0lo + 1 - k0g, G.KG - 1);
                if (kk_lo > kk_hi) continue;
                const int zin_lo = max(kk_lo, (z0 + k0g == 0) ? 1 : 0);
                const int zin_hi = min(nz + kk_hi, G.D0 - (z0 + k0g - 1));
                if (zin_lo >= zin_hi) continue;
                mbar_wait(emptyB + sb, pb ^ 1);
                mbar_expect_tx(fullB + sb, (uint32_t)((kk_hi - kk_lo + 1) * 2 * G.NT * 128));
                for (int kk = kk_lo; kk <= kk_hi; ++kk)
                  for (int k1 = k1lo; k1 <= k1lo + 1; ++k1)
                    tma_load_2d(&map_w, fullB + sb, sB + (size_t)sb * bgroup_bytes + (size_t)(kk * 3 + k1) * G.NT * 128, 0,
                                brow + ((k2 * 3 + (k0g + kk)) * 3 + k1) * G.Npad);
                if (++sb == SB) { sb = 0; pb ^= 1; }
                for (int zin = zin_lo; zin < zin_hi; ++zin) {
                  mbar_wait(emptyA + sa, pa ^ 1);
                  mbar_expect_tx(fullA + sa, SLAB_BYTES);
                  tma_load_5d(mx, fullA + sa, sA + (size_t)sa * SLAB_BYTES, c0, x0 + k2 - 1, y0 - 1, z0 + k0g + zin - 1, b);
                  if (++sa == G.SA) { sa = 0; pa ^= 1; }
                }
              }
            }
          }
        }
      }
    }
  } else if (warp <= G.TZ) {
    // ================================ MMA issuers: warp w owns accumulator (output plane) zo = w - 1 ================
    const int zo = warp - 1;
    const uint32_t idesc32 = make_idesc_tf32(G.NT), idesc16 = make_idesc_bf16(G.NT);
    const int KG = G.KG, SA = G.SA, nchunks = G.nchunks, D0 = G.D0;
    const uint32_t NT = (uint32_t)G.NT;
    const uint32_t btile16 = (NT * 128u) >> 4;
    const uint32_t a_base = desc_lo(smem_u32(sA), 16), b_base = desc_lo(smem_u32(sB), 16);
    int sa = 0, pa = 0, sb = 0, pb = 0, it = 0;
    uint32_t kind_phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      UP_DECODE_TILE(tile)
      const int set = it & 1;
      const uint32_t dcol = tmem_base + (uint32_t)set * set_cols + (uint32_t)zo * NT;
      const bool active = zo < nz;
      mbar_wait(accEmpty + set, ((it >> 1) & 1) ^ 1);          // epilogue has drained this accumulator set
      tc_fence_after();
      uint32_t acc = 0u;                         // first MMA of the tile into this accumulator overwrites
      bool prev_f16 = false;
      for (int pk = 0; pk < NPK; ++pk) {
        const int par = MODE == 1 ? tpar : pk;
        UP_TAPS(par)
        (void)k2lo;
        for (int ch = 0; ch < nchunks; ++ch) {
          const int nks = G.chunk_ks[ch];
          const bool f16 = G.chunk_f16[ch] != 0;
          const uint32_t idesc = f16 ? idesc16 : idesc32;
          if (f16 != prev_f16) {                   // kind switch: see conv3d_tc_kernel
            if (elect_one()) umma_commit(kindBar + zo);
            __syncwarp();
            mbar_wait(kindBar + zo, kind_phase);
            kind_phase ^= 1u;
            tc_fence_after();
            prev_f16 = f16;
          }
          for (int k2i = 0; k2i < 2; ++k2i) {
            for (int k0g = 0; k0g < 3; k0g += KG) {
              const int kk_lo = max(k0lo - k0g, 0), kk_hi = min(k0lo + 1 - k0g, KG - 1);
              if (kk_lo > kk_hi) continue;
              const int zin_lo = max(kk_lo, (z0 + k0g == 0) ? 1 : 0);
              const int zin_hi = min(nz + kk_hi, D0 - (z0 + k0g - 1));
              if (zin_lo >= zin_hi) continue;
              mbar_wait(fullB + sb, pb);
              const uint32_t blo0 = b_base + (uint32_t)sb * ((uint32_t)bgroup_bytes >> 4);
              for (int zin = zin_lo; zin < zin_hi; ++zin) {
                mbar_wait(fullA + sa, pa);
                const int kk = zin - zo;
                if (active && kk >= kk_lo && kk <= kk_hi) {            // warp-uniform
                  if (elect_one()) {
                    uint32_t alo = a_base + (uint32_t)sa * (SLAB_BYTES >> 4) + (uint32_t)k1lo * (uint32_t)(TM2 * 128 >> 4);
                    uint32_t blo = blo0 + (uint32_t)(kk * 3 + k1lo) * btile16;
#pragma unroll
                    for (int k1i = 0; k1i < 2; ++k1i) {
                      umma_chain_any(f16, nks, dcol, alo, blo, idesc, acc);
                      acc = 1u;
                      alo += (uint32_t)(TM2 * 128 >> 4);
                      blo += btile16;
                    }
                    umma_commit(emptyA + sa);        // this warp's reads of the slab are done once these MMAs complete
                  }
                  acc = 1u;
                } else if (lane == 0) {
                  mbar_arrive(emptyA + sa);          // not my tap: release immediately
                }
                __syncwarp();
                if (++sa == SA) { sa = 0; pa ^= 1; }
              }
              if (elect_one()) umma_commit(emptyB + sb);   // arrives when this warp's MMAs on the group are complete
              __syncwarp();
              if (++sb == SB) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
      if (elect_one()) umma_commit(accFull + set);
      __syncwarp();
    }
  } else {
    // ================================ epilogue (last four warps) ================================
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                    // GEMM row = voxel inside the tile
    const bool vec_ok = (G.Cout & 3) == 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      UP_DECODE_TILE(tile)
      const int set = it & 1;
      const uint32_t acc_base = tmem_base + (uint32_t)set * set_cols;
      const int i1 = y0 + (r >> 3), i2 = x0 + (r & 7);
      mbar_wait(accFull + set, (it >> 1) & 1);
      tc_fence_after();
      const bool vox_ok = i1 < G.D1 && i2 < G.D2;
      for (int zo = 0; zo < nz; ++zo) {
        const int i0 = z0 + zo;
        float* orow;
        if (MODE == 1) {      // parity sub-lattice of the 2x finer output grid
          const long long o0 = 2 * i0 + ((tpar >> 2) & 1), o1 = 2 * i1 + ((tpar >> 1) & 1), o2 = 2 * i2 + (tpar & 1);
          orow = y + ((((long long)b * (2 * G.D0) + o0) * (2 * G.D1) + o1) * (2 * G.D2) + o2) * G.Cout + n0;
        } else {
          orow = y + ((((long long)b * G.D0 + i0) * G.D1 + i1) * G.D2 + i2) * G.Cout + n0;
        }
        for (int cb = 0; cb < G.NT; cb += 16) {
          uint32_t v[16];
          tmem_ld16(acc_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(zo * G.NT + cb), v);
          tmem_ld_wait();
          if (!vox_ok) continue;
          const int nvalid = G.Cout - (n0 + cb);       // channels of this 16-block that exist
          if (nvalid >= 16 && vec_ok) {
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              *reinterpret_cast<float4*>(orow + cb + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                      __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
          } else if (nvalid >= 8 && vec_ok) {
            *reinterpret_cast<float4*>(orow + cb) = make_float4(__uint_as_float(v[0]), __uint_as_float(v[1]),
                                                                __uint_as_float(v[2]), __uint_as_float(v[3]));
            *reinterpret_cast<float4*>(orow + cb + 4) = make_float4(__uint_as_float(v[4]), __uint_as_float(v[5]),
                                                                    __uint_as_float(v[6]), __uint_as_float(v[7]));
#pragma unroll
            for (int e = 8; e < 16; ++e)
              if (e < nvalid) orow[cb + e] = __uint_as_float(v[e]);
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (e < nvalid) orow[cb + e] = __uint_as_float(v[e]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(accEmpty + set);
    }
  }
#undef UP_DECODE_TILE
#undef UP_TAPS
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)G.tmem_cols);
}

// effective weights of the upsampled part of a decoder convolution: for parity class p (bit 2: d0, bit 1: d1, bit 0: d2)
//   weff[p][k0][k1][k2][ci][co] = sum over the original taps t_a that land on effective tap k_a under parity p_a
//       p_a = 0:  k = 0 <- {0},  k = 1 <- {1, 2},  k = 2 <- {}        p_a = 1:  k = 0 <- {},  k = 1 <- {0, 1},  k = 2 <- {2}
//   of w[t0][t1][t2][Cskip + ci][co];  wskip[t][ci][co] = w[t][ci][co] for ci < Cskip (contiguous copy of the skip part).
__global__ void up_weights_kernel(const float* __restrict__ w, int Cskip, int Cup, int Cout, float* __restrict__ wskip,
                                  float* __restrict__ weff) {
  const int Cin = Cskip + Cup;
  const long long n_eff = 8LL * 27 * Cup * Cout, n_skip = 27LL * Cskip * Cout;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n_eff + n_skip;
       t += (long long)gridDim.x * blockDim.x) {
    if (t >= n_eff) {
      const long long u = t - n_eff;
      const int co = (int)(u % Cout);
      const int ci = (int)((u / Cout) % Cskip);
      const int tap = (int)(u / ((long long)Cout * Cskip));
      wskip[u] = w[((long long)tap * Cin + ci) * Cout + co];
      continue;
    }
    long long r = t;
    const int co = (int)(r % Cout); r /= Cout;
    const int ci = (int)(r % Cup); r /= Cup;
    const int k2 = (int)(r % 3); r /= 3;
    const int k1 = (int)(r % 3); r /= 3;
    const int k0 = (int)(r % 3);
    const int par = (int)(r / 3);
    const int pp[3] = {(par >> 2) & 1, (par >> 1) & 1, par & 1}, kk[3] = {k0, k1, k2};
    int lo[3], hi[3];
    bool any = true;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (pp[a] == 0) { lo[a] = kk[a] == 0 ? 0 : 1; hi[a] = kk[a] == 0 ? 0 : (kk[a] == 1 ? 2 : -1); }
      else { lo[a] = kk[a] == 2 ? 2 : 0; hi[a] = kk[a] == 2 ? 2 : (kk[a] == 1 ? 1 : -1); }
      if (hi[a] < lo[a]) any = false;
    }
    float acc = 0.f;
    if (any)
      for (int t0 = lo[0]; t0 <= hi[0]; ++t0)
        for (int t1 = lo[1]; t1 <= hi[1]; ++t1)
          for (int t2 = lo[2]; t2 <= hi[2]; ++t2)
            acc += w[((long long)((t0 * 3 + t1) * 3 + t2) * Cin + Cskip + ci) * Cout + co];
    weff[t] = acc;
  }
}

// dw[t][cin_off + ci][co] += sum_p geff[p][k(p, t)][ci][co]: the original tap t_a is part of the effective tap
// k_a = (p_a == 0 ? (t_a == 0 ? 0 : 1) : (t_a == 2 ? 2 : 1)) of parity class p_a (transpose of up_weights_kernel)
__global__ void up_wgrad_combine_kernel(const float* __restrict__ geff, float* __restrict__ dw, int cin_total, int cin_off,
                                        int Cup, int Cout) {
  const long long n = 27LL * Cup * Cout;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    long long r = t;
    const int co = (int)(r % Cout); r /= Cout;
    const int ci = (int)(r % Cup); r /= Cup;
    const int t2 = (int)(r % 3); r /= 3;
    const int t1 = (int)(r % 3);
    const int t0 = (int)(r / 3);
    float acc = 0.f;
    for (int par = 0; par < 8; ++par) {
      const int p0 = (par >> 2) & 1, p1 = (par >> 1) & 1, p2 = par & 1;
      const int k0 = p0 == 0 ? (t0 == 0 ? 0 : 1) : (t0 == 2 ? 2 : 1);
      const int k1 = p1 == 0 ? (t1 == 0 ? 0 : 1) : (t1 == 2 ? 2 : 1);
      const int k2 = p2 == 0 ? (t2 == 0 ? 0 : 1) : (t2 == 2 ? 2 : 1);
      acc += geff[(((long long)par * 27 + (k0 * 3 + k1) * 3 + k2) * Cup + ci) * Cout + co];
    }
    dw[((long long)((t0 * 3 + t1) * 3 + t2) * cin_total + cin_off + ci) * Cout + co] += acc;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Forward / data-gradient convolution for Cin <= 32, Cout <= 32 (the full-resolution 24-channel layers: half of the
// forward + dgrad time).  With N = 32 the MMA is bound by reading the 128 x 8 A operand from shared memory (40 cycles
// for 16 cycles of math), so the three d2 taps ride in N instead:
//     P_j[v'] = sum_{k0,k1,ci} X[v' + (k0-1, k1-1, 0)][ci] * W[k0][k1][j][ci][:]        one MMA, N = 96 = (j, co)
//     out[z, y, x] = P_0[z, y, x-1] + P_1[z, y, x] + P_2[z, y, x+1]                       (epilogue: lane shuffles)
// One N = 96 MMA (56 cycles) replaces three N = 32 MMAs (120 cycles) and every X slab is loaded once, not three times.
//   tile      8 (d1) x 16 (d2) input columns = 128 GEMM rows -> 8 x 14 outputs; the CTA walks a range of d0 planes
//   A         slab = TMA box 32ch x 16 x 10 of plane p (K-major, SWIZZLE_128B); d1 tap k1 = descriptor + 2048 B
//   B         all 27 taps of the layer (9 tiles of 96 rows x 128 B = 108 KB) stay resident in shared memory
//   D         ring of four accumulators (96 TMEM columns each), one per output plane in flight; plane z takes the d0
//             taps from slabs z-1, z, z+1.  MMA warp w owns ring slot w (one issuing warp per accumulator), the four
//             epilogue warps drain finished planes while the next ones are being accumulated.
// ---------------------------------------------------------------------------------------------------------
constexpr int KF_TM1 = 8, KF_TM2 = 16, KF_OUT2 = KF_TM2 - 2;
constexpr int KF_SLAB_BYTES = (KF_TM1 + 2) * KF_TM2 * 128;       // 20480
constexpr int KF_BTILE_BYTES = 96 * 128;                           // 12288 per (k0, k1)
constexpr int KF_SA = 5, KF_NACC = 4, KF_N = 96;

struct KfGeom {
  int B, D0, D1, D2, Cout, act, nks;
  uint16_t* y2;    // optional (final part): also write [bf16(y_lo) | bf16(y_hi)] (2 Cout bf16 per voxel), the input of the
                   // NEXT layer's hybrid forward -- saves that layer's ssr_tf32_split_bf16 pass over y
  int f16;         // the source holds bf16 channels (64 per chunk) and the weights bf16 pairs: kind::f16 MMAs
  int c0;          // first channel of this part in the source tensor (TMA coordinate)
  int accumulate;  // epilogue adds the partial result already in y (earlier channel parts of a concatenated input)
  int final;       // last part: bias + activation are applied
  int n1tiles, n2tiles, nzr, zlen;
};

// Epilogue fusions (EPI): the full-resolution tensors these layers write are the largest of the net, so the two
// memory-bound passes that used to follow them ride in the epilogue instead (per-thread fp32 partial sums over the CTA's
// whole item list, one warp-shuffle + shared-memory reduction per CTA at the end, one atomic per channel per CTA):
//   EPI 1 (data gradient):  out *= elu'(h)  (h = forward output of the previous convolution, prefetched before the
//                           accumulator wait) and  dbias[c] += sum_v out[v][c]        -- replaces elu_bwd_kernel
//   EPI 2 (forward):        sums[c] += sum_v out[v][c], sums[C + c] += sum_v out^2    -- replaces colsum2_vec_kernel<0>
// NB = channel blocks of 8 (compile time so the accumulators stay in registers).
// ACC (with EPI 0, Cout == NB * 8): the partial sums already in y (G.accumulate) are prefetched before the accumulator wait
// like h in EPI 1, instead of being loaded after it.
template <int EPI, int NB, bool ACC = false>
__global__ void __launch_bounds__(416, 1)
conv3d_tc_k2n_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                     const float* __restrict__ bias, float* __restrict__ y, const KfGeom G,
                     const float* __restrict__ elu_h, float* __restrict__ dbias, double* __restrict__ sums) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;
  uint8_t* sA = sB + 9 * KF_BTILE_BYTES;
  uint64_t* bars = (uint64_t*)(sA + (size_t)KF_SA * KF_SLAB_BYTES);
  uint64_t* fullA = bars;
  uint64_t* emptyA = bars + KF_SA;
  uint64_t* accFull = bars + 2 * KF_SA;
  uint64_t* accEmpty = accFull + KF_NACC;
  uint64_t* fullB = accEmpty + KF_NACC;
  uint32_t* tmem_slot = (uint32_t*)(fullB + 1);
  float* sbias = (float*)(bars + 24);                    // 32 floats, zero padded
  double* sred = (double*)(bars + 40);                   // 64 doubles: per-CTA channel sums of the fused epilogues
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (EPI != 0 && threadIdx.x < 64) sred[threadIdx.x] = 0.0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < KF_SA; ++i) { mbar_init(fullA + i, 1); mbar_init(emptyA + i, 3); }   // planes p-1, p, p+1 read slab p
    for (int i = 0; i < KF_NACC; ++i) { mbar_init(accFull + i, 1); mbar_init(accEmpty + i, 4); }
    mbar_init(fullB, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (threadIdx.x < 32) sbias[threadIdx.x] = (bias && threadIdx.x < G.Cout) ? bias[threadIdx.x] : 0.f;
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_w); }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nitems = G.B * G.n1tiles * G.n2tiles * G.nzr;

#define KF_DECODE(item)                                                          \
  int t_ = (item);                                                                \
  const int zr = t_ % G.nzr; t_ /= G.nzr;                                         \
  const int t2 = t_ % G.n2tiles; t_ /= G.n2tiles;                                 \
  const int t1 = t_ % G.n1tiles;                                                  \
  const int b = t_ / G.n1tiles;                                                   \
  const int x0 = t2 * KF_OUT2, y0 = t1 * KF_TM1;                                  \
  const int zs = zr * G.zlen, ze = min(G.D0, zs + G.zlen);                        \
  const int pmin = max(zs - 1, 0), pmax = min(ze, G.D0 - 1);                      \
  (void)x0; (void)y0; (void)b; (void)pmax;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      mbar_expect_tx(fullB, 9 * KF_BTILE_BYTES);
      for (int t = 0; t < 9; ++t) tma_load_2d(&map_w, fullB, sB + (size_t)t * KF_BTILE_BYTES, 0, t * KF_N);
      int seq = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        KF_DECODE(item)
        for (int p = pmin; p <= pmax; ++p, ++seq) {
          const int slot = seq % KF_SA;
          mbar_wait(emptyA + slot, ((seq / KF_SA) & 1) ^ 1);
          mbar_expect_tx(fullA + slot, KF_SLAB_BYTES);
          tma_load_5d(&map_x, fullA + slot, sA + (size_t)slot * KF_SLAB_BYTES, G.c0, x0 - 1, y0 - 1, p, b);
        }
      }
    }
  } else if (warp <= KF_NACC) {
    // ================================ MMA issuers: warp w owns accumulator ring slot w - 1 ================================
    const int w = warp - 1;
    const bool f16 = G.f16 != 0;
    const uint32_t idesc = f16 ? make_idesc_bf16(KF_N) : make_idesc_tf32(KF_N);
    const uint32_t a_base = desc_lo(smem_u32(sA), 16), b_base = desc_lo(smem_u32(sB), 16);
    const uint32_t dcol = tmem_base + (uint32_t)(w * KF_N);
    const int nks = G.nks;
    mbar_wait(fullB, 0);
    int seq_base = 0;
    uint32_t uses = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      KF_DECODE(item)
      for (int z = zs + w; z < ze; z += KF_NACC) {
        mbar_wait(accEmpty + w, (uses & 1u) ^ 1u);            // epilogue has drained this ring slot
        tc_fence_after();
        uint32_t acc = 0u;
#pragma unroll
        for (int k0 = 0; k0 < 3; ++k0) {
          const int p = z + k0 - 1;
          if (p < 0 || p >= G.D0) continue;                     // zero padding along d0: tap contributes nothing
          const int seq = seq_base + (p - pmin), slot = seq % KF_SA;
          mbar_wait(fullA + slot, (seq / KF_SA) & 1);
          // slab p is released by three arrivals (planes p-1, p, p+1); planes outside this item's range never come
          const int extra = k0 == 0 ? (z == zs ? 2 : 0) : k0 == 1 ? ((z == zs) + (z == ze - 1)) : (z == ze - 1 ? 2 : 0);
          if (elect_one()) {
            uint32_t alo = a_base + (uint32_t)slot * (KF_SLAB_BYTES >> 4);
            uint32_t blo = b_base + (uint32_t)(k0 * 3) * (KF_BTILE_BYTES >> 4);
#pragma unroll
            for (int k1 = 0; k1 < 3; ++k1) {
              umma_chain_any(f16, nks, dcol, alo, blo, idesc, acc);
              acc = 1u;
              alo += (uint32_t)(KF_TM2 * 128 >> 4);
              blo += (uint32_t)(KF_BTILE_BYTES >> 4);
            }
            umma_commit(emptyA + slot);
          }
          acc = 1u;
          if (lane == 0)
            for (int e = 0; e < extra; ++e) mbar_arrive(emptyA + slot);
          __syncwarp();
        }
        if (elect_one()) umma_commit(accFull + w);
        __syncwarp();
        ++uses;
      }
      seq_base += pmax - pmin + 1;
    }
  } else {
    // ================================ epilogue (last eight warps: two sets of four) ================================
    // Set h drains the planes with (z - zs) % 2 == h, so two planes are in the epilogue at any time; each warp of a set
    // owns one TMEM lane quarter.  Channels are handled in blocks of 8 (24 = 3 blocks, no padded work).
    const int q = warp & 3;                         // TMEM lane quarter of this warp
    const int h = (warp - (KF_NACC + 1)) >> 2;      // epilogue set 0 / 1
    const int r = q * 32 + lane;                    // GEMM row = (d1 row r / 16, input column r % 16)
    const int xin = r & 15, yl = r >> 4;
    const bool vec_ok = (G.Cout & 3) == 0;
    const int nblk8 = (G.Cout + 7) >> 3;
    uint32_t par = 0;                               // phase bit per ring slot
    float s1[EPI != 0 ? NB * 8 : 1], s2[EPI == 2 ? NB * 8 : 1];
#pragma unroll
    for (int i = 0; i < (EPI != 0 ? NB * 8 : 1); ++i) s1[i] = 0.f;
#pragma unroll
    for (int i = 0; i < (EPI == 2 ? NB * 8 : 1); ++i) s2[i] = 0.f;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      KF_DECODE(item)
      const int i1 = y0 + yl, i2 = x0 - 1 + xin;
      const bool store_ok = xin >= 1 && xin <= KF_OUT2 && i1 < G.D1 && i2 < G.D2;
      for (int z = zs + h; z < ze; z += 2) {
        const int slot = (z - zs) & (KF_NACC - 1);
        const long long voff = ((((long long)b * G.D0 + z) * G.D1 + i1) * G.D2 + i2) * G.Cout;
        float4 hp[(EPI == 1 || ACC) ? NB * 2 : 1];
        if constexpr (EPI == 1) {                     // h of this thread's voxel: in flight while the MMAs finish
#pragma unroll
          for (int i = 0; i < NB * 2; ++i)
            hp[i] = store_ok ? __ldg(reinterpret_cast<const float4*>(elu_h + voff) + i) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
        if constexpr (ACC) {                          // partial sums of the earlier parts (plain loads: y is written below)
#pragma unroll
          for (int i = 0; i < NB * 2; ++i)
            hp[i] = store_ok ? *(reinterpret_cast<const float4*>(y + voff) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(accFull + slot, (par >> slot) & 1u);
        par ^= 1u << slot;
        tc_fence_after();
        float* orow = y + voff;
        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(slot * KF_N);
        auto block = [&](const int cb, const int blk) {
          (void)blk;
          uint32_t v0[8], v1[8], v2[8];
          tmem_ld8(tbase + (uint32_t)cb, v0);
          tmem_ld8(tbase + (uint32_t)(32 + cb), v1);
          tmem_ld8(tbase + (uint32_t)(64 + cb), v2);
          tmem_ld_wait();
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(v0[e]), 1);      // P_0 at input column x - 1
            const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[e]), 1);   // P_2 at input column x + 1
            o[e] = left + __uint_as_float(v1[e]) + right;
          }
          if constexpr (ACC) {
            const float4 pa = hp[2 * blk], pb = hp[2 * blk + 1];
            o[0] += pa.x; o[1] += pa.y; o[2] += pa.z; o[3] += pa.w; o[4] += pb.x; o[5] += pb.y; o[6] += pb.z; o[7] += pb.w;
          } else
          if (G.accumulate && store_ok) {             // partial sums of the earlier channel parts (fp32, pre-activation)
            const int nv8 = G.Cout - cb;
            if (nv8 >= 8 && vec_ok) {
              const float4 p0 = *reinterpret_cast<const float4*>(orow + cb), p1 = *reinterpret_cast<const float4*>(orow + cb + 4);
              o[0] += p0.x; o[1] += p0.y; o[2] += p0.z; o[3] += p0.w; o[4] += p1.x; o[5] += p1.y; o[6] += p1.z; o[7] += p1.w;
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (e < nv8) o[e] += orow[cb + e];
            }
          }
          if (G.final) {
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] += sbias[cb + e];
          }
          if (G.act && G.final) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float neg = __expf(fminf(o[e], 0.f)) - 1.f;
              o[e] = o[e] > 0.f ? o[e] : neg;
            }
          }
          if constexpr (EPI == 1) {                   // elu'(pre) from the ELU output: 1 where h > 0, h + 1 elsewhere
            const float4 ha = hp[2 * blk], hb = hp[2 * blk + 1];
            const float hv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] *= hv[e] > 0.f ? 1.f : hv[e] + 1.f;
          }
          if (!store_ok) return;
          if constexpr (EPI != 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) s1[blk * 8 + e] += o[e];
          }
          if constexpr (EPI == 2) {
#pragma unroll
            for (int e = 0; e < 8; ++e) s2[blk * 8 + e] += o[e] * o[e];
          }
          const int nvalid = G.Cout - cb;
          if (G.y2 != nullptr && G.final && nvalid >= 8) {          // Cout % 8 == 0 (host-checked)
            uint32_t lo[8], hi[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split_bf16(o[e], lo[e], hi[e]);
            uint16_t* r2 = G.y2 + 2 * voff + cb;
            *reinterpret_cast<uint4*>(r2) = make_uint4(lo[0] | (lo[1] << 16), lo[2] | (lo[3] << 16), lo[4] | (lo[5] << 16),
                                                       lo[6] | (lo[7] << 16));
            *reinterpret_cast<uint4*>(r2 + G.Cout) = make_uint4(hi[0] | (hi[1] << 16), hi[2] | (hi[3] << 16),
                                                                hi[4] | (hi[5] << 16), hi[6] | (hi[7] << 16));
          }
          if (nvalid >= 8 && vec_ok) {
            *reinterpret_cast<float4*>(orow + cb) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(orow + cb + 4) = make_float4(o[4], o[5], o[6], o[7]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (e < nvalid) orow[cb + e] = o[e];
          }
        };
        if constexpr (EPI == 0 && !ACC) {
          for (int cb = 0; cb < nblk8 * 8; cb += 8) block(cb, 0);
        } else {                                      // Cout == NB * 8 (host-checked): constant indices into s1 / s2 / hp
#pragma unroll
          for (int blk = 0; blk < NB; ++blk) block(blk * 8, blk);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(accEmpty + slot);
      }
    }
    if constexpr (EPI != 0) {                         // per-CTA reduction: warp butterfly, then shared-memory doubles
#pragma unroll
      for (int i = 0; i < NB * 8; ++i) {
        float a = s1[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        if (lane == 0) atomicAdd(sred + i, (double)a);
        if constexpr (EPI == 2) {
          float c = s2[i];
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
          if (lane == 0) atomicAdd(sred + 32 + i, (double)c);
        }
      }
    }
  }
#undef KF_DECODE
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
  if constexpr (EPI == 1) {
    if (threadIdx.x < NB * 8) atomicAdd(dbias + threadIdx.x, (float)sred[threadIdx.x]);
  }
  if constexpr (EPI == 2) {
    if (threadIdx.x < NB * 8) {
      atomicAdd(sums + threadIdx.x, sred[threadIdx.x]);
      atomicAdd(sums + NB * 8 + threadIdx.x, sred[32 + threadIdx.x]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Forward of the upsampled part of the LAST decoder convolution (Cup <= 64 -> 24 channels at full resolution) in the
// k2n layout: conv3d_tc_up_kernel<1> with N = 24 is bound by L2 -> shared-memory slab traffic (844 B per output,
// profiles/r01_fwd_up72_ncu_full.txt), so the d2 parity p2 and its two taps ride in N instead:
//     Q_g[v'] = sum_{k0, k1, ci} low[v' + (k0 - 1, k1 - 1, 0)][ci] * Weff_(p0,p1,p2)[k0][k1][k2][ci][:]   g = (p2, k2) in
//               {(0,0), (0,1), (1,1), (1,2)}  ->  one MMA with N = 4 x 24 = 96
//     y[2 i0 + p0, 2 i1 + p1, 2 i2]     = Q_(0,0)[i2 - 1] + Q_(0,1)[i2]          (epilogue: lane shuffles, two adjacent
//     y[2 i0 + p0, 2 i1 + p1, 2 i2 + 1] = Q_(1,1)[i2]     + Q_(1,2)[i2 + 1]       full-resolution voxels = 192 B per thread)
//   class     a CTA serves ONE (p0, p1) class (blockIdx.x & 3): its 2 (d0 tap) x 2 (d1 tap) x 2 (chunk) kernel tiles of
//             96 rows (96 KB) stay resident in shared memory
//   A         slab = TMA box 32 ch x 16 x 10 of low-resolution plane q and chunk ch (as conv3d_tc_k2n_kernel); output plane
//             z takes planes z + p0 - 1 and z + p0, so a slab is read by two output planes (ring of 6 = 3 planes x 2 chunks)
//   D         ring of four accumulators (96 TMEM columns), MMA warp w owns slot w, eight epilogue warps in two sets
// ---------------------------------------------------------------------------------------------------------
constexpr int KU_SA = 6;

struct KuGeom {
  int B, D0, D1, D2;       // LOW-resolution grid
  int Cout;                // 24
  int nch;                 // channel chunks of the low-resolution tensor (1 or 2)
  int nks[2];              // K-steps of 8 channels per chunk
  int n1tiles, n2tiles, nzr, zlen;
};

__global__ void __launch_bounds__(416, 1)
conv3d_tc_up_k2n_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                        float* __restrict__ y, const KuGeom G) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;                                            // [t][k1i][ch] x 96 rows x 128 B
  uint8_t* sA = sB + 8 * KF_BTILE_BYTES;
  uint64_t* bars = (uint64_t*)(sA + (size_t)KU_SA * KF_SLAB_BYTES);
  uint64_t* fullA = bars;
  uint64_t* emptyA = bars + KU_SA;
  uint64_t* accFull = bars + 2 * KU_SA;
  uint64_t* accEmpty = accFull + KF_NACC;
  uint64_t* fullB = accEmpty + KF_NACC;
  uint32_t* tmem_slot = (uint32_t*)(fullB + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cls = blockIdx.x & 3, p0 = cls >> 1, p1 = cls & 1;   // parity class of this CTA
  const int cta = blockIdx.x >> 2, ncta = gridDim.x >> 2;        // index / count of the CTAs serving this class
  const int nch = G.nch;

  if (threadIdx.x == 0) {
    for (int i = 0; i < KU_SA; ++i) { mbar_init(fullA + i, 1); mbar_init(emptyA + i, KF_NACC); }   // every issuing warp releases every slab
    for (int i = 0; i < KF_NACC; ++i) { mbar_init(accFull + i, 1); mbar_init(accEmpty + i, 4); }
    mbar_init(fullB, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_w); }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nitems = G.B * G.n1tiles * G.n2tiles * G.nzr;

  // input planes of output plane z: z + da and z + da + 1 with da = p0 - 1 (d0 taps k0 = p0 and p0 + 1)
  const int da = p0 - 1;
#define KU_DECODE(item)                                                          \
  int t_ = (item);                                                                \
  const int zr = t_ % G.nzr; t_ /= G.nzr;                                         \
  const int t2 = t_ % G.n2tiles; t_ /= G.n2tiles;                                 \
  const int t1 = t_ % G.n1tiles;                                                  \
  const int b = t_ / G.n1tiles;                                                   \
  const int x0 = t2 * KF_OUT2, y0 = t1 * KF_TM1;                                  \
  const int zs = zr * G.zlen, ze = min(G.D0, zs + G.zlen);                        \
  const int pmin = max(zs + da, 0), pmax = min(ze + da, G.D0 - 1);                \
  (void)x0; (void)y0; (void)b; (void)pmax;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      mbar_expect_tx(fullB, (uint32_t)(4 * nch * KF_BTILE_BYTES));
      for (int t = 0; t < 4; ++t)
        for (int ch = 0; ch < nch; ++ch)
          tma_load_2d(&map_w, fullB, sB + (size_t)(t * 2 + ch) * KF_BTILE_BYTES, 0, ((cls * 4 + t) * 2 + ch) * KF_N);
      int seq = 0;
      for (int item = cta; item < nitems; item += ncta) {
        KU_DECODE(item)
        for (int q = pmin; q <= pmax; ++q)
          for (int ch = 0; ch < nch; ++ch, ++seq) {
            const int slot = seq % KU_SA;
            mbar_wait(emptyA + slot, ((seq / KU_SA) & 1) ^ 1);
            mbar_expect_tx(fullA + slot, KF_SLAB_BYTES);
            tma_load_5d(&map_x, fullA + slot, sA + (size_t)slot * KF_SLAB_BYTES, ch * 32, x0 - 1, y0 - 1, q, b);
          }
      }
    }
  } else if (warp <= KF_NACC) {
    // ================================ MMA issuers: warp w owns accumulator ring slot w - 1 ================================
    // Every issuing warp walks EVERY slab in producer order (and releases the ones it does not read): a slab is read by
    // only two of the four warps here, and a warp that skipped a slot's earlier phases could not tell them apart from the
    // one it needs (mbarrier parity waits alias every other phase).
    const int w = warp - 1;
    const uint32_t idesc = make_idesc_tf32(KF_N);
    const uint32_t a_base = desc_lo(smem_u32(sA), 16), b_base = desc_lo(smem_u32(sB), 16);
    const uint32_t dcol = tmem_base + (uint32_t)(w * KF_N);
    mbar_wait(fullB, 0);
    int seq = 0;
    uint32_t uses = 0, acc = 0u;
    for (int item = cta; item < nitems; item += ncta) {
      KU_DECODE(item)
      for (int q = pmin; q <= pmax; ++q) {
        // readers of plane q: output plane q - da through its first d0 tap (t = 0), q - da - 1 through its second (t = 1)
        int z = -1, t = 0;
        const int z1 = q - da, z2 = q - da - 1;
        if (z1 >= zs && z1 < ze && ((z1 - zs) & (KF_NACC - 1)) == w) { z = z1; t = 0; }
        if (z2 >= zs && z2 < ze && ((z2 - zs) & (KF_NACC - 1)) == w) { z = z2; t = 1; }
        for (int ch = 0; ch < nch; ++ch, ++seq) {
          const int slot = seq % KU_SA;
          mbar_wait(fullA + slot, (seq / KU_SA) & 1);
          if (z >= 0) {                                           // warp-uniform
            // first chain of output plane z: its t = 0 slab, or the t = 1 slab when plane z + da lies outside the volume
            if (ch == 0 && (t == 0 || z + da < 0)) {
              mbar_wait(accEmpty + w, (uses & 1u) ^ 1u);          // epilogue has drained this ring slot
              tc_fence_after();
              acc = 0u;
            }
            const int nks = G.nks[ch];
            if (elect_one()) {
              // d1 taps k1 = p1 + k1i: operand view starts k1 rows of 16 voxels into the slab
              uint32_t alo = a_base + (uint32_t)slot * (KF_SLAB_BYTES >> 4) + (uint32_t)p1 * (uint32_t)(KF_TM2 * 128 >> 4);
              uint32_t blo = b_base + (uint32_t)((t * 2) * 2 + ch) * (KF_BTILE_BYTES >> 4);
#pragma unroll
              for (int k1i = 0; k1i < 2; ++k1i) {
                if (nks == 4) umma_chain_k<4>(dcol, alo, blo, DESC_HI_K_SW128, idesc, acc);
                else if (nks == 3) umma_chain_k<3>(dcol, alo, blo, DESC_HI_K_SW128, idesc, acc);
                else if (nks == 2) umma_chain_k<2>(dcol, alo, blo, DESC_HI_K_SW128, idesc, acc);
                else umma_chain_k<1>(dcol, alo, blo, DESC_HI_K_SW128, idesc, acc);
                acc = 1u;
                alo += (uint32_t)(KF_TM2 * 128 >> 4);
                blo += (uint32_t)(2 * KF_BTILE_BYTES >> 4);       // tile index (t * 2 + k1i) * 2 + ch
              }
              umma_commit(emptyA + slot);
            }
            acc = 1u;
            // last chain of output plane z: its t = 1 slab, or the t = 0 slab when plane z + da + 1 lies outside the volume
            if (ch == nch - 1 && (t == 1 || z + da + 1 > G.D0 - 1)) {
              if (elect_one()) umma_commit(accFull + w);
              ++uses;
            }
          } else if (lane == 0) {
            mbar_arrive(emptyA + slot);                           // not read by this warp: release immediately
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================ epilogue (last eight warps: two sets of four) ================================
    const int q4 = warp & 3;                        // TMEM lane quarter of this warp
    const int h = (warp - (KF_NACC + 1)) >> 2;      // epilogue set 0 / 1
    const int r = q4 * 32 + lane;                   // GEMM row = (d1 row r / 16, input column r % 16)
    const int xin = r & 15, yl = r >> 4;
    const long long F1 = 2LL * G.D1, F2 = 2LL * G.D2;
    uint32_t par = 0;                               // phase bit per ring slot
    for (int item = cta; item < nitems; item += ncta) {
      KU_DECODE(item)
      const int i1 = y0 + yl, i2 = x0 - 1 + xin;
      const bool store_ok = xin >= 1 && xin <= KF_OUT2 && i1 < G.D1 && i2 < G.D2;
      for (int z = zs + h; z < ze; z += 2) {
        const int slot = (z - zs) & (KF_NACC - 1);
        mbar_wait(accFull + slot, (par >> slot) & 1u);
        par ^= 1u << slot;
        tc_fence_after();
        // two adjacent full-resolution voxels (2 i2, 2 i2 + 1) of row (2 z + p0, 2 i1 + p1): 48 contiguous floats
        float* orow = y + ((((long long)b * (2 * G.D0) + 2 * z + p0) * F1 + 2 * i1 + p1) * F2 + 2 * i2) * 24;
        const uint32_t tbase = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(slot * KF_N);
#pragma unroll
        for (int cb = 0; cb < 24; cb += 8) {
          uint32_t v0[8], v1[8], v2[8], v3[8];
          tmem_ld8(tbase + (uint32_t)cb, v0);              // (p2, k2) = (0, 0)
          tmem_ld8(tbase + (uint32_t)(24 + cb), v1);       // (0, 1)
          tmem_ld8(tbase + (uint32_t)(48 + cb), v2);       // (1, 1)
          tmem_ld8(tbase + (uint32_t)(72 + cb), v3);       // (1, 2)
          tmem_ld_wait();
          float o0[8], o1[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(v0[e]), 1);      // Q_(0,0) at input column x - 1
            const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(v3[e]), 1);   // Q_(1,2) at input column x + 1
            o0[e] = left + __uint_as_float(v1[e]);
            o1[e] = __uint_as_float(v2[e]) + right;
          }
          if (store_ok) {
            *reinterpret_cast<float4*>(orow + cb) = make_float4(o0[0], o0[1], o0[2], o0[3]);
            *reinterpret_cast<float4*>(orow + cb + 4) = make_float4(o0[4], o0[5], o0[6], o0[7]);
            *reinterpret_cast<float4*>(orow + 24 + cb) = make_float4(o1[0], o1[1], o1[2], o1[3]);
            *reinterpret_cast<float4*>(orow + 24 + cb + 4) = make_float4(o1[4], o1[5], o1[6], o1[7]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(accEmpty + slot);
      }
    }
  }
#undef KU_DECODE
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// packed kernels of conv3d_tc_up_k2n_kernel from the effective kernels weff (8, 27, Cup, 24) of up_weights_kernel:
//   wp[cls = p0 * 2 + p1][t][k1i][ch][row = g * 24 + co][s]  =  weff[p0 * 4 + p1 * 2 + p2(g)][k0 = p0 + t][k1 = p1 + k1i][k2(g)]
//   [ch * 32 + s][co],  g = 0..3 <-> (p2, k2) = (0,0), (0,1), (1,1), (1,2);  zero for channels >= Cup; TF32-rounded
__global__ void pack_up_k2n_kernel(const float* __restrict__ weff, float* __restrict__ wp, int Cup, int round_rn) {
  const long long total = 4LL * 8 * 96 * 32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(i & 31);
    long long r = i >> 5;
    const int row = (int)(r % 96); r /= 96;
    const int ch = (int)(r % 2); r /= 2;
    const int k1i = (int)(r % 2); r /= 2;
    const int t = (int)(r % 2);
    const int cls = (int)(r / 2);
    const int p0 = cls >> 1, p1 = cls & 1;
    const int g = row / 24, co = row % 24;
    const int p2 = g >> 1, k2 = g == 0 ? 0 : (g == 3 ? 2 : 1);
    const int ci = ch * 32 + s;
    float val = 0.f;
    if (ci < Cup) {
      const int par = p0 * 4 + p1 * 2 + p2, tap = ((p0 + t) * 3 + (p1 + k1i)) * 3 + k2;
      val = weff[(((long long)par * 27 + tap) * Cup + ci) * 24 + co];
    }
    if (round_rn) {
      uint32_t u;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(val));
      val = __uint_as_float(u);
    }
    wp[i] = val;
  }
}

// ---------------------------------------------------------------------------------------------------------
// weight gradient on tcgen05:  dW[k0][k1][k2][ci][co] += sum_v X[v + (k0,k1,k2) - 1][ci] * dY[v][co]
//
//   GEMM   D[M x N] += A[M x K] * B[N x K]^T   with K = 8 voxels per instruction, both operands MN-major:
//   A      the SAME activation slab as the forward kernel (TMA box 32ch x 8 x 18, SWIZZLE_128B) read as an
//          MN-major operand: the 128 B row of a voxel is the M axis (32 channels), 8 consecutive rows are K.
//          M = 128 = 4 atoms at +1024 B = the three d1 taps (k1 = 0,1,2) of the slab + one ignored atom.
//   B      dY tile (TMA box 32ch x 8 x 16) per 32 output channels, MN-major, N = NT (<= 96) output channels.
//   D      KG accumulators (one per d0 tap in the CTA's group) of NT columns in TMEM; at the end the CTA adds its
//          partial sums to dW with fp32 atomics (split over spatial tiles / d0 ranges).
//   CTA    = (input-channel chunk, d2 tap, d0-tap group, N tile, (d1,d2) tile, d0 range); it walks the d0 planes of
//          its range keeping the last KG dY planes resident.
// ---------------------------------------------------------------------------------------------------------
constexpr int WG_BTILE_BYTES = TM1 * TM2 * 128;     // 16384: dY tile of 128 voxels x 32 channels

struct WgGeom {
  int B, D0, D1, D2;
  int Cin, C1, Cout;
  int NT, nNtiles, KG, SA, SBT;
  int nchunks, n1tiles, n2tiles, n0splits, zlen;
  int tmem_cols, exp_flags;
  int cin_off;       // first input channel of this launch inside dw's Cin axis (G.Cin = dw's total Cin)
  unsigned char chunk_src[MAX_CHUNKS];
  unsigned char chunk_valid[MAX_CHUNKS];
  short chunk_c0[MAX_CHUNKS];
};

// MN-major SWIZZLE_128B descriptor: atoms (32 x 8) of 1024 B, `lbo` bytes between atoms along M/N
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;              // K-group stride (one group per instruction)
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(192, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_x2,
                const __grid_constant__ CUtensorMap map_dy, float* __restrict__ dw, const WgGeom G) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  const int bstage = (G.NT / 32) * WG_BTILE_BYTES;
  uint8_t* sB = sA + (size_t)G.SA * SLAB_BYTES;
  uint64_t* bars = (uint64_t*)(sB + (size_t)G.SBT * bstage);
  uint64_t* fullA = bars;
  uint64_t* emptyA = bars + G.SA;
  uint64_t* fullB = bars + 2 * G.SA;
  uint64_t* emptyB = fullB + G.SBT;
  uint64_t* accFull = emptyB + G.SBT;
  uint32_t* tmem_slot = (uint32_t*)(accFull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* dbg = (g_dbg && (blockIdx.x % 13) == 0 && blockIdx.x / 13 < 128) ? g_dbg + (blockIdx.x / 13) * 16 : nullptr;
  if (threadIdx.x == 0) DBG_STAMP(0);

  int t = blockIdx.x;
  const int zs_i = t % G.n0splits; t /= G.n0splits;
  const int t2 = t % G.n2tiles; t /= G.n2tiles;
  const int t1 = t % G.n1tiles; t /= G.n1tiles;
  const int b = t % G.B; t /= G.B;
  const int nt = t % G.nNtiles; t /= G.nNtiles;
  const int ngroups = 3 / G.KG;
  const int k0g = (t % ngroups) * G.KG; t /= ngroups;
  const int k2 = t % 3;
  const int ch = t / 3;
  const int x0 = t2 * TM2, y0 = t1 * TM1, n0 = nt * G.NT;
  const int zs = zs_i * G.zlen, ze = min(G.D0, zs + G.zlen);
  // X planes visited: zin = zo + k0 - 1 for zo in [zs,ze), k0 in [k0g, k0g+KG)
  const int zin_lo = zs + k0g - 1, zin_hi = ze - 1 + k0g + G.KG - 2;   // inclusive

  if (threadIdx.x == 0) {
    for (int i = 0; i < G.SA; ++i) { mbar_init(fullA + i, 1); mbar_init(emptyA + i, 1); }
    for (int i = 0; i < G.SBT; ++i) { mbar_init(fullB + i, 1); mbar_init(emptyB + i, 1); }
    mbar_init(accFull, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)G.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) DBG_STAMP(1);

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* mx = G.chunk_src[ch] ? &map_x2 : &map_x1;
      const int c0 = G.chunk_c0[ch];
      int sa = 0, pa = 0;
      for (int zin = zin_lo; zin <= zin_hi; ++zin) {
        const int zo_new = zin - k0g + 1;                      // dY plane first needed at this step
        if (zo_new >= zs && zo_new < ze) {
          const int p = zo_new - zs, sb = p % G.SBT, pb = (p / G.SBT) & 1;
          mbar_wait(emptyB + sb, pb ^ 1);
          mbar_expect_tx(fullB + sb, (uint32_t)bstage);
          for (int a = 0; a < G.NT / 32; ++a)
            tma_load_5d(&map_dy, fullB + sb, sB + (size_t)sb * bstage + (size_t)a * WG_BTILE_BYTES, n0 + a * 32, x0, y0,
                        zo_new, b);
        }
        if (zin >= 0 && zin < G.D0) {
          mbar_wait(emptyA + sa, pa ^ 1);
          mbar_expect_tx(fullA + sa, SLAB_BYTES);
          tma_load_5d(mx, fullA + sa, sA + (size_t)sa * SLAB_BYTES, c0, x0 + k2 - 1, y0 - 1, zin, b);
          if (++sa == G.SA) { sa = 0; pa ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {
      // whole warp converged, one elected lane issues.  D=F32, A=B=TF32, both MN-major (bits 15,16), M=128, N=NT
      const uint32_t idesc = make_idesc_tf32(G.NT) | (1u << 15) | (1u << 16);
      int sa = 0, pa = 0;
      uint32_t started = 0;
      long long wait_a = 0, wait_b = 0;
      if (lane == 0) DBG_STAMP(3);
      for (int zin = zin_lo; zin <= zin_hi; ++zin) {
        const int zo_new = zin - k0g + 1;
        if (zo_new >= zs && zo_new < ze) {
          const int p = zo_new - zs;
          const long long w0 = dbg ? clock64() : 0;
          mbar_wait(fullB + (p % G.SBT), (p / G.SBT) & 1);
          if (dbg) wait_b += clock64() - w0;
        }
        if (zin >= 0 && zin < G.D0) {
          { const long long w0 = dbg ? clock64() : 0; mbar_wait(fullA + sa, pa); if (dbg) wait_a += clock64() - w0; }
          if (G_FENCE_IN_LOOP) tc_fence_after();   // not needed for TMA-written operands (mbarrier complete_tx orders them); costs a pipe drain
          const uint32_t alo0 = desc_lo(smem_u32(sA + (size_t)sa * SLAB_BYTES), 1024);      // M atoms = d1 taps
          if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            if (kk >= G.KG) break;
            const int zo = zin - (k0g + kk) + 1;
            if (zo < zs || zo >= ze) continue;
            const uint32_t blo0 = desc_lo(smem_u32(sB + (size_t)((zo - zs) % G.SBT) * bstage), WG_BTILE_BYTES);
            const uint32_t dcol = tmem_base + (uint32_t)(kk * G.NT);
            uint32_t acc = (started >> kk) & 1u;
            umma_chain_mn16(dcol, alo0, blo0, DESC_HI_MN_SW128_32B, idesc, acc);   // 16 K-steps of 8 voxels
          }
            umma_commit(emptyA + sa);
          }
          __syncwarp();
          for (int kk = 0; kk < G.KG; ++kk) {
            const int zo = zin - (k0g + kk) + 1;
            if (zo >= zs && zo < ze) started |= 1u << kk;
          }
          if (++sa == G.SA) { sa = 0; pa ^= 1; }
        }
        const int zo_old = zin - k0g - (G.KG - 1) + 1;         // dY plane whose last use was this step
        if (zo_old >= zs && zo_old < ze) {
          if (elect_one()) umma_commit(emptyB + ((zo_old - zs) % G.SBT));
          __syncwarp();
        }
      }
      if (elect_one()) umma_commit(accFull);
      __syncwarp();
      if (lane == 0) { DBG_STAMP(4); if (dbg) { dbg[9] = wait_a; dbg[10] = wait_b; } }
    }
  } else {
    const int q = warp & 3;                       // rows 32q..32q+31 <-> d1 tap k1 = q (q == 3: unused atom)
    mbar_wait(accFull, 0);
    tc_fence_after();
    if (warp == 2 && lane == 0) DBG_STAMP(5);
    // accumulator kk received MMAs iff some dY plane zo of the range pairs with an in-bounds X plane zo + k0 - 1
    uint32_t started = 0;
    for (int kk = 0; kk < G.KG; ++kk)
      for (int zo = zs; zo < ze; ++zo) {
        const int zin = zo + k0g + kk - 1;
        if (zin >= 0 && zin < G.D0) { started |= 1u << kk; break; }
      }
    const int valid = G.chunk_valid[ch];
    const int cin_idx = (G.chunk_src[ch] ? G.C1 : 0) + G.chunk_c0[ch] + lane;
    for (int kk = 0; kk < G.KG; ++kk) {
      if (!((started >> kk) & 1u)) continue;      // uniform
      const int k0 = k0g + kk;
      for (int cb = 0; cb < G.NT; cb += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kk * G.NT + cb), v);
        tmem_ld_wait();
        if (q < 3 && lane < valid) {
          float* o = dw + ((long long)(((k0 * 3 + q) * 3 + k2)) * G.Cin + cin_idx) * G.Cout + n0 + cb;
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (n0 + cb + e < G.Cout) atomicAdd(o + e, __uint_as_float(v[e]));
        }
      }
    }
    if (warp == 2 && lane == 0) DBG_STAMP(6);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)G.tmem_cols);
  if (threadIdx.x == 0) DBG_STAMP(7);
}

// ---------------------------------------------------------------------------------------------------------
// weight gradient, Cout <= 32 (the full-resolution layers = most of the wgrad time): the three d2 taps ride in N.
//   dW[k0][k1][k2][ci][co] = sum_v X[v0+k0-1, v1+k1-1, v2'][ci] * dY[v0, v1, v2'-k2+1][co]          (v2' = v2 + k2 - 1)
//   A = X slab of plane (zo + k0 - 1), unshifted in d2;  M = (k1, ci) as before
//   B = THREE dY tiles of plane zo, TMA boxes shifted by (1 - k2) along d2 (OOB zero-fill)  ->  N = (k2, co) = 96
// One MMA (N = 96: 56 cycles) does the work of three N = 32 MMAs (3 x 40 cycles), X is read once instead of 3x.
// The CTA walks the dY planes of its range; X slabs live in a rolling ring (planes zo-1, zo, zo+1 + one prefetched),
// dY stages are double buffered: nothing the MMA warp needs is ever loaded late.
// ---------------------------------------------------------------------------------------------------------
constexpr int WK_SA = 5, WK_SB = 4, WK_N = 96;
constexpr int WK_BROWS = (TM2 + 2) * TM1;                // dY tile with a one-voxel d2 halo: 16 rows of 10 voxels
constexpr int WK_BSTAGE = WK_BROWS * 128;                // 20480 (multiple of 1024)
constexpr int WK_THREADS = 256;                          // warp 0 TMA, warps 1-3 MMA (one per d0 tap), warps 4-7 epilogue

// The thread that issues tcgen05.mma is stalled while the tensor pipe is busy (the MMA queue is shallow), so all the
// scalar work between two MMA chains (descriptor arithmetic, barrier waits, commits) is exposed when a single warp
// issues everything: measured 74 cycles per N=96 MMA against the 56-cycle pipe rate.  Here each d0 tap (= accumulator)
// has its own issuing warp; their chains are independent, so one warp's scalar work hides behind the others' MMAs.
__global__ void __launch_bounds__(WK_THREADS, 1)
wgrad_tc_k2n_kernel(const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_x2,
                    const __grid_constant__ CUtensorMap map_dy, float* __restrict__ dw, const WgGeom G) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)WK_SA * SLAB_BYTES;
  uint64_t* bars = (uint64_t*)(sB + (size_t)WK_SB * WK_BSTAGE);
  uint64_t* fullA = bars;
  uint64_t* emptyA = bars + WK_SA;
  uint64_t* fullB = bars + 2 * WK_SA;
  uint64_t* emptyB = fullB + WK_SB;
  uint64_t* accFull = emptyB + WK_SB;
  uint32_t* tmem_slot = (uint32_t*)(accFull + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* dbg = (g_dbg && (blockIdx.x % 9) == 0 && blockIdx.x / 9 < 128) ? g_dbg + (blockIdx.x / 9) * 16 : nullptr;
  if (threadIdx.x == 0) DBG_STAMP(0);

  int t = blockIdx.x;
  const int zs_i = t % G.n0splits; t /= G.n0splits;
  const int t2 = t % G.n2tiles; t /= G.n2tiles;
  const int t1 = t % G.n1tiles; t /= G.n1tiles;
  const int b = t % G.B; t /= G.B;
  const int nt = t % G.nNtiles;
  const int ch = t / G.nNtiles;
  const int x0 = t2 * TM2, y0 = t1 * TM1, n0 = nt * 32;
  const int nrows = min(TM1, G.D1 - y0);          // d1 rows of the tile inside the volume = K-steps that carry data
  const int zs = zs_i * G.zlen, ze = min(G.D0, zs + G.zlen);
  const int pmin = max(zs - 1, 0), pmax = min(ze, G.D0 - 1);      // X planes used: [pmin, pmax]

  if (threadIdx.x == 0) {
    // every X slab / dY stage is released by all three MMA warps (a warp that never reads a slab arrives for it up front)
    for (int i = 0; i < WK_SA; ++i) { mbar_init(fullA + i, 1); mbar_init(emptyA + i, 3); }
    for (int i = 0; i < WK_SB; ++i) { mbar_init(fullB + i, 1); mbar_init(emptyB + i, 3); }
    mbar_init(accFull, 3);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) DBG_STAMP(1);

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* mx = G.chunk_src[ch] ? &map_x2 : &map_x1;
      const int c0 = G.chunk_c0[ch];
      int next_a = pmin;
      for (int zo = zs; zo < ze; ++zo) {
        const int need = min(zo + 1, pmax);
        for (; next_a <= need; ++next_a) {
          const int i = next_a - pmin, sa = i % WK_SA;
          mbar_wait(emptyA + sa, ((i / WK_SA) & 1) ^ 1);
          mbar_expect_tx(fullA + sa, SLAB_BYTES);
          tma_load_5d(mx, fullA + sa, sA + (size_t)sa * SLAB_BYTES, c0, x0, y0 - 1, next_a, b);
        }
        const int j = zo - zs, sb = j % WK_SB;
        mbar_wait(emptyB + sb, ((j / WK_SB) & 1) ^ 1);
        mbar_expect_tx(fullB + sb, WK_BSTAGE);
        tma_load_5d(&map_dy, fullB + sb, sB + (size_t)sb * WK_BSTAGE, n0, x0 - 1, y0, zo, b);
      }
      DBG_STAMP(2);
    }
  } else if (warp <= 3) {
    // ================================ MMA issuer for d0 tap k0 = warp - 1 (accumulator k0) ================================
    const int k0 = warp - 1;
    const uint32_t idesc = make_idesc_tf32(WK_N) | (1u << 15) | (1u << 16);
    // B: ONE dY tile with a d2 halo; the three N atoms are the same rows shifted by one voxel (LBO = 128 B): atom j
    // starts at voxel x0 - 1 + j = the d2 tap k2 = 2 - j.  Unaligned / overlapping operand views are read consistently
    // with the TMA swizzle (profiles/r01_desc_probe_unaligned_views.txt).
    const uint32_t a_base = desc_lo(smem_u32(sA), 1024), b_base = desc_lo(smem_u32(sB), 128);
    const uint32_t dcol = tmem_base + (uint32_t)(k0 * WK_N);
    // X plane read at output plane zo: p = zo + k0 - 1.  Loaded planes this warp never reads (p < zs + k0 - 1): arrive now.
    if (lane == 0)
      for (int p = pmin; p < min(zs + k0 - 1, pmax + 1); ++p) mbar_arrive(emptyA + ((p - pmin) % WK_SA));
    __syncwarp();
    uint32_t acc = 0;
    long long wait_a = 0, wait_b = 0;
    if (warp == 1 && lane == 0) DBG_STAMP(3);
    int sb = 0, pb = 0;
    int ia = zs + k0 - 1 - pmin;                       // slab index (relative to pmin) of plane p; may start at -1
    int sa = ia < 0 ? 0 : ia % WK_SA, pa = ia < 0 ? 0 : (ia / WK_SA) & 1;
    for (int zo = zs; zo < ze; ++zo) {
      const int p = zo + k0 - 1;
      { const long long w0 = dbg ? clock64() : 0; mbar_wait(fullB + sb, pb); if (dbg) wait_b += clock64() - w0; }
      const bool use = p >= 0 && p < G.D0;             // warp-uniform
      if (use) {
        { const long long w0 = dbg ? clock64() : 0; mbar_wait(fullA + sa, pa); if (dbg) wait_a += clock64() - w0; }
        if (elect_one()) {
          const uint32_t alo = a_base + (uint32_t)sa * (SLAB_BYTES >> 4), blo = b_base + (uint32_t)sb * (WK_BSTAGE >> 4);
          if (nrows == TM1) umma_chain_mn16_ab(dcol, alo, blo, DESC_HI_MN_SW128_32B, idesc, acc);
          else umma_chain_mn_ab(nrows, dcol, alo, blo, DESC_HI_MN_SW128_32B, idesc, acc);
          umma_commit(emptyB + sb);
          umma_commit(emptyA + sa);                    // this warp is done with plane p once the chain has completed
        }
        __syncwarp();
        acc = 1u;
        if (++sa == WK_SA) { sa = 0; pa ^= 1; }
      } else {
        if (lane == 0) mbar_arrive(emptyB + sb);
        __syncwarp();
        if (p >= 0 && ++sa == WK_SA) { sa = 0; pa ^= 1; }
      }
      if (++sb == WK_SB) { sb = 0; pb ^= 1; }
    }
    if (elect_one()) umma_commit(accFull);
    __syncwarp();
    if (lane == 0 && dbg) { if (warp == 1) { DBG_STAMP(4); dbg[9] = wait_a; dbg[10] = wait_b; } }
  } else {
    const int q = warp & 3;                       // rows 32q..32q+31 <-> d1 tap k1 = q (q == 3: unused atom)
    mbar_wait(accFull, 0);
    tc_fence_after();
    if (warp == 4 && lane == 0) DBG_STAMP(5);
    uint32_t started = 0;
    for (int k0 = 0; k0 < 3; ++k0)
      for (int zo = zs; zo < ze; ++zo) {
        const int p = zo + k0 - 1;
        if (p >= 0 && p < G.D0) { started |= 1u << k0; break; }
      }
    const int valid = G.chunk_valid[ch];
    const int cin_idx = (G.chunk_src[ch] ? G.C1 : 0) + G.chunk_c0[ch] + lane;
    for (int k0 = 0; k0 < 3; ++k0) {
      if (!((started >> k0) & 1u)) continue;      // uniform
      for (int cb = 0; cb < WK_N; cb += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(k0 * WK_N + cb), v);
        tmem_ld_wait();
        const int k2 = 2 - (cb >> 5), cobase = n0 + (cb & 31);  // column = (2 - k2) * 32 + co
        if (q < 3 && lane < valid) {
          float* o = dw + ((long long)(((k0 * 3 + q) * 3 + k2)) * G.Cin + cin_idx) * G.Cout + cobase;
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (cobase + e < G.Cout) atomicAdd(o + e, __uint_as_float(v[e]));
        }
      }
    }
    if (warp == 4 && lane == 0) DBG_STAMP(6);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
  if (threadIdx.x == 0) DBG_STAMP(7);
}

// ---------------------------------------------------------------------------------------------------------
// Persistent weight gradient (same MMA scheme as wgrad_tc_k2n_kernel).  A *unit* is (32-channel input chunk, 32 output
// channels) = one 3 x 96 x 128 accumulator set; its work is the list of (spatial tile, d0 plane) pairs.  The global list
// units x tiles x planes is cut into gridDim.x equal contiguous ranges, so every CTA does the same number of plane
// steps (no tail wave), keeps accumulating in TMEM across tiles, and flushes to dW with atomics once per unit it touches
// (normally once or twice per CTA instead of once per 8-80 planes: the flush was up to half of the CTA time on the
// 40^3 and deeper layers).
// ---------------------------------------------------------------------------------------------------------
struct WpSeg { int u, tile, za, zb; bool unit_last; long long gnext; };

__device__ __forceinline__ WpSeg wp_segment(long long g, long long g1, long long W, int D0) {
  WpSeg s;
  s.u = (int)(g / W);
  const long long w = g - (long long)s.u * W;
  s.tile = (int)(w / D0);
  s.za = (int)(w - (long long)s.tile * D0);
  long long e = (long long)s.u * W + (long long)(s.tile + 1) * D0;
  if (e > g1) e = g1;
  s.zb = s.za + (int)(e - g);
  s.unit_last = e == g1 || e == (long long)(s.u + 1) * W;
  s.gnext = e;
  return s;
}

// PAR: weight gradient of the upsampled part of a decoder convolution from the LOW-resolution tensor (see
// conv3d_tc_up_kernel): the 8 output parity classes are 8x more units; class p reads its own strided view of dy
// (maps_dy.x[p]), accumulates the gradient of its effective kernel into dw + p * 27 * Cin * Cout (combined into the
// 3x3x3 gradient by up_wgrad_combine_kernel) and skips the d0 tap its effective kernel does not have.
template <bool PAR>
__global__ void __launch_bounds__(WK_THREADS, 1)
wgrad_tc_persistent_kernel(const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_x2,
                           const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ UpMaps maps_dy,
                           float* __restrict__ dw, const WgGeom G) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)WK_SA * SLAB_BYTES;
  uint64_t* bars = (uint64_t*)(sB + (size_t)WK_SB * WK_BSTAGE);
  uint64_t* fullA = bars;
  uint64_t* emptyA = bars + WK_SA;
  uint64_t* fullB = bars + 2 * WK_SA;
  uint64_t* emptyB = fullB + WK_SB;
  uint64_t* accFull = emptyB + WK_SB;
  uint64_t* accEmpty = accFull + 1;
  uint32_t* tmem_slot = (uint32_t*)(accEmpty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int ntile = G.B * G.n1tiles * G.n2tiles;
  const long long W = (long long)ntile * G.D0;                    // plane steps per unit
  const int upp = G.nchunks * G.nNtiles;                          // units per parity class
  const long long T = W * upp * (PAR ? 8 : 1);
  // this launch covers slice `zlen` of `n0splits` of the global list (several short launches instead of one long one let
  // higher-priority kernels of the backward chain get SMs between them)
  const long long sl0 = T * G.zlen / G.n0splits, sl1 = T * (G.zlen + 1) / G.n0splits;
  const long long g0 = sl0 + (sl1 - sl0) * blockIdx.x / gridDim.x, g1 = sl0 + (sl1 - sl0) * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < WK_SA; ++i) { mbar_init(fullA + i, 1); mbar_init(emptyA + i, 3); }
    for (int i = 0; i < WK_SB; ++i) { mbar_init(fullB + i, 1); mbar_init(emptyB + i, 3); }
    mbar_init(accFull, 3);
    mbar_init(accEmpty, 4);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

#define WP_TILE(tile)                                                      \
  int tt_ = (tile);                                                         \
  const int t2 = tt_ % G.n2tiles; tt_ /= G.n2tiles;                         \
  const int t1 = tt_ % G.n1tiles;                                           \
  const int b = tt_ / G.n1tiles;                                            \
  const int x0 = t2 * TM2, y0 = t1 * TM1;                                   \
  (void)x0; (void)y0; (void)b;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int seqA = 0, seqB = 0;
      for (long long g = g0; g < g1;) {
        const WpSeg sg = wp_segment(g, g1, W, G.D0);
        g = sg.gnext;
        const int par = PAR ? sg.u / upp : 0, uu = PAR ? sg.u % upp : sg.u;
        const int ch = uu / G.nNtiles, n0 = (uu % G.nNtiles) * 32;
        const CUtensorMap* mx = G.chunk_src[ch] ? &map_x2 : &map_x1;
        const CUtensorMap* my = PAR ? &maps_dy.x[par] : &map_dy;
        const int c0 = G.chunk_c0[ch];
        WP_TILE(sg.tile)
        const int pmin = max(sg.za - 1, 0), pmax = min(sg.zb, G.D0 - 1);
        int next_a = pmin;
        for (int zo = sg.za; zo < sg.zb; ++zo) {
          const int need = min(zo + 1, pmax);
          for (; next_a <= need; ++next_a, ++seqA) {
            const int sa = seqA % WK_SA;
            mbar_wait(emptyA + sa, ((seqA / WK_SA) & 1) ^ 1);
            mbar_expect_tx(fullA + sa, SLAB_BYTES);
            tma_load_5d(mx, fullA + sa, sA + (size_t)sa * SLAB_BYTES, c0, x0, y0 - 1, next_a, b);
          }
          const int sb = seqB % WK_SB;
          mbar_wait(emptyB + sb, ((seqB / WK_SB) & 1) ^ 1);
          mbar_expect_tx(fullB + sb, WK_BSTAGE);
          tma_load_5d(my, fullB + sb, sB + (size_t)sb * WK_BSTAGE, n0, x0 - 1, y0, zo, b);
          ++seqB;
        }
      }
    }
  } else if (warp <= 3) {
    // ================================ MMA issuer for d0 tap k0 = warp - 1 (accumulator k0) ================================
    const int k0 = warp - 1;
    const uint32_t idesc = make_idesc_tf32(WK_N) | (1u << 15) | (1u << 16);
    const uint32_t a_base = desc_lo(smem_u32(sA), 1024), b_base = desc_lo(smem_u32(sB), 128);
    const uint32_t dcol = tmem_base + (uint32_t)(k0 * WK_N);
    int seqA_base = 0, seqB = 0;
    uint32_t acc = 0u, flushes = 0u;
    for (long long g = g0; g < g1;) {
      const WpSeg sg = wp_segment(g, g1, W, G.D0);
      g = sg.gnext;
      WP_TILE(sg.tile)
      const int nrows = min(TM1, G.D1 - y0);
      const int pmin = max(sg.za - 1, 0), pmax = min(sg.zb, G.D0 - 1);
      // parity class p0 of the output planes: the effective kernel has the d0 taps {p0, p0 + 1} only
      const int p0par = PAR ? ((sg.u / upp) >> 2) & 1 : 0;
      const bool my_tap = !PAR || k0 == p0par || k0 == p0par + 1;
      for (int zo = sg.za; zo < sg.zb; ++zo, ++seqB) {
        const int sb = seqB % WK_SB;
        mbar_wait(fullB + sb, (seqB / WK_SB) & 1);
        const int p = zo + k0 - 1;
        if (p >= 0 && p < G.D0) {                                  // warp-uniform
          const int seq = seqA_base + (p - pmin), sa = seq % WK_SA;
          mbar_wait(fullA + sa, (seq / WK_SA) & 1);
          // slab p is released by three arrivals (taps k0 = 0, 1, 2 at planes p+1, p, p-1); readers that fall outside
          // this segment's plane range are accounted for by the in-range reader at the range end they fall off
          const int extra = (zo == sg.za ? 2 - k0 : 0) + (zo == sg.zb - 1 ? k0 : 0);
          if (!my_tap) {                                           // tap absent from this parity class: release only
            if (lane == 0) {
              mbar_arrive(emptyB + sb);
              for (int e = 0; e <= extra; ++e) mbar_arrive(emptyA + sa);
            }
            __syncwarp();
            continue;
          }
          if (elect_one()) {
            const uint32_t alo = a_base + (uint32_t)sa * (SLAB_BYTES >> 4), blo = b_base + (uint32_t)sb * (WK_BSTAGE >> 4);
            if (nrows == TM1) umma_chain_mn16_ab(dcol, alo, blo, DESC_HI_MN_SW128_32B, idesc, acc);
            else umma_chain_mn_ab(nrows, dcol, alo, blo, DESC_HI_MN_SW128_32B, idesc, acc);
            umma_commit(emptyB + sb);
            umma_commit(emptyA + sa);
          }
          if (lane == 0)
            for (int e = 0; e < extra; ++e) mbar_arrive(emptyA + sa);
          __syncwarp();
          acc = 1u;
        } else {
          if (lane == 0) mbar_arrive(emptyB + sb);
          __syncwarp();
        }
      }
      seqA_base += pmax - pmin + 1;
      if (sg.unit_last) {
        if (elect_one()) umma_commit(accFull);
        __syncwarp();
        if (g < g1) {                                              // another unit follows: wait for the flush
          mbar_wait(accEmpty, flushes & 1u);
          tc_fence_after();
        }
        ++flushes;
        acc = 0u;
      }
    }
  } else {
    // ================================ flush (warps 4-7): TMEM -> atomics on dW, once per unit ================================
    const int q = warp & 3;                       // rows 32q..32q+31 <-> d1 tap k1 = q (q == 3: unused atom)
    const bool vec = (((uintptr_t)dw) & 15) == 0 && (G.Cout & 3) == 0;
    uint32_t flushes = 0u, touched = 0u;
    for (long long g = g0; g < g1;) {
      const WpSeg sg = wp_segment(g, g1, W, G.D0);
      g = sg.gnext;
      for (int k0 = 0; k0 < 3; ++k0)               // accumulator k0 received MMAs iff some plane pairs with an in-volume slab
        if (max(sg.za, 1 - k0) <= min(sg.zb - 1, G.D0 - k0)) touched |= 1u << k0;
      if (!sg.unit_last) continue;
      const int par = PAR ? sg.u / upp : 0, uu = PAR ? sg.u % upp : sg.u;
      const int ch = uu / G.nNtiles, n0 = (uu % G.nNtiles) * 32;
      float* dwp = dw + (PAR ? (long long)par * 27 * G.Cin * G.Cout : 0);
      mbar_wait(accFull, flushes & 1u);
      tc_fence_after();
      const int valid = G.chunk_valid[ch];
      const int cin_idx = G.cin_off + (G.chunk_src[ch] ? G.C1 : 0) + G.chunk_c0[ch] + lane;
      if (q < 3) {
        for (int k0 = 0; k0 < 3; ++k0) {
          if (!((touched >> k0) & 1u)) continue;      // uniform
          if (PAR && k0 != ((par >> 2) & 1) && k0 != ((par >> 2) & 1) + 1) continue;   // tap absent from the class
          for (int cb = 0; cb < WK_N; cb += 16) {
            const int k2 = 2 - (cb >> 5), cobase = n0 + (cb & 31);  // column = (2 - k2) * 32 + co
            if (cobase >= G.Cout) continue;           // padded output channels (uniform)
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(k0 * WK_N + cb), v);
            tmem_ld_wait();
            if (lane < valid) {
              float* o = dwp + ((long long)(((k0 * 3 + q) * 3 + k2)) * G.Cin + cin_idx) * G.Cout + cobase;
              const int nv = G.Cout - cobase;
              if (vec && nv >= 16) {
#pragma unroll
                for (int e = 0; e < 16; e += 4)
                  red_add_v4(o + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
              } else if (vec && nv >= 8) {
                red_add_v4(o, __uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]), __uint_as_float(v[3]));
                red_add_v4(o + 4, __uint_as_float(v[4]), __uint_as_float(v[5]), __uint_as_float(v[6]), __uint_as_float(v[7]));
#pragma unroll
                for (int e = 8; e < 16; ++e)
                  if (e < nv) atomicAdd(o + e, __uint_as_float(v[e]));
              } else {
#pragma unroll
                for (int e = 0; e < 16; ++e)
                  if (e < nv) atomicAdd(o + e, __uint_as_float(v[e]));
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(accEmpty);
      ++flushes;
      touched = 0u;
    }
  }
#undef WP_TILE
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------
// descriptor probe: one CTA, operands A and B are [rows][32] fp32 matrices TMA-loaded whole (128 B rows, swizzled by
// the TMA unit), then `nk` MMAs with fully caller-specified shared-memory descriptors.  Used to establish which
// (unaligned start, base offset, LBO/SBO) combinations the tensor core reads consistently with the TMA swizzle, e.g.
// row-shifted views of one tile.  d[128][N] receives the accumulator.
// ---------------------------------------------------------------------------------------------------------
struct ProbeArgs {
  int rows_a, rows_b, N, mn_major, nk;
  uint32_t a_off, a_lbo, a_sbo, a_bo, a_step, b_off, b_lbo, b_sbo, b_bo, b_step, layout;
};

__global__ void __launch_bounds__(128, 1)
desc_probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  float* __restrict__ d, const ProbeArgs P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 256 * 128;
  uint64_t* bars = (uint64_t*)(sB + 256 * 128);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bars, 1); mbar_init(bars + 1, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bars, (uint32_t)(P.rows_a + P.rows_b) * 128u);
    tma_load_2d(&map_a, bars, sA, 0, 0);
    tma_load_2d(&map_b, bars, sB, 0, 0);
  }
  mbar_wait(bars, 0);
  tc_fence_after();
  if (warp == 0) {
    if (elect_one()) {
      uint32_t idesc = make_idesc_tf32(P.N);
      if (P.mn_major) idesc |= (1u << 15) | (1u << 16);
      const uint32_t hi_a = (P.a_sbo >> 4) | (1u << 14) | (P.a_bo << 17) | (P.layout << 29);
      const uint32_t hi_b = (P.b_sbo >> 4) | (1u << 14) | (P.b_bo << 17) | (P.layout << 29);
      for (int k = 0; k < P.nk; ++k) {
        const uint64_t da = ((uint64_t)hi_a << 32) | desc_lo(smem_u32(sA) + P.a_off + k * P.a_step, P.a_lbo);
        const uint64_t db = ((uint64_t)hi_b << 32) | desc_lo(smem_u32(sB) + P.b_off + k * P.b_step, P.b_lbo);
        umma_tf32(tmem_base, da, db, idesc, k > 0 ? 1u : 0u);
      }
      umma_commit(bars + 1);
    }
    __syncwarp();
  }
  mbar_wait(bars + 1, 0);
  tc_fence_after();
  const int r = warp * 32 + lane;
  for (int cb = 0; cb < P.N; cb += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb, v);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 16; ++e) d[(size_t)r * P.N + cb + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------------------
// microbenchmark: issue cost of tcgen05.mma.kind::tf32 (M=128, N, K=8, both operands from shared memory) as a
// function of N, of the number of accumulators cycled through and of the chain length on one accumulator.
// No TMA, shared memory contents are irrelevant.  out[blockIdx.x] = cycles per MMA.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
mma_microbench_kernel(float* __restrict__ out, int N, int nacc, int chain, int iters, int kmajor, int commit_every,
                      int cycle_addr) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;
  __shared__ uint64_t bar3;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (190 * 1024) / 4; i += blockDim.x) ((float*)smem)[i] = 1.0f + (float)(i & 255) * 0.001f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); fence_barrier_init(); fence_proxy_async(); }
  if (warp == 1) tmem_alloc(&tslot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_tf32(N) | (kmajor ? 0u : ((1u << 15) | (1u << 16)));
    const uint32_t hi = kmajor ? DESC_HI_K_SW128 : DESC_HI_MN_SW128_32B;
    const uint32_t alo = desc_lo(smem_u32(smem), kmajor ? 16 : 1024);
    const uint32_t blo = desc_lo(smem_u32(smem + 150 * 1024), kmajor ? 16 : 16384);
    long long t0 = 0, t1 = 0;
    if (kmajor >= 2) {
      // asm-chained issue exactly as the kernels do it: 2 = MN-major chains of 16 (+1024 B per K-step, A atoms 1024 B
      // apart, B atoms 16 KB apart), 3 = K-major chains of 4 (+32 B per K-step)
      const uint32_t idesc2 = make_idesc_tf32(N) | (kmajor == 2 ? ((1u << 15) | (1u << 16)) : 0u);
      const uint32_t a2 = desc_lo(smem_u32(smem), kmajor == 2 ? 1024 : 16);
      // mode 2 re-uses the scalar arguments: chain = B base (KB), commit_every = number of A slabs cycled through,
      // nacc = accumulators, cycle_addr = parked warps; B alternates between two 48 KB stages like the k2n kernel
      const uint32_t b_kb = kmajor == 2 ? (uint32_t)chain : 72u;
      const int nslab = kmajor == 2 ? (commit_every > 0 ? commit_every : 3) : 3;
      const uint32_t b2 = desc_lo(smem_u32(smem + b_kb * 1024), kmajor == 2 ? 16384 : 16);
      if (elect_one()) {
        t0 = clock64();
        int acc = 0, slab = 0, plane = 0;
        for (int i = 0; i < iters; i += (kmajor == 2 ? 16 : 4)) {
          if (kmajor == 2) {
            umma_chain_mn16(tb + (uint32_t)(acc * N), a2 + (uint32_t)slab * (SLAB_BYTES >> 4),
                            b2 + (uint32_t)(plane & 1) * (49152u >> 4), DESC_HI_MN_SW128_32B, idesc2, 1u);
            if (++slab == nslab) slab = 0;
          } else {
            umma_chain_k<4>(tb + (uint32_t)(acc * N), a2 + (uint32_t)((i >> 2) % 3) * 64, b2, DESC_HI_K_SW128, idesc2, 1u);
          }
          if (++acc == nacc) { acc = 0; ++plane; }
        }
        umma_commit(&bar);
      }
    } else
    if (elect_one()) {
      t0 = clock64();
      int acc = 0, c = 0;
      for (int i = 0; i < iters; ++i) {
        uint32_t ao = (uint32_t)((i & 3) * 2), bo = ao;
        if (cycle_addr) {        // like the convolution: 8 slab stages x 3 d1 taps for A, 9 taps for B
          ao += (uint32_t)(((i >> 2) % 3) * 64 + ((i / 27) % 8) * (SLAB_BYTES >> 4));
          bo += (uint32_t)(((i >> 2) % 9) * ((N * 128) >> 4));
        }
        umma_tf32_lh(tb + (uint32_t)(acc * N), alo + ao, blo + bo, hi, idesc, 1u);
        if (++c == chain) { c = 0; if (++acc == nacc) acc = 0; }
        if (commit_every > 0 && (i % commit_every) == commit_every - 1) umma_commit(&bar2);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    const long long tt = __shfl_sync(0xffffffffu, t0, 0) ;
    if (threadIdx.x == 0) out[blockIdx.x] = (float)(t1 - (t0 ? t0 : tt)) / (float)iters;
  }
  else if (kmajor >= 2 && warp <= cycle_addr) {
    mbar_wait(&bar, 0);        // asm modes: `cycle_addr` extra warps park on the barrier like the kernels' epilogue warps do
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tb, 512);
}

// y[v][c] = act(y[v][c] + bias[c]) in place: the follow-up pass of a split-K convolution (C % 4 == 0)
__global__ void bias_act_kernel(float* __restrict__ y, const float* __restrict__ bias, long long nvox, int C, int act) {
  const int c4 = C >> 2;
  const long long total = nvox * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) << 2;
    float4 v = reinterpret_cast<float4*>(y)[i];
    if (bias) { v.x += bias[c]; v.y += bias[c + 1]; v.z += bias[c + 2]; v.w += bias[c + 3]; }
    if (act) {
      v.x = v.x > 0.f ? v.x : __expf(v.x) - 1.f; v.y = v.y > 0.f ? v.y : __expf(v.y) - 1.f;
      v.z = v.z > 0.f ? v.z : __expf(v.z) - 1.f; v.w = v.w > 0.f ? v.w : __expf(v.w) - 1.f;
    }
    reinterpret_cast<float4*>(y)[i] = v;
  }
}

__device__ __forceinline__ float tf32_lo(float v) {
  uint32_t u = __float_as_uint(v);
  u = (u + 0xFFFu + ((u >> 13) & 1u)) & ~0x1FFFu;      // round to nearest even on the 13 dropped bits (what the TMA does)
  return v - __uint_as_float(u);
}
__global__ void tf32_residual_kernel(const float* __restrict__ x, float* __restrict__ lo, long long n) {
  const long long n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float4* l4 = reinterpret_cast<float4*>(lo);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x4 + i);
    l4[i] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) lo[(n4 << 2) + threadIdx.x] = tf32_lo(x[(n4 << 2) + threadIdx.x]);
}

// x2[v][0:C] = bf16(x - rne_tf32(x)),  x2[v][C:2C] = bf16(rne_tf32(x)): the two operands of the bf16 correction term of
// the compensated forward, written as ONE 2C-channel bf16 tensor (same bytes per voxel as C fp32 channels)
// bf16x3 (mode 1): x2[v][0:C] = x1 = bf16(x),  x2[v][C:2C] = bf16(x - x1): the operands of x1 w1 + x2 w1 + x1 w2
__global__ void tf32_split_bf16_kernel(const float* __restrict__ x, uint16_t* __restrict__ x2, long long nvox, int C,
                                       int mode) {
  const int c4 = C >> 2;                                      // C % 4 == 0 (host-checked)
  const long long total = nvox * c4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long v = i / c4;
    const int c = (int)(i - v * c4) << 2;
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + v * C + c));
    const float in[4] = {a.x, a.y, a.z, a.w};
    uint32_t lo[4], hi[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (mode == 1) {
        const uint32_t b1 = bf16_bits(in[e]);
        lo[e] = b1;                                              // first half of the row: x1
        hi[e] = bf16_bits(in[e] - __uint_as_float(b1 << 16));    // second half: x2 (the subtraction is exact)
        continue;
      }
      uint32_t u = __float_as_uint(in[e]);
      u = (u + 0xFFFu + ((u >> 13) & 1u)) & ~0x1FFFu;
      const float h = __uint_as_float(u);
      hi[e] = bf16_bits(h);
      lo[e] = bf16_bits(in[e] - h);
    }
    uint16_t* row = x2 + v * 2 * C;
    *reinterpret_cast<uint2*>(row + c) = make_uint2(lo[0] | (lo[1] << 16), lo[2] | (lo[3] << 16));
    *reinterpret_cast<uint2*>(row + C + c) = make_uint2(hi[0] | (hi[1] << 16), hi[2] | (hi[3] << 16));
  }
}

// ---------------------------------------------------------------------------------------------------------
// weight packing:  wp[chunk][k2][k0][k1][n][32]   (K-major rows of 32 input channels, zero padded)
//   mode 0 (forward):       value = w[k0][k1][k2][cin(chunk, s)][n]
//   mode 1 (data gradient): value = w[2-k0][2-k1][2-k2][n][cout(chunk, s)]   (n runs over the layer's Cin)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pack_weights_body(const float* __restrict__ w, float* __restrict__ wp, int C1, int C2,
                                                  int Cout, int mode, int Npad, int nchunks, int nch1, int round_rn) {
  const bool k2n_layout = mode >= 2 && mode != 5 && mode != 7 && mode != 9;
  const long long total = k2n_layout ? 9LL * 96 * 32 : (long long)nchunks * 27 * Npad * 32;
  const int Cin = (mode == 5 || mode == 7 || mode == 9) ? C1 : C1 + C2;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(t & 31);
    long long r = t >> 5;
    const int n = (int)(r % Npad); r /= Npad;
    const int k1 = (int)(r % 3); r /= 3;
    const int k0 = (int)(r % 3); r /= 3;
    const int k2 = (int)(r % 3);
    const int ch = (int)(r / 3);
    float val = 0.f;
    bool lo_part = mode == 6;      // compensated forward: this entry holds w - rna_tf32(w) (then rounded itself)
    if (mode == 7 && ch >= nch1) {
      // bf16 chunks of the hybrid compensated forward: K' = [w_hi (cn) ; w_lo (cn)] in 64-channel bf16 chunks, two bf16 per
      // float slot; they meet x2 = [x_lo | x_hi] (tf32_split_bf16_kernel).  C2 = (coff << 12) | cn, C1 = total Cin
      const int coff = C2 >> 12, cn = C2 & 4095;
      uint32_t pair = 0;
      for (int e = 0; e < 2; ++e) {
        const int kk = (ch - nch1) * 64 + 2 * s + e;
        float v = 0.f;
        if (kk < 2 * cn && n < Cout) {
          const float wv = w[((long long)((k0 * 3 + k1) * 3 + k2) * C1 + coff + (kk < cn ? kk : kk - cn)) * Cout + n];
          uint32_t u;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(wv));
          v = kk < cn ? __uint_as_float(u) : wv - __uint_as_float(u);
        }
        pair |= bf16_bits(v) << (16 * e);
      }
      wp[t] = __uint_as_float(pair);
      continue;
    }
    if (mode == 9) {
      // bf16x3 compensated forward: 64-channel bf16 chunks (two bf16 per float slot) of the input channels [coff, coff + cn)
      // of a (27, C1, Cout) kernel; chunks [0, nch1) hold w1 = bf16(w), chunks [nch1, 2 nch1) hold w2 = bf16(w - w1).
      // Rows past cn are zero (a partial chunk of x1 reads on into the x2 half of the activation row).
      const int coff = C2 >> 12, cn = C2 & 4095;
      const bool second = ch >= nch1;
      uint32_t pair = 0;
      for (int e = 0; e < 2; ++e) {
        const int cl = (second ? ch - nch1 : ch) * 64 + 2 * s + e;
        uint32_t bits = 0;
        if (cl < cn && n < Cout) {
          const float wv = w[((long long)((k0 * 3 + k1) * 3 + k2) * C1 + coff + cl) * Cout + n];
          const uint32_t b1 = bf16_bits(wv);
          bits = second ? bf16_bits(wv - __uint_as_float(b1 << 16)) : b1;
        }
        pair |= bits << (16 * e);
      }
      wp[t] = __uint_as_float(pair);
      continue;
    }
    if (mode == 8) {
      // k2n layout (one 64-channel bf16 chunk, 2 C1 <= 64): row ((k0, k1), k2 * 32 + n), K' = [w_hi (C1) ; w_lo (C1)]
      long long r2 = t >> 5;
      const int nn = (int)(r2 % 96); r2 /= 96;
      const int q1 = (int)(r2 % 3);
      const int q0 = (int)(r2 / 3);
      const int q2 = nn >> 5, no = nn & 31;
      uint32_t pair = 0;
      for (int e = 0; e < 2; ++e) {
        const int kk = 2 * s + e;
        float v = 0.f;
        if (kk < 2 * C1 && no < Cout) {
          const float wv = w[((long long)((q0 * 3 + q1) * 3 + q2) * C1 + (kk < C1 ? kk : kk - C1)) * Cout + no];
          uint32_t u;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(wv));
          v = kk < C1 ? __uint_as_float(u) : wv - __uint_as_float(u);
        }
        pair |= bf16_bits(v) << (16 * e);
      }
      wp[t] = __uint_as_float(pair);
      continue;
    }
    if (k2n_layout) {
      // d2-taps-in-N layout of conv3d_tc_k2n_kernel: t = ((k0 * 3 + k1) * 96 + (k2 * 32 + n)) * 32 + s, one 32-channel chunk
      long long r2 = t >> 5;
      const int nn = (int)(r2 % 96); r2 /= 96;
      const int q1 = (int)(r2 % 3);
      const int q0 = (int)(r2 / 3);
      const int q2 = nn >> 5, no = nn & 31;
      if (mode == 4) {      // channel part of a (concatenated) input: C1 = total Cin, C2 = (first channel << 8) | channels
        const int coff = C2 >> 8, cn = C2 & 255;
        if (s < cn && no < Cout) val = w[((long long)((q0 * 3 + q1) * 3 + q2) * C1 + coff + s) * Cout + no];
      } else
      if (mode == 2 || mode == 6) { if (s < C1 && no < Cout) val = w[((long long)((q0 * 3 + q1) * 3 + q2) * Cin + s) * Cout + no]; }
      else { if (s < Cout && no < Cin) val = w[((long long)(((2 - q0) * 3 + (2 - q1)) * 3 + (2 - q2)) * Cin + no) * Cout + s]; }
    } else
    if (mode == 5 || mode == 7) {
      // hi / lo split of the input channels [coff, coff + cn) of a (27, C1, Cout) kernel: chunks [0, nch1) hold rna(w),
      // chunks [nch1, 2 nch1) hold rna(w - rna(w)); C2 = (coff << 12) | cn   (mode 7: only the hi chunks get here)
      const int coff = C2 >> 12, cn = C2 & 4095;
      lo_part = ch >= nch1;
      const int cl = (lo_part ? ch - nch1 : ch) * 32 + s;
      if (cl < cn && n < Cout) val = w[((long long)((k0 * 3 + k1) * 3 + k2) * Cin + coff + cl) * Cout + n];
    } else
    if (mode == 0) {
      int c;   // concat channel index
      if (ch < nch1) c = ch * 32 + s < C1 ? ch * 32 + s : -1;
      else c = (ch - nch1) * 32 + s < C2 ? C1 + (ch - nch1) * 32 + s : -1;
      if (c >= 0 && n < Cout) val = w[((long long)((k0 * 3 + k1) * 3 + k2) * Cin + c) * Cout + n];
    } else {
      const int co = ch * 32 + s;      // K runs over the layer's output channels
      if (co < Cout && n < Cin) val = w[((long long)(((2 - k0) * 3 + (2 - k1)) * 3 + (2 - k2)) * Cin + n) * Cout + co];
    }
    if (lo_part) {
      uint32_t u;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(val));
      val -= __uint_as_float(u);                   // exact in fp32
    }
    if (round_rn || mode >= 5) {   // round-to-nearest TF32 (the tensor core would otherwise truncate the low 13 mantissa bits)
      uint32_t u;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(val));
      val = __uint_as_float(u);
    }
    wp[t] = val;
  }
}
__global__ void pack_weights_kernel(const float* __restrict__ w, float* __restrict__ wp, int C1, int C2, int Cout,
                                    int mode, int Npad, int nchunks, int nch1, int round_rn) {
  pack_weights_body(w, wp, C1, C2, Cout, mode, Npad, nchunks, nch1, round_rn);
}
// all layers of a network in one launch: blockIdx.y = job, jobs[j] = {w, wp, C1, C2, Cout, mode} (device array)
__global__ void pack_weights_batch_kernel(const long long* __restrict__ jobs, int round_rn) {
  const long long* j = jobs + (long long)blockIdx.y * 6;
  const int C1 = (int)j[2], C2 = (int)j[3], Cout = (int)j[4], mode = (int)j[5];
  int Npad, nch, nch1;
  if (mode == 5) { Npad = (Cout + 15) / 16 * 16; nch1 = ((C2 & 4095) + 31) / 32; nch = 2 * nch1; }
  else if (mode == 7) { Npad = (Cout + 15) / 16 * 16; nch1 = ((C2 & 4095) + 31) / 32; nch = nch1 + (2 * (C2 & 4095) + 63) / 64; }
  else if (mode == 9) { Npad = (Cout + 15) / 16 * 16; nch1 = ((C2 & 4095) + 63) / 64; nch = 2 * nch1; }
  else if (mode >= 2) { Npad = 96; nch = 1; nch1 = 1; }
  else if (mode == 0) { Npad = (Cout + 15) / 16 * 16; nch1 = (C1 + 31) / 32; nch = nch1 + (C2 + 31) / 32; }
  else { Npad = (C1 + C2 + 15) / 16 * 16; nch = (Cout + 31) / 32; nch1 = nch; }
  pack_weights_body(reinterpret_cast<const float*>(j[0]), reinterpret_cast<float*>(j[1]), C1, C2, Cout, mode, Npad, nch,
                    nch1, round_rn);
}

// ---------------------------------------------------------------------------------------------------------
// host side: TMA descriptors
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// Activations are fp32 in HBM; CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 makes the TMA unit round them to nearest TF32 on the
// way into shared memory (measured: conv error 4.7e-4 -> 2.9e-4 rel. L2 vs feeding raw fp32 bits, which the tensor
// core truncates; profiles/r01_tf32_precision.txt).  SSR_TMA_DTYPE=f32 restores the raw-bits behaviour.
CUtensorMapDataType tma_dtype() {
  const char* e = getenv("SSR_TMA_DTYPE");
  if (e && strcmp(e, "f32") == 0) return CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  return CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
}

int make_map_act(CUtensorMap* m, const float* ptr, int C, int B, int D0, int D1, int D2, int box_d1 = TM1 + 2,
                 CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, int box_d2 = TM2) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { ssr_set_error("cuTensorMapEncodeTiled not available"); return SSR_ERR_CUDA; }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)D2, (cuuint64_t)D1, (cuuint64_t)D0, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)C * 4, (cuuint64_t)D2 * C * 4, (cuuint64_t)D1 * D2 * C * 4,
                           (cuuint64_t)D0 * D1 * D2 * C * 4};
  cuuint32_t box[5] = {32, (cuuint32_t)box_d2, (cuuint32_t)box_d1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, tma_dtype(), 5, (void*)ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { ssr_set_error("cuTensorMapEncodeTiled(activation C=%d %dx%dx%d) failed: %d", C, D0, D1, D2, (int)r); return SSR_ERR_CUDA; }
  return SSR_OK;
}

// bf16 activation tensor [B, D0, D1, D2, C2] (C2 = 2 C: [x_lo | x_hi], tf32_split_bf16_kernel): boxes of 64 channels = 128 B
int make_map_act16(CUtensorMap* m, const void* ptr, int C2, int B, int D0, int D1, int D2, int box_d1 = TM1 + 2,
                   int box_d2 = TM2) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { ssr_set_error("cuTensorMapEncodeTiled not available"); return SSR_ERR_CUDA; }
  cuuint64_t dims[5] = {(cuuint64_t)C2, (cuuint64_t)D2, (cuuint64_t)D1, (cuuint64_t)D0, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)C2 * 2, (cuuint64_t)D2 * C2 * 2, (cuuint64_t)D1 * D2 * C2 * 2,
                           (cuuint64_t)D0 * D1 * D2 * C2 * 2};
  cuuint32_t box[5] = {64, (cuuint32_t)box_d2, (cuuint32_t)box_d1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { ssr_set_error("cuTensorMapEncodeTiled(bf16 activation C=%d %dx%dx%d) failed: %d", C2, D0, D1, D2, (int)r); return SSR_ERR_CUDA; }
  return SSR_OK;
}

int make_map_w(CUtensorMap* m, const float* ptr, long long rows, int NT) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { ssr_set_error("cuTensorMapEncodeTiled not available"); return SSR_ERR_CUDA; }
  cuuint64_t dims[2] = {32, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {32, (cuuint32_t)NT};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { ssr_set_error("cuTensorMapEncodeTiled(weights rows=%lld NT=%d) failed: %d", rows, NT, (int)r); return SSR_ERR_CUDA; }
  return SSR_OK;
}

int round_up(int a, int b) { return (a + b - 1) / b * b; }

int pick_nt(int Npad) {
  if (Npad <= 192) return Npad;
  for (int parts = 2; parts <= 8; ++parts)
    if (Npad % parts == 0 && (Npad / parts) % 16 == 0 && Npad / parts <= 192) return Npad / parts;
  return 16;
}

template <int EPI, int NB, bool ACC = false>
static int launch_k2n(unsigned grid, size_t smem, cudaStream_t st, const CUtensorMap& mx, const CUtensorMap& mw,
                      const float* bias, float* y, const KfGeom& G, const float* elu_h, float* dbias, double* sums) {
  static bool attr_set = false;
  if (!attr_set) {
    SSR_CHECK_CUDA(cudaFuncSetAttribute(conv3d_tc_k2n_kernel<EPI, NB, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  conv3d_tc_k2n_kernel<EPI, NB, ACC><<<grid, 416, smem, st>>>(mx, mw, bias, y, G, elu_h, dbias, sums);   // TMA + 4 MMA + 8 epilogue warps
  return SSR_OK;
}

}  // namespace

extern "C" {

// jobs: DEVICE array of njobs x {w pointer, wp pointer, Cin1, Cin2, Cout, mode} (int64); one launch packs them all
int ssr_conv3d_pack_weights_batch(const long long* jobs, int njobs, void* stream) {
  SSR_CHECK_ARG(jobs && njobs > 0 && njobs <= 65535, "pack batch args");
  pack_weights_batch_kernel<<<dim3(148, (unsigned)njobs), 256, 0, (cudaStream_t)stream>>>(jobs, getenv("SSR_PACK_TRUNC") ? 0 : 1);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}
long long ssr_conv3d_packed_size(int Cin1, int Cin2, int Cout, int mode) {
  if (mode == 5) return 2LL * (((Cin2 & 4095) + 31) / 32) * 27 * round_up(Cout, 16) * 32;
  if (mode == 7) return (long long)(((Cin2 & 4095) + 31) / 32 + (2 * (Cin2 & 4095) + 63) / 64) * 27 * round_up(Cout, 16) * 32;
  if (mode == 9) return 2LL * (((Cin2 & 4095) + 63) / 64) * 27 * round_up(Cout, 16) * 32;
  if (mode >= 2) return 9LL * 96 * 32;
  if (mode == 0) {
    const int nch = (Cin1 + 31) / 32 + (Cin2 + 31) / 32;
    return (long long)nch * 27 * round_up(Cout, 16) * 32;
  }
  const int nch = (Cout + 31) / 32;
  return (long long)nch * 27 * round_up(Cin1 + Cin2, 16) * 32;
}

int ssr_conv3d_pack_weights(const float* w, float* wp, int Cin1, int Cin2, int Cout, int mode, void* stream) {
  SSR_CHECK_ARG(w && wp && Cin1 > 0 && Cin2 >= 0 && Cout > 0 && mode >= 0 && mode <= 9, "pack args");
  SSR_CHECK_ARG(mode < 2 || mode == 4 || mode == 5 || mode == 7 || mode == 9 || (Cin2 == 0 && Cin1 <= 32 && Cout <= 32), "k2n packing needs Cin <= 32, Cout <= 32");
  SSR_CHECK_ARG((mode != 5 && mode != 7 && mode != 9) || ((Cin2 & 4095) > 0 && (Cin2 >> 12) + (Cin2 & 4095) <= Cin1),
                "hi/lo packing: Cin1 = total input channels, Cin2 = (first channel << 12) | channels");
  SSR_CHECK_ARG(mode != 4 || ((Cin2 & 255) <= 32 && (Cin2 >> 8) + (Cin2 & 255) <= Cin1 && Cout <= 32),
                "part packing: Cin1 = total input channels, Cin2 = (first channel << 8) | channels (<= 32)");
  int Npad, nch, nch1;
  if (mode == 5) { Npad = round_up(Cout, 16); nch1 = ((Cin2 & 4095) + 31) / 32; nch = 2 * nch1; }
  else if (mode == 7) { Npad = round_up(Cout, 16); nch1 = ((Cin2 & 4095) + 31) / 32; nch = nch1 + (2 * (Cin2 & 4095) + 63) / 64; }
  else if (mode == 9) { Npad = round_up(Cout, 16); nch1 = ((Cin2 & 4095) + 63) / 64; nch = 2 * nch1; }
  else if (mode >= 2) { Npad = 96; nch = 1; nch1 = 1; }
  else if (mode == 0) { Npad = round_up(Cout, 16); nch1 = (Cin1 + 31) / 32; nch = nch1 + (Cin2 + 31) / 32; }
  else { Npad = round_up(Cin1 + Cin2, 16); nch = (Cout + 31) / 32; nch1 = nch; }
  const long long total = (mode >= 2 && mode != 5 && mode != 7 && mode != 9) ? 9LL * 96 * 32 : (long long)nch * 27 * Npad * 32;
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  pack_weights_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(w, wp, Cin1, Cin2, Cout, mode, Npad, nch, nch1,
                                                                      getenv("SSR_PACK_TRUNC") ? 0 : 1);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// Tile shape of conv3d_tc_kernel for one layer: G.NT / G.TZ / G.ksplit and the plane-linearised geometry (G.pl_*).
// Needs G.Npad set.  epi != 0 (fused epilogue requested) excludes split-K.
static int tc_tile_shape(TcGeom& G, int C1, int C2, int Cout, int B, int D0, int D1, int D2, int epi, int comp) {
  int ks_total = 0;
  for (int c = 0; c < C1; c += 32) ks_total += ((C1 - c < 32 ? C1 - c : 32) + 7) / 8;
  for (int c = 0; c < C2; c += 32) ks_total += ((C2 - c < 32 ? C2 - c : 32) + 7) / 8;
  if (comp == 3) ks_total = ks_total / 2 * 3;
  if (comp == 4) ks_total = ks_total / 2 + (2 * C1 + 15) / 16;
  if (comp == 5) { ks_total = 0; for (int c = 0; c < C1; c += 64) ks_total += 3 * (((C1 - c < 64 ? C1 - c : 64) + 15) / 16); }
  int n1t = (D1 + TM1 - 1) / TM1, n2t = (D2 + TM2 - 1) / TM2;
  // plane-linearised tiling for the small deep levels: windows of 128 rows of the padded plane instead of 16 x 8 tiles
  const int pitch = D2 + 2;
  const int nwin = ((D1 - 1) * pitch + D2 + 127) / 128;
  const int pl_slab = round_up((nwin * 128 + 2 * pitch + 2) * 128, 1024);
  // measured (160^3 net): 10^3 layers 74 vs 80 us with plane tiles; at 20^3 the 71 KB slabs leave only two pipeline
  // stages and the layers get slower (82 vs 75 us), so plane tiles are used while a slab stays at the 18 x 8 slab's size
  const int pl_max = getenv("SSR_PLANE_TILES_MAX_KB") ? atoi(getenv("SSR_PLANE_TILES_MAX_KB")) * 1024 : 24 * 1024;
  if (nwin < n1t * n2t && pl_slab <= pl_max && D1 + 2 <= 256 && pitch <= 256 && !getenv("SSR_NO_PLANE_TILES")) {
    G.pl_pitch = pitch; G.pl_slab = pl_slab; G.pl_tx = (D1 + 2) * pitch * 128;
    n1t = 1; n2t = nwin;
  }
  // shared-memory fit of an N tile: two B groups (all 9 (k0,k1) taps of a d2 tap when they fit, else 3) + >= 2 slabs
  const int slab_b = G.pl_pitch > 0 ? G.pl_slab : SLAB_BYTES;
  const int avail = 227 * 1024 - 1024 - TC_TAIL_BYTES - (epi ? 2 * 4 * 576 : 0);
  auto fits = [&](int nt) { return SB * 3 * nt * 128 + 2 * slab_b <= avail; };
  double best = 1e300;
  int best_nt = 0, best_tz = 0, best_ks = 1;
  // split-K candidates: only where the fused epilogues are not requested, the output rows are float4-addressable and
  // the layer is small (the parts meet in y through atomics: 2 x ksplit passes over y instead of 1)
  const int nchunks_all = comp == 5 ? 3 * ((C1 + 63) / 64) : comp == 4 ? (C1 + 31) / 32 + (2 * C1 + 63) / 64 : comp ? comp * ((C1 + 31) / 32)
                                                                                : (C1 + 31) / 32 + (C2 + 31) / 32;
  const bool may_split = epi == 0 && Cout % 4 == 0 && !getenv("SSR_NO_SPLIT_K") &&
                         (long long)B * D0 * D1 * D2 <= 27000;
  // Default: <= 4 planes per tile in a ring of two tiles' planes, L2 term on the small (split-K) levels only.
  // SSR_TC_DEEP_TILES=1 lets the L2-traffic term choose on EVERY level (up to 5 planes of 48 channels): fewer operand bytes
  // per output, but measured slower at 80^3 (48 -> 48: 0.318 ms against 0.301; data gradient 0.193 against 0.172) -- 14.00
  // against 13.92 ms per step (scripts/gpu/r02_s.sh).  SSR_TC_MID_TILES=1 does so on the 40^3-class levels only (27k .. 100k
  // voxels), where single layers measured 5 - 8 % faster -- and the whole step 0.1 ms SLOWER (13.43 / 13.45 against 13.29 /
  // 13.37 ms, scripts/gpu/r02_w.sh).  Both stay experiments.
  const long long nvox_all = (long long)B * D0 * D1 * D2;
  const bool old_tiles = getenv("SSR_TC_DEEP_TILES") == nullptr &&
                         !(nvox_all > 27000 && nvox_all <= 100000 && getenv("SSR_TC_MID_TILES"));
  for (int ksp = 1; ksp <= (may_split ? 8 : 1) && ksp <= nchunks_all; ++ksp) {
    for (int nt = 16; nt <= 192 && nt <= G.Npad; nt += 16) {
      if (G.Npad % nt || !fits(nt)) continue;
      // plane accumulators that fit the TMEM (and the barrier arrays).  A tile's planes all complete within its last few
      // slabs, so the ring must hold TWO tiles' planes for the epilogue burst to hide behind the next tile: with 8 planes in
      // 10 slots the MMA warps of the next tile caught up with the four epilogue warps (48 -> 48 at 80^3: 0.23 ms against
      // 0.16 ms, gpurun_out/r02q_layer_times.txt)
      const int slots_max = 512 / nt < 10 ? 512 / nt : 10;
      for (int tz = 1; tz <= (old_tiles ? 4 : 8) && tz <= D0 && (old_tiles ? tz * nt <= 256 : 2 * tz <= slots_max); ++tz) {
        const long long tiles = (long long)B * ((D0 + tz - 1) / tz) * n1t * n2t * (G.Npad / nt) * ksp;
        const long long rounds = (tiles + 147) / 148;
        // + issue overhead: exposed with a single issuing warp (tz == 1), mostly hidden with one warp per accumulator
        const double mma = (nt / 2 > 32 + nt / 4 ? nt / 2 : 32 + nt / 4) + (tz == 1 ? 20.0 : 8.0);
        const double ks_part = (double)((ks_total + ksp - 1) / ksp);
        double cost = rounds * (tz * 27.0 * ks_part * mma * (old_tiles ? 1.0 + 0.04 * (4 - tz) : 1.0) + 2500.0);
        if (may_split || !old_tiles) {
          // L2 -> shared-memory traffic of the whole launch: every tile streams its weight slice (27 taps x 128-byte rows x
          // nt per chunk) and its activation slabs: with all 9 (k0, k1) taps of a d2 tap resident (KG = 3, nt <= 64) a slab
          // serves up to 3 output planes, tz + 2 slabs per (chunk, d2 tap); otherwise every plane is loaded once per d0 tap.
          // THIS bounds the kernel on every level, not the MMA chain: 48 -> 48 at 80^3 moves 984 MB in 158 us, 6.2 TB/s, with
          // the tensor pipe 42 % busy (profiles/r02_conv_schemes_ncu.txt); 192 -> 192 at 20^3 480 MB in 82 us.  More planes
          // per tile share the weights and shrink the d0 halo; split-K restores the CTA count on the small levels.
          const double nch_part = (double)((nchunks_all + ksp - 1) / ksp);
          const bool kg3 = 9 * nt * 128 <= 74 * 1024 && !(G.pl_pitch > 0 && SB * 9 * nt * 128 + 2 * slab_b > avail);
          const double slabs = 3.0 * (kg3 ? tz + 2 : 3 * tz);
          const double l2_bytes = old_tiles ? (double)tiles * (27.0 * ks_part * 8.0 * nt * 4.0 + (ks_part / 4.0) * 3.0 * (tz + 2) * slab_b)
                                            : (double)tiles * nch_part * (27.0 * nt * 128.0 + slabs * slab_b);
          const double l2_cycles = l2_bytes / 3000.0 + 2500.0;          // ~5.7 TB/s at 1.9 GHz
          if (l2_cycles > cost) cost = l2_cycles;
        }
        // memset + ksp red.add passes (read-modify-write in L2) + the bias / activation pass over y, at ~3 TB/s of L2
        // traffic shared by all CTAs (1.9 GHz: 6.5e-4 cycles per byte)
        if (ksp > 1) cost += 6000.0 + (2.0 * ksp + 3.0) * (double)B * D0 * D1 * D2 * Cout * 4.0 * 6.5e-4;
        if (cost < best * 0.999 || (cost < best * 1.001 && ksp == best_ks && (tz > best_tz || (tz == best_tz && nt > best_nt)))) {
          best = cost; best_nt = nt; best_tz = tz; best_ks = ksp;
        }
      }
    }
  }
  SSR_CHECK_ARG(best_nt > 0, "no tile shape");
  G.NT = best_nt; G.TZ = best_tz; G.ksplit = best_ks;
  return SSR_OK;
}

// y[B,D0,D1,D2,Cout] = act(conv3x3x3([x1,x2], wp) + bias) ; wp from ssr_conv3d_pack_weights.
// Used for the data gradient too (x1 = dy, wp packed with mode 1, Cout = layer's Cin, bias NULL, act 0).
static int conv3d_fwd_tc_impl(const float* x1, int C1, const float* x2, int C2, const float* wp, const float* bias, float* y,
                              int B, int D0, int D1, int D2, int Cout, int act, int accumulate, void* stream, int epi = 0,
                              const float* elu_h = nullptr, float* dbias = nullptr, double* sums = nullptr, int comp = 0) {
  // comp (compensated forward, "3xTF32"): x2 = x1 - rne_tf32(x1) (ssr_tf32_residual), wp = hi/lo packing (mode 5);
  // K = [x1 | x2 | x1] against [w_hi | w_hi | w_lo] (comp == 3), or [x1 | x2] against [w_hi | w_hi] (comp == 2)
  // comp == 4 (hybrid, 2 MMA chains instead of 3): x2 = [x_lo | x_hi] as 2 C1 bf16 channels (ssr_tf32_split_bf16), wp = pack mode
  // 7; K = [x (TF32) | x2 (bf16)] against [w_hi (TF32) | w_hi ; w_lo (bf16)]: the corrections x_lo w_hi + x_hi w_lo are
  // ~2^-11 of the result, so bf16's 8 bits on their operands leave ~2^-20 -- at twice the K per MMA of TF32
  // comp == 5 (bf16x3, 1.5 chains): x2 = [x1 | x2] as 2 C1 bf16 channels (ssr_bf16x3_split), wp = pack mode 9; three bf16 terms
  // x1 w1 + x2 w1 + x1 w2 per 64-channel chunk -- what is dropped (x2 w2 and the third bf16 pieces) is ~2^-17 of a product
  SSR_CHECK_ARG(comp == 0 || ((comp >= 2 && comp <= 5) && x2 && C2 == C1), "compensated forward: x2 = residual of x1");
  const float* const bias_in = bias;
  SSR_CHECK_ARG(x1 && wp && y && B > 0 && D0 > 0 && D1 > 0 && D2 > 0 && Cout > 0, "pointers/shape");
  SSR_CHECK_ARG(C1 > 0 && C1 % 4 == 0 && C2 >= 0 && C2 % 4 == 0 && (C2 == 0 || x2), "channel counts must be multiples of 4");
  SSR_CHECK_ARG(C1 % 8 == 0 && C2 % 8 == 0, "channel counts must be multiples of 8 (TF32 K-step)");
  SSR_CHECK_ARG(((uintptr_t)x1 & 15) == 0 && ((uintptr_t)wp & 127) == 0 && (!x2 || ((uintptr_t)x2 & 15) == 0), "alignment");
  TcGeom G;
  memset(&G, 0, sizeof(G));
  G.B = B; G.D0 = D0; G.D1 = D1; G.D2 = D2; G.Cout = Cout; G.act = act; G.accumulate = accumulate;
  G.Npad = round_up(Cout, 16);
  // Tile shape (NT output channels x TZ output planes per CTA tile): the kernel is persistent with a static round-robin
  // over tiles, so the cost is rounds x per-tile time.  Small layers (40^3 and below) have few tiles: a narrower N tile or
  // fewer planes per tile trades MMA efficiency for SM utilisation.  MMA cost per instruction from the measured
  // max(N/2, 32 + N/4) (+ issue overhead), profiles/r01_mma_issue_microbench.txt.
  { const int rc_shape = tc_tile_shape(G, C1, C2, Cout, B, D0, D1, D2, epi, comp); if (rc_shape) return rc_shape; }
  G.nNtiles = G.Npad / G.NT;
  G.KG = (3 * 3 * G.NT * 128 <= 74 * 1024) ? 3 : 1;
  if (G.pl_pitch > 0 && SB * 9 * G.NT * 128 + 2 * G.pl_slab > 227 * 1024 - 1024 - TC_TAIL_BYTES - (epi ? 2 * 4 * 576 : 0)) G.KG = 1;
  // ring of plane accumulators: as many slots as fit the TMEM, at most two tiles' worth (and the barrier arrays' 10)
  G.nslot = 512 / G.NT < 2 * G.TZ ? 512 / G.NT : 2 * G.TZ;
  if (G.nslot > 10) G.nslot = 10;
  SSR_CHECK_ARG(G.nslot >= G.TZ && G.TZ <= 8, "accumulator ring");
  int cols = G.nslot * G.NT, pc = 32;
  while (pc < cols) pc <<= 1;
  G.tmem_cols = pc;
  if (getenv("SSR_TC_PRINT_TILES"))
    fprintf(stderr, "conv3d_tc %dx%dx%d C=%d+%d->%d comp=%d epi=%d: NT=%d TZ=%d KG=%d nslot=%d ksplit=%d pl=%d\n", D0, D1, D2, C1,
            comp ? 0 : C2, Cout, comp, epi, G.NT, G.TZ, G.KG, G.nslot, G.ksplit, G.pl_pitch);
  int nch = 0, nwch = 0;                       // chunks of K, chunks of the packed weights
  if (comp == 5) {
    const int nchb = (C1 + 63) / 64;
    for (int term = 0; term < 3; ++term)
      for (int j = 0; j < nchb; ++j) {
        SSR_CHECK_ARG(nch < MAX_CHUNKS, "too many input channels");
        const int left = C1 - 64 * j;
        G.chunk_src[nch] = 1; G.chunk_c0[nch] = (short)((term == 1 ? C1 : 0) + 64 * j);
        G.chunk_ks[nch] = (unsigned char)(((left < 64 ? left : 64) + 15) / 16);
        G.chunk_w[nch] = (unsigned char)((term == 2 ? nchb : 0) + j); G.chunk_f16[nch] = 1; ++nch;
      }
    nwch = 2 * nchb;
  } else if (comp == 4) {
    const int nchc = (C1 + 31) / 32, nch2 = (2 * C1 + 63) / 64;
    for (int c = 0; c < C1; c += 32) {
      G.chunk_src[nch] = 0; G.chunk_c0[nch] = (short)c; G.chunk_ks[nch] = (unsigned char)(((C1 - c < 32 ? C1 - c : 32) + 7) / 8);
      G.chunk_w[nch] = (unsigned char)nch; ++nch;
    }
    for (int j = 0; j < nch2; ++j) {
      SSR_CHECK_ARG(nch < MAX_CHUNKS, "too many input channels");
      const int left = 2 * C1 - 64 * j;
      G.chunk_src[nch] = 1; G.chunk_c0[nch] = (short)(64 * j); G.chunk_ks[nch] = (unsigned char)(((left < 64 ? left : 64) + 15) / 16);
      G.chunk_w[nch] = (unsigned char)(nchc + j); G.chunk_f16[nch] = 1; ++nch;
    }
    nwch = nchc + nch2;
  } else if (comp) {
    const int nchc = (C1 + 31) / 32;
    for (int term = 0; term < comp; ++term)
      for (int c = 0; c < C1; c += 32) {
        SSR_CHECK_ARG(nch < MAX_CHUNKS, "too many input channels");
        G.chunk_src[nch] = (unsigned char)(term == 1); G.chunk_c0[nch] = (short)c;
        G.chunk_ks[nch] = (unsigned char)(((C1 - c < 32 ? C1 - c : 32) + 7) / 8);
        G.chunk_w[nch] = (unsigned char)((term == 2 ? nchc : 0) + c / 32); ++nch;
      }
    nwch = 2 * nchc;
  } else {
    for (int c = 0; c < C1; c += 32) {
      SSR_CHECK_ARG(nch < MAX_CHUNKS, "too many input channels");
      G.chunk_src[nch] = 0; G.chunk_c0[nch] = (short)c; G.chunk_ks[nch] = (unsigned char)(((C1 - c < 32 ? C1 - c : 32) + 7) / 8);
      G.chunk_w[nch] = (unsigned char)nch; ++nch;
    }
    for (int c = 0; c < C2; c += 32) {
      SSR_CHECK_ARG(nch < MAX_CHUNKS, "too many input channels");
      G.chunk_src[nch] = 1; G.chunk_c0[nch] = (short)c; G.chunk_ks[nch] = (unsigned char)(((C2 - c < 32 ? C2 - c : 32) + 7) / 8);
      G.chunk_w[nch] = (unsigned char)nch; ++nch;
    }
    nwch = nch;
  }
  G.nchunks = nch;
  const bool pl = G.pl_pitch > 0;
  G.n2tiles = (D2 + TM2 - 1) / TM2; G.n1tiles = (D1 + TM1 - 1) / TM1; G.n0tiles = (D0 + G.TZ - 1) / G.TZ;
  if (pl) { G.n1tiles = 1; G.n2tiles = ((D1 - 1) * G.pl_pitch + D2 + 127) / 128; }
  const int slab_bytes = pl ? G.pl_slab : SLAB_BYTES;
  const int bgroup = G.KG * 3 * G.NT * 128;
  const int tail = TC_TAIL_BYTES /*barriers + bias*/ + (epi ? 2 * 4 * 576 : 0) /*per-CTA channel sums of the fused epilogues*/;
  const int budget = 227 * 1024 - 1024 /*align slack*/ - tail - SB * bgroup;
  int sa = budget / slab_bytes; if (sa > 8) sa = 8;
  SSR_CHECK_ARG(sa >= 2, "shared memory budget");
  G.SA = sa;
  const size_t smem = 1024 + (size_t)G.SA * slab_bytes + (size_t)SB * bgroup + tail;
  SSR_CHECK_ARG(epi == 0 || (epi == 1 ? (elu_h && dbias) : (epi == 2 && sums)), "fused epilogue buffers");
  SSR_CHECK_ARG(epi == 0 || !accumulate, "fused epilogues do not combine with accumulate");

  CUtensorMap m1, m2, mw;
  const int bx1 = pl ? D1 + 2 : TM1 + 2, bx2 = pl ? G.pl_pitch : TM2;        // TMA box: whole padded plane / 18 x 8 slab
  int rc = make_map_act(&m1, x1, C1, B, D0, D1, D2, bx1, CU_TENSOR_MAP_SWIZZLE_128B, bx2);
  if (rc) return rc;
  if (comp == 4 || comp == 5) { rc = make_map_act16(&m2, x2, 2 * C1, B, D0, D1, D2, bx1, bx2); if (rc) return rc; }
  else if (C2 > 0) { rc = make_map_act(&m2, x2, C2, B, D0, D1, D2, bx1, CU_TENSOR_MAP_SWIZZLE_128B, bx2); if (rc) return rc; } else m2 = m1;
  rc = make_map_w(&mw, wp, (long long)nwch * 27 * G.Npad, G.NT);
  if (rc) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    SSR_CHECK_CUDA(cudaFuncSetAttribute(conv3d_tc_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SSR_CHECK_CUDA(cudaFuncSetAttribute(conv3d_tc_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SSR_CHECK_CUDA(cudaFuncSetAttribute(conv3d_tc_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SSR_CHECK_CUDA(cudaFuncSetAttribute(conv3d_tc_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SSR_CHECK_CUDA(cudaFuncSetAttribute(conv3d_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SSR_CHECK_CUDA(cudaFuncSetAttribute(conv3d_tc_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const long long ntiles = (long long)B * G.n0tiles * G.n1tiles * G.n2tiles * G.nNtiles * G.ksplit;
  SSR_CHECK_ARG(ntiles < (1LL << 31) && G.Npad <= 576, "grid / channel count too large");
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    SSR_CHECK_CUDA(cudaGetDevice(&dev));
    SSR_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const unsigned grid = (unsigned)(ntiles < num_sms ? ntiles : num_sms);       // persistent: one CTA per SM
  // TMA + TZ MMA + 4 epilogue warps
  const unsigned nthr = 32 * (5 + G.TZ);
  cudaStream_t cst = (cudaStream_t)stream;
  const float* kbias = bias;
  if (G.ksplit > 1) {
    // split-K: the parts add raw partial sums into y (zeroed here unless it already holds the earlier channel parts);
    // bias + activation follow in bias_act_kernel
    const long long nvox = (long long)B * D0 * D1 * D2;
    if (!accumulate) SSR_CHECK_CUDA(cudaMemsetAsync(y, 0, (size_t)nvox * Cout * sizeof(float), cst));
    G.act = 0; G.accumulate = 0; kbias = nullptr;
  }
  bias = kbias;
  if (pl) {
    if (epi == 1) conv3d_tc_kernel<1, true><<<grid, nthr, smem, cst>>>(m1, m2, mw, bias, y, G, elu_h, dbias, sums);
    else if (epi == 2) conv3d_tc_kernel<2, true><<<grid, nthr, smem, cst>>>(m1, m2, mw, bias, y, G, elu_h, dbias, sums);
    else conv3d_tc_kernel<0, true><<<grid, nthr, smem, cst>>>(m1, m2, mw, bias, y, G, nullptr, nullptr, nullptr);
  } else {
    if (epi == 1) conv3d_tc_kernel<1, false><<<grid, nthr, smem, cst>>>(m1, m2, mw, bias, y, G, elu_h, dbias, sums);
    else if (epi == 2) conv3d_tc_kernel<2, false><<<grid, nthr, smem, cst>>>(m1, m2, mw, bias, y, G, elu_h, dbias, sums);
    else conv3d_tc_kernel<0, false><<<grid, nthr, smem, cst>>>(m1, m2, mw, bias, y, G, nullptr, nullptr, nullptr);
  }
  SSR_COUNT_LAUNCH();
  if (G.ksplit > 1 && (bias_in || act)) {
    const long long n4 = (long long)B * D0 * D1 * D2 * (Cout / 4);
    long long g = (n4 + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    bias_act_kernel<<<(unsigned)g, 256, 0, cst>>>(y, bias_in, (long long)B * D0 * D1 * D2, Cout, act);
    SSR_COUNT_LAUNCH();
  }
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// split-K factor conv3d_fwd_tc_impl would pick for this shape without fused epilogues (1 = none): callers that would ask
// for a fused epilogue (BatchNorm sums, ELU') use the separate passes instead where this is > 1
int ssr_conv3d_fwd_tc_ksplit(int C1, int C2, int Cout, int B, int D0, int D1, int D2, int comp) {
  SSR_CHECK_ARG(C1 > 0 && C2 >= 0 && Cout > 0 && B > 0 && D0 > 0 && D1 > 0 && D2 > 0, "shape");
  TcGeom G;
  memset(&G, 0, sizeof(G));
  G.Npad = round_up(Cout, 16);
  const int rc = tc_tile_shape(G, C1, comp ? C1 : C2, Cout, B, D0, D1, D2, 0, comp);
  return rc ? rc : G.ksplit;
}

int ssr_conv3d_fwd_tc(const float* x1, int C1, const float* x2, int C2, const float* wp, const float* bias, float* y,
                      int B, int D0, int D1, int D2, int Cout, int act, void* stream) {
  return conv3d_fwd_tc_impl(x1, C1, x2, C2, wp, bias, y, B, D0, D1, D2, Cout, act, 0, stream);
}
// same, added to the partial result already in y (then bias + activation): the skip part of a decoder convolution whose
// upsampled part was written by ssr_conv3d_fwd_tc_up
int ssr_conv3d_fwd_tc_acc(const float* x1, int C1, const float* x2, int C2, const float* wp, const float* bias, float* y,
                          int B, int D0, int D1, int D2, int Cout, int act, void* stream) {
  return conv3d_fwd_tc_impl(x1, C1, x2, C2, wp, bias, y, B, D0, D1, D2, Cout, act, 1, stream);
}

// forward + BatchNorm sums of the output in the epilogue (sums: 2*Cout doubles, zeroed here; finish with ssr_bn_finalize)
int ssr_conv3d_fwd_tc_stats(const float* x1, int C1, const float* x2, int C2, const float* wp, const float* bias, float* y,
                            double* sums, int B, int D0, int D1, int D2, int Cout, int act, void* stream) {
  SSR_CHECK_ARG(sums, "sums");
  SSR_CHECK_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)Cout * sizeof(double), (cudaStream_t)stream));
  return conv3d_fwd_tc_impl(x1, C1, x2, C2, wp, bias, y, B, D0, D1, D2, Cout, act, 0, stream, 2, nullptr, nullptr, sums);
}
// data gradient (wp packed with mode 1) fused with the ELU backward of the layer below: dx = conv(dy, wp) * elu'(h),
// dbias[c] += sum_v dx[v][c]; Cout = channels of dx / h
int ssr_conv3d_dgrad_tc_elu(const float* dy, int C, const float* wp, const float* h, float* dx, float* dbias, int B, int D0,
                            int D1, int D2, int Cout, void* stream) {
  return conv3d_fwd_tc_impl(dy, C, nullptr, 0, wp, nullptr, dx, B, D0, D1, D2, Cout, 0, 0, stream, 1, h, dbias, nullptr);
}

// Compensated forward ("3xTF32", fp32-class accuracy on the TF32 tensor cores): with x = x_hi + x_lo, w = w_hi + w_lo
// (x_hi = the TF32 value the TMA load produces, x_lo from ssr_tf32_residual; weights from pack mode 5) the convolution is
//   x * w ~= x_hi * w_hi + x_lo * w_hi + x_hi * w_lo          (level 3; the dropped x_lo * w_lo term is ~2^-22 relative)
// evaluated as ONE implicit GEMM whose K dimension is the concatenation [x | x_lo | x] against [w_hi | w_hi | w_lo].
// level 2 keeps only the activation correction ([x | x_lo] against [w_hi | w_hi]).  sums != NULL: BatchNorm sums in the
// epilogue (as ssr_conv3d_fwd_tc_stats); accumulate: add to the partial result already in y before bias / activation.
int ssr_conv3d_fwd_tc_comp(const float* x, const float* xlo, int C, const float* wp, const float* bias, float* y,
                           double* sums, int B, int D0, int D1, int D2, int Cout, int act, int accumulate, int level,
                           void* stream) {
  SSR_CHECK_ARG(level >= 2 && level <= 5, "compensation level must be 2, 3, 4 (hybrid TF32 + bf16) or 5 (bf16x3)");
  SSR_CHECK_ARG(!(sums && accumulate), "BatchNorm sums do not combine with accumulate");
  if (sums) SSR_CHECK_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)Cout * sizeof(double), (cudaStream_t)stream));
  return conv3d_fwd_tc_impl(x, C, xlo, C, wp, bias, y, B, D0, D1, D2, Cout, act, accumulate, stream, sums ? 2 : 0, nullptr,
                            nullptr, sums, level);
}

// ---- convolution over a 2x nearest-upsampled tensor from its LOW-resolution source (conv3d_tc_up_kernel) ------------
static int up_tile_shape(int Npad, int B, int D0, int D1, int D2, int ks_total, int kparts, int tile_mult, int* NT, int* TZ) {
  const int n1t = (D1 + TM1 - 1) / TM1, n2t = (D2 + TM2 - 1) / TM2;
  double best = 1e300;
  int best_nt = 0, best_tz = 0;
  for (int nt = 16; nt <= 192 && nt <= Npad; nt += 16) {
    if (Npad % nt) continue;
    for (int tz = 1; tz <= 4 && tz <= D0 && tz * nt <= 256; ++tz) {
      const long long tiles = (long long)B * ((D0 + tz - 1) / tz) * n1t * n2t * (Npad / nt) * tile_mult;
      const long long rounds = (tiles + 147) / 148;
      const double mma = (nt / 2 > 32 + nt / 4 ? nt / 2 : 32 + nt / 4) + (tz == 1 ? 20.0 : 8.0);
      const double cost = rounds * (tz * 8.0 * kparts * ks_total * mma * (1.0 + 0.04 * (4 - tz)) + 2500.0);
      if (cost < best * 0.999 || (cost < best * 1.001 && (tz > best_tz || (tz == best_tz && nt > best_nt)))) {
        best = cost; best_nt = nt; best_tz = tz;
      }
    }
  }
  *NT = best_nt; *TZ = best_tz;
  return best_nt > 0 ? SSR_OK : SSR_ERR_ARG;
}

static int make_map_view(CUtensorMap* m, const float* ptr, int C, int B, int D0, int D1, int D2, long long s2, long long s1,
                         long long s0, long long sb, int box_d1 = TM1 + 2, int box_d2 = TM2,
                         CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {      // strides in floats; dims (C, D2, D1, D0, B)
  EncodeTiledFn enc = get_encode();
  if (!enc) { ssr_set_error("cuTensorMapEncodeTiled not available"); return SSR_ERR_CUDA; }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)D2, (cuuint64_t)D1, (cuuint64_t)D0, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)s2 * 4, (cuuint64_t)s1 * 4, (cuuint64_t)s0 * 4, (cuuint64_t)sb * 4};
  cuuint32_t box[5] = {32, (cuuint32_t)box_d2, (cuuint32_t)box_d1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, tma_dtype(), 5, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { ssr_set_error("cuTensorMapEncodeTiled(strided view C=%d %dx%dx%d) failed: %d", C, D0, D1, D2, (int)r); return SSR_ERR_CUDA; }
  return SSR_OK;
}

// mode 1: x = low-resolution input [B,d0,d1,d2,C], y = [B,2d0,2d1,2d2,Cout] partial sums (no bias / activation);
// mode 2: x = full-resolution dy [B,2d0,2d1,2d2,C], y = gradient w.r.t. the low-resolution tensor [B,d0,d1,d2,Cout].
// wp8: 8 parity classes x standard packed weights (mode 0 / mode 1 packing of the effective kernels).
static int conv3d_tc_up_impl(int mode, const float* x, int C, const float* wp8, float* y, int B, int D0, int D1, int D2,
                             int Cout, void* stream, const float* xlo = nullptr, int comp = 0) {
  // comp (mode 1 only): compensated forward, see ssr_conv3d_fwd_tc_comp; wp8 = 8 parity classes x hi/lo packing (mode 5)
  SSR_CHECK_ARG(x && wp8 && y && B > 0 && D0 > 0 && D1 > 0 && D2 > 0 && Cout > 0, "pointers/shape");
  SSR_CHECK_ARG(comp == 0 || (mode == 1 && xlo && comp >= 2 && comp <= 5), "compensated parity forward args");
  SSR_CHECK_ARG(C > 0 && C % 8 == 0 &&
                    (comp == 5 ? 3 * ((C + 63) / 64) : (C + 31) / 32 * (comp == 4 ? 2 : comp ? comp : 1)) <= UP_MAX_CHUNKS,
                "channel count must be a multiple of 8 (<= 768; <= 256 compensated)");
  SSR_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)wp8 & 127) == 0 && ((uintptr_t)xlo & 15) == 0, "alignment");
  UpGeom G;
  memset(&G, 0, sizeof(G));
  G.B = B; G.D0 = D0; G.D1 = D1; G.D2 = D2; G.Cout = Cout;
  G.Npad = round_up(Cout, 16);
  int ks_total = 0, nch = 0;
  const int nchc = (C + 31) / 32;
  for (int term = 0; term < (comp == 5 ? 0 : comp == 4 ? 1 : comp ? comp : 1); ++term)
    for (int c = 0; c < C; c += 32) {
      G.chunk_c0[nch] = (short)c; G.chunk_ks[nch] = (unsigned char)(((C - c < 32 ? C - c : 32) + 7) / 8);
      G.chunk_src[nch] = (unsigned char)(term == 1); G.chunk_w[nch] = (unsigned char)((term == 2 ? nchc : 0) + c / 32);
      ks_total += G.chunk_ks[nch]; ++nch;
    }
  int nwch = comp ? 2 * nchc : nchc;                // chunks of the packed weights per parity class
  if (comp == 4) {                                  // hybrid: bf16 chunks of x2 = [x_lo | x_hi] against [w_hi ; w_lo]
    const int nch2 = (2 * C + 63) / 64;
    for (int j = 0; j < nch2; ++j) {
      const int left = 2 * C - 64 * j;
      G.chunk_c0[nch] = (short)(64 * j); G.chunk_ks[nch] = (unsigned char)(((left < 64 ? left : 64) + 15) / 16);
      G.chunk_src[nch] = 1; G.chunk_w[nch] = (unsigned char)(nchc + j); G.chunk_f16[nch] = 1;
      ks_total += G.chunk_ks[nch]; ++nch;
    }
    nwch = nchc + nch2;
  }
  if (comp == 5) {                                  // bf16x3: [x1 | x2 | x1] against [w1 ; w1 ; w2], all bf16
    const int nchb = (C + 63) / 64;
    for (int term = 0; term < 3; ++term)
      for (int j = 0; j < nchb; ++j) {
        const int left = C - 64 * j;
        G.chunk_c0[nch] = (short)((term == 1 ? C : 0) + 64 * j); G.chunk_ks[nch] = (unsigned char)(((left < 64 ? left : 64) + 15) / 16);
        G.chunk_src[nch] = 1; G.chunk_w[nch] = (unsigned char)((term == 2 ? nchb : 0) + j); G.chunk_f16[nch] = 1;
        ks_total += G.chunk_ks[nch]; ++nch;
      }
    nwch = 2 * nchb;
  }
  G.nchunks = nch;
  int rc = up_tile_shape(G.Npad, B, D0, D1, D2, ks_total, mode == 2 ? 8 : 1, mode == 1 ? 8 : 1, &G.NT, &G.TZ);
  if (rc) { ssr_set_error("no tile shape"); return rc; }
  G.nNtiles = G.Npad / G.NT;
  G.KG = (3 * 3 * G.NT * 128 <= 74 * 1024) ? 3 : 1;
  int cols = 2 * G.TZ * G.NT, pc = 32;
  while (pc < cols) pc <<= 1;
  G.tmem_cols = pc;
  G.par_rows = nwch * 27 * G.Npad;
  G.n2tiles = (D2 + TM2 - 1) / TM2; G.n1tiles = (D1 + TM1 - 1) / TM1; G.n0tiles = (D0 + G.TZ - 1) / G.TZ;
  const int bgroup = G.KG * 3 * G.NT * 128;
  const int budget = 227 * 1024 - 1024 - 2816 - SB * bgroup;
  int sa = budget / SLAB_BYTES; if (sa > 8) sa = 8;
  SSR_CHECK_ARG(sa >= 2, "shared memory budget");
  G.SA = sa;
  const size_t smem = 1024 + (size_t)G.SA * SLAB_BYTES + (size_t)SB * bgroup + 2816;
  UpMaps maps;
  memset(&maps, 0, sizeof(maps));
  if (mode == 1) {
    rc = make_map_act(&maps.x[0], x, C, B, D0, D1, D2);
    if (rc) return rc;
    for (int i = 1; i < 8; ++i) maps.x[i] = maps.x[0];
    if (comp == 4 || comp == 5) { rc = make_map_act16(&maps.x[1], xlo, 2 * C, B, D0, D1, D2); if (rc) return rc; }
    else if (comp) { rc = make_map_act(&maps.x[1], xlo, C, B, D0, D1, D2); if (rc) return rc; }
  } else {
    const long long F0 = 2LL * D0, F1 = 2LL * D1, F2 = 2LL * D2;
    for (int par = 0; par < 8; ++par) {
      const long long off = ((((par >> 2) & 1) * F1 + ((par >> 1) & 1)) * F2 + (par & 1)) * C;
      rc = make_map_view(&maps.x[par], x + off, C, B, D0, D1, D2, 2LL * C, 2 * F2 * C, 2 * F1 * F2 * C, F0 * F1 * F2 * C);
      if (rc) return rc;
    }
  }
  CUtensorMap mw;
  rc = make_map_w(&mw, wp8, 8LL * G.par_rows, G.NT);
  if (rc) return rc;
  const long long ntiles = (long long)B * G.n0tiles * G.n1tiles * G.n2tiles * G.nNtiles * (mode == 1 ? 8 : 1);
  SSR_CHECK_ARG(ntiles < (1LL << 31) && G.Npad <= 576, "grid / channel count too large");
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    SSR_CHECK_CUDA(cudaGetDevice(&dev));
    SSR_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const unsigned grid = (unsigned)(ntiles < num_sms ? ntiles : num_sms);
  if (mode == 1) {
    static bool attr1 = false;
    if (!attr1) { SSR_CHECK_CUDA(cudaFuncSetAttribute(conv3d_tc_up_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr1 = true; }
    conv3d_tc_up_kernel<1><<<grid, 32 * (5 + G.TZ), smem, (cudaStream_t)stream>>>(maps, mw, y, G);
  } else {
    static bool attr2 = false;
    if (!attr2) { SSR_CHECK_CUDA(cudaFuncSetAttribute(conv3d_tc_up_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr2 = true; }
    conv3d_tc_up_kernel<2><<<grid, 32 * (5 + G.TZ), smem, (cudaStream_t)stream>>>(maps, mw, y, G);
  }
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

int ssr_conv3d_fwd_tc_up(const float* low, int Cup, const float* wp8, float* y, int B, int d0, int d1, int d2, int Cout,
                         void* stream) {
  return conv3d_tc_up_impl(1, low, Cup, wp8, y, B, d0, d1, d2, Cout, stream);
}
// compensated parity forward (see ssr_conv3d_fwd_tc_comp): lowlo = ssr_tf32_residual(low), wp8c = 8 x pack mode 5
int ssr_conv3d_fwd_tc_up_comp(const float* low, const float* lowlo, int Cup, const float* wp8c, float* y, int B, int d0,
                              int d1, int d2, int Cout, int level, void* stream) {
  return conv3d_tc_up_impl(1, low, Cup, wp8c, y, B, d0, d1, d2, Cout, stream, lowlo, level);
}
int ssr_conv3d_dgrad_tc_up(const float* dy, int Cout_layer, const float* wp8, float* dlow, int B, int d0, int d1, int d2,
                           int Cup, void* stream) {
  return conv3d_tc_up_impl(2, dy, Cout_layer, wp8, dlow, B, d0, d1, d2, Cup, stream);
}

// k2n layout of ssr_conv3d_fwd_tc_up for Cout == 24, Cup <= 64 (the last decoder level): wpk from
// ssr_conv3d_pack_up_k2n (4 * 8 * 96 * 32 floats).  Same result contract as ssr_conv3d_fwd_tc_up.
int ssr_conv3d_pack_up_k2n(const float* weff, float* wpk, int Cup, void* stream) {
  SSR_CHECK_ARG(weff && wpk && Cup > 0 && Cup <= 64 && Cup % 8 == 0, "pack_up_k2n args");
  pack_up_k2n_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(weff, wpk, Cup, getenv("SSR_PACK_TRUNC") ? 0 : 1);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}
int ssr_conv3d_fwd_tc_up_k2n(const float* low, int Cup, const float* wpk, float* y, int B, int D0, int D1, int D2, int Cout,
                             void* stream) {
  SSR_CHECK_ARG(low && wpk && y && B > 0 && D0 > 0 && D1 > 0 && D2 > 0, "pointers/shape");
  SSR_CHECK_ARG(Cout == 24 && Cup > 0 && Cup <= 64 && Cup % 8 == 0, "the k2n parity forward needs Cout == 24 and Cup <= 64");
  SSR_CHECK_ARG(((uintptr_t)low & 15) == 0 && ((uintptr_t)wpk & 127) == 0 && ((uintptr_t)y & 15) == 0, "alignment");
  KuGeom G;
  memset(&G, 0, sizeof(G));
  G.B = B; G.D0 = D0; G.D1 = D1; G.D2 = D2; G.Cout = Cout;
  G.nch = (Cup + 31) / 32;
  for (int ch = 0; ch < G.nch; ++ch) G.nks[ch] = ((Cup - ch * 32 < 32 ? Cup - ch * 32 : 32) + 7) / 8;
  G.n1tiles = (D1 + KF_TM1 - 1) / KF_TM1; G.n2tiles = (D2 + KF_OUT2 - 1) / KF_OUT2;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    SSR_CHECK_CUDA(cudaGetDevice(&dev));
    SSR_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int per_class = num_sms / 4;               // CTAs per parity class
  SSR_CHECK_ARG(per_class >= 1, "device too small");
  const long long cols = (long long)B * G.n1tiles * G.n2tiles;
  long long best = -1; int best_nzr = 1;
  for (int nzr = 1; nzr <= D0; ++nzr) {
    const int zlen = (D0 + nzr - 1) / nzr;
    if (zlen < 8 && nzr > 1) break;
    const long long items = cols * ((D0 + zlen - 1) / zlen);
    const long long cost = ((items + per_class - 1) / per_class) * (zlen + 1);
    if (best < 0 || cost < best) { best = cost; best_nzr = (D0 + zlen - 1) / zlen; G.zlen = zlen; }
  }
  G.nzr = best_nzr;
  CUtensorMap mx, mw;
  int rc = make_map_act(&mx, low, Cup, B, D0, D1, D2, KF_TM1 + 2, CU_TENSOR_MAP_SWIZZLE_128B, KF_TM2);
  if (rc) return rc;
  rc = make_map_w(&mw, wpk, 4LL * 8 * KF_N, KF_N);
  if (rc) return rc;
  const size_t smem = 1024 + 8 * (size_t)KF_BTILE_BYTES + (size_t)KU_SA * KF_SLAB_BYTES + 32 * 8;
  static bool attr_set = false;
  if (!attr_set) {
    SSR_CHECK_CUDA(cudaFuncSetAttribute(conv3d_tc_up_k2n_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const long long nitems = cols * G.nzr;
  SSR_CHECK_ARG(nitems < (1LL << 31), "grid too large");
  long long ncta = nitems < per_class ? nitems : per_class;
  conv3d_tc_up_k2n_kernel<<<(unsigned)(4 * ncta), 416, smem, (cudaStream_t)stream>>>(mx, mw, y, G);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// wskip (27, Cskip, Cout) = skip part of w (27, Cskip + Cup, Cout); weff (8, 27, Cup, Cout) = effective kernels of the
// upsampled part per parity class (see up_weights_kernel)
int ssr_conv3d_up_weights(const float* w, int Cskip, int Cup, int Cout, float* wskip, float* weff, void* stream) {
  SSR_CHECK_ARG(w && wskip && weff && Cskip > 0 && Cup > 0 && Cout > 0, "args");
  const long long n = 8LL * 27 * Cup * Cout + 27LL * Cskip * Cout;
  long long g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  up_weights_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(w, Cskip, Cup, Cout, wskip, weff);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// One channel part (<= 32 channels starting at c0) of a convolution with Cout <= 32 through conv3d_tc_k2n_kernel.
// x: tensor with Ctot channels; wp: the part's weights (pack mode 2 / 3, or ssr_conv3d_pack_weights_part).  accumulate:
// add to the partial result already in y; final: apply bias + activation.  A concatenated input [x1, x2] is the sum of
// its parts: first part accumulate = 0, last part final = 1.
// epi: 0 plain, 1 multiply by elu'(elu_h) and accumulate dbias, 2 accumulate BatchNorm sums (sum | sum of squares)
static int conv3d_fwd_tc_k2n_impl(const float* x, int Ctot, int c0, int C, const float* wp, const float* bias, float* y,
                                  int B, int D0, int D1, int D2, int Cout, int act, int accumulate, int final, void* stream,
                                  int epi = 0, const float* elu_h = nullptr, float* dbias = nullptr, double* sums = nullptr,
                                  int f16 = 0, void* y2 = nullptr) {
  // f16: x is the bf16 tensor [x_lo | x_hi] of Ctot = C = 2 * (layer channels) <= 64 channels, wp from pack mode 8
  SSR_CHECK_ARG(x && wp && y && B > 0 && D0 > 0 && D1 > 0 && D2 > 0, "pointers/shape");
  SSR_CHECK_ARG(C > 0 && C <= (f16 ? 64 : 32) && C % (f16 ? 16 : 8) == 0 && Cout > 0 && Cout <= 32 && c0 >= 0 && c0 + C <= Ctot &&
                Ctot % 4 == 0 && (!f16 || (c0 == 0 && C == Ctot)),
                "k2n forward needs <= 32 input channels per part (multiple of 8), Cout <= 32");
  SSR_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)wp & 127) == 0, "alignment");
  KfGeom G;
  memset(&G, 0, sizeof(G));
  G.B = B; G.D0 = D0; G.D1 = D1; G.D2 = D2; G.Cout = Cout; G.act = act; G.nks = f16 ? C / 16 : C / 8;
  G.c0 = c0; G.accumulate = accumulate; G.final = final; G.f16 = f16;
  SSR_CHECK_ARG(!y2 || (final && Cout % 8 == 0 && ((uintptr_t)y2 & 15) == 0), "split output needs the final part, Cout % 8 == 0");
  G.y2 = reinterpret_cast<uint16_t*>(y2);
  G.n1tiles = (D1 + KF_TM1 - 1) / KF_TM1; G.n2tiles = (D2 + KF_OUT2 - 1) / KF_OUT2;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    SSR_CHECK_CUDA(cudaGetDevice(&dev));
    SSR_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  // d0 ranges: long enough that the two halo slabs are a small overhead, short enough for an even static round-robin:
  // pick the split count with the fewest (rounds x planes per CTA)
  const long long cols = (long long)B * G.n1tiles * G.n2tiles;
  long long best = -1; int best_nzr = 1;
  for (int nzr = 1; nzr <= D0; ++nzr) {
    const int zlen = (D0 + nzr - 1) / nzr;
    if (zlen < 8 && nzr > 1) break;
    const long long items = cols * ((D0 + zlen - 1) / zlen);
    const long long cost = ((items + num_sms - 1) / num_sms) * (zlen + 2);
    if (best < 0 || cost < best) { best = cost; best_nzr = (D0 + zlen - 1) / zlen; G.zlen = zlen; }
  }
  G.nzr = best_nzr;
  CUtensorMap mx, mw;
  int rc = f16 ? make_map_act16(&mx, x, Ctot, B, D0, D1, D2, KF_TM1 + 2, KF_TM2)
               : make_map_act(&mx, x, Ctot, B, D0, D1, D2, KF_TM1 + 2, CU_TENSOR_MAP_SWIZZLE_128B, KF_TM2);
  if (rc) return rc;
  rc = make_map_w(&mw, wp, 9 * KF_N, KF_N);
  if (rc) return rc;
  const size_t smem = 1024 + 9 * (size_t)KF_BTILE_BYTES + (size_t)KF_SA * KF_SLAB_BYTES + 24 * 8 + 128 + 64 * 8;
  const long long nitems = cols * G.nzr;
  SSR_CHECK_ARG(nitems < (1LL << 31), "grid too large");
  const unsigned grid = (unsigned)(nitems < num_sms ? nitems : num_sms);
  cudaStream_t st = (cudaStream_t)stream;
  if (epi != 0) {
    SSR_CHECK_ARG(final == 1 && (Cout == 24 || Cout == 32), "fused k2n epilogues need the final part and Cout = 24 or 32");
    SSR_CHECK_ARG(epi == 1 ? (elu_h && dbias && ((uintptr_t)elu_h & 15) == 0) : (epi == 2 && sums), "fused epilogue buffers");
  }
  if (epi == 1 && Cout == 24) rc = launch_k2n<1, 3>(grid, smem, st, mx, mw, bias, y, G, elu_h, dbias, sums);
  else if (epi == 1) rc = launch_k2n<1, 4>(grid, smem, st, mx, mw, bias, y, G, elu_h, dbias, sums);
  else if (epi == 2 && accumulate && Cout == 24) rc = launch_k2n<2, 3, true>(grid, smem, st, mx, mw, bias, y, G, elu_h, dbias, sums);
  else if (epi == 2 && accumulate) rc = launch_k2n<2, 4, true>(grid, smem, st, mx, mw, bias, y, G, elu_h, dbias, sums);
  else if (epi == 2 && Cout == 24) rc = launch_k2n<2, 3>(grid, smem, st, mx, mw, bias, y, G, elu_h, dbias, sums);
  else if (epi == 2) rc = launch_k2n<2, 4>(grid, smem, st, mx, mw, bias, y, G, elu_h, dbias, sums);
  else if (accumulate && Cout == 24 && !getenv("SSR_K2N_NO_ACC_PREFETCH")) rc = launch_k2n<0, 3, true>(grid, smem, st, mx, mw, bias, y, G, nullptr, nullptr, nullptr);
  else if (accumulate && Cout == 32 && !getenv("SSR_K2N_NO_ACC_PREFETCH")) rc = launch_k2n<0, 4, true>(grid, smem, st, mx, mw, bias, y, G, nullptr, nullptr, nullptr);
  else rc = launch_k2n<0, 4>(grid, smem, st, mx, mw, bias, y, G, nullptr, nullptr, nullptr);
  if (rc) return rc;
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

int ssr_conv3d_fwd_tc_k2n(const float* x, int C, const float* wp, const float* bias, float* y, int B, int D0, int D1,
                          int D2, int Cout, int act, void* stream) {
  return conv3d_fwd_tc_k2n_impl(x, C, 0, C, wp, bias, y, B, D0, D1, D2, Cout, act, 0, 1, stream);
}
int ssr_conv3d_fwd_tc_k2n_part(const float* x, int Ctot, int c0, int C, const float* wp, const float* bias, float* y, int B,
                               int D0, int D1, int D2, int Cout, int act, int accumulate, int final, void* stream) {
  return conv3d_fwd_tc_k2n_impl(x, Ctot, c0, C, wp, bias, y, B, D0, D1, D2, Cout, act, accumulate, final, stream);
}
// final part that ALSO writes y2 = [bf16(y_lo) | bf16(y_hi)] (what ssr_tf32_split_bf16(y) would produce) from its epilogue
int ssr_conv3d_fwd_tc_k2n_part_split(const float* x, int Ctot, int c0, int C, const float* wp, const float* bias, float* y,
                                     void* y2, int B, int D0, int D1, int D2, int Cout, int act, int accumulate,
                                     void* stream) {
  return conv3d_fwd_tc_k2n_impl(x, Ctot, c0, C, wp, bias, y, B, D0, D1, D2, Cout, act, accumulate, 1, stream, 0, nullptr,
                                nullptr, nullptr, 0, y2);
}

// Data gradient of a Cin, Cout <= 32 layer fused with the ELU backward of the convolution below it (replaces
// ssr_conv3d_fwd_tc_k2n + ssr_elu_bwd):  dx = conv(dy, wp) * elu'(h),  dbias[c] += sum_v dx[v][c].
// h: forward OUTPUT (post-ELU) of the layer whose pre-activation gradient dx is; Cout = channels of dx (24 or 32).
int ssr_conv3d_dgrad_tc_k2n_elu(const float* dy, int C, const float* wp, const float* h, float* dx, float* dbias, int B,
                                int D0, int D1, int D2, int Cout, void* stream) {
  return conv3d_fwd_tc_k2n_impl(dy, C, 0, C, wp, nullptr, dx, B, D0, D1, D2, Cout, 0, 0, 1, stream, 1, h, dbias, nullptr);
}
// Forward of a Cin, Cout <= 32 layer that also accumulates the BatchNorm statistics of its output (replaces the
// reduction pass of ssr_bn_stats): sums[0..Cout) += sum_v y, sums[Cout..2 Cout) += sum_v y^2 (zeroed here; finish with
// ssr_bn_finalize).
int ssr_conv3d_fwd_tc_k2n_stats(const float* x, int C, const float* wp, const float* bias, float* y, double* sums, int B,
                                int D0, int D1, int D2, int Cout, int act, void* stream) {
  SSR_CHECK_ARG(sums, "sums");
  SSR_CHECK_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)Cout * sizeof(double), (cudaStream_t)stream));
  return conv3d_fwd_tc_k2n_impl(x, C, 0, C, wp, bias, y, B, D0, D1, D2, Cout, act, 0, 1, stream, 2, nullptr, nullptr, sums);
}

// final channel part of a k2n convolution that also accumulates the BatchNorm sums of the finished output (the last
// term of a compensated convolution: x * w_hi and x_lo * w_hi are already in y)
int ssr_conv3d_fwd_tc_k2n_part_stats(const float* x, int Ctot, int c0, int C, const float* wp, const float* bias, float* y,
                                     double* sums, int B, int D0, int D1, int D2, int Cout, int act, int accumulate,
                                     void* stream) {
  SSR_CHECK_ARG(sums, "sums");
  SSR_CHECK_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)Cout * sizeof(double), (cudaStream_t)stream));
  return conv3d_fwd_tc_k2n_impl(x, Ctot, c0, C, wp, bias, y, B, D0, D1, D2, Cout, act, accumulate, 1, stream, 2, nullptr,
                                nullptr, sums);
}

// second (and last) pass of a compensated k2n convolution in the hybrid scheme: x2 = [x_lo | x_hi] (2 C bf16 channels,
// ssr_tf32_split_bf16), wp = pack mode 8; adds x_lo w_hi + x_hi w_lo to the partial result x_hi w_hi already in y, then
// bias + activation (+ the BatchNorm sums of the finished output when sums != NULL)
int ssr_conv3d_fwd_tc_k2n_bf16(const void* x2, int C2, const float* wp, const float* bias, float* y, double* sums, int B,
                               int D0, int D1, int D2, int Cout, int act, void* stream) {
  if (sums) SSR_CHECK_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)Cout * sizeof(double), (cudaStream_t)stream));
  return conv3d_fwd_tc_k2n_impl(reinterpret_cast<const float*>(x2), C2, 0, C2, wp, bias, y, B, D0, D1, D2, Cout, act, 1, 1,
                                stream, sums ? 2 : 0, nullptr, nullptr, sums, 1);
}

// x2[v][0:C] = bf16(x - rne_tf32(x)), x2[v][C:2C] = bf16(rne_tf32(x))  (x: [nvox, C] fp32; x2: [nvox, 2C] bf16)
int ssr_tf32_split_bf16(const float* x, void* x2, long long nvox, int C, void* stream) {
  SSR_CHECK_ARG(x && x2 && nvox > 0 && C > 0 && C % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)x2 & 15) == 0,
                "tf32_split_bf16 args (C must be a multiple of 4)");
  long long g = (nvox * (C / 4) + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  tf32_split_bf16_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<uint16_t*>(x2), nvox, C, 0);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}
// x2[v][0:C] = x1 = bf16(x), x2[v][C:2C] = bf16(x - x1): the activation operand of the bf16x3 compensated forward (level 5)
int ssr_bf16x3_split(const float* x, void* x2, long long nvox, int C, void* stream) {
  SSR_CHECK_ARG(x && x2 && nvox > 0 && C > 0 && C % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)x2 & 15) == 0,
                "bf16x3_split args (C must be a multiple of 4)");
  long long g = (nvox * (C / 4) + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  tf32_split_bf16_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<uint16_t*>(x2), nvox, C, 1);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// lo[i] = x[i] - rne_tf32(x[i]): the part of an fp32 activation the TMA's TFLOAT32 load rounds away (round to nearest
// even on the 13 dropped mantissa bits, profiles/r01_tma_tfloat32_rounding.txt); exact in fp32
int ssr_tf32_residual(const float* x, float* lo, long long n, void* stream) {
  SSR_CHECK_ARG(x && lo && n > 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)lo & 15) == 0, "tf32_residual args");
  long long g = (n / 4 + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (g < 1) g = 1;
  tf32_residual_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(x, lo, n);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

long long ssr_conv3d_wgrad_scratch_bytes(int, int, int, int, int, int, int) { return 0; }

// dw (3,3,3,C1+C2,Cout) += [x1,x2] (*) dy  ;  db[Cout] += sum_v dy   (tcgen05, see wgrad_tc_kernel)
// cin_total / cin_off: dw is (27, cin_total, Cout) and this launch covers its input channels [cin_off, cin_off + C1 + C2)
// (cin_total <= 0: dw is exactly (27, C1 + C2, Cout)).  parity: dy is the FULL-resolution gradient [B,2D0,2D1,2D2,Cout],
// x1 the low-resolution tensor, dw receives the 8 effective-kernel gradients (8, 27, C1, Cout) (see conv3d_tc_up_kernel).
static int wgrad_tc_impl(const float* x1, int C1, const float* x2, int C2, const float* dy, float* dw, float* db,
                         int B, int D0, int D1, int D2, int Cout, void* stream, int cin_total, int cin_off, int parity) {
  SSR_CHECK_ARG(x1 && dy && dw && B > 0 && D0 > 0 && D1 > 0 && D2 > 0 && Cout > 0, "pointers/shape");
  SSR_CHECK_ARG(C1 > 0 && C1 % 8 == 0 && C2 >= 0 && C2 % 8 == 0 && (C2 == 0 || x2) && Cout % 8 == 0,
                "channel counts must be multiples of 8");
  WgGeom G;
  memset(&G, 0, sizeof(G));
  G.B = B; G.D0 = D0; G.D1 = D1; G.D2 = D2; G.Cin = cin_total > 0 ? cin_total : C1 + C2; G.Cout = Cout;
  G.cin_off = cin_total > 0 ? cin_off : 0;
  SSR_CHECK_ARG(G.cin_off >= 0 && G.cin_off + C1 + C2 <= G.Cin, "channel range");
  const int Npad = round_up(Cout, 32);
  int ntile = Npad;
  if (ntile > 96) { ntile = 96; while (Npad % ntile) ntile -= 32; }
  G.NT = ntile; G.nNtiles = Npad / ntile;
  G.KG = G.NT <= 64 ? 3 : 1;
  G.SBT = G.KG + 1; if (G.KG == 1) G.SBT = 3;
  const int bstage = (G.NT / 32) * WG_BTILE_BYTES;
  int sa = (227 * 1024 - 1024 - 512 - G.SBT * bstage) / SLAB_BYTES; if (sa > 4) sa = 4;
  SSR_CHECK_ARG(sa >= 2, "shared memory budget");
  G.SA = sa;
  int cols = G.KG * G.NT, pc = 32; while (pc < cols) pc <<= 1;
  G.tmem_cols = pc;
  int nch = 0;
  for (int c = 0; c < C1; c += 32) { SSR_CHECK_ARG(nch < MAX_CHUNKS, "too many channels"); G.chunk_src[nch] = 0; G.chunk_c0[nch] = (short)c; G.chunk_valid[nch] = (unsigned char)(C1 - c < 32 ? C1 - c : 32); ++nch; }
  for (int c = 0; c < C2; c += 32) { SSR_CHECK_ARG(nch < MAX_CHUNKS, "too many channels"); G.chunk_src[nch] = 1; G.chunk_c0[nch] = (short)c; G.chunk_valid[nch] = (unsigned char)(C2 - c < 32 ? C2 - c : 32); ++nch; }
  G.nchunks = nch; G.C1 = C1;
  G.n2tiles = (D2 + TM2 - 1) / TM2; G.n1tiles = (D1 + TM1 - 1) / TM1;
  const bool k2n = !getenv("SSR_WGRAD_NO_K2N");    // d2 taps in the MMA N dimension, 32 output channels per CTA
  if (k2n) G.nNtiles = Npad / 32;
  const long long base_units = k2n ? (long long)nch * G.nNtiles * G.n1tiles * G.n2tiles * B
                                   : (long long)nch * 3 * (3 / G.KG) * G.nNtiles * G.n1tiles * G.n2tiles * B;
  // enough CTAs for >= ~8 waves of 148 (tail-wave loss < ~6 %), but at least 8 planes per CTA (halo planes are re-read)
  int S = (int)((8 * 148 + base_units - 1) / base_units); if (S < 1) S = 1; if (S > (D0 + 7) / 8) S = (D0 + 7) / 8; if (S < 1) S = 1;
  G.zlen = (D0 + S - 1) / S; G.n0splits = (D0 + G.zlen - 1) / G.zlen;
  const size_t smem = k2n ? 1024 + (size_t)WK_SA * SLAB_BYTES + (size_t)WK_SB * WK_BSTAGE + 512
                          : 1024 + (size_t)G.SA * SLAB_BYTES + (size_t)G.SBT * bstage + 512;
  CUtensorMap m1, m2, my;
  const CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;     // MN-major tf32 operands (UMMA 128B_BASE32B)
  int rc = make_map_act(&m1, x1, C1, B, D0, D1, D2, TM1 + 2, swz); if (rc) return rc;
  if (C2 > 0) { rc = make_map_act(&m2, x2, C2, B, D0, D1, D2, TM1 + 2, swz); if (rc) return rc; } else m2 = m1;
  UpMaps mys;
  memset(&mys, 0, sizeof(mys));
  if (parity) {
    SSR_CHECK_ARG(k2n && !getenv("SSR_WGRAD_NO_PERSISTENT") && C2 == 0, "the parity weight gradient needs the persistent kernel");
    const long long F0 = 2LL * D0, F1 = 2LL * D1, F2 = 2LL * D2;
    for (int par = 0; par < 8; ++par) {
      const long long off = ((((par >> 2) & 1) * F1 + ((par >> 1) & 1)) * F2 + (par & 1)) * Cout;
      rc = make_map_view(&mys.x[par], dy + off, Cout, B, D0, D1, D2, 2LL * Cout, 2 * F2 * Cout, 2 * F1 * F2 * Cout,
                         F0 * F1 * F2 * Cout, TM1, TM2 + 2, swz);
      if (rc) return rc;
    }
    my = mys.x[0];
  } else {
    rc = make_map_act(&my, dy, Cout, B, D0, D1, D2, TM1, swz, k2n ? TM2 + 2 : TM2); if (rc) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    SSR_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SSR_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_k2n_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const long long nblk = base_units * G.n0splits;
  SSR_CHECK_ARG(nblk < (1LL << 31), "grid too large");
  cudaStream_t st = (cudaStream_t)stream;
  if (k2n && !getenv("SSR_WGRAD_NO_PERSISTENT")) {
    static int num_sms = 0;
    if (!num_sms) {
      int dev = 0;
      SSR_CHECK_CUDA(cudaGetDevice(&dev));
      SSR_CHECK_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
      SSR_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_persistent_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      SSR_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_persistent_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    // plane steps: units x tiles x planes, cut into equal contiguous ranges (at least ~4 planes per CTA)
    const long long T = (long long)nch * G.nNtiles * G.n1tiles * G.n2tiles * B * D0 * (parity ? 8 : 1);
    // one launch; SSR_WGRAD_SLICES > 1 cuts it into several shorter launches (measured slower: 14.2 / 14.5 / 15.1 ms per
    // step for 1 / 2 / 4 slices -- the extra flushes cost more than the earlier SM hand-over to the dgrad chain gains)
    int nslices = 1;
    if (const char* e = getenv("SSR_WGRAD_SLICES")) nslices = atoi(e);
    if (nslices < 1) nslices = 1;
    if (nslices > 8) nslices = 8;
    G.n0splits = nslices;
    for (int sl = 0; sl < nslices; ++sl) {
      G.zlen = sl;
      const long long Ts = T * (sl + 1) / nslices - T * sl / nslices;
      long long grid = num_sms;
      if (Ts / 4 < grid) grid = Ts / 4 > 0 ? Ts / 4 : 1;
      if (parity) wgrad_tc_persistent_kernel<true><<<(unsigned)grid, WK_THREADS, smem, st>>>(m1, m2, my, mys, dw, G);
      else wgrad_tc_persistent_kernel<false><<<(unsigned)grid, WK_THREADS, smem, st>>>(m1, m2, my, mys, dw, G);
      if (sl + 1 < nslices) SSR_COUNT_LAUNCH();
    }
  } else if (k2n) wgrad_tc_k2n_kernel<<<(unsigned)nblk, WK_THREADS, smem, st>>>(m1, m2, my, dw, G);
  else wgrad_tc_kernel<<<(unsigned)nblk, 192, smem, st>>>(m1, m2, my, dw, G);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  if (db) {
    const long long nvox = (long long)B * D0 * D1 * D2 * (parity ? 8 : 1);
    int rc2 = ssr_channel_sum(dy, nvox, Cout, db, stream);
    if (rc2) return rc2;
  }
  return SSR_OK;
}

int ssr_conv3d_wgrad_tc(const float* x1, int C1, const float* x2, int C2, const float* dy, float* dw, float* db,
                        float* scratch, long long scratch_bytes, int B, int D0, int D1, int D2, int Cout,
                        void* stream) {
  (void)scratch; (void)scratch_bytes;
  return wgrad_tc_impl(x1, C1, x2, C2, dy, dw, db, B, D0, D1, D2, Cout, stream, 0, 0, 0);
}
// weight gradient w.r.t. the input channels [cin_off, cin_off + C) of a kernel dw (27, cin_total, Cout)
int ssr_conv3d_wgrad_tc_part(const float* x, int C, const float* dy, float* dw, int cin_total, int cin_off, int B, int D0,
                             int D1, int D2, int Cout, void* stream) {
  return wgrad_tc_impl(x, C, nullptr, 0, dy, dw, nullptr, B, D0, D1, D2, Cout, stream, cin_total, cin_off, 0);
}
// Upsampled part of a decoder convolution: low [B,d0,d1,d2,Cup] (the tensor that is upsampled), dy [B,2d0,2d1,2d2,Cout];
// dw (27, cin_total, Cout) += gradient w.r.t. the input channels [cin_off, cin_off + Cup).  scratch: 8*27*Cup*Cout floats
// (zeroed here) for the gradients of the 8 effective kernels.
int ssr_conv3d_wgrad_tc_up(const float* low, int Cup, const float* dy, float* dw, int cin_total, int cin_off, float* scratch,
                           int B, int d0, int d1, int d2, int Cout, void* stream) {
  SSR_CHECK_ARG(scratch && dw && cin_off >= 0 && cin_off + Cup <= cin_total, "args");
  cudaStream_t st = (cudaStream_t)stream;
  SSR_CHECK_CUDA(cudaMemsetAsync(scratch, 0, 8ULL * 27 * Cup * Cout * sizeof(float), st));
  int rc = wgrad_tc_impl(low, Cup, nullptr, 0, dy, scratch, nullptr, B, d0, d1, d2, Cout, stream, 0, 0, 1);
  if (rc) return rc;
  const long long n = 27LL * Cup * Cout;
  long long g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  up_wgrad_combine_kernel<<<(unsigned)g, 256, 0, st>>>(scratch, dw, cin_total, cin_off, Cup, Cout);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// out: `nblocks` floats (cycles per MMA measured by each CTA, one CTA per SM)
int ssr_tc_microbench(float* out, int nblocks, int N, int nacc, int chain, int iters, int kmajor, int commit_every,
                      int cycle_addr, void* stream) {
  SSR_CHECK_ARG(out && nblocks > 0 && N % 16 == 0 && N >= 16 && N <= 256 && nacc >= 1 && nacc * N <= 512 && chain >= 1 &&
                iters > 0, "microbench args");
  SSR_CHECK_CUDA(cudaFuncSetAttribute(mma_microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  SSR_CHECK_ARG(!cycle_addr || N <= 32 || kmajor >= 2, "address cycling needs 9*N*128 B of B tiles");
  mma_microbench_kernel<<<nblocks, 128, 200 * 1024, (cudaStream_t)stream>>>(out, N, nacc, chain, iters, kmajor, commit_every,
                                                                              cycle_addr);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

// profiling aid: phase timestamps of sampled CTAs of conv3d_tc_kernel are written to `buf` (>= 128*16 int64), NULL disables
// see desc_probe_kernel.  swz: 0 = SWIZZLE_128B, 1 = SWIZZLE_128B_ATOM_32B.  p = 16 ints: N, mn_major, nk, layout,
// a_off, a_lbo, a_sbo, a_bo, a_step, b_off, b_lbo, b_sbo, b_bo, b_step (HOST array).
int ssr_tc_desc_probe(const float* a, int rows_a, const float* b, int rows_b, float* d, int swz, const int* p,
                      void* stream) {
  SSR_CHECK_ARG(a && b && d && p && rows_a > 0 && rows_a <= 256 && rows_b > 0 && rows_b <= 256, "probe arguments");
  EncodeTiledFn enc = get_encode();
  if (!enc) { ssr_set_error("cuTensorMapEncodeTiled not available"); return SSR_ERR_CUDA; }
  CUtensorMap ma, mb;
  const CUtensorMapSwizzle sw = swz ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  for (int i = 0; i < 2; ++i) {
    cuuint64_t dims[2] = {32, (cuuint64_t)(i ? rows_b : rows_a)};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {32, (cuuint32_t)(i ? rows_b : rows_a)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(i ? &mb : &ma, tma_dtype(), 2, (void*)(i ? b : a), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ssr_set_error("cuTensorMapEncodeTiled(probe) failed: %d", (int)r); return SSR_ERR_CUDA; }
  }
  ProbeArgs P;
  P.rows_a = rows_a; P.rows_b = rows_b; P.N = p[0]; P.mn_major = p[1]; P.nk = p[2]; P.layout = (uint32_t)p[3];
  P.a_off = p[4]; P.a_lbo = p[5]; P.a_sbo = p[6]; P.a_bo = p[7]; P.a_step = p[8];
  P.b_off = p[9]; P.b_lbo = p[10]; P.b_sbo = p[11]; P.b_bo = p[12]; P.b_step = p[13];
  SSR_CHECK_ARG(P.N >= 16 && P.N <= 256 && P.N % 16 == 0 && P.nk >= 1, "probe N/nk");
  const size_t smem = 1024 + 2 * 256 * 128 + 64;
  SSR_CHECK_CUDA(cudaFuncSetAttribute(desc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  desc_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(ma, mb, d, P);
  SSR_COUNT_LAUNCH();
  SSR_CHECK_LAUNCH();
  return SSR_OK;
}

int ssr_tc_set_debug(long long* buf) {
  SSR_CHECK_CUDA(cudaMemcpyToSymbol(g_dbg, &buf, sizeof(buf)));
  return SSR_OK;
}

int ssr_tc_selftest(void* stream) {
  (void)stream;
  int arch = 0, dev = 0;
  SSR_CHECK_CUDA(cudaGetDevice(&dev));
  SSR_CHECK_CUDA(cudaDeviceGetAttribute(&arch, cudaDevAttrComputeCapabilityMajor, dev));
  if (arch != 10) { ssr_set_error("tcgen05 path needs compute capability 10.x, found %d.x", arch); return SSR_ERR_UNSUPPORTED; }
  if (!get_encode()) { ssr_set_error("cuTensorMapEncodeTiled unavailable"); return SSR_ERR_CUDA; }
  return SSR_OK;
}

}  // extern "C"
