// Dense projections on tcgen05 + TMEM, TMA-fed (the preferred path; gemm_tc.cu is the fallback for operands
// that TMA cannot address: unaligned base pointers or leading dimensions).
//
//   C[M,N] = op(A) . op(B)   NT / NN / TN as in gemm_tc.cu, fp32-accurate through the TF32 hi/lo split.
//
// Warp roles of the persistent CTA (one per SM, 14 warps):
//   warps 0-3   epilogue  : tcgen05.ld -> padded smem -> row-contiguous 128 B stores (+bias / +C); straight-line code
//                           for interior tiles
//   warp  4     MMA       : warp-uniform loop, tcgen05.mma.kind::tf32 (lo.hi, hi.lo, hi.hi per k-step) issued under
//                           elect.sync so descriptors / TMEM addresses stay in uniform registers; the A operand is
//                           read from TENSOR MEMORY (lane = row), B through a shared-memory descriptor
//   warp  5     TMA       : one thread issues cp.async.bulk.tensor loads of the raw fp32 A/B tiles straight
//                           into their final positions (B: 128B-swizzled UMMA tiles -- K-major one 128x32 box
//                           with SWIZZLE_128B, MN-major four 32x32 boxes with SWIZZLE_128B_ATOM_32B; raw A: the same
//                           K-major box, or four unswizzled 32x32 boxes when MN-major), completing
//                           on an mbarrier with expect_tx; it runs up to STAGES k-blocks ahead, so the HBM
//                           stream is never exposed to thread-level latency and costs no registers
//   warps 6-13  split     : A: raw smem tile -> registers -> hi = rn_tf32(x), lo = rn_tf32(x - hi) -> tcgen05.st into the
//                           stage's TMEM columns (a thread owns one row; the two warps of a TMEM lane quarter take
//                           half of the k-columns each), which removes A_hi/A_lo from shared memory altogether
//                           (a 128x128 tf32 MMA fed from smem alone needs ~120 B/clk of the 128 B/clk smem port);
//                           B (only when it is not a pre-split weight): smem -> smem in place + twin tile,
//                           fence.proxy.async; then arrive on the stage's "full" barrier
// Out-of-range rows / columns / K tail are zero-filled by TMA itself, so there is no edge-tile code path.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>

namespace {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int STAGES = 4;
constexpr int ACC_STAGES = 2;
constexpr int TILE_BYTES = BM * BK * 4;                 // 16 KB
constexpr int STAGE_BYTES = 3 * TILE_BYTES;             // raw A, B_hi, B_lo
constexpr int EPI_WARPS = 4, SPLIT_WARPS = 8;
constexpr int MMA_WARP = EPI_WARPS, TMA_WARP = EPI_WARPS + 1, SPLIT_WARP0 = EPI_WARPS + 2;
constexpr int THREADS = (EPI_WARPS + 2 + SPLIT_WARPS) * 32;   // 448
constexpr int SPLIT_THREADS = SPLIT_WARPS * 32;
constexpr int A_TMEM_COL0 = ACC_STAGES * BN;            // A operand stages live behind the accumulators:
constexpr int A_TMEM_STAGE_COLS = 2 * BK;               //   per stage 32 columns of A_hi, 32 of A_lo (lane = row)
constexpr int TMEM_COLS = 512;                          // 256 (accumulators) + STAGES * 64, rounded to a power of two
static_assert(A_TMEM_COL0 + STAGES * A_TMEM_STAGE_COLS <= TMEM_COLS, "TMEM budget");
constexpr int EPI_PITCH = 36;
constexpr int EPI_STAGE_BYTES = 32 * EPI_PITCH * 4;
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + (size_t)EPI_WARPS * EPI_STAGE_BYTES + 256;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n }"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
// same with a suspend-time hint: the hardware parks the polling thread until the phase completes (or the hint expires), so a
// waiting role does not take issue slots from the epilogue / split warps of its scheduler
__device__ __forceinline__ void mbar_wait_t(uint32_t ticks, uint64_t* bar, uint32_t parity) {
  if (ticks == 0) { mbar_wait(bar, parity); return; }
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n }"
                 : "=r"(ok) : "r"(addr), "r"(parity), "r"(ticks) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// one lane of a converged warp; ptxas then knows the guarded region runs on a single thread and feeds the
// uniform-register operands of UTCHMMA without a per-instruction waterfall loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n }" : "=r"(pred));
  return pred != 0;
}

// A operand read from tensor memory (lane = row of the 128-row tile, one tf32 per 32-bit column)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n }"
               ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64): SWIZZLE_128B = 2 (K-major tiles),
// SWIZZLE_128B_BASE32B = 1 (the only layout the hardware offers for MN-major tf32 operands).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// Same result as cvt.rna.tf32.f32 (nearest, ties away from zero) with two full-rate integer ops.
__device__ __forceinline__ float tf32_rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 2-D tiled TMA load: box at element coordinates (c0 = contiguous dim, c1 = row) -> swizzled smem tile
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1) : "memory");
}

constexpr int TRACE_N = 512;
#define TRACE(slot, idx) do { if (trace_buf != nullptr && blockIdx.x == 0 && (idx) < TRACE_N) trace_buf[(slot) * TRACE_N + (idx)] = clock64(); } while (0)

// one 32-column chunk of the accumulator: lane = row, r[j] = column j of the chunk (asynchronous: tcgen05.wait::ld before use)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

struct Params {
  int64_t M, N, K;
  const float* bias;
  float* C; int64_t ldc;
  float* partial;
  int accumulate;
  int a_mn_major, b_mn_major;
  int tiles_m, tiles_n, splits;
  int64_t k_per_split;
  int c_vec;
  int passes;              // 4: lo.lo + lo.hi + hi.lo + hi.hi, 3: without lo.lo, 1: plain TF32
  int b_presplit;          // B arrives already split (tmB = hi image, tmB2 = lo image): only A is split in-kernel
  uint32_t idesc;
  float* colsum_part;       // TN only: [splits][2][M] partial column sums of A (the bias gradient), or null
  int cta_acc;              // split-K: every CTA accumulates its splits into ITS OWN partial tile (read-modify-write, fp32
                            // round-to-nearest) instead of one partial tile per split
  int fused;                // the epilogue is epi (a tagan_epilogue mode), not the plain store through C
  long long* trace;         // debug: CTA 0 writes clock64() stamps of its first TRACE_N k-blocks / tiles ([8][TRACE_N]) or null
  int prefetch;             // the TMA producer asks L2 for the streamed operand tiles this many k-blocks ahead of the smem ring
                            // (the ring holds 4 x 16 KB per operand per SM: too few bytes in flight to cover HBM latency)
  int st256;                // interior plain stores as 256-bit STG (C and bias 32-byte aligned, ldc % 8 == 0)
  uint32_t ticks;           // suspend-time hint of every mbarrier wait (0 = plain try_wait polling)
  int epi_pipe;             // epilogue: issue the TMEM load of chunk cc+1 before the read-back / stores of chunk cc
  int early_release;        // resident mode: the A smem slot is released by the split warps, the A TMEM slot by the MMA commit
  int b_resident;           // the pre-split B panel of this CTA (K <= STAGES*BK) stays in the B slots of the stages for the
                            // whole kernel: k-block kb in stage slot kb, loaded once; only A streams
  int64_t K1;               // > 0: A is the column concatenation [A (k < K1) | A2 (k >= K1)] (tmA / tmA2), NT only
  tagan_epilogue epi;       // fused epilogue (mode 0 = plain store / bias / accumulate through C)
};


// ---------------------------------------------------------------------------------------------------------
// Fused epilogues (struct tagan_epilogue, include/tagan_b200.h).  An epilogue warp owns 32 rows of the tile; one
// 32-column chunk of the accumulator goes TMEM -> registers (lane = row) -> padded smem -> registers in the
// row-contiguous layout: lane holds rows 4*i + (lane >> 3), i = 0..7, columns 4*(lane & 7)..+3 of the chunk, so
// every global access below is a 128-byte row segment per 8 lanes.
// ---------------------------------------------------------------------------------------------------------
// FAST: MUFU-based exp / reciprocal (relative error ~1e-6, far inside the 1e-5 parity band): the four epilogue warps
// are the only threads that see the accumulator, so their instruction count bounds the whole GEMM
template <bool FAST> __device__ __forceinline__ float sigmoid_e(float x) {
  return FAST ? __fdividef(1.f, 1.f + __expf(-x)) : 1.f / (1.f + expf(-x));
}
template <bool FAST> __device__ __forceinline__ float tanh_e(float x) {
  if (!FAST) return tanhf(x);
  const float a = fminf(fabsf(x), 15.f);                 // tanh(|x|) = 1 - 2 / (1 + e^{2|x|}), odd extension
  const float t = 1.f - __fdividef(2.f, 1.f + __expf(2.f * a));
  return copysignf(t, x);
}
__device__ __forceinline__ float4 ld4g(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4g(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ void epi_load_chunk(uint32_t taddr, uint32_t stg, int lane, float4 (&v)[8]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  __syncwarp();                                              // the previous chunk's reads of the staging tile are done
#pragma unroll
  for (int j = 0; j < 32; j += 4)
    sts128(stg + (uint32_t)((lane * EPI_PITCH + j) * 4),
           make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
  __syncwarp();
  const uint32_t src0 = stg + (uint32_t)(((lane >> 3) * EPI_PITCH + (lane & 7) * 4) * 4);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = lds128(src0 + (uint32_t)(i * 4 * EPI_PITCH * 4));
}

// One 128x128 tile through a fused epilogue; arrives on the accumulator's "empty" barrier as soon as the last
// tcgen05.ld of the tile has completed (before the second LayerNorm phase), so the MMA warp never waits on the
// normalisation pass.
template <bool FAST>
__device__ __forceinline__ void epilogue_fused(const Params& p, int warp, int lane, int acc, int mt, int64_t n0,
                                               uint32_t epi_u32, uint64_t* tmem_empty_bar) {
  const tagan_epilogue& e = p.epi;
  const uint32_t stg = epi_u32 + (uint32_t)(warp * EPI_STAGE_BYTES);
  const int64_t row0 = (int64_t)mt * BM + warp * 32 + (lane >> 3);
  const int cl = (lane & 7) * 4;
  const uint32_t tbase = ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BN);
  if (e.mode == TAGAN_EPI_STORE_BF16) {
    // acc + bias rounded to bf16 (RNE): the fused QKV projection of the bf16-storage mode; a lane's 4 columns are one 8-byte store
    uint16_t* o16 = reinterpret_cast<uint16_t*>(e.out0);
#pragma unroll 1
    for (int cc = 0; cc < BN / 32; ++cc) {
      float4 v[8];
      epi_load_chunk(tbase + (uint32_t)(cc * 32), stg, lane, v);
      const int64_t c0 = n0 + cc * 32 + cl;
      const bool cok = c0 < p.N;
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cok && p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t grow = row0 + 4 * i;
        if (!(cok && grow < p.M)) continue;
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v[i].x + b4.x, v[i].y + b4.y);
        const __nv_bfloat162 hi = __floats2bfloat162_rn(v[i].z + b4.z, v[i].w + b4.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo);
        pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(o16 + grow * e.ld_out0 + c0) = pk;
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tmem_empty_bar);
    return;
  }
  if (e.mode == TAGAN_EPI_RES_LN || e.mode == TAGAN_EPI_STORE) {
    const bool ln = e.mode == TAGAN_EPI_RES_LN && e.gamma != nullptr;
    float* tmp = e.out1 ? e.out1 : e.out0;
    const int64_t ldt = e.out1 ? e.ld_out1 : e.ld_out0;
    float sd[8], sq[8], ks[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sd[i] = 0.f; sq[i] = 0.f; ks[i] = 0.f; }
#pragma unroll 1
    for (int cc = 0; cc < BN / 32; ++cc) {
      float4 v[8];
      epi_load_chunk(tbase + (uint32_t)(cc * 32), stg, lane, v);
      const int64_t c0 = n0 + cc * 32 + cl;
      const bool cok = c0 < p.N;
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cok && p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float4 rr[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int64_t grow = row0 + 4 * (half * 4 + q);
          rr[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e.in0 != nullptr && cok && grow < p.M) rr[q] = ld4g(e.in0 + grow * e.ld_in0 + c0);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = half * 4 + q;
          const int64_t grow = row0 + 4 * i;
          const bool ok = cok && grow < p.M;
          v[i].x += b4.x + rr[q].x; v[i].y += b4.y + rr[q].y; v[i].z += b4.z + rr[q].z; v[i].w += b4.w + rr[q].w;
          if (ok) st4g(tmp + grow * ldt + c0, v[i]);
          if (ln) {
            // shifted single-pass moments: the shift is the row's first element, so the subtraction below cancels the
            // row mean up to O(std) and var = E[d^2] - E[d]^2 loses no precision
            if (cc == 0) ks[i] = __shfl_sync(FULL_MASK, v[i].x, lane & ~7);
            if (ok) {
              const float dx = v[i].x - ks[i], dy = v[i].y - ks[i], dz = v[i].z - ks[i], dw = v[i].w - ks[i];
              sd[i] += (dx + dy) + (dz + dw);
              sq[i] = fmaf(dx, dx, sq[i]); sq[i] = fmaf(dy, dy, sq[i]); sq[i] = fmaf(dz, dz, sq[i]); sq[i] = fmaf(dw, dw, sq[i]);
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tmem_empty_bar);
    if (!ln) return;
    const float inv_n = 1.f / (float)p.N;
    float (&mean)[8] = sd;
    float (&rstd)[8] = sq;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        sd[i] += __shfl_xor_sync(FULL_MASK, sd[i], o);
        sq[i] += __shfl_xor_sync(FULL_MASK, sq[i], o);
      }
      const float md = sd[i] * inv_n;
      rstd[i] = 1.f / sqrtf(fmaxf(sq[i] * inv_n - md * md, 0.f) + 1e-5f);
      mean[i] = ks[i] + md;
    }
#pragma unroll 1
    for (int cc = 0; cc < BN / 32; ++cc) {
      const int64_t c0 = n0 + cc * 32 + cl;
      if (c0 >= p.N) continue;
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(e.gamma + c0));
      const float4 be4 = __ldg(reinterpret_cast<const float4*>(e.beta + c0));
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t grow = row0 + 4 * i;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grow < p.M) v[i] = __ldcg(reinterpret_cast<const float4*>(tmp + grow * ldt + c0));   // written above by this thread
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t grow = row0 + 4 * i;
        if (grow >= p.M) continue;
        float4 y;
        y.x = (v[i].x - mean[i]) * rstd[i] * g4.x + be4.x;
        y.y = (v[i].y - mean[i]) * rstd[i] * g4.y + be4.y;
        y.z = (v[i].z - mean[i]) * rstd[i] * g4.z + be4.z;
        y.w = (v[i].w - mean[i]) * rstd[i] * g4.w + be4.w;
        st4g(e.out0 + grow * e.ld_out0 + c0, y);
      }
    }
    if ((lane & 7) == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t grow = row0 + 4 * i;
        if (grow < p.M) {
          if (e.mean) e.mean[grow] = mean[i];
          if (e.rstd) e.rstd[grow] = rstd[i];
        }
      }
    }
    return;
  }
  // ---- element-wise modes ----
#pragma unroll 1
  for (int cc = 0; cc < BN / 32; ++cc) {
    float4 v[8];
    epi_load_chunk(tbase + (uint32_t)(cc * 32), stg, lane, v);
    const int64_t c0 = n0 + cc * 32 + cl;
    const bool cok = c0 < p.N;
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cok && p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float4 a0[4], a1[4];
      if (e.mode == TAGAN_EPI_GATES) {
        const bool rpart = c0 < e.split;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int64_t grow = row0 + 4 * (half * 4 + q);
          a0[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (rpart && cok && grow < p.M) a0[q] = ld4g(e.in0 + grow * e.ld_in0 + c0);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = half * 4 + q;
          const int64_t grow = row0 + 4 * i;
          if (!(cok && grow < p.M)) continue;
          float4 s;
          s.x = sigmoid_e<FAST>(v[i].x + b4.x); s.y = sigmoid_e<FAST>(v[i].y + b4.y);
          s.z = sigmoid_e<FAST>(v[i].z + b4.z); s.w = sigmoid_e<FAST>(v[i].w + b4.w);
          if (rpart) {
            st4g(e.out0 + grow * e.ld_out0 + c0, s);
            st4g(e.out1 + grow * e.ld_out1 + c0, make_float4(s.x * a0[q].x, s.y * a0[q].y, s.z * a0[q].z, s.w * a0[q].w));
          } else {
            st4g(e.out2 + grow * e.ld_out2 + (c0 - e.split), s);
          }
        }
      } else if (e.mode == TAGAN_EPI_BLEND) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int64_t grow = row0 + 4 * (half * 4 + q);
          a0[q] = a1[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cok && grow < p.M) { a0[q] = ld4g(e.in0 + grow * e.ld_in0 + c0); a1[q] = ld4g(e.in1 + grow * e.ld_in1 + c0); }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = half * 4 + q;
          const int64_t grow = row0 + 4 * i;
          if (!(cok && grow < p.M)) continue;
          float4 t, o;
          t.x = tanh_e<FAST>(v[i].x + b4.x); t.y = tanh_e<FAST>(v[i].y + b4.y);
          t.z = tanh_e<FAST>(v[i].z + b4.z); t.w = tanh_e<FAST>(v[i].w + b4.w);
          o.x = (1.f - a0[q].x) * a1[q].x + a0[q].x * t.x;
          o.y = (1.f - a0[q].y) * a1[q].y + a0[q].y * t.y;
          o.z = (1.f - a0[q].z) * a1[q].z + a0[q].z * t.z;
          o.w = (1.f - a0[q].w) * a1[q].w + a0[q].w * t.w;
          st4g(e.out0 + grow * e.ld_out0 + c0, t);
          st4g(e.out1 + grow * e.ld_out1 + c0, o);
        }
      } else {                                               // TAGAN_EPI_GATES_BWD
        float4 a2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int64_t grow = row0 + 4 * (half * 4 + q);
          a0[q] = a1[q] = a2[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cok && grow < p.M) {
            a0[q] = ld4g(e.in0 + grow * e.ld_in0 + c0);
            a1[q] = ld4g(e.in1 + grow * e.ld_in1 + c0);
            a2[q] = ld4g(e.out1 + grow * e.ld_out1 + c0);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = half * 4 + q;
          const int64_t grow = row0 + 4 * i;
          if (!(cok && grow < p.M)) continue;
          const float4 d = v[i], r = a0[q], sc = a1[q];
          st4g(e.out0 + grow * e.ld_out0 + c0, make_float4(d.x * sc.x * r.x * (1.f - r.x), d.y * sc.y * r.y * (1.f - r.y),
                                                          d.z * sc.z * r.z * (1.f - r.z), d.w * sc.w * r.w * (1.f - r.w)));
          st4g(e.out1 + grow * e.ld_out1 + c0, make_float4(a2[q].x + d.x * r.x, a2[q].y + d.y * r.y,
                                                          a2[q].z + d.z * r.z, a2[q].w + d.w * r.w));
        }
      }
    }
  }
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(tmem_empty_bar);
}

// FUSED = false leaves the fused epilogues (struct tagan_epilogue modes) out of the instantiation: the plain projections are
// sensitive to the register allocation of the epilogue role, and the fused code paths are what pushes it into spills
// SIMPLE = the shape of almost every projection of a step (pre-split K-major weights, no split-K, plain store, no profiling
// knobs): the instantiation drops the MN-major / in-kernel-split / split-K / trace / prefetch code, which leaves a kernel a
// quarter of the general one's size.
// KIND 2 = the dW products (TN: both operands MN-major activations split in-kernel, split-K into CTA-private partial tiles):
// the same specialisation for the other frequent shape.  KIND 0 = everything else.
template <bool FUSED, int KIND>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tma_kernel(const Params p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmA2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* raw_bar = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES + (size_t)EPI_WARPS * EPI_STAGE_BYTES);
  uint64_t* full_bar = raw_bar + STAGES;
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);
  uint64_t* bres_bar = tmem_empty + ACC_STAGES + 1;
  uint64_t* sfree_bar = bres_bar + 1;                       // [STAGES] early release of the A smem slots (resident mode)
  const uint32_t epi_u32 = smem_u32(smem + (size_t)STAGES * STAGE_BYTES);      // 1024-byte aligned

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool SIMPLE = KIND == 1, TNK = KIND == 2;
  const bool a_mn = SIMPLE ? false : (TNK ? true : p.a_mn_major != 0);
  const bool b_mn = SIMPLE ? false : (TNK ? true : p.b_mn_major != 0);
  const bool b_pre = SIMPLE ? true : (TNK ? false : p.b_presplit != 0);
  float* const colsum_part = SIMPLE ? nullptr : p.colsum_part;
  float* const partial = SIMPLE ? nullptr : p.partial;
  const int cta_acc = SIMPLE ? 0 : (TNK ? 1 : p.cta_acc);
  const int accum_c = (SIMPLE || TNK) ? 0 : p.accumulate;
  const int pf_dist = (SIMPLE || TNK) ? 0 : p.prefetch;
  const bool early_rel = (SIMPLE || TNK) ? false : p.early_release != 0;
  const bool epi_pipe_on = (SIMPLE || TNK) ? false : p.epi_pipe != 0;
  long long* const trace_buf = (SIMPLE || TNK) ? nullptr : p.trace;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&raw_bar[s], 1); mbar_init(&full_bar[s], SPLIT_WARPS); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], EPI_WARPS); }
    mbar_init(bres_bar, 1);
    for (int s = 0; s < STAGES; ++s) mbar_init(&sfree_bar[s], SPLIT_WARPS);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == TMA_WARP && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // The CTA allocates all 512 columns, so the allocation can only start at lane 0 / column 0.  Treating the base
  // as the constant 0 keeps every TMEM address of the MMA loop in uniform registers (no R2UR per instruction).
  if (*tmem_slot != 0u) __trap();
  constexpr uint32_t tmem_base = 0u;
  const int64_t num_work = (int64_t)p.tiles_m * p.tiles_n * p.splits;

  if (warp == TMA_WARP) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (p.b_resident && (int64_t)blockIdx.x < num_work) {
        // the grid is a multiple of tiles_n, so this CTA's N panel never changes: its B_hi / B_lo k-blocks go into the B slots
        // of stage 0..nkb-1 once (weights: L2 hits) and stay
        const int n0 = (int)(blockIdx.x % p.tiles_n) * BN;
        const int nkb = (int)((p.K + BK - 1) / BK);
        mbar_arrive_expect_tx(bres_bar, (uint32_t)(nkb * 2 * TILE_BYTES));
        for (int kb = 0; kb < nkb; ++kb) {
          uint8_t* st = smem + (size_t)kb * STAGE_BYTES;
          if (!b_mn) {
            tma_load_2d(st + TILE_BYTES, &tmB, kb * BK, n0, bres_bar);
            tma_load_2d(st + 2 * TILE_BYTES, &tmB2, kb * BK, n0, bres_bar);
          } else {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              tma_load_2d(st + TILE_BYTES + b * 4096, &tmB, n0 + 32 * b, kb * BK, bres_bar);
              tma_load_2d(st + 2 * TILE_BYTES + b * 4096, &tmB2, n0 + 32 * b, kb * BK, bres_bar);
            }
          }
        }
      }
      int tr_kb = 0;
      // L2 prefetch iterator: the same (work item, k-block) sequence, pf_dist k-blocks ahead
      const bool pf_b = !b_pre;                       // pre-split weights are L2 hits anyway
      int64_t pw = blockIdx.x, pk0 = 0, pkend = 0;
      auto pf_open = [&]() {
        if (pw >= num_work) return;
        const int ks = (int)(pw / ((int64_t)p.tiles_n * p.tiles_m));
        pk0 = (int64_t)ks * p.k_per_split;
        pkend = pk0 + p.k_per_split < p.K ? pk0 + p.k_per_split : p.K;
      };
      auto pf_step = [&]() {
        while (pw < num_work && pk0 >= pkend) { pw += gridDim.x; pf_open(); }
        if (pw >= num_work) return;
        const int m0 = (int)((pw / p.tiles_n) % p.tiles_m) * BM, n0 = (int)(pw % p.tiles_n) * BN;
        if (!a_mn) {
          if (p.K1 > 0 && pk0 >= p.K1) tma_prefetch_2d(&tmA2, (int)(pk0 - p.K1), m0);
          else tma_prefetch_2d(&tmA, (int)pk0, m0);
        } else {
#pragma unroll
          for (int b = 0; b < 4; ++b) tma_prefetch_2d(&tmA, m0 + 32 * b, (int)pk0);
        }
        if (pf_b) {
          if (!b_mn) tma_prefetch_2d(&tmB, (int)pk0, n0);
          else {
#pragma unroll
            for (int b = 0; b < 4; ++b) tma_prefetch_2d(&tmB, n0 + 32 * b, (int)pk0);
          }
        }
        pk0 += BK;
      };
      if (pf_dist > 0) {
        pf_open();
        for (int i = 0; i < pf_dist; ++i) pf_step();
      }
      for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int nt = (int)(w % p.tiles_n);
        const int mt = (int)((w / p.tiles_n) % p.tiles_m);
        const int ks = (int)(w / ((int64_t)p.tiles_n * p.tiles_m));
        const int m0 = mt * BM, n0 = nt * BN;
        const int64_t kbeg = (int64_t)ks * p.k_per_split;
        const int64_t kend = kbeg + p.k_per_split < p.K ? kbeg + p.k_per_split : p.K;
        for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
          if (pf_dist > 0) pf_step();
          mbar_wait_t(p.ticks, early_rel ? &sfree_bar[stage] : &empty_bar[stage], phase ^ 1);
          TRACE(0, tr_kb); ++tr_kb;
          uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
          mbar_arrive_expect_tx(&raw_bar[stage], (p.b_resident ? 1 : b_pre ? 3 : 2) * TILE_BYTES);
          if (!a_mn) {
            if (p.K1 > 0 && k0 >= p.K1) tma_load_2d(st, &tmA2, (int)(k0 - p.K1), m0, &raw_bar[stage]);
            else tma_load_2d(st, &tmA, (int)k0, m0, &raw_bar[stage]);
          } else {
#pragma unroll
            for (int b = 0; b < 4; ++b) tma_load_2d(st + b * 4096, &tmA, m0 + 32 * b, (int)k0, &raw_bar[stage]);
          }
          if (p.b_resident) {
            // B is in place
          } else if (!b_mn) {
            tma_load_2d(st + TILE_BYTES, &tmB, (int)k0, n0, &raw_bar[stage]);
            if (b_pre) tma_load_2d(st + 2 * TILE_BYTES, &tmB2, (int)k0, n0, &raw_bar[stage]);
          } else {
#pragma unroll
            for (int b = 0; b < 4; ++b) tma_load_2d(st + TILE_BYTES + b * 4096, &tmB, n0 + 32 * b, (int)k0, &raw_bar[stage]);
            if (b_pre) {
#pragma unroll
              for (int b = 0; b < 4; ++b) tma_load_2d(st + 2 * TILE_BYTES + b * 4096, &tmB2, n0 + 32 * b, (int)k0, &raw_bar[stage]);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= SPLIT_WARP0) {
    // ============================== hi/lo split (smem -> smem) ==============================
    const int tt = threadIdx.x - SPLIT_WARP0 * 32;           // 0..255
    int stage = 0, tr_kb = 0;
    uint32_t phase = 0;
    for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x) {
      const int ks = (int)(w / ((int64_t)p.tiles_n * p.tiles_m));
      const int64_t kbeg = (int64_t)ks * p.k_per_split;
      const int64_t kend = kbeg + p.k_per_split < p.K ? kbeg + p.k_per_split : p.K;
      float csum = 0.f;                                      // this thread's row of A summed over its k-columns
      for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
        mbar_wait_t(p.ticks, &raw_bar[stage], phase);                   // TMA bytes have landed
        if (tt == 0) TRACE(1, tr_kb);
        const uint32_t st_u32 = smem_u32(smem + (size_t)stage * STAGE_BYTES);
        {
          // ---- A: raw smem tile -> registers -> hi/lo -> tensor memory.  A warp owns TMEM lanes 32*(warp%4)..+31, so
          // the two split warps of a lane quarter take 16 of the 32 k-columns each; a thread handles its own row.
          const int quarter = warp & 3, half = (warp - SPLIT_WARP0) >> 2;
          const int row = quarter * 32 + lane;
          uint32_t hi[16], lo[16];
          if (!a_mn) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int c = half * 4 + i;                                   // 16-byte chunk of the 128-byte row
              const float4 v = lds128(st_u32 + row * 128 + ((c ^ (row & 7)) << 4));
              const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float h = tf32_rn(x[q]);
                hi[i * 4 + q] = __float_as_uint(h);
                lo[i * 4 + q] = __float_as_uint(tf32_rn(x[q] - h));
              }
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {                               // boxes [32 k][32 m], unswizzled
              const float x = lds32(st_u32 + (row >> 5) * 4096 + (half * 16 + kk) * 128 + (row & 31) * 4);
              const float h = tf32_rn(x);
              hi[kk] = __float_as_uint(h);
              lo[kk] = __float_as_uint(tf32_rn(x - h));
              csum += x;
            }
          }
          if (early_rel) {
            // the raw tile is in registers: hand the smem slot back to the TMA producer now (the arrive is a release: the
            // shared-memory reads above are ordered before it), then wait for the MMAs that read this TMEM slot four
            // k-blocks ago -- the smem ring now only holds bytes in flight, the TMEM ring the operands waiting for the MMA
            __syncwarp();
            if (lane == 0) mbar_arrive(&sfree_bar[stage]);
            mbar_wait_t(p.ticks, &empty_bar[stage], phase ^ 1);
            tc_fence_after();
          }
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                                 (uint32_t)(A_TMEM_COL0 + stage * A_TMEM_STAGE_COLS + half * 16);
          tmem_st16(taddr, hi);
          tmem_st16(taddr + BK, lo);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          if (!b_pre) {
#pragma unroll 4
            for (int i = 0; i < 1024 / SPLIT_THREADS; ++i) {
              const uint32_t bh = st_u32 + TILE_BYTES + (tt + SPLIT_THREADS * i) * 16;
              const float4 v = lds128(bh);
              float4 h, l;
              h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
              l.x = tf32_rn(v.x - h.x); l.y = tf32_rn(v.y - h.y); l.z = tf32_rn(v.z - h.z); l.w = tf32_rn(v.w - h.w);
              sts128(bh, h);
              sts128(bh + TILE_BYTES, l);
            }
          }
          tc_fence_before();
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[stage]);
        if (tt == 0) TRACE(2, tr_kb);
        ++tr_kb;
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (colsum_part != nullptr) {                        // bias gradient for free: A has just been read anyway
        const int nt = (int)(w % p.tiles_n);
        const int mt = (int)((w / p.tiles_n) % p.tiles_m);
        const int64_t m = (int64_t)mt * BM + (warp & 3) * 32 + lane;
        if (nt == 0 && m < p.M) colsum_part[((int64_t)ks * 2 + ((warp - SPLIT_WARP0) >> 2)) * p.M + m] = csum;
      }
    }
  } else if (warp == MMA_WARP) {
    // ============================== MMA issuer ==============================
    // The whole warp runs the (warp-uniform) control flow so that stage indices, descriptors and TMEM addresses
    // live in uniform registers; only the tcgen05 instructions themselves are issued by lane 0.
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    // per k-step (8 tf32 = 32 bytes) descriptor advance: +32 B inside the swizzle row (K-major),
    // +1024 B = two 4-row k-atoms further (MN-major)
    const uint32_t b_step = b_mn ? 1024u : 32u;
    const uint64_t b_desc0 = make_desc(0, b_mn ? 4096u : 16u, b_mn ? 512u : 1024u, b_mn ? 1u : 2u);
    const uint32_t smem0 = smem_u32(smem);
    const uint32_t idesc = p.idesc;
    const int passes = p.passes;
    const bool bres = p.b_resident != 0;
    int tr_kb = 0, tr_tile = 0;
    if (bres && (int64_t)blockIdx.x < num_work) mbar_wait_t(p.ticks, bres_bar, 0);
    for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x) {
      const int ks = (int)(w / ((int64_t)p.tiles_n * p.tiles_m));
      const int64_t kbeg = (int64_t)ks * p.k_per_split;
      const int64_t kend = kbeg + p.k_per_split < p.K ? kbeg + p.k_per_split : p.K;
      const int nkb = kend > kbeg ? (int)((kend - kbeg + BK - 1) / BK) : 0;
      mbar_wait_t(p.ticks, &tmem_empty[acc], acc_phase ^ 1);          // epilogue has drained this accumulator stage
      tc_fence_after();
      if (lane == 0) TRACE(5, tr_tile);
      ++tr_tile;
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait_t(p.ticks, &full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) TRACE(3, tr_kb);
        const uint32_t sbase = smem0 + (uint32_t)(bres ? kb : stage) * STAGE_BYTES;     // B slot: resident k-block or the stage
        const uint32_t ta0 = tmem_base + (uint32_t)(A_TMEM_COL0 + stage * A_TMEM_STAGE_COLS);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t db_hi = b_desc0 | (uint64_t)(((sbase + TILE_BYTES + kk * b_step) >> 4) & 0x3FFF);
            const uint64_t db_lo = b_desc0 | (uint64_t)(((sbase + 2 * TILE_BYTES + kk * b_step) >> 4) & 0x3FFF);
            const uint32_t ta_hi = ta0 + kk * 8, ta_lo = ta0 + BK + kk * 8;
            const uint32_t accum = (kb | kk) != 0;
            if (passes == 3) {
              umma_tf32_ts(tmem_d, ta_lo, db_hi, idesc, accum);
              umma_tf32_ts(tmem_d, ta_hi, db_lo, idesc, 1);
              umma_tf32_ts(tmem_d, ta_hi, db_hi, idesc, 1);
            } else if (passes == 4) {
              umma_tf32_ts(tmem_d, ta_lo, db_lo, idesc, accum);   // small terms first
              umma_tf32_ts(tmem_d, ta_lo, db_hi, idesc, 1);
              umma_tf32_ts(tmem_d, ta_hi, db_lo, idesc, 1);
              umma_tf32_ts(tmem_d, ta_hi, db_hi, idesc, 1);
            } else {
              umma_tf32_ts(tmem_d, ta_hi, db_hi, idesc, accum);
            }
          }
          umma_commit(&empty_bar[stage]);                    // frees the stage when the MMAs retire
        }
        __syncwarp();
        if (lane == 0) TRACE(4, tr_kb);
        ++tr_kb;
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tmem_full[acc]);         // accumulator complete -> epilogue
      __syncwarp();
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ============================== epilogue ==============================
    int acc = 0, tr_tile = 0;
    uint32_t acc_phase = 0;
    for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x) {
      const int nt = (int)(w % p.tiles_n);
      const int mt = (int)((w / p.tiles_n) % p.tiles_m);
      const int ks = (int)(w / ((int64_t)p.tiles_n * p.tiles_m));
      const int64_t kbeg = (int64_t)ks * p.k_per_split;
      const bool has_k = kbeg < p.K;
      const int64_t n0 = (int64_t)nt * BN;
      mbar_wait_t(p.ticks, &tmem_full[acc], acc_phase);
      tc_fence_after();
      if (threadIdx.x == 0) TRACE(6, tr_tile);
      if (FUSED && p.fused) {                              // warp-uniform
        if (p.fused == 2) epilogue_fused<true>(p, warp, lane, acc, mt, n0, epi_u32, &tmem_empty[acc]);
        else epilogue_fused<false>(p, warp, lane, acc, mt, n0, epi_u32, &tmem_empty[acc]);
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      float* out = partial ? partial + (int64_t)(cta_acc ? (int)blockIdx.x : ks) * p.M * p.N : p.C;
      const int64_t ldo = partial ? p.N : p.ldc;
      const bool interior = p.c_vec && ((int64_t)(mt + 1) * BM <= p.M) && (n0 + BN <= p.N);
      // C += A.B (the GRU scan's per-step GEMMs), or this CTA's partial tile += its next split
      const bool rmw = partial == nullptr ? (accum_c != 0) : (cta_acc != 0);
      const bool bias_vec = partial == nullptr && p.bias != nullptr;       // host guarantees 16-byte alignment when c_vec
      // The TMEM load of chunk cc+1 is issued as soon as the registers of chunk cc have gone to shared memory, so its latency
      // runs under the shared-memory read-back and the global stores of chunk cc (the epilogue warps are the slowest stage of
      // the pipeline for the [T*N, 128] projections: tools/trace_gemm.py).
      uint32_t r[32];
      const uint32_t taddr0 = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BN);
      const bool pipe = epi_pipe_on;
      if (has_k && pipe) tmem_ld32(taddr0, r);
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        if (threadIdx.x == 0 && cc == 1) TRACE(8, tr_tile);
        if (has_k) {
          if (!pipe) tmem_ld32(taddr0 + (uint32_t)(cc * 32), r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        if (threadIdx.x == 0 && cc == 1) TRACE(9, tr_tile);
        // registers (lane = row, 32 consecutive columns) -> padded smem tile -> row-contiguous 128-byte stores
        const uint32_t stg = epi_u32 + (uint32_t)(warp * EPI_STAGE_BYTES);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          sts128(stg + (uint32_t)((lane * EPI_PITCH + j) * 4),
                 make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
        if (has_k && pipe && cc + 1 < BN / 32) tmem_ld32(taddr0 + (uint32_t)((cc + 1) * 32), r);
        __syncwarp();
        const int64_t c0 = n0 + cc * 32 + (lane & 7) * 4;
        const uint32_t src0 = stg + (uint32_t)(((lane >> 3) * EPI_PITCH + (lane & 7) * 4) * 4);
        if (interior && p.st256 && !rmw) {
          // 256-bit stores (sm_100 STG.256): lane = row (lane >> 2) + 8 i, columns 8 (lane & 3) .. +7 -- half as many store
          // instructions per chunk as the 128-bit form below
          const int64_t c8 = n0 + cc * 32 + (lane & 3) * 8;
          float4 b4a = make_float4(0.f, 0.f, 0.f, 0.f), b4b = b4a;
          if (bias_vec) { b4a = __ldg(reinterpret_cast<const float4*>(p.bias + c8)); b4b = __ldg(reinterpret_cast<const float4*>(p.bias + c8 + 4)); }
          const uint32_t srow = stg + (uint32_t)((((lane >> 2) * EPI_PITCH) + (lane & 3) * 8) * 4);
          float* orow = out + ((int64_t)mt * BM + warp * 32 + (lane >> 2)) * ldo + c8;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 a = lds128(srow + (uint32_t)(i * 8 * EPI_PITCH * 4));
            float4 b = lds128(srow + (uint32_t)(i * 8 * EPI_PITCH * 4 + 16));
            a.x += b4a.x; a.y += b4a.y; a.z += b4a.z; a.w += b4a.w;
            b.x += b4b.x; b.y += b4b.y; b.z += b4b.z; b.w += b4b.w;
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         ::"l"(orow + (int64_t)(i * 8) * ldo), "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)),
                           "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)), "r"(__float_as_uint(b.x)),
                           "r"(__float_as_uint(b.y)), "r"(__float_as_uint(b.z)), "r"(__float_as_uint(b.w)) : "memory");
          }
        } else if (interior) {
          // whole tile inside the matrix, 16-byte aligned rows, plain store: straight-line code
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias_vec) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
          float4 v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = lds128(src0 + (uint32_t)(i * 4 * EPI_PITCH * 4));
          if (threadIdx.x == 0 && cc == 1) TRACE(10, tr_tile);
          float* orow = out + ((int64_t)mt * BM + warp * 32 + (lane >> 3)) * ldo + c0;
          if (rmw) {
            float4 o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = *reinterpret_cast<const float4*>(orow + (int64_t)(i * 4) * ldo);
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i].x += o[i].x; v[i].y += o[i].y; v[i].z += o[i].z; v[i].w += o[i].w; }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[i].x += b4.x; v[i].y += b4.y; v[i].z += b4.z; v[i].w += b4.w;
            *reinterpret_cast<float4*>(orow + (int64_t)(i * 4) * ldo) = v[i];
          }
          if (threadIdx.x == 0 && cc == 1) TRACE(11, tr_tile);
        } else {
          const bool direct = partial == nullptr;
          const bool acc_here = direct ? (accum_c != 0) : (cta_acc != 0);
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (direct && p.bias) {
            if (c0 < p.N) b4.x = p.bias[c0];
            if (c0 + 1 < p.N) b4.y = p.bias[c0 + 1];
            if (c0 + 2 < p.N) b4.z = p.bias[c0 + 2];
            if (c0 + 3 < p.N) b4.w = p.bias[c0 + 3];
          }
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + (lane >> 3);
            const int64_t grow = (int64_t)mt * BM + warp * 32 + rr;
            if (grow >= p.M || c0 >= p.N) continue;
            float4 v = lds128(src0 + (uint32_t)(i * 4 * EPI_PITCH * 4));
            v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
            float* orow = out + grow * ldo + c0;
            if (c0 + 3 < p.N && p.c_vec) {
              if (acc_here) {
                float4 o4 = *reinterpret_cast<const float4*>(orow);
                v.x += o4.x; v.y += o4.y; v.z += o4.z; v.w += o4.w;
              }
              *reinterpret_cast<float4*>(orow) = v;
            } else {
              const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (c0 + q < p.N) orow[q] = acc_here ? orow[q] + vv[q] : vv[q];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (threadIdx.x == 0) TRACE(7, tr_tile);
      if (lane == 0) TRACE(12 + warp, tr_tile);
      ++tr_tile;
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// C (+)= sum over `parts` partial tiles (+ bias).  Up to ~1600 parts: a block takes 32 consecutive elements x 8 interleaved
// slices of the parts, four independent accumulators per thread (loads in flight), then the slices in order -- a fixed
// summation order (deterministic) with enough parallelism that the reduction streams instead of chasing latency.
__global__ void __launch_bounds__(256)
tma_splitk_reduce(const float* __restrict__ partial, int parts, int64_t M, int64_t N,
                  const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int accumulate) {
  __shared__ float red[8][32];
  const int64_t i = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  const int slice = threadIdx.x >> 5;
  const int64_t MN = M * N;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (i < MN) {
    int q = slice;
    for (; q + 24 < parts; q += 32) {
      a0 += partial[(int64_t)q * MN + i];
      a1 += partial[(int64_t)(q + 8) * MN + i];
      a2 += partial[(int64_t)(q + 16) * MN + i];
      a3 += partial[(int64_t)(q + 24) * MN + i];
    }
    for (; q < parts; q += 8) a0 += partial[(int64_t)q * MN + i];
  }
  red[slice][threadIdx.x & 31] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (slice == 0 && i < MN) {
    float s = red[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 8; ++k) s += red[k][threadIdx.x];
    const int64_t m = i / N, n = i - m * N;
    if (bias) s += bias[n];
    float* o = C + m * ldc + n;
    *o = accumulate ? *o + s : s;
  }
}

// out[m] = sum over the [parts] partial rows: 32 columns x 32 interleaved slices per block (M is only 128..768 wide, so
// the parallelism has to come from the parts: up to ~3000 of them), slices combined in order (fixed => deterministic)
__global__ void __launch_bounds__(1024)
colsum_parts_reduce(const float* __restrict__ part, int parts, int64_t M, float* __restrict__ out) {
  __shared__ float red[32][32];
  const int64_t m = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  const int slice = threadIdx.x >> 5;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (m < M) {
    int q = slice;
    for (; q + 96 < parts; q += 128) {
      a0 += part[(int64_t)q * M + m];
      a1 += part[(int64_t)(q + 32) * M + m];
      a2 += part[(int64_t)(q + 64) * M + m];
      a3 += part[(int64_t)(q + 96) * M + m];
    }
    for (; q < parts; q += 32) a0 += part[(int64_t)q * M + m];
  }
  red[slice][threadIdx.x & 31] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (slice == 0 && m < M) {
    float s = red[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 32; ++k) s += red[k][threadIdx.x];
    out[m] = s;
  }
}

struct Plan { int tiles_m, tiles_n, splits, cta_acc, parts; int64_t k_per_split; };
Plan make_plan(int64_t M, int64_t N, int64_t K) {
  Plan pl;
  pl.tiles_m = (int)((M + BM - 1) / BM);
  pl.tiles_n = (int)((N + BN - 1) / BN);
  int64_t tiles = (int64_t)pl.tiles_m * pl.tiles_n;
  int64_t splits = 1;
  if (tiles < 148 && K >= 4 * 1024) {            // few output tiles, long reduction (the dW GEMMs): split K
    splits = (296 + tiles - 1) / tiles;
    int64_t maxs = K / 1024;
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
  }
  pl.cta_acc = 0;
  if (K >= 65536) {
    // Long reductions (the dW GEMMs: K = T*N rows).  tcgen05 accumulates in TMEM with round-toward-zero, a bias that grows
    // linearly with the number of sequential k-steps (measured: 1e-4 of the gradient's magnitude at 5400 rows per split, K = 1.6M).
    // Keep every TMEM accumulation to ~1024 rows; each CTA adds its splits into its own partial tile in fp32 round-to-nearest
    // (an L2-resident read-modify-write), and the <= 148 CTA tiles are reduced in a fixed order.
    const int64_t want = K / 1024;
    if (want > splits) splits = want;
    pl.cta_acc = 1;
  }
  int64_t kps = (K + splits - 1) / splits;
  kps = (kps + BK - 1) / BK * BK;
  if (kps < BK) kps = BK;
  pl.k_per_split = kps;
  pl.splits = (int)((K + kps - 1) / kps);
  if (pl.splits < 1) pl.splits = 1;
  if (pl.splits == 1) pl.cta_acc = 0;
  const int64_t work = tiles * pl.splits;
  pl.parts = pl.cta_acc ? (int)(work < 148 ? work : 148) : pl.splits;
  return pl;
}



// weights: hi = rn_tf32(w), lo = rn_tf32(w - hi), same [rows, ld] layout, done once per call (tiny)
// hi / lo images of a weight matrix w[rows, cols]; TRANSPOSE writes them as [cols, rows] (an NN product's B = W[K,N] becomes
// the K-major [N,K] operand of the NT path: the MN-major shared-memory layout costs the MMA half its rate)
template <bool TRANSPOSE>
__global__ void presplit_kernel(const float* __restrict__ w, int64_t rows, int64_t cols, int64_t ld, float* __restrict__ hi,
                                float* __restrict__ lo, int64_t ldo) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  int64_t r, c;
  if (TRANSPOSE) { c = i / rows; r = i - c * rows; }       // consecutive threads write consecutive elements of the output
  else { r = i / cols; c = i - r * cols; }
  const float x = w[r * ld + c];
  const float h = tf32_rn(x);
  const int64_t o = TRANSPOSE ? c * ldo + r : r * ldo + c;
  hi[o] = h;
  lo[o] = tf32_rn(x - h);
}

constexpr int64_t PRESPLIT_MAX_ELEMS = 1 << 22;   // B operands up to 16 MB (weights) are pre-split

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(ptr);
  }
  return fn;
}

// operand stored [rows, cols] row-major with leading dimension ld (floats).  K-major use: cols = K, box 32 x 128,
// SWIZZLE_128B.  MN-major use: cols = M or N, rows = K, box 32 x 32, SWIZZLE_128B_ATOM_32B.
bool make_map(CUtensorMap* tm, const float* base, int64_t rows, int64_t cols, int64_t ld, bool mn_major, bool plain = false) {
  EncodeFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, mn_major ? 32u : 128u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   plain ? CU_TENSOR_MAP_SWIZZLE_NONE
                         : (mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace

bool tagan_gemm_tma_supported(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb) {
  if (K <= 0 || M >= (1LL << 31) || N >= (1LL << 31) || K >= (1LL << 31)) return false;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (lda % 4) || (ldb % 4)) return false;
  return get_encode() != nullptr;
}

static int g_st256 = 1, g_simple = 3;      // bit 0: KIND 1 (pre-split K-major weights), bit 1: KIND 2 (dW products)
static int g_b_resident = 1, g_prefetch = 0, g_epi_pipe = 0, g_early_release = 0, g_wait_ticks = 0x989680;
void tagan_gemm_tma_set_tuning(int key, int value) {
  if (key == 0) g_b_resident = value;
  else if (key == 1) g_prefetch = value < 0 ? 0 : value;
  else if (key == 2) g_epi_pipe = value;
  else if (key == 3) g_early_release = value;
  else if (key == 4) g_wait_ticks = value;
  else if (key == 5) g_st256 = value;
  else if (key == 6) g_simple = value;
}
static long long* g_trace = nullptr;
void tagan_gemm_tma_set_trace(void* buf) { g_trace = static_cast<long long*>(buf); }

static inline int64_t presplit_ld(int64_t cols) { return (cols + 3) / 4 * 4; }
static inline bool want_presplit(int32_t op, int64_t N, int64_t K) { return op != 2 && N * K <= PRESPLIT_MAX_ELEMS; }

size_t tagan_gemm_tma_colsum_bytes(int64_t M, int64_t N, int64_t K) {
  Plan pl = make_plan(M, N, K);
  return 256 + 2 * (size_t)pl.splits * (size_t)M * sizeof(float);
}

size_t tagan_gemm_tma_workspace_bytes(int32_t op, int64_t M, int64_t N, int64_t K) {
  Plan pl = make_plan(M, N, K);
  size_t b = pl.splits > 1 ? (size_t)pl.parts * (size_t)M * (size_t)N * sizeof(float) : 0;
  if (want_presplit(op, N, K)) {
    b = (b + 255) / 256 * 256 + 2 * (size_t)N * presplit_ld(K) * sizeof(float) + 256;
  }
  return b;
}

int tagan_gemm_tma(int32_t op, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                   int64_t ldb, const float* bias, float* C, int64_t ldc, int32_t accumulate, int32_t passes,
                   void* workspace, size_t workspace_bytes, cudaStream_t st, float* colsum_a /* TN only, [M] or null */,
                   const float* A2, int64_t lda2, int64_t K1, const tagan_epilogue* epi, int32_t epi_fast) {
  // the opt-in to > 48 KB of dynamic shared memory is per device: remember it per device ordinal
  static bool attr_set[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = -1;
  if (dev < 0 || !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tma_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_tma_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_tma_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_tma_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    if (dev >= 0) attr_set[dev] = true;
  }
  Plan pl = make_plan(M, N, K);
  Params p;
  p.fused = 0;
  p.K1 = 0;
  p.epi = tagan_epilogue{};
  // a STORE epilogue without extra inputs / outputs is the plain epilogue writing to out0 (resident weights, 256-bit stores)
  if (epi != nullptr && epi->mode == TAGAN_EPI_STORE && epi->in0 == nullptr && epi->in1 == nullptr && epi->out1 == nullptr &&
      epi->out2 == nullptr && op != 2 && pl.splits == 1 && colsum_a == nullptr && !accumulate) {
    C = epi->out0;
    ldc = epi->ld_out0;
    epi = nullptr;
  }
  if (epi != nullptr) {
    if (op == 2 || pl.splits != 1 || colsum_a != nullptr || accumulate) return TAGAN_E_UNSUPPORTED;
    if (epi->mode == TAGAN_EPI_RES_LN && epi->gamma != nullptr && pl.tiles_n != 1) return TAGAN_E_UNSUPPORTED;
    p.fused = epi_fast ? 2 : 1;
    p.epi = *epi;
  }
  if (A2 != nullptr) {
    if (op != 0 || K1 <= 0 || K1 >= K || (K1 % BK) != 0) return TAGAN_E_UNSUPPORTED;
    p.K1 = K1;
  }
  p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.C = C; p.ldc = ldc;
  p.accumulate = accumulate;
  p.passes = passes;
  p.a_mn_major = (op == 2);
  p.b_mn_major = (op != 0);
  p.tiles_m = pl.tiles_m; p.tiles_n = pl.tiles_n; p.splits = pl.splits; p.k_per_split = pl.k_per_split;
  p.partial = nullptr;
  p.cta_acc = pl.cta_acc;
  if (pl.splits > 1) {
    if (!workspace || workspace_bytes < (size_t)pl.parts * M * N * sizeof(float)) return TAGAN_E_WORKSPACE;
    p.partial = static_cast<float*>(workspace);
    if (pl.cta_acc) {
      cudaError_t me = cudaMemsetAsync(p.partial, 0, (size_t)pl.parts * M * N * sizeof(float), st);
      if (me != cudaSuccess) return (int)me;
    }
  }
  p.colsum_part = nullptr;
  if (colsum_a != nullptr) {
    if (op != 2) return TAGAN_E_INVALID;
    const size_t off = pl.splits > 1 ? ((size_t)pl.parts * M * N * sizeof(float) + 255) / 256 * 256 : 0;
    if (!workspace || workspace_bytes < off + 2 * (size_t)pl.splits * M * sizeof(float)) return TAGAN_E_WORKSPACE;
    p.colsum_part = reinterpret_cast<float*>(static_cast<char*>(workspace) + off);
  }
  if (p.partial) p.c_vec = (N % 4 == 0);
  else p.c_vec = ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(bias) & 15) == 0);
  CUtensorMap tmA, tmB, tmB2, tmA2;
  // NT: A[M,K], B[N,K] (K-major).  NN: A[M,K], B[K,N] (MN-major).  TN: A[K,M], B[K,N] (both MN-major).
  bool okA = p.a_mn_major ? make_map(&tmA, A, K, M, lda, true, true)
                          : make_map(&tmA, A, M, p.K1 > 0 ? p.K1 : K, lda, false);
  if (p.K1 > 0) okA = okA && make_map(&tmA2, A2, M, K - p.K1, lda2, false);
  else tmA2 = tmA;
  p.b_presplit = 0;
  bool okB;
  if (want_presplit(op, N, K) && passes != 1) {
    // images are always [N, K] (K-major): an NN product's W[K,N] is transposed on the way
    const int64_t rows = op == 0 ? N : K, cols = op == 0 ? K : N, ldo = presplit_ld(K);
    size_t off = pl.splits > 1 ? ((size_t)pl.parts * M * N * sizeof(float) + 255) / 256 * 256 : 0;
    if (!workspace || workspace_bytes < off + 2 * (size_t)N * ldo * sizeof(float)) return TAGAN_E_WORKSPACE;
    float* hi = reinterpret_cast<float*>(static_cast<char*>(workspace) + off);
    float* lo = hi + N * ldo;
    p.b_mn_major = 0;
    okB = make_map(&tmB, hi, N, K, ldo, false) && make_map(&tmB2, lo, N, K, ldo, false);
    p.b_presplit = 1;
    // the tensor maps only need addresses: encode them first so that nothing is enqueued when encoding fails
    if (okA && okB) {
      if (op == 0) presplit_kernel<false><<<ceil_div_i64(rows * cols, 256), 256, 0, st>>>(B, rows, cols, ldb, hi, lo, ldo);
      else presplit_kernel<true><<<ceil_div_i64(rows * cols, 256), 256, 0, st>>>(B, rows, cols, ldb, hi, lo, ldo);
    }
  } else {
    okB = p.b_mn_major ? make_map(&tmB, B, K, N, ldb, true) : make_map(&tmB, B, N, K, ldb, false);
    tmB2 = tmB;
  }
  if (!okA || !okB) return TAGAN_E_UNSUPPORTED;
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | ((uint32_t)p.b_mn_major << 16) |
            ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  int64_t work = (int64_t)pl.tiles_m * pl.tiles_n * pl.splits;
  int grid = (int)(work < 148 ? work : 148);
  // resident weights: every work item of a CTA has the same N panel when the grid is a multiple of tiles_n
  p.prefetch = g_prefetch;
  p.epi_pipe = g_epi_pipe;
  p.ticks = (uint32_t)g_wait_ticks;
  p.st256 = (g_st256 && p.partial == nullptr && (reinterpret_cast<uintptr_t>(C) & 31) == 0 && (ldc % 8) == 0 &&
             (reinterpret_cast<uintptr_t>(bias) & 31) == 0) ? 1 : 0;
  p.early_release = 0;
  p.trace = g_trace;
  p.b_resident = 0;
  if (g_b_resident && p.b_presplit && pl.splits == 1 && K <= (int64_t)STAGES * BK && pl.tiles_n <= 74 && work >= 2 * 148) {
    p.b_resident = 1;
    p.early_release = g_early_release;
    grid = 148 / pl.tiles_n * pl.tiles_n;
  }
  const bool simple = (g_simple & 1) && !p.fused && p.b_presplit && !p.a_mn_major && !p.b_mn_major && p.partial == nullptr && !accumulate &&
                      p.colsum_part == nullptr && p.prefetch == 0 && !p.early_release && !p.epi_pipe && p.trace == nullptr;
  const bool knobs_off = p.prefetch == 0 && !p.early_release && !p.epi_pipe && p.trace == nullptr;
  const bool tnk = (g_simple & 2) && !p.fused && p.a_mn_major && p.b_mn_major && !p.b_presplit && p.partial != nullptr && p.cta_acc &&
                   knobs_off;
  if (p.fused) gemm_tma_kernel<true, 0><<<grid, THREADS, SMEM_BYTES, st>>>(p, tmA, tmB, tmB2, tmA2);
  else if (simple) gemm_tma_kernel<false, 1><<<grid, THREADS, SMEM_BYTES, st>>>(p, tmA, tmB, tmB2, tmA2);
  else if (tnk) gemm_tma_kernel<false, 2><<<grid, THREADS, SMEM_BYTES, st>>>(p, tmA, tmB, tmB2, tmA2);
  else gemm_tma_kernel<false, 0><<<grid, THREADS, SMEM_BYTES, st>>>(p, tmA, tmB, tmB2, tmA2);
  if (p.partial)
    tma_splitk_reduce<<<ceil_div_i64(M * N, 32), 256, 0, st>>>(p.partial, pl.parts, M, N, bias, C, ldc, accumulate);
  if (p.colsum_part)
    colsum_parts_reduce<<<ceil_div_i64(M, 32), 1024, 0, st>>>(p.colsum_part, 2 * pl.splits, M, colsum_a);
  return tagan_launch_status();
}
