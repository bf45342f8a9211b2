// Dense projections on the 5th-gen tensor cores (tcgen05 + TMEM), fp32-accurate via 3xTF32.
//
//   C[M,N] = op(A) . op(B)   NT: A[M,K],B[N,K] (nn.Linear forward)   NN: A[M,K],B[K,N] (dX = dY.W)
//                            TN: A[K,M],B[K,N] (dW = dY^T.X, split-K, fixed-order reduction)
//
// The reference's projections are fp32 nn.Linear; parity is rtol 1e-4 / atol 1e-5, which plain TF32
// (10-bit mantissa) misses.  Every fp32 operand is split on the fly into hi = rn_tf32(x) and lo = rn_tf32(x - hi) (both exactly
// representable in TF32, unbiased), and D += Alo.Blo + Alo.Bhi + Ahi.Blo + Ahi.Bhi is
// accumulated in fp32 in TMEM (the MMA pipe is far from being the bottleneck, so the lo.lo term is kept).
//
// Persistent, warp-specialised CTA (one per SM), 128x128 output tiles, BK = 32 floats = one 128-byte
// swizzle row, 3-stage smem ring, 2 accumulator stages in TMEM (256 columns):
//   warps 0-3   epilogue : tcgen05.ld 32 lanes x 32 columns -> +bias/+C -> global (row-contiguous 128 B runs)
//   warp  4     MMA      : one elected thread issues tcgen05.mma.kind::tf32 (3 per k-step) and tcgen05.commit
//   warps 5-20  producers (two groups of 8 warps alternating k-blocks): coalesced 128-bit global loads -> hi/lo split -> st.shared into the canonical
//                          SWIZZLE_128B UMMA layouts (K-major or MN-major, so no transposes are ever
//                          materialised) -> fence.proxy.async -> mbarrier arrive
// With K <= 512 and N <= 768 these GEMMs sit at ~140 flop/byte even at 3x work, i.e. they are HBM-bound
// once on tensor cores; the producers are sized to stream A at HBM rate while B (weights) stays in L2.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int STAGES = 3;
constexpr int ACC_STAGES = 2;
constexpr int TILE_BYTES = BM * BK * 4;                 // 16 KB (A and B tiles have the same size)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;             // A_hi, A_lo, B_hi, B_lo
constexpr int EPI_WARPS = 4, PROD_WARPS = 16;
constexpr int MMA_WARP = EPI_WARPS;                     // warp 4
constexpr int THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32;   // 416
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int PROD_GROUPS = 2;                          // groups alternate k-blocks: TLP hides the load latency
constexpr int GROUP_WARPS = PROD_WARPS / PROD_GROUPS;   // 8 warps = 256 threads per group
constexpr int TMEM_COLS = ACC_STAGES * BN;              // 256
constexpr int EPI_PITCH = 36;                           // floats; 144-byte rows keep the staging tile conflict-free
constexpr int EPI_STAGE_BYTES = 32 * EPI_PITCH * 4;     // per epilogue warp
constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + (size_t)STAGES * STAGE_BYTES + 256 /*barriers*/ +
                              (size_t)EPI_WARPS * EPI_STAGE_BYTES;

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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n }"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accum) : "memory");
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

// Canonical SWIZZLE_128B tile offsets (bytes) for a [128 mn x 32 k] fp32 tile.
//   K-major : 8-row groups of 128-byte rows (k contiguous), 16-byte chunk index XOR (row % 8); SBO = 1024.
//   MN-major (SWIZZLE_128B_BASE32B, Swizzle<2,5,2>): atoms of 4 k-rows x 128 bytes (32 mn contiguous), the
//             32-byte chunk index XOR (k % 4); k-atoms 512 B apart (SBO), mn-blocks of 32 floats 4096 B apart (LBO).
__device__ __forceinline__ uint32_t off_kmajor(int mn, int kchunk /*k/4*/) {
  return (uint32_t)((mn >> 3) * 1024 + (mn & 7) * 128 + ((kchunk ^ (mn & 7)) << 4));
}
__device__ __forceinline__ uint32_t off_mnmajor(int mnchunk /*mn/4*/, int k) {
  const int c16 = mnchunk & 7;                     // 16-byte chunk inside the 128-byte row
  return (uint32_t)((mnchunk >> 3) * 4096 + (k >> 2) * 512 + (k & 3) * 128 + ((((c16 >> 1) ^ (k & 3)) << 5) | ((c16 & 1) << 4)));
}

// Round-to-nearest TF32 (unbiased: a truncating split would accumulate a systematic error ~K*2^-22).
// Same result as cvt.rna.tf32.f32 (nearest, ties away from zero) with two full-rate integer ops.
__device__ __forceinline__ float tf32_rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
// hi = rn_tf32(x), lo = rn_tf32(x - hi): both exactly representable, so the tensor core's own
// fp32->tf32 conversion is the identity and every product is exact.
__device__ __forceinline__ void split_store(char* hi_tile, char* lo_tile, uint32_t off, float4 v) {
  float4 h, l;
  h.x = tf32_rn(v.x); h.y = tf32_rn(v.y); h.z = tf32_rn(v.z); h.w = tf32_rn(v.w);
  l.x = tf32_rn(v.x - h.x); l.y = tf32_rn(v.y - h.y); l.z = tf32_rn(v.z - h.z); l.w = tf32_rn(v.w - h.w);
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

// element-wise guarded load of 4 consecutive floats along the contiguous dimension
__device__ __forceinline__ float4 load4(const float* __restrict__ base, int64_t row, int64_t ld, int64_t col, int64_t nrows,
                                        int64_t ncols, bool vec_ok) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < nrows) {
    const float* p = base + row * ld + col;
    if (vec_ok && col + 3 < ncols) {
      v = __ldg(reinterpret_cast<const float4*>(p));
    } else {
      if (col < ncols) v.x = __ldg(p);
      if (col + 1 < ncols) v.y = __ldg(p + 1);
      if (col + 2 < ncols) v.z = __ldg(p + 2);
      if (col + 3 < ncols) v.w = __ldg(p + 3);
    }
  }
  return v;
}

struct Params {
  int64_t M, N, K;
  const float* A; int64_t lda;
  const float* B; int64_t ldb;
  const float* bias;
  float* C; int64_t ldc;
  float* partial;          // split-K partial tiles [splits][M][N] or null
  int accumulate;
  int a_mn_major, b_mn_major;
  int tiles_m, tiles_n, splits;
  int64_t k_per_split;
  int a_vec, b_vec, c_vec;
  int passes;              // 4 / 3 (without lo.lo) / 1 (plain TF32)
  uint32_t idesc;
};

__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC_STAGES);
  float* epi_stage = reinterpret_cast<float*>(smem + (size_t)STAGES * STAGE_BYTES + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], GROUP_WARPS); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < ACC_STAGES; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t num_work = (int64_t)p.tiles_m * p.tiles_n * p.splits;

  if (warp > MMA_WARP) {
    // ============================== producers ==============================
    // 512 threads share the 2048 float4 chunks of a k-block (A tile + B tile).  Global loads run two
    // k-blocks ahead of the shared-memory stores through a 3-deep register ring (across tile
    // boundaries too), so ~64 KB per SM are in flight and the A stream is not latency-bound.
    // Two groups of 8 warps; group g produces k-blocks g, g+2, g+4, ... of this CTA's sequence (across
    // tile boundaries).  A group issues its 8 loads per thread, waits for the stage to be free, splits and
    // stores, then arrives on the stage's full barrier.  While one group waits on HBM the other one stores,
    // so two k-blocks (64 KB) are always in flight without any register-resident prefetch ring.
    const int group = (warp - (MMA_WARP + 1)) / GROUP_WARPS;
    const int pt = (threadIdx.x - (MMA_WARP + 1) * 32) % (GROUP_WARPS * 32);      // 0..255 inside the group
    const int kb_per_split = (int)(p.k_per_split / BK);
    const int kb_total = (int)((p.K + BK - 1) / BK);
    const int nwork = (int)num_work;
    // chunk i (0..3) of a tile for this thread: K-major: row (pt>>3)+32i, 16-byte k-chunk pt&7;
    // MN-major: k row (pt>>5)+8i, 16-byte mn-chunk pt&31.  smem / global offsets are affine in i.
    const uint32_t sa0 = p.a_mn_major ? off_mnmajor(pt & 31, pt >> 5) : off_kmajor(pt >> 3, pt & 7);
    const uint32_t sb0 = 2u * TILE_BYTES + (p.b_mn_major ? off_mnmajor(pt & 31, pt >> 5) : off_kmajor(pt >> 3, pt & 7));
    const uint32_t sas = p.a_mn_major ? 1024u : 4096u, sbs = p.b_mn_major ? 1024u : 4096u;
    const int64_t ga0 = p.a_mn_major ? (int64_t)(pt >> 5) * p.lda + (pt & 31) * 4 : (int64_t)(pt >> 3) * p.lda + (pt & 7) * 4;
    const int64_t gb0 = p.b_mn_major ? (int64_t)(pt >> 5) * p.ldb + (pt & 31) * 4 : (int64_t)(pt >> 3) * p.ldb + (pt & 7) * 4;
    const int64_t gas = (p.a_mn_major ? 8 : 32) * p.lda, gbs = (p.b_mn_major ? 8 : 32) * p.ldb;
    const int64_t a_kstep = p.a_mn_major ? (int64_t)BK * p.lda : BK;      // elements per k-block
    const int64_t b_kstep = p.b_mn_major ? (int64_t)BK * p.ldb : BK;
    int w = blockIdx.x, kb = 0, nkb = 0, mt = 0, nt = 0, kbeg = 0;
    auto setup = [&]() {
      if (w >= nwork) { nkb = 0; return; }
      nt = w % p.tiles_n;
      const int q = w / p.tiles_n;
      mt = q % p.tiles_m;
      kbeg = (q / p.tiles_m) * kb_per_split;
      const int rem = kb_total - kbeg;
      nkb = rem < kb_per_split ? rem : kb_per_split;
      kb = 0;
    };
    auto advance = [&]() {
      if (++kb >= nkb) { w += gridDim.x; setup(); }
    };
    setup();
    int j = 0;                                            // index in the CTA's k-block sequence
    for (int g = 0; g < group && w < nwork; ++g) { advance(); ++j; }
    while (w < nwork) {
      const int64_t m0 = (int64_t)mt * BM, n0 = (int64_t)nt * BN;
      const int64_t k0 = (int64_t)(kbeg + kb) * BK;
      const int64_t klim = (int64_t)(kbeg + nkb) * BK < p.K ? (int64_t)(kbeg + nkb) * BK : p.K;
      float4 r[8];
      if (p.a_vec && p.b_vec && m0 + BM <= p.M && n0 + BN <= p.N && k0 + BK <= klim) {
        const float* pa = p.A + (p.a_mn_major ? m0 : m0 * p.lda) + (int64_t)(kbeg + kb) * a_kstep + ga0;
        const float* pb = p.B + (p.b_mn_major ? n0 : n0 * p.ldb) + (int64_t)(kbeg + kb) * b_kstep + gb0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          r[i] = __ldg(reinterpret_cast<const float4*>(pa + i * gas));
          r[4 + i] = __ldg(reinterpret_cast<const float4*>(pb + i * gbs));
        }
      } else {             // edge tiles: guarded element-wise loads with zero fill
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!p.a_mn_major) r[i] = load4(p.A, m0 + (pt >> 3) + 32 * i, p.lda, k0 + (pt & 7) * 4, p.M, klim, p.a_vec);
          else r[i] = load4(p.A, k0 + (pt >> 5) + 8 * i, p.lda, m0 + (pt & 31) * 4, klim, p.M, p.a_vec);
          if (!p.b_mn_major) r[4 + i] = load4(p.B, n0 + (pt >> 3) + 32 * i, p.ldb, k0 + (pt & 7) * 4, p.N, klim, p.b_vec);
          else r[4 + i] = load4(p.B, k0 + (pt >> 5) + 8 * i, p.ldb, n0 + (pt & 31) * 4, klim, p.N, p.b_vec);
        }
      }
      const int stage = j % STAGES;
      mbar_wait(&empty_bar[stage], ((j / STAGES) & 1) ^ 1);
      char* st = reinterpret_cast<char*>(smem) + (size_t)stage * STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        split_store(st, st + TILE_BYTES, sa0 + i * sas, r[i]);
        split_store(st, st + TILE_BYTES, sb0 + i * sbs, r[4 + i]);
      }
      fence_proxy_async();                   // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
      // skip the other groups' k-blocks
      for (int g = 0; g < PROD_GROUPS && w < nwork; ++g) advance();
      j += PROD_GROUPS;
    }
  } else if (warp == MMA_WARP) {
    // ============================== MMA issuer ==============================
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      // per k-step (8 tf32 = 32 bytes) descriptor advance: +32 B inside the swizzle row (K-major),
      // +1024 B = two 4-row k-atoms further (MN-major)
      const uint32_t a_step = p.a_mn_major ? 1024u : 32u, b_step = p.b_mn_major ? 1024u : 32u;
      const uint32_t a_lbo = p.a_mn_major ? 4096u : 16u, b_lbo = p.b_mn_major ? 4096u : 16u;
      const uint32_t a_sbo = p.a_mn_major ? 512u : 1024u, b_sbo = p.b_mn_major ? 512u : 1024u;
      const uint32_t a_lay = p.a_mn_major ? 1u : 2u, b_lay = p.b_mn_major ? 1u : 2u;
      for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int ks = (int)(w / ((int64_t)p.tiles_n * p.tiles_m));
        const int64_t kbeg = (int64_t)ks * p.k_per_split;
        const int64_t kend = kbeg + p.k_per_split < p.K ? kbeg + p.k_per_split : p.K;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);        // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        uint32_t accum = 0;
        for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sbase = smem_u32(smem + (size_t)stage * STAGE_BYTES);
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t da_hi = make_desc(sbase + kk * a_step, a_lbo, a_sbo, a_lay);
            const uint64_t da_lo = make_desc(sbase + TILE_BYTES + kk * a_step, a_lbo, a_sbo, a_lay);
            const uint64_t db_hi = make_desc(sbase + 2 * TILE_BYTES + kk * b_step, b_lbo, b_sbo, b_lay);
            const uint64_t db_lo = make_desc(sbase + 3 * TILE_BYTES + kk * b_step, b_lbo, b_sbo, b_lay);
            if (p.passes == 4) {
              umma_tf32(tmem_d, da_lo, db_lo, p.idesc, accum);   // small terms first
              umma_tf32(tmem_d, da_lo, db_hi, p.idesc, 1);
              umma_tf32(tmem_d, da_hi, db_lo, p.idesc, 1);
              umma_tf32(tmem_d, da_hi, db_hi, p.idesc, 1);
            } else if (p.passes == 3) {
              umma_tf32(tmem_d, da_lo, db_hi, p.idesc, accum);
              umma_tf32(tmem_d, da_hi, db_lo, p.idesc, 1);
              umma_tf32(tmem_d, da_hi, db_hi, p.idesc, 1);
            } else {
              umma_tf32(tmem_d, da_hi, db_hi, p.idesc, accum);
            }
            accum = 1;
          }
          umma_commit(&empty_bar[stage]);                    // frees the smem stage when the MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);                        // accumulator complete -> epilogue
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ============================== epilogue ==============================
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t w = blockIdx.x; w < num_work; w += gridDim.x) {
      const int nt = (int)(w % p.tiles_n);
      const int mt = (int)((w / p.tiles_n) % p.tiles_m);
      const int ks = (int)(w / ((int64_t)p.tiles_n * p.tiles_m));
      const int64_t kbeg = (int64_t)ks * p.k_per_split;
      const bool has_k = kbeg < p.K;
      const int64_t n0 = (int64_t)nt * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      float* out = p.partial ? p.partial + (int64_t)ks * p.M * p.N : p.C;
      const int64_t ldo = p.partial ? p.N : p.ldc;
#pragma unroll 1
      for (int cc = 0; cc < BN / 32; ++cc) {
        uint32_t r[32];
        if (has_k) {
          const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BN + cc * 32);
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
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        // registers (lane = row, 32 consecutive columns) -> padded smem tile -> row-contiguous 128-byte stores
        float* stg = epi_stage + warp * (32 * EPI_PITCH);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stg + lane * EPI_PITCH + j) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        __syncwarp();
        const int64_t c0 = n0 + cc * 32 + (lane & 7) * 4;
        const bool direct = p.partial == nullptr;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (direct && p.bias) {
          if (c0 < p.N) b4.x = p.bias[c0];
          if (c0 + 1 < p.N) b4.y = p.bias[c0 + 1];
          if (c0 + 2 < p.N) b4.z = p.bias[c0 + 2];
          if (c0 + 3 < p.N) b4.w = p.bias[c0 + 3];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + (lane >> 3);
          const int64_t grow = (int64_t)mt * BM + warp * 32 + rr;
          if (grow >= p.M || c0 >= p.N) continue;
          float4 v = *reinterpret_cast<const float4*>(stg + rr * EPI_PITCH + (lane & 7) * 4);
          v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
          float* orow = out + grow * ldo + c0;
          if (c0 + 3 < p.N && p.c_vec) {
            if (direct && p.accumulate) {
              float4 o4 = *reinterpret_cast<const float4*>(orow);
              v.x += o4.x; v.y += o4.y; v.z += o4.z; v.w += o4.w;
            }
            *reinterpret_cast<float4*>(orow) = v;
          } else {
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (c0 + q < p.N) orow[q] = (direct && p.accumulate) ? orow[q] + vv[q] : vv[q];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
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

__global__ void tc_splitk_reduce(const float* __restrict__ partial, int parts, int64_t M, int64_t N,
                                 const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int accumulate) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  float s = 0.f;
  for (int q = 0; q < parts; ++q) s += partial[(int64_t)q * M * N + i];
  const int64_t m = i / N, n = i - m * N;
  if (bias) s += bias[n];
  float* o = C + m * ldc + n;
  *o = accumulate ? *o + s : s;
}

struct Plan { int tiles_m, tiles_n, splits; int64_t k_per_split; };
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
  int64_t kps = (K + splits - 1) / splits;
  kps = (kps + BK - 1) / BK * BK;
  if (kps < BK) kps = BK;
  pl.k_per_split = kps;
  pl.splits = (int)((K + kps - 1) / kps);
  if (pl.splits < 1) pl.splits = 1;
  return pl;
}

}  // namespace

bool tagan_gemm_tc_supported(int32_t op, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                             int64_t ldb) {
  (void)op; (void)M; (void)N; (void)A; (void)B; (void)lda; (void)ldb;
  return K > 0;
}

size_t tagan_gemm_tc_workspace_bytes(int32_t op, int64_t M, int64_t N, int64_t K) {
  (void)op;
  Plan pl = make_plan(M, N, K);
  return pl.splits > 1 ? (size_t)pl.splits * (size_t)M * (size_t)N * sizeof(float) : 0;
}

int tagan_gemm_tc(int32_t op, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                  int64_t ldb, const float* bias, float* C, int64_t ldc, int32_t accumulate, int32_t passes,
                  void* workspace, size_t workspace_bytes, cudaStream_t st) {
  static SmemOptIn opt_in;
  {
    const cudaError_t e = opt_in.ensure(gemm_tc_kernel, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
  }
  Plan pl = make_plan(M, N, K);
  Params p;
  p.M = M; p.N = N; p.K = K;
  p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.bias = bias; p.C = C; p.ldc = ldc;
  p.accumulate = accumulate;
  p.passes = passes;
  p.a_mn_major = (op == 2);
  p.b_mn_major = (op != 0);
  p.tiles_m = pl.tiles_m; p.tiles_n = pl.tiles_n; p.splits = pl.splits; p.k_per_split = pl.k_per_split;
  p.partial = nullptr;
  if (pl.splits > 1) {
    if (!workspace || workspace_bytes < (size_t)pl.splits * M * N * sizeof(float)) return TAGAN_E_WORKSPACE;
    p.partial = static_cast<float*>(workspace);
  }
  p.a_vec = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (lda % 4 == 0);
  p.b_vec = ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (ldb % 4 == 0);
  if (p.partial) p.c_vec = (N % 4 == 0);
  else p.c_vec = ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (ldc % 4 == 0) &&
                 (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0);
  // cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b_format TF32 (2) [7,10)/[10,13),
  // a_major [15], b_major [16] (1 = MN-major), n_dim = N>>3 [17,23), m_dim = M>>4 [24,29)
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn_major << 15) | ((uint32_t)p.b_mn_major << 16) |
            ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  int64_t work = (int64_t)pl.tiles_m * pl.tiles_n * pl.splits;
  int grid = (int)(work < 148 ? work : 148);
  gemm_tc_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(p);
  if (p.partial)
    tc_splitk_reduce<<<ceil_div_i64(M * N, 256), 256, 0, st>>>(p.partial, pl.splits, M, N, bias, C, ldc, accumulate);
  return tagan_launch_status();
}
