// (a2-a4) Fused geometric attention over a destination-sorted CSR.
// Replaces GeometricAttention._get_attention_weights + `attn @ v` of the reference
// (src/tagan/layers/geometric_attention.py:332-516, :579), which materialises [1,h,N,N].
//
// Layout: one warp per destination row.  A row of H floats is spread over the warp as NCHUNK
// chunks of 32*VEC floats, lane l owning VEC contiguous floats of each chunk, so every gather
// of a K/V/Q/dCtx row is NCHUNK fully coalesced 128-bit (VEC=4) warp loads.  A head of D floats
// is owned by `group = D/VEC` adjacent lanes; per-head dot products are xor-shuffle reductions
// over the group.  Softmax is online (running max / sum per head in registers); nothing is
// accumulated with atomics.  Backward recomputes the scores: a row pass over the CSR produces
// dQ and delta = dctx.ctx, a column pass over the transposed CSR produces dK and dV.
//
// HBM-bound: algorithmic bytes are nnz*(2*H*4+4) + N*(2*H*4+8h+8) forward.
#include "common.cuh"

// Storage type of the gathered operands Q, K, V.  This file is compiled twice: as is (fp32, the parity mode) and through
// geo_attn_bf16.cu with TAGAN_GEO_BF16 defined (bf16 rows: half the gather bytes; scores, softmax and every accumulation stay
// fp32, outputs and gradients are fp32).  The bf16 build exports the same entry points with a `_bf16` suffix.
#ifdef TAGAN_GEO_BF16
#include <cuda_bf16.h>
typedef __nv_bfloat16 qkv_t;
typedef uint16_t qkv_api_t;           // the C ABI carries bf16 rows as raw 16-bit words
#define GEO_NAME(n) n##_bf16
#else
typedef float qkv_t;
typedef float qkv_api_t;
#define GEO_NAME(n) n
#endif

namespace {

constexpr int WARPS_PER_BLOCK = 8;
#ifndef FWD_U4
#define FWD_U4 4            // gathered entries in flight per warp in the forward walk at H = 128 (8: 2.87 -> 3.60 ms, spills)
#endif

template <int METRIC> struct MetricTraits {
  static constexpr bool kDot = METRIC == TAGAN_METRIC_SCALED_DOT || METRIC == TAGAN_METRIC_DOT;
  static constexpr bool kCos = METRIC == TAGAN_METRIC_COSINE_SIM || METRIC == TAGAN_METRIC_COSINE_DIST;
  static constexpr bool kManhattan = METRIC == TAGAN_METRIC_MANHATTAN;
  static constexpr bool kParam = METRIC == TAGAN_METRIC_GAUSSIAN || METRIC == TAGAN_METRIC_RBF;
};

// The three kernels are instruction-issue-bound at config 3 (ncu: 69 % issue-active, ~100 instructions per gathered entry,
// L2 36 %, DRAM 45 %: profiles/r02_ncu_geo_batched_raw.csv), so the per-entry arithmetic uses the one-instruction SFU forms:
// exp(x) = ex2.approx(x * log2 e), sqrt.approx, approximate division.  Each is good to ~2 ulp (|s - max| <= 30 adds
// ~2e-6 relative to a probability), two orders of magnitude inside the rtol 1e-4 parity bar.
__device__ __forceinline__ float fast_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// group sum with the group size known at compile time (GROUP = 0: runtime loop); the walks below are instantiated for the
// common sizes 4 and 8 (D = 16 / 32 with 4 floats per lane) and fall back to the runtime form otherwise
template <int GROUP>
__device__ __forceinline__ float group_sum_t(float v, int group) {
  if (GROUP == 0) return group_sum(v, group);
#pragma unroll
  for (int o = GROUP >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}

// Per-(entry, head) score and the partial sums its backward needs.
//   p1: q.k (dot/cos) | sum (q-k)^2 (sq family) | sum |q-k| (manhattan);  p2: k.k (cos only)
template <int METRIC, int VEC, int GROUP = 0>
__device__ __forceinline__ float score_from_vectors(const float* q, const float* k, int group, float par, float qnorm,
                                                    float& p1, float& p2) {
  using MT = MetricTraits<METRIC>;
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    if (MT::kDot) a = fmaf(q[i], k[i], a);
    else if (MT::kCos) { a = fmaf(q[i], k[i], a); b = fmaf(k[i], k[i], b); }
    else if (MT::kManhattan) a += fabsf(q[i] - k[i]);
    else { float d = q[i] - k[i]; a = fmaf(d, d, a); }
  }
  a = group_sum_t<GROUP>(a, group);
  if (MT::kCos) b = group_sum_t<GROUP>(b, group);
  p1 = a; p2 = b;
  if (METRIC == TAGAN_METRIC_SCALED_DOT) return a * par;            // par = 1/sqrt(D)
  if (METRIC == TAGAN_METRIC_DOT) return a;
  if (MT::kCos) {
    float kn = sqrtf(b);
    kn = kn == 0.f ? 1e-8f : kn;
    float u = a / (qnorm * kn);
    u = fminf(fmaxf(u, -1.f), 1.f);
    return METRIC == TAGAN_METRIC_COSINE_SIM ? u : -(1.f - u);
  }
  if (METRIC == TAGAN_METRIC_EUCLIDEAN) return -fast_sqrt(a + 1e-8f);
  if (METRIC == TAGAN_METRIC_SQ_EUCLIDEAN) return -a;
  if (METRIC == TAGAN_METRIC_MANHATTAN) return -a;
  if (METRIC == TAGAN_METRIC_GAUSSIAN) return expf(-a / (2.f * par * par));   // par = sigma
  return expf(-par * a);                                                       // rbf, par = gamma
}

// out += g * d(score)/d(q)  (WRT_Q) or g * d(score)/d(k); returns g * d(score)/d(param).
template <int METRIC, int VEC, bool WRT_Q>
__device__ __forceinline__ float accum_score_grad(float g, const float* q, const float* k, float s, float p1, float p2,
                                                  float par, float qnorm, float* out) {
  using MT = MetricTraits<METRIC>;
  if (MT::kDot) {
    float c = METRIC == TAGAN_METRIC_SCALED_DOT ? g * par : g;
#pragma unroll
    for (int i = 0; i < VEC; ++i) out[i] = fmaf(c, WRT_Q ? k[i] : q[i], out[i]);
    return 0.f;
  }
  if (MT::kCos) {
    float kn = sqrtf(p2);
    const bool kzero = kn == 0.f;
    kn = kzero ? 1e-8f : kn;
    float inv = 1.f / (qnorm * kn);
    float u = p1 * inv;
    float gg = (u >= -1.f && u <= 1.f) ? g : 0.f;   // clamp passes gradient on the closed interval
    float self_c;
    if (WRT_Q) self_c = (qnorm == 1e-8f) ? 0.f : -gg * u / (qnorm * qnorm);
    else self_c = kzero ? 0.f : -gg * u / (kn * kn);
    float cross_c = gg * inv;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float self_v = WRT_Q ? q[i] : k[i];
      float cross_v = WRT_Q ? k[i] : q[i];
      out[i] = fmaf(cross_c, cross_v, fmaf(self_c, self_v, out[i]));
    }
    return 0.f;
  }
  if (MT::kManhattan) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      float d = q[i] - k[i];
      float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      out[i] += (WRT_Q ? -g : g) * sg;
    }
    return 0.f;
  }
  // squared-distance family: d(score)/dq = coef * (q - k), d/dk = -coef * (q - k)
  float coef, dpar = 0.f;
  if (METRIC == TAGAN_METRIC_EUCLIDEAN) coef = __fdividef(g, s);      // s = -sqrt(sq+eps)
  else if (METRIC == TAGAN_METRIC_SQ_EUCLIDEAN) coef = -2.f * g;
  else if (METRIC == TAGAN_METRIC_GAUSSIAN) {
    float is2 = 1.f / (par * par);
    coef = -g * s * is2;
    dpar = g * s * p1 * is2 / par;
  } else {
    coef = -2.f * par * g * s;
    dpar = -g * p1 * s;
  }
  if (!WRT_Q) coef = -coef;
#pragma unroll
  for (int i = 0; i < VEC; ++i) out[i] = fmaf(coef, q[i] - k[i], out[i]);
  return dpar;
}

template <int METRIC>
__device__ __forceinline__ float head_param(const float* metric_param, int head, int D) {
  if (METRIC == TAGAN_METRIC_SCALED_DOT) return 1.f / sqrtf((float)D);
  if (MetricTraits<METRIC>::kParam) return metric_param ? __ldg(metric_param + head) : 1.f;
  return 0.f;
}

template <int VEC, int NCHUNK>
__device__ __forceinline__ void load_row(float (&dst)[NCHUNK][VEC], const float* base, int lane) {
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) VecIO<VEC>::load(dst[c], base + c * 32 * VEC + lane * VEC);
}
#ifdef TAGAN_GEO_BF16
__device__ __forceinline__ float bf16_bits_to_float(uint32_t b) { return __uint_as_float(b << 16); }
// bf16 row: VEC contiguous bf16 per lane and chunk = one 2 / 4 / 8-byte load, widened to fp32 in registers
template <int VEC, int NCHUNK>
__device__ __forceinline__ void load_row(float (&dst)[NCHUNK][VEC], const __nv_bfloat16* base, int lane) {
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    const __nv_bfloat16* p = base + c * 32 * VEC + lane * VEC;
    if (VEC == 4) {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
      dst[c][0] = bf16_bits_to_float(u.x & 0xffffu); dst[c][1 % VEC] = bf16_bits_to_float(u.x >> 16);
      dst[c][2 % VEC] = bf16_bits_to_float(u.y & 0xffffu); dst[c][3 % VEC] = bf16_bits_to_float(u.y >> 16);
    } else if (VEC == 2) {
      const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(p));
      dst[c][0] = bf16_bits_to_float(u & 0xffffu); dst[c][1 % VEC] = bf16_bits_to_float(u >> 16);
    } else {
      dst[c][0] = bf16_bits_to_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p)));
    }
  }
}
#endif

// ---- L2 residency hints.  At config 3 one snapshot's K|V (100 MB) competes for the L2 with the kernel's own streams (Q rows,
// ctx / gradient rows, column ids: another ~115 MB per snapshot) and the gathers hit only 52 % of the time (ncu).  With mode 1
// the gathered rows are loaded with an evict_last policy and the streamed rows / results with evict_first
// (createpolicy + .L2::cache_hint: the policy rides in the instruction's descriptor, no extra instructions per access).
__constant__ int c_l2_mode = 1;
struct L2Policies { uint64_t keep, stream; };
__device__ __forceinline__ L2Policies make_l2_policies() {
  L2Policies p;
  if (c_l2_mode == 1) {
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p.keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p.stream));
  } else {
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p.keep));
    p.stream = p.keep;
  }
  return p;
}
template <int VEC> struct HintIO;
template <> struct HintIO<1> {
  static __device__ __forceinline__ void load(float* d, const float* p, uint64_t pol) {
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(d[0]) : "l"(p), "l"(pol));
  }
  static __device__ __forceinline__ void store(float* p, const float* s, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(s[0]), "l"(pol) : "memory");
  }
};
template <> struct HintIO<2> {
  static __device__ __forceinline__ void load(float* d, const float* p, uint64_t pol) {
    asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(d[0]), "=f"(d[1]) : "l"(p), "l"(pol));
  }
  static __device__ __forceinline__ void store(float* p, const float* s, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(s[0]), "f"(s[1]), "l"(pol) : "memory");
  }
};
template <> struct HintIO<4> {
  static __device__ __forceinline__ void load(float* d, const float* p, uint64_t pol) {
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]) : "l"(p), "l"(pol));
  }
  static __device__ __forceinline__ void store(float* p, const float* s, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 ::"l"(p), "f"(s[0]), "f"(s[1]), "f"(s[2]), "f"(s[3]), "l"(pol) : "memory");
  }
};
template <int VEC, int NCHUNK>
__device__ __forceinline__ void load_row(float (&dst)[NCHUNK][VEC], const float* base, int lane, uint64_t pol) {
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) HintIO<VEC>::load(dst[c], base + c * 32 * VEC + lane * VEC, pol);
}
#ifdef TAGAN_GEO_BF16
template <int VEC, int NCHUNK>
__device__ __forceinline__ void load_row(float (&dst)[NCHUNK][VEC], const __nv_bfloat16* base, int lane, uint64_t pol) {
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    const __nv_bfloat16* p = base + c * 32 * VEC + lane * VEC;
    if (VEC == 4) {
      uint32_t x, y;
      asm volatile("ld.global.nc.L2::cache_hint.v2.b32 {%0,%1}, [%2], %3;" : "=r"(x), "=r"(y) : "l"(p), "l"(pol));
      dst[c][0] = bf16_bits_to_float(x & 0xffffu); dst[c][1 % VEC] = bf16_bits_to_float(x >> 16);
      dst[c][2 % VEC] = bf16_bits_to_float(y & 0xffffu); dst[c][3 % VEC] = bf16_bits_to_float(y >> 16);
    } else if (VEC == 2) {
      uint32_t x;
      asm volatile("ld.global.nc.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(x) : "l"(p), "l"(pol));
      dst[c][0] = bf16_bits_to_float(x & 0xffffu); dst[c][1 % VEC] = bf16_bits_to_float(x >> 16);
    } else {
      unsigned short x;
      asm volatile("ld.global.nc.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(x) : "l"(p), "l"(pol));
      dst[c][0] = bf16_bits_to_float((uint32_t)x);
    }
  }
}
#endif

// Rows (or, in the column pass, source nodes) with more than HEAVY_THRESH entries are "heavy": the
// warp-per-row kernels skip them and a second launch gives each of them a whole CTA -- its 8 warps take
// contiguous slices of the entries and their partial states are merged through shared memory in warp
// order, so the result stays deterministic and independent of scheduling (power-law graphs, config 2).
// resident CTAs per SM the register allocator must allow: memory-level parallelism is the limiter of these
// gather kernels (profiles/r01_SUMMARY.md), so occupancy is worth a few spilled scalars
constexpr int min_ctas(int vec, int nchunk) { return vec * nchunk >= 16 ? 2 : (vec * nchunk >= 8 ? 3 : 4); }
constexpr int HEAVY_THRESH = 128;
constexpr int HEAVY_GRID = 148 * 2;

// Calls f(row) with the whole CTA for every row whose entry count exceeds HEAVY_THRESH.
template <typename F>
__device__ __forceinline__ void for_each_heavy_row(const int* __restrict__ ptr, int N, F f) {
  __shared__ int heavy_list[WARPS_PER_BLOCK * 32];
  __shared__ int heavy_count;
  // Rows are dealt to the CTAs round-robin (thread t of CTA b looks at row chunk + t*gridDim + b): hubs of generated and
  // real power-law graphs have neighbouring (small) ids, and a contiguous block of rows per CTA would serialise all of
  // them on one CTA.
  for (int64_t chunk = 0; chunk < N; chunk += (int64_t)gridDim.x * (WARPS_PER_BLOCK * 32)) {
    if (threadIdx.x == 0) heavy_count = 0;
    __syncthreads();
    const int64_t r = chunk + (int64_t)threadIdx.x * gridDim.x + blockIdx.x;
    if (r < N && ptr[r + 1] - ptr[r] > HEAVY_THRESH) heavy_list[atomicAdd(&heavy_count, 1)] = (int)r;
    __syncthreads();
    const int cnt = heavy_count;
    for (int i = 0; i < cnt; ++i) {
      f(heavy_list[i]);
      __syncthreads();
    }
  }
}

// slice of [beg,end) owned by warp w of WARPS_PER_BLOCK (multiples of 32 entries so index loads stay aligned)
__device__ __forceinline__ void warp_slice(int beg, int end, int w, int& b, int& e) {
  int per = (end - beg + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
  per = (per + 31) & ~31;
  b = min(end, beg + w * per);
  e = min(end, b + per);
}

template <int METRIC, int VEC, int NCHUNK>
struct RowCtx {
  float q[NCHUNK][VEC], par[NCHUNK], qn[NCHUNK];
  int head[NCHUNK];
  int group;
};

template <int METRIC, int VEC, int NCHUNK>
__device__ __forceinline__ void init_row(RowCtx<METRIC, VEC, NCHUNK>& rc, const qkv_t* qrow, const float* metric_param,
                                         int D, int lane, uint64_t pol) {
  rc.group = D / VEC;
  load_row<VEC, NCHUNK>(rc.q, qrow, lane, pol);
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
    rc.head[c] = (c * 32 * VEC + lane * VEC) / D;
    rc.par[c] = head_param<METRIC>(metric_param, rc.head[c], D);
    rc.qn[c] = 0.f;
    if (MetricTraits<METRIC>::kCos) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) t = fmaf(rc.q[c][i], rc.q[c][i], t);
      t = sqrtf(group_sum(t, rc.group));
      rc.qn[c] = t == 0.f ? 1e-8f : t;
    }
  }
}

// online-softmax walk over entries [beg,end) of one row.  U gathered entries are in flight per warp; the running maximum is
// updated once per group of U entries (one rescale of l / acc per group instead of one per entry).
template <int METRIC, int VEC, int NCHUNK, int GROUP>
__device__ __forceinline__ void fwd_walk_t(const RowCtx<METRIC, VEC, NCHUNK>& rc, const qkv_t* __restrict__ K,
                                           const qkv_t* __restrict__ V, int64_t ld, const int* __restrict__ col, int beg,
                                           int end, int lane, float (&m)[NCHUNK], float (&l)[NCHUNK],
                                           float (&acc)[NCHUNK][VEC], uint64_t pk) {
  constexpr int U = (VEC * NCHUNK >= 8) ? 2 : (VEC * NCHUNK >= 4 ? FWD_U4 : 4);
  for (int base = beg; base < end; base += 32) {
    const int n = min(32, end - base);
    const int mycol = lane < n ? __ldg(col + base + lane) : 0;
    for (int j = 0; j < n; j += U) {
      float kk[U][NCHUNK][VEC], vv[U][NCHUNK][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int cj = __shfl_sync(FULL_MASK, mycol, min(j + u, n - 1));
        load_row<VEC, NCHUNK>(kk[u], K + (int64_t)cj * ld, lane, pk);
        load_row<VEC, NCHUNK>(vv[u], V + (int64_t)cj * ld, lane, pk);
      }
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        float sv[U];
        float mn = m[c];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float p1, p2;
          sv[u] = score_from_vectors<METRIC, VEC, GROUP>(rc.q[c], kk[u][c], rc.group, rc.par[c], rc.qn[c], p1, p2);
          if (j + u >= n) sv[u] = -INFINITY;                 // padding slot of the last group (a duplicate of entry n-1)
          mn = fmaxf(mn, sv[u]);
        }
        const float sc = fast_exp(m[c] - mn);                 // m = -inf on the first group: exp(-inf) = 0
        l[c] *= sc;
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[c][i] *= sc;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float p = fast_exp(sv[u] - mn);
          l[c] += p;
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[c][i] = fmaf(p, vv[u][c][i], acc[c][i]);
        }
        m[c] = mn;
      }
    }
  }
}
template <int METRIC, int VEC, int NCHUNK>
__device__ __forceinline__ void fwd_walk(const RowCtx<METRIC, VEC, NCHUNK>& rc, const qkv_t* __restrict__ K,
                                         const qkv_t* __restrict__ V, int64_t ld, const int* __restrict__ col, int beg,
                                         int end, int lane, float (&m)[NCHUNK], float (&l)[NCHUNK],
                                         float (&acc)[NCHUNK][VEC], uint64_t pk) {
  if (rc.group == 4) fwd_walk_t<METRIC, VEC, NCHUNK, 4>(rc, K, V, ld, col, beg, end, lane, m, l, acc, pk);
  else if (rc.group == 8) fwd_walk_t<METRIC, VEC, NCHUNK, 8>(rc, K, V, ld, col, beg, end, lane, m, l, acc, pk);
  else fwd_walk_t<METRIC, VEC, NCHUNK, 0>(rc, K, V, ld, col, beg, end, lane, m, l, acc, pk);
}

template <int METRIC, int VEC, int NCHUNK>
__device__ __forceinline__ void fwd_attn_pass(const RowCtx<METRIC, VEC, NCHUNK>& rc, const qkv_t* __restrict__ K,
                                              int64_t ld, const int* __restrict__ col, int beg, int end, int lane,
                                              int heads, const float (&lse_c)[NCHUNK], float* __restrict__ attn) {
  for (int e = beg; e < end; ++e) {
    const int cj = __ldg(col + e);
    float kr[NCHUNK][VEC];
    load_row<VEC, NCHUNK>(kr, K + (int64_t)cj * ld, lane);
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      float p1, p2;
      const float s = score_from_vectors<METRIC, VEC>(rc.q[c], kr[c], rc.group, rc.par[c], rc.qn[c], p1, p2);
      if ((lane & (rc.group - 1)) == 0) attn[(int64_t)e * heads + rc.head[c]] = expf(s - lse_c[c]);
    }
  }
}

template <int METRIC, int VEC, int NCHUNK, bool HEAVY>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, min_ctas(VEC, NCHUNK))
geo_attn_fwd_kernel(const qkv_t* __restrict__ Q, int64_t ldq, const qkv_t* __restrict__ K, const qkv_t* __restrict__ V, int64_t ld,
                    const int* __restrict__ rowptr, const int* __restrict__ col, int N, int heads, int D,
                    const float* __restrict__ metric_param, float* __restrict__ ctx, float* __restrict__ lse,
                    float* __restrict__ attn) {
  constexpr int H = 32 * VEC * NCHUNK;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const L2Policies pol = make_l2_policies();
  auto finish = [&](const RowCtx<METRIC, VEC, NCHUNK>& rc, int row, const float (&m)[NCHUNK], const float (&l)[NCHUNK],
                    const float (&acc)[NCHUNK][VEC], float (&lse_c)[NCHUNK]) {
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      const float inv = 1.f / l[c];
      float o[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) o[i] = acc[c][i] * inv;
      HintIO<VEC>::store(ctx + (int64_t)row * H + c * 32 * VEC + lane * VEC, o, pol.stream);
      lse_c[c] = m[c] + logf(l[c]);
      if ((lane & (rc.group - 1)) == 0) lse[(int64_t)row * heads + rc.head[c]] = lse_c[c];   // group leader
    }
  };
  if (!HEAVY) {
    const int row = blockIdx.x * WARPS_PER_BLOCK + w;
    if (row >= N) return;
    const int beg = rowptr[row], end = rowptr[row + 1];
    if (end - beg > HEAVY_THRESH) return;
    RowCtx<METRIC, VEC, NCHUNK> rc;
    init_row<METRIC, VEC, NCHUNK>(rc, Q + (int64_t)row * ldq, metric_param, D, lane, pol.stream);
    float m[NCHUNK], l[NCHUNK], acc[NCHUNK][VEC], lse_c[NCHUNK];
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      m[c] = -INFINITY; l[c] = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[c][i] = 0.f;
    }
    fwd_walk<METRIC, VEC, NCHUNK>(rc, K, V, ld, col, beg, end, lane, m, l, acc, pol.keep);
    finish(rc, row, m, l, acc, lse_c);
    if (attn != nullptr) fwd_attn_pass<METRIC, VEC, NCHUNK>(rc, K, ld, col, beg, end, lane, heads, lse_c, attn);
  } else {
    __shared__ float sm_m[WARPS_PER_BLOCK][NCHUNK][32], sm_l[WARPS_PER_BLOCK][NCHUNK][32];
    __shared__ float sm_acc[WARPS_PER_BLOCK][NCHUNK][VEC][32];
    __shared__ float sm_lse[NCHUNK][32];
    for_each_heavy_row(rowptr, N, [&](int row) {
      const int beg = rowptr[row], end = rowptr[row + 1];
      int b, e;
      warp_slice(beg, end, w, b, e);
      RowCtx<METRIC, VEC, NCHUNK> rc;
      init_row<METRIC, VEC, NCHUNK>(rc, Q + (int64_t)row * ldq, metric_param, D, lane, pol.stream);
      float m[NCHUNK], l[NCHUNK], acc[NCHUNK][VEC], lse_c[NCHUNK];
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        m[c] = -INFINITY; l[c] = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[c][i] = 0.f;
      }
      fwd_walk<METRIC, VEC, NCHUNK>(rc, K, V, ld, col, b, e, lane, m, l, acc, pol.keep);
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        sm_m[w][c][lane] = m[c]; sm_l[w][c][lane] = l[c];
#pragma unroll
        for (int i = 0; i < VEC; ++i) sm_acc[w][c][i][lane] = acc[c][i];
      }
      __syncthreads();
      if (w == 0) {            // merge the 8 partial softmax states in warp order
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
          float M = sm_m[0][c][lane];
          for (int k = 1; k < WARPS_PER_BLOCK; ++k) M = fmaxf(M, sm_m[k][c][lane]);
          float L = 0.f, A[VEC];
#pragma unroll
          for (int i = 0; i < VEC; ++i) A[i] = 0.f;
          for (int k = 0; k < WARPS_PER_BLOCK; ++k) {
            const float sc = expf(sm_m[k][c][lane] - M);      // exp(-inf) = 0 for empty slices
            L = fmaf(sm_l[k][c][lane], sc, L);
#pragma unroll
            for (int i = 0; i < VEC; ++i) A[i] = fmaf(sm_acc[k][c][i][lane], sc, A[i]);
          }
          m[c] = M; l[c] = L;
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[c][i] = A[i];
        }
        finish(rc, row, m, l, acc, lse_c);
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) sm_lse[c][lane] = lse_c[c];
      }
      if (attn != nullptr) {
        __syncthreads();
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) lse_c[c] = sm_lse[c][lane];
        fwd_attn_pass<METRIC, VEC, NCHUNK>(rc, K, ld, col, b, e, lane, heads, lse_c, attn);
      }
    });
  }
}

// ---- row pass: dQ[i] = sum_e ds_e * dscore/dq, delta[i,h] = dctx_i . ctx_i, optional dparam partials.
template <int METRIC, int VEC, int NCHUNK, int GROUP>
__device__ __forceinline__ void bwd_row_walk_t(const RowCtx<METRIC, VEC, NCHUNK>& rc, const float (&go)[NCHUNK][VEC],
                                             const float (&ls)[NCHUNK], const float (&dl)[NCHUNK],
                                             const qkv_t* __restrict__ K, const qkv_t* __restrict__ V, int64_t ld,
                                             const int* __restrict__ col, int beg, int end, int lane,
                                             float (&dq)[NCHUNK][VEC], float (&dpar)[NCHUNK], uint64_t pk) {
  constexpr int U = (VEC * NCHUNK >= 8) ? 2 : 4;
  for (int base = beg; base < end; base += 32) {
    const int n = min(32, end - base);
    const int mycol = lane < n ? __ldg(col + base + lane) : 0;
    for (int j = 0; j < n; j += U) {
      float kk[U][NCHUNK][VEC], vv[U][NCHUNK][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int cj = __shfl_sync(FULL_MASK, mycol, min(j + u, n - 1));
        load_row<VEC, NCHUNK>(kk[u], K + (int64_t)cj * ld, lane, pk);
        load_row<VEC, NCHUNK>(vv[u], V + (int64_t)cj * ld, lane, pk);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j + u < n) {
#pragma unroll
          for (int c = 0; c < NCHUNK; ++c) {
            float p1, p2;
            const float s = score_from_vectors<METRIC, VEC, GROUP>(rc.q[c], kk[u][c], rc.group, rc.par[c], rc.qn[c], p1, p2);
            const float a = fast_exp(s - ls[c]);
            float dp = 0.f;
#pragma unroll
            for (int i = 0; i < VEC; ++i) dp = fmaf(go[c][i], vv[u][c][i], dp);
            dp = group_sum_t<GROUP>(dp, rc.group);
            const float ds = a * (dp - dl[c]);
            dpar[c] += accum_score_grad<METRIC, VEC, true>(ds, rc.q[c], kk[u][c], s, p1, p2, rc.par[c], rc.qn[c], dq[c]);
          }
        }
      }
    }
  }
}

template <int METRIC, int VEC, int NCHUNK>
__device__ __forceinline__ void bwd_row_walk(const RowCtx<METRIC, VEC, NCHUNK>& rc, const float (&go)[NCHUNK][VEC],
                                             const float (&ls)[NCHUNK], const float (&dl)[NCHUNK],
                                             const qkv_t* __restrict__ K, const qkv_t* __restrict__ V, int64_t ld,
                                             const int* __restrict__ col, int beg, int end, int lane,
                                             float (&dq)[NCHUNK][VEC], float (&dpar)[NCHUNK], uint64_t pk) {
  if (rc.group == 4) bwd_row_walk_t<METRIC, VEC, NCHUNK, 4>(rc, go, ls, dl, K, V, ld, col, beg, end, lane, dq, dpar, pk);
  else if (rc.group == 8) bwd_row_walk_t<METRIC, VEC, NCHUNK, 8>(rc, go, ls, dl, K, V, ld, col, beg, end, lane, dq, dpar, pk);
  else bwd_row_walk_t<METRIC, VEC, NCHUNK, 0>(rc, go, ls, dl, K, V, ld, col, beg, end, lane, dq, dpar, pk);
}

template <int METRIC, int VEC, int NCHUNK, bool HEAVY>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, min_ctas(VEC, NCHUNK))
geo_attn_bwd_row_kernel(const qkv_t* __restrict__ Q, int64_t ldq, const qkv_t* __restrict__ K, const qkv_t* __restrict__ V, int64_t ld,
                        const int* __restrict__ rowptr, const int* __restrict__ col, int N, int heads, int D,
                        const float* __restrict__ metric_param, const float* __restrict__ ctx,
                        const float* __restrict__ lse, const float* __restrict__ dctx, float* __restrict__ dQ,
                        int64_t ldd, float* __restrict__ delta, float* __restrict__ dparam_rows) {
  constexpr int H = 32 * VEC * NCHUNK;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const L2Policies pol = make_l2_policies();
  // per-row prologue shared by both modes: q, dctx row, lse, delta = dctx . ctx
  auto prologue = [&](int row, RowCtx<METRIC, VEC, NCHUNK>& rc, float (&go)[NCHUNK][VEC], float (&ls)[NCHUNK],
                      float (&dl)[NCHUNK], bool write_delta) {
    init_row<METRIC, VEC, NCHUNK>(rc, Q + (int64_t)row * ldq, metric_param, D, lane, pol.stream);
    load_row<VEC, NCHUNK>(go, dctx + (int64_t)row * H, lane, pol.stream);
    float cx[NCHUNK][VEC];
    load_row<VEC, NCHUNK>(cx, ctx + (int64_t)row * H, lane, pol.stream);
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) t = fmaf(go[c][i], cx[c][i], t);
      dl[c] = group_sum(t, rc.group);
      ls[c] = __ldg(lse + (int64_t)row * heads + rc.head[c]);
      if (write_delta && (lane & (rc.group - 1)) == 0) delta[(int64_t)row * heads + rc.head[c]] = dl[c];
    }
  };
  auto store = [&](const RowCtx<METRIC, VEC, NCHUNK>& rc, int row, const float (&dq)[NCHUNK][VEC], const float (&dpar)[NCHUNK]) {
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      HintIO<VEC>::store(dQ + (int64_t)row * ldd + c * 32 * VEC + lane * VEC, dq[c], pol.stream);
      if (MetricTraits<METRIC>::kParam && dparam_rows != nullptr && (lane & (rc.group - 1)) == 0)
        dparam_rows[(int64_t)row * heads + rc.head[c]] = dpar[c];
    }
  };
  if (!HEAVY) {
    const int row = blockIdx.x * WARPS_PER_BLOCK + w;
    if (row >= N) return;
    const int beg = rowptr[row], end = rowptr[row + 1];
    if (end - beg > HEAVY_THRESH) return;
    RowCtx<METRIC, VEC, NCHUNK> rc;
    float go[NCHUNK][VEC], ls[NCHUNK], dl[NCHUNK], dq[NCHUNK][VEC], dpar[NCHUNK];
    prologue(row, rc, go, ls, dl, true);
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      dpar[c] = 0.f;
#pragma unroll
      for (int i = 0; i < VEC; ++i) dq[c][i] = 0.f;
    }
    bwd_row_walk<METRIC, VEC, NCHUNK>(rc, go, ls, dl, K, V, ld, col, beg, end, lane, dq, dpar, pol.keep);
    store(rc, row, dq, dpar);
  } else {
    __shared__ float sm_dq[WARPS_PER_BLOCK][NCHUNK][VEC][32];
    __shared__ float sm_dp[WARPS_PER_BLOCK][NCHUNK][32];
    for_each_heavy_row(rowptr, N, [&](int row) {
      const int beg = rowptr[row], end = rowptr[row + 1];
      int b, e;
      warp_slice(beg, end, w, b, e);
      RowCtx<METRIC, VEC, NCHUNK> rc;
      float go[NCHUNK][VEC], ls[NCHUNK], dl[NCHUNK], dq[NCHUNK][VEC], dpar[NCHUNK];
      prologue(row, rc, go, ls, dl, w == 0);
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        dpar[c] = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) dq[c][i] = 0.f;
      }
      bwd_row_walk<METRIC, VEC, NCHUNK>(rc, go, ls, dl, K, V, ld, col, b, e, lane, dq, dpar, pol.keep);
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c) {
        sm_dp[w][c][lane] = dpar[c];
#pragma unroll
        for (int i = 0; i < VEC; ++i) sm_dq[w][c][i][lane] = dq[c][i];
      }
      __syncthreads();
      if (w == 0) {
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
          dpar[c] = sm_dp[0][c][lane];
#pragma unroll
          for (int i = 0; i < VEC; ++i) dq[c][i] = sm_dq[0][c][i][lane];
          for (int k = 1; k < WARPS_PER_BLOCK; ++k) {
            dpar[c] += sm_dp[k][c][lane];
#pragma unroll
            for (int i = 0; i < VEC; ++i) dq[c][i] += sm_dq[k][c][i][lane];
          }
        }
        store(rc, row, dq, dpar);
      }
    });
  }
}

// ---- column pass over the transposed CSR: for source node j, dV[j] = sum_e a_e dctx[row_e],
// dK[j] = sum_e ds_e * dscore/dk.  Gathers Q[row], dctx[row], lse[row,h], delta[row,h].
template <int METRIC, int VEC, int NCHUNK, int GROUP>
__device__ __forceinline__ void bwd_col_walk_t(const float (&k)[NCHUNK][VEC], const float (&v)[NCHUNK][VEC],
                                             const float (&par)[NCHUNK], const int (&head)[NCHUNK], int group,
                                             const qkv_t* __restrict__ Q, int64_t ldq, const float* __restrict__ dctx,
                                             const float* __restrict__ lse, const float* __restrict__ delta, int heads,
                                             const int* __restrict__ row_t, int beg, int end, int lane,
                                             float (&dk)[NCHUNK][VEC], float (&dv)[NCHUNK][VEC], uint64_t pk) {
  constexpr int U = (VEC * NCHUNK >= 8) ? 2 : 4;
  constexpr int H = 32 * VEC * NCHUNK;
  for (int base = beg; base < end; base += 32) {
    const int n = min(32, end - base);
    const int myrow = lane < n ? __ldg(row_t + base + lane) : 0;
    for (int j = 0; j < n; j += U) {
      float qq[U][NCHUNK][VEC], gg[U][NCHUNK][VEC], ls[U][NCHUNK], dl[U][NCHUNK];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int r = __shfl_sync(FULL_MASK, myrow, min(j + u, n - 1));
        load_row<VEC, NCHUNK>(qq[u], Q + (int64_t)r * ldq, lane, pk);
        load_row<VEC, NCHUNK>(gg[u], dctx + (int64_t)r * H, lane, pk);
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
          ls[u][c] = __ldg(lse + (int64_t)r * heads + head[c]);
          dl[u][c] = __ldg(delta + (int64_t)r * heads + head[c]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j + u < n) {
#pragma unroll
          for (int c = 0; c < NCHUNK; ++c) {
            float qn = 0.f;
            if (MetricTraits<METRIC>::kCos) {
              float t = 0.f;
#pragma unroll
              for (int i = 0; i < VEC; ++i) t = fmaf(qq[u][c][i], qq[u][c][i], t);
              t = sqrtf(group_sum_t<GROUP>(t, group));
              qn = t == 0.f ? 1e-8f : t;
            }
            float p1, p2;
            const float s = score_from_vectors<METRIC, VEC, GROUP>(qq[u][c], k[c], group, par[c], qn, p1, p2);
            const float a = fast_exp(s - ls[u][c]);
            float dp = 0.f;
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              dv[c][i] = fmaf(a, gg[u][c][i], dv[c][i]);
              dp = fmaf(gg[u][c][i], v[c][i], dp);
            }
            dp = group_sum_t<GROUP>(dp, group);
            const float ds = a * (dp - dl[u][c]);
            accum_score_grad<METRIC, VEC, false>(ds, qq[u][c], k[c], s, p1, p2, par[c], qn, dk[c]);
          }
        }
      }
    }
  }
}

template <int METRIC, int VEC, int NCHUNK>
__device__ __forceinline__ void bwd_col_walk(const float (&k)[NCHUNK][VEC], const float (&v)[NCHUNK][VEC],
                                             const float (&par)[NCHUNK], const int (&head)[NCHUNK], int group,
                                             const qkv_t* __restrict__ Q, int64_t ldq, const float* __restrict__ dctx,
                                             const float* __restrict__ lse, const float* __restrict__ delta, int heads,
                                             const int* __restrict__ row_t, int beg, int end, int lane,
                                             float (&dk)[NCHUNK][VEC], float (&dv)[NCHUNK][VEC], uint64_t pk) {
  if (group == 4) bwd_col_walk_t<METRIC, VEC, NCHUNK, 4>(k, v, par, head, group, Q, ldq, dctx, lse, delta, heads, row_t, beg, end, lane, dk, dv, pk);
  else if (group == 8) bwd_col_walk_t<METRIC, VEC, NCHUNK, 8>(k, v, par, head, group, Q, ldq, dctx, lse, delta, heads, row_t, beg, end, lane, dk, dv, pk);
  else bwd_col_walk_t<METRIC, VEC, NCHUNK, 0>(k, v, par, head, group, Q, ldq, dctx, lse, delta, heads, row_t, beg, end, lane, dk, dv, pk);
}

// the column pass keeps two gathered rows per entry plus dK and dV: at 4 CTAs per SM (64 registers) it spills
constexpr int min_ctas_col(int vec, int nchunk) { return vec * nchunk >= 16 ? 2 : 3; }   // H = 256 / 512 as before
template <int METRIC, int VEC, int NCHUNK, bool HEAVY>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, min_ctas_col(VEC, NCHUNK))
geo_attn_bwd_col_kernel(const qkv_t* __restrict__ Q, int64_t ldq, const qkv_t* __restrict__ K, const qkv_t* __restrict__ V, int64_t ld,
                        const int* __restrict__ rowptr_t, const int* __restrict__ row_t, int N, int heads, int D,
                        const float* __restrict__ metric_param, const float* __restrict__ lse,
                        const float* __restrict__ delta, const float* __restrict__ dctx, float* __restrict__ dK,
                        float* __restrict__ dV, int64_t ldd) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const L2Policies pol = make_l2_policies();
  const int group = D / VEC;
  auto load_node = [&](int node, float (&k)[NCHUNK][VEC], float (&v)[NCHUNK][VEC], float (&par)[NCHUNK], int (&head)[NCHUNK],
                       float (&dk)[NCHUNK][VEC], float (&dv)[NCHUNK][VEC]) {
    load_row<VEC, NCHUNK>(k, K + (int64_t)node * ld, lane, pol.stream);
    load_row<VEC, NCHUNK>(v, V + (int64_t)node * ld, lane, pol.stream);
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      head[c] = (c * 32 * VEC + lane * VEC) / D;
      par[c] = head_param<METRIC>(metric_param, head[c], D);
#pragma unroll
      for (int i = 0; i < VEC; ++i) { dk[c][i] = 0.f; dv[c][i] = 0.f; }
    }
  };
  auto store = [&](int node, const float (&dk)[NCHUNK][VEC], const float (&dv)[NCHUNK][VEC]) {
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
      HintIO<VEC>::store(dK + (int64_t)node * ldd + c * 32 * VEC + lane * VEC, dk[c], pol.stream);
      HintIO<VEC>::store(dV + (int64_t)node * ldd + c * 32 * VEC + lane * VEC, dv[c], pol.stream);
    }
  };
  if (!HEAVY) {
    const int node = blockIdx.x * WARPS_PER_BLOCK + w;
    if (node >= N) return;
    const int beg = rowptr_t[node], end = rowptr_t[node + 1];
    if (end - beg > HEAVY_THRESH) return;
    float k[NCHUNK][VEC], v[NCHUNK][VEC], dk[NCHUNK][VEC], dv[NCHUNK][VEC], par[NCHUNK];
    int head[NCHUNK];
    load_node(node, k, v, par, head, dk, dv);
    bwd_col_walk<METRIC, VEC, NCHUNK>(k, v, par, head, group, Q, ldq, dctx, lse, delta, heads, row_t, beg, end, lane, dk, dv, pol.keep);
    store(node, dk, dv);
  } else {
    __shared__ float sm_dk[WARPS_PER_BLOCK][NCHUNK][VEC][32], sm_dv[WARPS_PER_BLOCK][NCHUNK][VEC][32];
    for_each_heavy_row(rowptr_t, N, [&](int node) {
      const int beg = rowptr_t[node], end = rowptr_t[node + 1];
      int b, e;
      warp_slice(beg, end, w, b, e);
      float k[NCHUNK][VEC], v[NCHUNK][VEC], dk[NCHUNK][VEC], dv[NCHUNK][VEC], par[NCHUNK];
      int head[NCHUNK];
      load_node(node, k, v, par, head, dk, dv);
      bwd_col_walk<METRIC, VEC, NCHUNK>(k, v, par, head, group, Q, ldq, dctx, lse, delta, heads, row_t, b, e, lane, dk, dv, pol.keep);
#pragma unroll
      for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
        for (int i = 0; i < VEC; ++i) { sm_dk[w][c][i][lane] = dk[c][i]; sm_dv[w][c][i][lane] = dv[c][i]; }
      __syncthreads();
      if (w == 0) {
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            float a = sm_dk[0][c][i][lane], bb = sm_dv[0][c][i][lane];
            for (int kk = 1; kk < WARPS_PER_BLOCK; ++kk) { a += sm_dk[kk][c][i][lane]; bb += sm_dv[kk][c][i][lane]; }
            dk[c][i] = a; dv[c][i] = bb;
          }
        store(node, dk, dv);
      }
    });
  }
}

// dparam[h] = sum over rows of dparam_rows[row,h]; one block per head, fixed-order tree.
__global__ void reduce_rows_per_head(const float* __restrict__ rows, int64_t n, int heads, float* __restrict__ out) {
  __shared__ float sm[256];
  const int h = blockIdx.x;
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 256) s += rows[i * heads + h];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[h] = sm[0];
}

struct Shape { int vec, nchunk; };
bool pick_shape(int H, int heads, Shape* s) {
  if (H <= 0 || heads <= 0 || H % heads) return false;
  int vec, nchunk;
  if (H == 32) { vec = 1; nchunk = 1; }
  else if (H == 64) { vec = 2; nchunk = 1; }
  else if (H == 128) { vec = 4; nchunk = 1; }
  else if (H == 256) { vec = 4; nchunk = 2; }
  else if (H == 512) { vec = 4; nchunk = 4; }
  else return false;
  int D = H / heads;
  if (D % vec) return false;
  int group = D / vec;
  if (group < 1 || group > 32 || (group & (group - 1))) return false;
  s->vec = vec; s->nchunk = nchunk;
  return true;
}

#define DISPATCH_SHAPE(METRIC, FN, ...)                                                     \
  if (sh.vec == 1 && sh.nchunk == 1) { FN<METRIC, 1, 1, false> GRID_L __VA_ARGS__; FN<METRIC, 1, 1, true> GRID_H __VA_ARGS__; }      \
  else if (sh.vec == 2 && sh.nchunk == 1) { FN<METRIC, 2, 1, false> GRID_L __VA_ARGS__; FN<METRIC, 2, 1, true> GRID_H __VA_ARGS__; } \
  else if (sh.vec == 4 && sh.nchunk == 1) { FN<METRIC, 4, 1, false> GRID_L __VA_ARGS__; FN<METRIC, 4, 1, true> GRID_H __VA_ARGS__; } \
  else if (sh.vec == 4 && sh.nchunk == 2) { FN<METRIC, 4, 2, false> GRID_L __VA_ARGS__; FN<METRIC, 4, 2, true> GRID_H __VA_ARGS__; } \
  else { FN<METRIC, 4, 4, false> GRID_L __VA_ARGS__; FN<METRIC, 4, 4, true> GRID_H __VA_ARGS__; }

// GRID_L: one warp per row; GRID_H: the heavy-row launch (fixed grid, CTA per heavy row)
#define GRID_L <<<grid, block, 0, st>>>
#define GRID_H <<<dim3(HEAVY_GRID), block, 0, st>>>
#define DISPATCH_METRIC(FN, ...)                                                            \
  switch (metric) {                                                                         \
    case TAGAN_METRIC_SCALED_DOT: { DISPATCH_SHAPE(TAGAN_METRIC_SCALED_DOT, FN, __VA_ARGS__) } break;   \
    case TAGAN_METRIC_DOT: { DISPATCH_SHAPE(TAGAN_METRIC_DOT, FN, __VA_ARGS__) } break;                 \
    case TAGAN_METRIC_COSINE_SIM: { DISPATCH_SHAPE(TAGAN_METRIC_COSINE_SIM, FN, __VA_ARGS__) } break;   \
    case TAGAN_METRIC_EUCLIDEAN: { DISPATCH_SHAPE(TAGAN_METRIC_EUCLIDEAN, FN, __VA_ARGS__) } break;     \
    case TAGAN_METRIC_SQ_EUCLIDEAN: { DISPATCH_SHAPE(TAGAN_METRIC_SQ_EUCLIDEAN, FN, __VA_ARGS__) } break; \
    case TAGAN_METRIC_MANHATTAN: { DISPATCH_SHAPE(TAGAN_METRIC_MANHATTAN, FN, __VA_ARGS__) } break;     \
    case TAGAN_METRIC_COSINE_DIST: { DISPATCH_SHAPE(TAGAN_METRIC_COSINE_DIST, FN, __VA_ARGS__) } break; \
    case TAGAN_METRIC_GAUSSIAN: { DISPATCH_SHAPE(TAGAN_METRIC_GAUSSIAN, FN, __VA_ARGS__) } break;       \
    case TAGAN_METRIC_RBF: { DISPATCH_SHAPE(TAGAN_METRIC_RBF, FN, __VA_ARGS__) } break;                 \
    default: return TAGAN_E_INVALID;                                                        \
  }

}  // namespace

/* 1 (default): gathered rows evict_last, streamed rows / results evict_first; 0: no L2 hints.  Synchronous (profiling knob). */
TAGAN_API int GEO_NAME(tagan_geo_attn_set_l2_policy)(int32_t mode) {
  const int m = mode ? 1 : 0;
  return (int)cudaMemcpyToSymbol(c_l2_mode, &m, sizeof(int));
}

TAGAN_API int GEO_NAME(tagan_geo_attn_fwd_part)(const qkv_api_t* Q_, int64_t ldq, const qkv_api_t* K_, const qkv_api_t* V_, int64_t ldkv,
                                      const int32_t* rowptr, const int32_t* col, int32_t n_rows, int32_t H,
                                      int32_t heads, int32_t metric, const float* metric_param, float* ctx, float* lse,
                                      float* attn, tagan_stream_t stream) {
  const qkv_t* Q = reinterpret_cast<const qkv_t*>(Q_);
  const qkv_t* K = reinterpret_cast<const qkv_t*>(K_);
  const qkv_t* V = reinterpret_cast<const qkv_t*>(V_);
  if (!Q || !K || !V || !rowptr || !col || !ctx || !lse || n_rows < 0 || ldq < H || ldkv < H) return TAGAN_E_INVALID;
  Shape sh;
  if (!pick_shape(H, heads, &sh) || (ldq % sh.vec) || (ldkv % sh.vec)) return TAGAN_E_UNSUPPORTED;
  if (n_rows == 0) return 0;
  const int N = n_rows;
  const int64_t ld = ldkv;
  const int D = H / heads;
  dim3 grid((N + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK), block(WARPS_PER_BLOCK * 32);
  cudaStream_t st = as_stream(stream);
  DISPATCH_METRIC(geo_attn_fwd_kernel, (Q, ldq, K, V, ld, rowptr, col, N, heads, D, metric_param, ctx, lse, attn))
  return tagan_launch_status();
}

TAGAN_API int GEO_NAME(tagan_geo_attn_fwd)(const qkv_api_t* Q, const qkv_api_t* K, const qkv_api_t* V, int64_t ld, const int32_t* rowptr,
                                 const int32_t* col, int32_t N, int32_t H, int32_t heads, int32_t metric,
                                 const float* metric_param, float* ctx, float* lse, float* attn,
                                 tagan_stream_t stream) {
  return GEO_NAME(tagan_geo_attn_fwd_part)(Q, ld, K, V, ld, rowptr, col, N, H, heads, metric, metric_param, ctx, lse, attn, stream);
}

TAGAN_API int GEO_NAME(tagan_geo_attn_bwd_part)(const qkv_api_t* Q_, int64_t ldq, const qkv_api_t* K_, const qkv_api_t* V_, int64_t ldkv,
                                      const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t,
                                      const int32_t* row_t, int32_t n_rows, int32_t n_src, int32_t H, int32_t heads,
                                      int32_t metric, const float* metric_param, const float* ctx, const float* lse,
                                      const float* dctx, float* dQ, int64_t lddq, float* dK, float* dV, int64_t lddkv,
                                      float* delta_ws, float* dparam_ws, float* dparam, tagan_stream_t stream) {
  const qkv_t* Q = reinterpret_cast<const qkv_t*>(Q_);
  const qkv_t* K = reinterpret_cast<const qkv_t*>(K_);
  const qkv_t* V = reinterpret_cast<const qkv_t*>(V_);
  if (!Q || !K || !V || !rowptr || !col || !rowptr_t || !row_t || !ctx || !lse || !dctx || !dQ || !dK || !dV ||
      !delta_ws || n_rows < 0 || n_src < 0 || ldq < H || ldkv < H || lddq < H || lddkv < H)
    return TAGAN_E_INVALID;
  Shape sh;
  if (!pick_shape(H, heads, &sh) || (ldq % sh.vec) || (ldkv % sh.vec) || (lddq % sh.vec) || (lddkv % sh.vec))
    return TAGAN_E_UNSUPPORTED;
  const bool want_dparam = metric_param != nullptr && (metric == TAGAN_METRIC_GAUSSIAN || metric == TAGAN_METRIC_RBF);
  if (want_dparam && (!dparam_ws || !dparam)) return TAGAN_E_INVALID;
  const int D = H / heads;
  const int64_t ld = ldkv;
  cudaStream_t st = as_stream(stream);
  float* dpr = want_dparam ? dparam_ws : nullptr;
  dim3 block(WARPS_PER_BLOCK * 32);
  if (n_rows > 0) {
    const int N = n_rows;
    const int64_t ldd = lddq;
    dim3 grid((N + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    DISPATCH_METRIC(geo_attn_bwd_row_kernel, (Q, ldq, K, V, ld, rowptr, col, N, heads, D, metric_param, ctx, lse, dctx, dQ, ldd, delta_ws, dpr))
  }
  if (n_src > 0) {
    const int N = n_src;
    const int64_t ldd = lddkv;
    dim3 grid((N + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    DISPATCH_METRIC(geo_attn_bwd_col_kernel, (Q, ldq, K, V, ld, rowptr_t, row_t, N, heads, D, metric_param, lse, delta_ws, dctx, dK, dV, ldd))
  }
  if (want_dparam) reduce_rows_per_head<<<heads, 256, 0, st>>>(dparam_ws, n_rows, heads, dparam);
  return tagan_launch_status();
}

TAGAN_API int GEO_NAME(tagan_geo_attn_bwd)(const qkv_api_t* Q, const qkv_api_t* K, const qkv_api_t* V, int64_t ld, const int32_t* rowptr,
                                 const int32_t* col, const int32_t* rowptr_t, const int32_t* row_t, int32_t N,
                                 int32_t H, int32_t heads, int32_t metric, const float* metric_param,
                                 const float* ctx, const float* lse, const float* dctx, float* dQ, float* dK,
                                 float* dV, int64_t ldd, float* delta_ws, float* dparam_ws, float* dparam,
                                 tagan_stream_t stream) {
  return GEO_NAME(tagan_geo_attn_bwd_part)(Q, ld, K, V, ld, rowptr, col, rowptr_t, row_t, N, N, H, heads, metric, metric_param,
                                 ctx, lse, dctx, dQ, ldd, dK, dV, ldd, delta_ws, dparam_ws, dparam, stream);
}
