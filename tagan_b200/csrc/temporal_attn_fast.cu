// Temporal attention, fast path for T <= 32 (see temporal_attn.cu for the generic kernels and the C ABI).
#include "temporal_attn_shared.cuh"

namespace {

// ---------------------------------------------------------------------------------------
// fast path: T <= 32, every head of a node resident in the CTA at once, bias shared by all nodes.
// A lane keeps the same (head, row) for the whole grid-stride loop over nodes, so its bias row lives in
// registers (no global loads in the inner loop), scores of a row are kept in registers (two-pass softmax:
// one exp2 per key, no running rescale), and in the backward pass the bias gradient is accumulated in
// registers over all nodes of the CTA and written once.
// ---------------------------------------------------------------------------------------
constexpr float LOG2E = 1.4426950408889634f;

template <int D, int TP>
__global__ void __launch_bounds__(MAX_WARPS * 32)
tattn_fwd_fast_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, int64_t ld,
                      int64_t B, int T, int heads, int64_t rsb, int64_t rst, const float* __restrict__ bias,
                      MaskSpec ms, float* __restrict__ ctx, float* __restrict__ lse, float* __restrict__ attn) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PPW = 32 / TP;
  const int H = heads * D;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / TP, li = lane - sub * TP;
  const int slot_floats = 2 * T * D + ((T + 3) & ~3);
  float* slot = smem + (size_t)(w * PPW + sub) * slot_floats;
  float* Ks = slot;
  float* Vs = slot + T * D;
  float* ts_s = slot + 2 * T * D;
  const float scale = 1.f / sqrtf((float)D);
  const bool causal = (ms.flags & 1) || ((ms.flags & 4) && ms.allones_flag && *ms.allones_flag != 0);
  const int hd = w * PPW + sub;
  const bool pv = hd < heads;
  const int i = li;
  const bool rv = pv && i < T;
  float brow[TP];
#pragma unroll
  for (int j = 0; j < TP; ++j) brow[j] = (bias && rv && j < T) ? bias[((int64_t)hd * T + i) * T + j] : 0.f;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    __syncwarp();
    if (pv) {
      const int64_t base = b * rsb * ld + (int64_t)hd * D;
      load_tile<D>(Ks, K + base, rst * ld, T, li, TP);
      load_tile<D>(Vs, V + base, rst * ld, T, li, TP);
      if (ms.ts) for (int t = li; t < T; t += TP) ts_s[t] = ms.ts[b * T + t];
    }
    float q[D];
    if (rv) {
      const float* qp = Q + (b * rsb + i * rst) * ld + (int64_t)hd * D;
#pragma unroll
      for (int c = 0; c < D; c += 4) {
        float4 t4 = __ldg(reinterpret_cast<const float4*>(qp + c));
        q[c] = t4.x; q[c + 1] = t4.y; q[c + 2] = t4.z; q[c + 3] = t4.w;
      }
    } else {
#pragma unroll
      for (int c = 0; c < D; ++c) q[c] = 0.f;
    }
    __syncwarp();
    const uint8_t* mbase = nullptr;
    if (ms.mask && pv)
      mbase = ms.mask + ((int64_t)(ms.mask_b > 1 ? b : 0) * ms.mask_h + (ms.mask_h > 1 ? hd : 0)) * T * T;
    float sc[TP];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < TP; ++j) {
      float v = -INFINITY;
      if (j < T) {
        const float d = dot_smem<D>(q, Ks + j * D) * scale + brow[j];
        if (rv && key_valid(ms, causal, ms.ts ? ts_s : nullptr, mbase, T, i, j)) v = d;
      }
      sc[j] = v;
      m = fmaxf(m, v);
    }
    float l = 0.f, acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) acc[c] = 0.f;
#pragma unroll
    for (int j = 0; j < TP; ++j) {
      if (j < T) {
        const float p = sc[j] > -INFINITY ? exp2f((sc[j] - m) * LOG2E) : 0.f;
        sc[j] = p;
        l += p;
        const float* vr = Vs + j * D;
#pragma unroll
        for (int c = 0; c < D; c += 4) {
          float4 v4 = *reinterpret_cast<const float4*>(vr + c);
          acc[c] = fmaf(p, v4.x, acc[c]); acc[c + 1] = fmaf(p, v4.y, acc[c + 1]);
          acc[c + 2] = fmaf(p, v4.z, acc[c + 2]); acc[c + 3] = fmaf(p, v4.w, acc[c + 3]);
        }
      }
    }
    if (rv) {
      const float inv = 1.f / l;
      float* op = ctx + (b * rsb + i * rst) * (int64_t)H + (int64_t)hd * D;
#pragma unroll
      for (int c = 0; c < D; c += 4)
        *reinterpret_cast<float4*>(op + c) = make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
      lse[(b * heads + hd) * T + i] = m + logf(l);
      if (attn != nullptr) {
        float* ap = attn + ((b * heads + hd) * T + i) * (int64_t)T;
#pragma unroll
        for (int j = 0; j < TP; ++j)
          if (j < T) ap[j] = sc[j] * inv;
      }
    }
  }
}

// Backward: phase 1 (lane = query row i) recomputes the probabilities from the saved log-sum-exp, forms
// dS = P o (dP - delta), accumulates dQ in registers and parks P and dS*scale in shared memory (pitch TP+1, so a
// lane's row-wise stores and the column-wise loads of phase 2 are both conflict-free); phase 2 (lane = key row j)
// reads its column of P / dS back and accumulates dV = P^T dO and dK = dS^T Q -- no second exp or dot product.
template <int D, int TP>
__global__ void __launch_bounds__(MAX_WARPS * 32, (D <= 16 && D * TP <= 256) ? 2 : 1)
tattn_bwd_fast_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, int64_t ld,
                      int64_t B, int T, int heads, int64_t rsb, int64_t rst, const float* __restrict__ bias, MaskSpec ms,
                      const float* __restrict__ ctx, const float* __restrict__ lse, const float* __restrict__ dctx,
                      float* __restrict__ dQ, float* __restrict__ dK, float* __restrict__ dV, int64_t ldd,
                      float* __restrict__ dbias_partial /* [grid, h, T, T] or null */) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PPW = 32 / TP;
  constexpr int PT = TP + 1;
  const int H = heads * D;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / TP, li = lane - sub * TP;
  const int tpad = (T + 3) & ~3;
  const int slot_floats = tattn_bwd_fast_slot_floats(T, D, TP);
  float* slot = smem + (size_t)(w * PPW + sub) * slot_floats;
  float* Qs = slot;
  float* Ks = Qs + T * D;
  float* Vs = Ks + T * D;
  float* Gs = Vs + T * D;
  float* Ps = Gs + T * D;            // [T][PT] probabilities
  float* Ss = Ps + TP * PT;          // [T][PT] dS * scale
  float* ts_s = Ss + TP * PT;
  (void)tpad;
  const float scale = 1.f / sqrtf((float)D);
  const bool causal = (ms.flags & 1) || ((ms.flags & 4) && ms.allones_flag && *ms.allones_flag != 0);
  const int hd = w * PPW + sub;
  const bool pv = hd < heads;
  const bool rv = pv && li < T;
  float brow[TP], dbrow[TP];
#pragma unroll
  for (int j = 0; j < TP; ++j) {
    brow[j] = (bias && rv && j < T) ? bias[((int64_t)hd * T + li) * T + j] : 0.f;   // lane = query i, over keys j
    dbrow[j] = 0.f;
  }
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    __syncwarp();
    if (pv) {
      const int64_t base = b * rsb * ld + (int64_t)hd * D;
      load_tile<D>(Qs, Q + base, rst * ld, T, li, TP);
      load_tile<D>(Ks, K + base, rst * ld, T, li, TP);
      load_tile<D>(Vs, V + base, rst * ld, T, li, TP);
      load_tile<D>(Gs, dctx + b * rsb * (int64_t)H + (int64_t)hd * D, rst * H, T, li, TP);
      if (ms.ts) for (int t = li; t < T; t += TP) ts_s[t] = ms.ts[b * T + t];
    }
    float dl = 0.f, ls2 = 0.f;
    if (rv) {
      const float* cp = ctx + (b * rsb + li * rst) * (int64_t)H + (int64_t)hd * D;
      const float* gp = dctx + (b * rsb + li * rst) * (int64_t)H + (int64_t)hd * D;
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int c = 0; c < D; c += 4) {
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(cp + c));
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gp + c));
        d0 = fmaf(g4.x, c4.x, d0); d1 = fmaf(g4.y, c4.y, d1); d0 = fmaf(g4.z, c4.z, d0); d1 = fmaf(g4.w, c4.w, d1);
      }
      dl = d0 + d1;
      ls2 = lse[(b * heads + hd) * T + li] * LOG2E;
    }
    __syncwarp();
    const uint8_t* mbase = nullptr;
    if (ms.mask && pv)
      mbase = ms.mask + ((int64_t)(ms.mask_b > 1 ? b : 0) * ms.mask_h + (ms.mask_h > 1 ? hd : 0)) * T * T;
    {   // ---- phase 1: lane = query row i
      const int i = li;
      float q[D], g[D], dq[D];
#pragma unroll
      for (int c = 0; c < D; c += 4) {
        const float4 q4 = *reinterpret_cast<const float4*>(Qs + (rv ? i : 0) * D + c);
        const float4 g4 = *reinterpret_cast<const float4*>(Gs + (rv ? i : 0) * D + c);
        q[c] = q4.x; q[c + 1] = q4.y; q[c + 2] = q4.z; q[c + 3] = q4.w;
        g[c] = g4.x; g[c + 1] = g4.y; g[c + 2] = g4.z; g[c + 3] = g4.w;
        dq[c] = dq[c + 1] = dq[c + 2] = dq[c + 3] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < TP; ++j) {
        if (j < T) {
          const float* kr = Ks + j * D;
          const float* vr = Vs + j * D;
          float s0 = 0.f, s1 = 0.f, p0 = 0.f, p1 = 0.f;
          float kreg[D];
#pragma unroll
          for (int c = 0; c < D; c += 4) {
            const float4 k4 = *reinterpret_cast<const float4*>(kr + c);
            const float4 v4 = *reinterpret_cast<const float4*>(vr + c);
            kreg[c] = k4.x; kreg[c + 1] = k4.y; kreg[c + 2] = k4.z; kreg[c + 3] = k4.w;
            s0 = fmaf(q[c], k4.x, s0); s1 = fmaf(q[c + 1], k4.y, s1); s0 = fmaf(q[c + 2], k4.z, s0); s1 = fmaf(q[c + 3], k4.w, s1);
            p0 = fmaf(g[c], v4.x, p0); p1 = fmaf(g[c + 1], v4.y, p1); p0 = fmaf(g[c + 2], v4.z, p0); p1 = fmaf(g[c + 3], v4.w, p1);
          }
          const float sv = (s0 + s1) * scale + brow[j];
          const bool kv = rv && key_valid(ms, causal, ms.ts ? ts_s : nullptr, mbase, T, i, j) && sv > -INFINITY;
          const float pr = kv ? exp2f(sv * LOG2E - ls2) : 0.f;
          const float ds = pr * ((p0 + p1) - dl);
          dbrow[j] += ds;
          const float dss = ds * scale;
          if (rv) { Ps[i * PT + j] = pr; Ss[i * PT + j] = dss; }
#pragma unroll
          for (int c = 0; c < D; ++c) dq[c] = fmaf(dss, kreg[c], dq[c]);
        }
      }
      if (rv) {
        float* op = dQ + (b * rsb + i * rst) * ldd + (int64_t)hd * D;
#pragma unroll
        for (int c = 0; c < D; c += 4) *reinterpret_cast<float4*>(op + c) = make_float4(dq[c], dq[c + 1], dq[c + 2], dq[c + 3]);
      }
    }
    __syncwarp();
    {   // ---- phase 2: lane = key row j
      const int j = li;
      float dk[D], dv[D];
#pragma unroll
      for (int c = 0; c < D; ++c) { dk[c] = 0.f; dv[c] = 0.f; }
#pragma unroll
      for (int i = 0; i < TP; ++i) {
        if (i < T) {
          const float pr = rv ? Ps[i * PT + j] : 0.f;
          const float dss = rv ? Ss[i * PT + j] : 0.f;
          const float* gr = Gs + i * D;
          const float* qr = Qs + i * D;
#pragma unroll
          for (int c = 0; c < D; c += 4) {
            const float4 g4 = *reinterpret_cast<const float4*>(gr + c);
            const float4 q4 = *reinterpret_cast<const float4*>(qr + c);
            dv[c] = fmaf(pr, g4.x, dv[c]); dv[c + 1] = fmaf(pr, g4.y, dv[c + 1]);
            dv[c + 2] = fmaf(pr, g4.z, dv[c + 2]); dv[c + 3] = fmaf(pr, g4.w, dv[c + 3]);
            dk[c] = fmaf(dss, q4.x, dk[c]); dk[c + 1] = fmaf(dss, q4.y, dk[c + 1]);
            dk[c + 2] = fmaf(dss, q4.z, dk[c + 2]); dk[c + 3] = fmaf(dss, q4.w, dk[c + 3]);
          }
        }
      }
      if (rv) {
        float* okp = dK + (b * rsb + j * rst) * ldd + (int64_t)hd * D;
        float* ovp = dV + (b * rsb + j * rst) * ldd + (int64_t)hd * D;
#pragma unroll
        for (int c = 0; c < D; c += 4) {
          *reinterpret_cast<float4*>(okp + c) = make_float4(dk[c], dk[c + 1], dk[c + 2], dk[c + 3]);
          *reinterpret_cast<float4*>(ovp + c) = make_float4(dv[c], dv[c + 1], dv[c + 2], dv[c + 3]);
        }
      }
    }
  }
  if (dbias_partial != nullptr && rv) {
    float* o = dbias_partial + ((int64_t)blockIdx.x * heads + hd) * T * T + (int64_t)li * T;
#pragma unroll
    for (int j = 0; j < TP; ++j)
      if (j < T) o[j] = dbrow[j];
  }
}


// ---------------------------------------------------------------------------------------
// Tensor-core formulation for 8 < T <= 16 (one 16x16 score tile per (node, head)), D = 16 or 32.
// The lane-per-row forward kernel above spends its time on broadcast LDS.128 of K/V/Q/dO rows (L1/TEX 66-80 % busy, issue
// slots ~40 %); here the five small products run on mma.sync.m16n8k8 (TF32 hi/lo split, 3 MMAs per product term, so
// fp32-accurate like the projections), every operand element is loaded from shared memory once per warp as a
// fragment, and the softmax / dS algebra happens on the accumulator fragments.  P feeds the second product straight
// from its accumulator registers (k-slot t <-> key 2t, k-slot t+4 <-> key 2t+1, applied to both operands).
// One warp per head (heads <= 8), persistent CTAs over nodes, bias fragment in registers.
// Forward only: the same formulation of the backward pass (five products, P / dS transposed through shared memory)
// was built and measured at 3.5 ms against 2.9 ms for the lane-per-row kernel above (its dependent mma.sync chains
// at 16 warps/SM are latency-bound), so the backward keeps the lane-per-row kernel.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  const float r = x - __uint_as_float(h);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
  hi = h; lo = l;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
struct FragA { uint32_t hi[4], lo[4]; };
struct FragB { uint32_t hi[2], lo[2]; };
__device__ __forceinline__ void mma3(float (&c)[4], const FragA& a, const FragB& b) {
  mma_tf32(c, a.lo, b.hi);
  mma_tf32(c, a.hi, b.lo);
  mma_tf32(c, a.hi, b.hi);
}
__device__ __forceinline__ FragA make_a(float a0, float a1, float a2, float a3) {
  FragA f;
  split_tf32(a0, f.hi[0], f.lo[0]); split_tf32(a1, f.hi[1], f.lo[1]);
  split_tf32(a2, f.hi[2], f.lo[2]); split_tf32(a3, f.hi[3], f.lo[3]);
  return f;
}
__device__ __forceinline__ FragB make_b(float b0, float b1) {
  FragB f;
  split_tf32(b0, f.hi[0], f.lo[0]); split_tf32(b1, f.hi[1], f.lo[1]);
  return f;
}
// A fragment of a row-major [16][pitch] tile, k-step ks (columns 8ks..8ks+7): rows g / g+8, columns t / t+4
__device__ __forceinline__ FragA load_a_rows(const float* tile, int pitch, int ks, int g, int t) {
  const float* p0 = tile + g * pitch + 8 * ks + t;
  return make_a(p0[0], p0[8 * pitch], p0[4], p0[8 * pitch + 4]);
}
// B fragment B(k, n) = tile[8nt + n][8ks + k]   (the "K^T" pattern: n = g, k = t / t+4)
__device__ __forceinline__ FragB load_b_kt(const float* tile, int pitch, int nt, int ks, int g, int t) {
  const float* p0 = tile + (8 * nt + g) * pitch + 8 * ks + t;
  return make_b(p0[0], p0[4]);
}
// B fragment B(kslot, n) = tile[row(kslot)][8ntc + n] with row(t) = r0, row(t+4) = r1
__device__ __forceinline__ FragB load_b_rows(const float* tile, int pitch, int r0, int r1, int ntc, int g) {
  return make_b(tile[r0 * pitch + 8 * ntc + g], tile[r1 * pitch + 8 * ntc + g]);
}
// cooperative [T][D] -> smem [16][pitch] by a full warp, rows >= T zero-filled
template <int D>
__device__ __forceinline__ void load_tile_mma(float* dst, const float* src, int64_t ld, int T, int lane) {
  constexpr int C4 = D / 4, PITCH = D + 4;
  for (int idx = lane; idx < 16 * C4; idx += 32) {
    const int r = idx / C4, c = idx - r * C4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < T) v = __ldg(reinterpret_cast<const float4*>(src + (int64_t)r * ld + c * 4));
    *reinterpret_cast<float4*>(dst + r * PITCH + c * 4) = v;
  }
}

template <int D>
__global__ void __launch_bounds__(256)
tattn_fwd_mma_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, int64_t ld,
                     int64_t B, int T, int heads, int64_t rsb, int64_t rst, const float* __restrict__ bias, MaskSpec ms,
                     float* __restrict__ ctx, float* __restrict__ lse, float* __restrict__ attn) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PITCH = D + 4, KS = D / 8, TILE = 16 * PITCH;
  const int hd = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  if (hd >= heads) return;
  const int H = heads * D;
  float* Qs = smem + (size_t)hd * (3 * TILE + 16);
  float* Ks = Qs + TILE;
  float* Vs = Ks + TILE;
  float* ts_s = Vs + TILE;
  const float scale = 1.f / sqrtf((float)D);
  const bool causal = (ms.flags & 1) || ((ms.flags & 4) && ms.allones_flag && *ms.allones_flag != 0);
  // element r of n-tile nt: row i = g + 8*(r>>1), key j = 8nt + 2t + (r&1)
  float bs[2][4];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = g + 8 * (r >> 1), j = 8 * nt + 2 * t + (r & 1);
      bs[nt][r] = (bias && i < T && j < T) ? bias[((int64_t)hd * T + i) * T + j] : 0.f;
    }
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    __syncwarp();
    const int64_t base = b * rsb * ld + (int64_t)hd * D;
    load_tile_mma<D>(Qs, Q + base, rst * ld, T, lane);
    load_tile_mma<D>(Ks, K + base, rst * ld, T, lane);
    load_tile_mma<D>(Vs, V + base, rst * ld, T, lane);
    if (ms.ts && lane < T) ts_s[lane] = ms.ts[b * T + lane];
    __syncwarp();
    const uint8_t* mbase = nullptr;
    if (ms.mask) mbase = ms.mask + ((int64_t)(ms.mask_b > 1 ? b : 0) * ms.mask_h + (ms.mask_h > 1 ? hd : 0)) * T * T;
    float sc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const FragA qa = load_a_rows(Qs, PITCH, ks, g, t);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) mma3(sc[nt], qa, load_b_kt(Ks, PITCH, nt, ks, g, t));
    }
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = g + 8 * (r >> 1), j = 8 * nt + 2 * t + (r & 1);
        float v = -INFINITY;
        if (i < T && j < T && key_valid(ms, causal, ms.ts ? ts_s : nullptr, mbase, T, i, j)) v = sc[nt][r] * scale + bs[nt][r];
        sc[nt][r] = v;
        if (r < 2) m0 = fmaxf(m0, v); else m1 = fmaxf(m1, v);
      }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float mm = r < 2 ? m0 : m1;
        const float pv = sc[nt][r] > -INFINITY ? exp2f((sc[nt][r] - mm) * LOG2E) : 0.f;
        sc[nt][r] = pv;
        if (r < 2) l0 += pv; else l1 += pv;
      }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    float o[KS][4];
#pragma unroll
    for (int c = 0; c < KS; ++c) { o[c][0] = o[c][1] = o[c][2] = o[c][3] = 0.f; }
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {                       // k-step over keys 8nt..8nt+7: slot t <-> key 2t, slot t+4 <-> 2t+1
      const FragA pa = make_a(sc[nt][0], sc[nt][2], sc[nt][1], sc[nt][3]);
#pragma unroll
      for (int c = 0; c < KS; ++c) mma3(o[c], pa, load_b_rows(Vs, PITCH, 8 * nt + 2 * t, 8 * nt + 2 * t + 1, c, g));
    }
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int i = g + 8 * half;
      if (i < T) {
        const float inv = half ? inv1 : inv0;
        float* op = ctx + (b * rsb + i * rst) * (int64_t)H + (int64_t)hd * D + 2 * t;
#pragma unroll
        for (int c = 0; c < KS; ++c)
          *reinterpret_cast<float2*>(op + 8 * c) = make_float2(o[c][2 * half] * inv, o[c][2 * half + 1] * inv);
        if (t == 0) lse[(b * heads + hd) * T + i] = (half ? m1 : m0) + logf(half ? l1 : l0);
        if (attn != nullptr) {
          float* ap = attn + ((b * heads + hd) * T + i) * (int64_t)T;
#pragma unroll
          for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int j = 8 * nt + 2 * t + q;
              if (j < T) ap[j] = sc[nt][2 * half + q] * inv;
            }
        }
      }
    }
  }
}

// instantiated (D, TP) pairs: head dims 8/16/32 x padded lengths 8/16/32; everything else takes the generic kernels
template <int D, int TP>
void tattn_bwd_fast_launch_one(int grid, int threads, size_t smem, cudaStream_t st, const float* Q, const float* K,
                               const float* V, int64_t ld, int64_t B, int T, int heads, int64_t rsb, int64_t rst,
                               const float* bias, MaskSpec ms, const float* ctx, const float* lse, const float* dctx,
                               float* dQ, float* dK, float* dV, int64_t ldd, float* dbias_partial) {
  static SmemOptIn opt_in;                      // per instantiation and device
  opt_in.ensure(tattn_bwd_fast_kernel<D, TP>, smem);
  tattn_bwd_fast_kernel<D, TP><<<grid, threads, smem, st>>>(Q, K, V, ld, B, T, heads, rsb, rst, bias, ms, ctx, lse, dctx,
                                                           dQ, dK, dV, ldd, dbias_partial);
}

#define FAST_CASES(FN, ...)                                    \
  switch (D * 64 + TP) {                                       \
    case 8 * 64 + 8: FN<8, 8> __VA_ARGS__; return true;               \
    case 8 * 64 + 16: FN<8, 16> __VA_ARGS__; return true;             \
    case 8 * 64 + 32: FN<8, 32> __VA_ARGS__; return true;             \
    case 16 * 64 + 8: FN<16, 8> __VA_ARGS__; return true;             \
    case 16 * 64 + 16: FN<16, 16> __VA_ARGS__; return true;           \
    case 16 * 64 + 32: FN<16, 32> __VA_ARGS__; return true;           \
    case 32 * 64 + 8: FN<32, 8> __VA_ARGS__; return true;             \
    case 32 * 64 + 16: FN<32, 16> __VA_ARGS__; return true;           \
    case 32 * 64 + 32: FN<32, 32> __VA_ARGS__; return true;           \
    default: return false;                                     \
  }

}  // namespace

bool tagan_tattn_fwd_fast_launch(int D, int TP, int grid, int threads, size_t smem, cudaStream_t st, const float* Q,
                                 const float* K, const float* V, int64_t ld, int64_t B, int T, int heads, int64_t rsb,
                                 int64_t rst, const float* bias, MaskSpec ms, float* ctx, float* lse, float* attn) {
  FAST_CASES(tattn_fwd_fast_kernel, <<<grid, threads, smem, st>>>(Q, K, V, ld, B, T, heads, rsb, rst, bias, ms, ctx, lse, attn))
}

bool tagan_tattn_bwd_fast_launch(int D, int TP, int grid, int threads, size_t smem, cudaStream_t st, const float* Q,
                                 const float* K, const float* V, int64_t ld, int64_t B, int T, int heads, int64_t rsb,
                                 int64_t rst, const float* bias, MaskSpec ms, const float* ctx, const float* lse,
                                 const float* dctx, float* dQ, float* dK, float* dV, int64_t ldd, float* dbias_partial) {
  FAST_CASES(tattn_bwd_fast_launch_one, (grid, threads, smem, st, Q, K, V, ld, B, T, heads, rsb, rst, bias, ms, ctx, lse, dctx, dQ, dK, dV, ldd, dbias_partial))
}



// ---- tensor-core kernels: 8 < T <= 16, D in {16, 32}, heads <= 8, bias shared by all nodes ----
static size_t mma_fwd_smem(int D, int heads) { return (size_t)heads * (3 * 16 * (D + 4) + 16) * sizeof(float); }


bool tagan_tattn_fwd_mma_launch(int D, int grid, cudaStream_t st, const float* Q, const float* K, const float* V, int64_t ld,
                                int64_t B, int T, int heads, int64_t rsb, int64_t rst, const float* bias, MaskSpec ms,
                                float* ctx, float* lse, float* attn) {
  if (T <= 8 || T > 16 || heads > 8 || (D != 16 && D != 32)) return false;
  const size_t smem = mma_fwd_smem(D, heads);
  static SmemOptIn opted16, opted32;
  if (D == 16) {
    if (opted16.ensure(tattn_fwd_mma_kernel<16>, smem) != cudaSuccess) return false;
    tattn_fwd_mma_kernel<16><<<grid, heads * 32, smem, st>>>(Q, K, V, ld, B, T, heads, rsb, rst, bias, ms, ctx, lse, attn);
  } else {
    if (opted32.ensure(tattn_fwd_mma_kernel<32>, smem) != cudaSuccess) return false;
    tattn_fwd_mma_kernel<32><<<grid, heads * 32, smem, st>>>(Q, K, V, ld, B, T, heads, rsb, rst, bias, ms, ctx, lse, attn);
  }
  return true;
}
