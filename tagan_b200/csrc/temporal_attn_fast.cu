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


// instantiated (D, TP) pairs: head dims 8/16/32 x padded lengths 8/16/32 (except 32x32, whose unrolled body is
// too large to pay off); everything else takes the generic kernels
template <int D, int TP>
void tattn_bwd_fast_launch_one(int grid, int threads, size_t smem, cudaStream_t st, const float* Q, const float* K,
                               const float* V, int64_t ld, int64_t B, int T, int heads, int64_t rsb, int64_t rst,
                               const float* bias, MaskSpec ms, const float* ctx, const float* lse, const float* dctx,
                               float* dQ, float* dK, float* dV, int64_t ldd, float* dbias_partial) {
  static size_t opted = 0;                      // per instantiation; raising the limit is idempotent
  if (smem > 48 * 1024 && smem > opted) {
    cudaFuncSetAttribute(tattn_bwd_fast_kernel<D, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    opted = smem;
  }
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

