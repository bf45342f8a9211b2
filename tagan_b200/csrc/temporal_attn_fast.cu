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

template <int D, int TP>
__global__ void __launch_bounds__(MAX_WARPS * 32)
tattn_bwd_fast_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, int64_t ld,
                      int64_t B, int T, int heads, int64_t rsb, int64_t rst, const float* __restrict__ bias, MaskSpec ms,
                      const float* __restrict__ ctx, const float* __restrict__ lse, const float* __restrict__ dctx,
                      float* __restrict__ dQ, float* __restrict__ dK, float* __restrict__ dV, int64_t ldd,
                      float* __restrict__ dbias_partial /* [grid, h, T, T] or null */) {
  extern __shared__ __align__(16) float smem[];
  constexpr int PPW = 32 / TP;
  const int H = heads * D;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / TP, li = lane - sub * TP;
  const int tpad = (T + 3) & ~3;
  const int slot_floats = 4 * T * D + 3 * tpad;
  float* slot = smem + (size_t)(w * PPW + sub) * slot_floats;
  float* Qs = slot;
  float* Ks = Qs + T * D;
  float* Vs = Ks + T * D;
  float* Gs = Vs + T * D;
  float* lse_s = Gs + T * D;
  float* del_s = lse_s + tpad;
  float* ts_s = del_s + tpad;
  const float scale = 1.f / sqrtf((float)D);
  const bool causal = (ms.flags & 1) || ((ms.flags & 4) && ms.allones_flag && *ms.allones_flag != 0);
  const int hd = w * PPW + sub;
  const bool pv = hd < heads;
  const bool rv = pv && li < T;
  float brow[TP], bcol[TP], dbrow[TP];
#pragma unroll
  for (int j = 0; j < TP; ++j) {
    brow[j] = (bias && rv && j < T) ? bias[((int64_t)hd * T + li) * T + j] : 0.f;   // lane = query i, over keys j
    bcol[j] = (bias && rv && j < T) ? bias[((int64_t)hd * T + j) * T + li] : 0.f;   // lane = key j, over queries i
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
    __syncwarp();
    if (rv) {
      const float* cp = ctx + (b * rsb + li * rst) * (int64_t)H + (int64_t)hd * D;
      float dl = 0.f;
#pragma unroll
      for (int c = 0; c < D; c += 4) {
        float4 c4 = __ldg(reinterpret_cast<const float4*>(cp + c));
        const float* g = Gs + li * D + c;
        dl = fmaf(g[0], c4.x, dl); dl = fmaf(g[1], c4.y, dl); dl = fmaf(g[2], c4.z, dl); dl = fmaf(g[3], c4.w, dl);
      }
      del_s[li] = dl;
      lse_s[li] = lse[(b * heads + hd) * T + li];
    }
    __syncwarp();
    const uint8_t* mbase = nullptr;
    if (ms.mask && pv)
      mbase = ms.mask + ((int64_t)(ms.mask_b > 1 ? b : 0) * ms.mask_h + (ms.mask_h > 1 ? hd : 0)) * T * T;
    {   // ---- phase 1: lane = query row i
      const int i = li;
      float q[D], g[D], dq[D];
#pragma unroll
      for (int c = 0; c < D; ++c) { q[c] = rv ? Qs[i * D + c] : 0.f; g[c] = rv ? Gs[i * D + c] : 0.f; dq[c] = 0.f; }
      const float ls2 = rv ? lse_s[i] * LOG2E : 0.f, dl = rv ? del_s[i] : 0.f;
#pragma unroll
      for (int j = 0; j < TP; ++j) {
        if (j < T) {
          const float sv = dot_smem<D>(q, Ks + j * D) * scale + brow[j];
          const bool kv = rv && key_valid(ms, causal, ms.ts ? ts_s : nullptr, mbase, T, i, j) && sv > -INFINITY;
          const float pr = kv ? exp2f(sv * LOG2E - ls2) : 0.f;
          const float dp = dot_smem<D>(g, Vs + j * D);
          const float ds = pr * (dp - dl);
          dbrow[j] += ds;
          const float dss = ds * scale;
          const float* kr = Ks + j * D;
#pragma unroll
          for (int c = 0; c < D; ++c) dq[c] = fmaf(dss, kr[c], dq[c]);
        }
      }
      if (rv) {
        float* op = dQ + (b * rsb + i * rst) * ldd + (int64_t)hd * D;
#pragma unroll
        for (int c = 0; c < D; c += 4) *reinterpret_cast<float4*>(op + c) = make_float4(dq[c], dq[c + 1], dq[c + 2], dq[c + 3]);
      }
    }
    {   // ---- phase 2: lane = key row j
      const int j = li;
      float k[D], v[D], dk[D], dv[D];
#pragma unroll
      for (int c = 0; c < D; ++c) { k[c] = rv ? Ks[j * D + c] : 0.f; v[c] = rv ? Vs[j * D + c] : 0.f; dk[c] = 0.f; dv[c] = 0.f; }
#pragma unroll
      for (int i = 0; i < TP; ++i) {
        if (i < T) {
          const float sv = dot_smem<D>(k, Qs + i * D) * scale + bcol[i];
          const bool kv = rv && key_valid(ms, causal, ms.ts ? ts_s : nullptr, mbase, T, i, j) && sv > -INFINITY;
          const float pr = kv ? exp2f((sv - lse_s[i]) * LOG2E) : 0.f;
          const float* gr = Gs + i * D;
          const float dp = dot_smem<D>(v, gr);
          const float dss = pr * (dp - del_s[i]) * scale;
          const float* qr = Qs + i * D;
#pragma unroll
          for (int c = 0; c < D; ++c) { dv[c] = fmaf(pr, gr[c], dv[c]); dk[c] = fmaf(dss, qr[c], dk[c]); }
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
  FAST_CASES(tattn_bwd_fast_kernel, <<<grid, threads, smem, st>>>(Q, K, V, ld, B, T, heads, rsb, rst, bias, ms, ctx, lse, dctx, dQ, dK, dV, ldd, dbias_partial))
}
