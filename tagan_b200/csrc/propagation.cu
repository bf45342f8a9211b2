// (b3-b6) Element-wise stages of temporal propagation.  The Linears run through tagan_gemm, the
// LayerNorms (with the exp(-dt) state decay as a fused row scale) through tagan_layernorm_*;
// what is left are the gate non-linearities of TemporalGRUCell / TemporalGatingUnit and the
// sliding-window aggregation of TemporalSkipConnection -- pure HBM streams.
//   GRU  (src/tagan/layers/temporal_propagation.py:531-539): r,z = sigmoid(W[x^,h^]); rh = r*h^;
//        h~ = tanh(Wc[x^,rh]); hn = (1-z)*h^ + z*h~
//   Gate (:1043-1060): u,r = sigmoid(W[c,p]); rp = r*p; cand = tanh(Wo[c,rp]);
//        out = (1-u)*c + u*cand (+ c if residual)
// Both share the kernels below: `second` is the gated half (h^ / p), `base` the blended one (h^ / c).
#include "common.cuh"

namespace {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// g [rows,2H] pre-activations (first H: blend gate z/u ... see `gate_first`): the reference
// concatenates nothing here -- two separate Linears -- so the caller packs [W_a; W_b] as it likes.
// Layout used by the host: g[:, :H] = reset pre-activation, g[:, H:] = update pre-activation.
__global__ void gates_fwd_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ second, int64_t lds,
                                 float* __restrict__ r, float* __restrict__ z, float* __restrict__ rs, int64_t ldrs,
                                 int64_t rows, int H) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H) return;
  int64_t row = i / H;
  int c = (int)(i - row * H);
  float rv = sigmoidf_(g[row * ldg + c]);
  float zv = sigmoidf_(g[row * ldg + H + c]);
  r[i] = rv;
  z[i] = zv;
  rs[row * ldrs + c] = rv * second[row * lds + c];
}

// dg[:, :H] = d(rs)*second * r(1-r);  dg[:, H:] = dz * z(1-z);  dsecond (+)= d(rs)*r
__global__ void gates_bwd_kernel(const float* __restrict__ drs, int64_t lddrs, const float* __restrict__ dz,
                                 const float* __restrict__ r, const float* __restrict__ z,
                                 const float* __restrict__ second, int64_t lds, float* __restrict__ dg, int64_t lddg,
                                 float* __restrict__ dsecond, int64_t ldds, int accumulate, int64_t rows, int H) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H) return;
  int64_t row = i / H;
  int c = (int)(i - row * H);
  float rv = r[i], zv = z[i], d = drs[row * lddrs + c];
  dg[row * lddg + c] = d * second[row * lds + c] * rv * (1.f - rv);
  dg[row * lddg + H + c] = dz[i] * zv * (1.f - zv);
  float* o = dsecond + row * ldds + c;
  float v = d * rv;
  *o = accumulate ? *o + v : v;
}

// cand = tanh(cpre); out = (1-z)*base + z*cand (+ base if residual)
__global__ void blend_fwd_kernel(const float* __restrict__ cpre, int64_t ldc, const float* __restrict__ z,
                                 const float* __restrict__ base, int64_t ldb, float* __restrict__ cand,
                                 float* __restrict__ out, int residual, int64_t rows, int H) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H) return;
  int64_t row = i / H;
  int c = (int)(i - row * H);
  float t = tanhf(cpre[row * ldc + c]);
  float zv = z[i], b = base[row * ldb + c];
  cand[i] = t;
  float o = (1.f - zv) * b + zv * t;
  out[i] = residual ? o + b : o;
}

// dcpre = dout*z*(1-cand^2); dz = dout*(cand-base); dbase (+)= dout*(1-z) (+ dout if residual)
__global__ void blend_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ z,
                                 const float* __restrict__ cand, const float* __restrict__ base, int64_t ldb,
                                 float* __restrict__ dcpre, int64_t lddc, float* __restrict__ dz, float* __restrict__ dbase,
                                 int64_t lddb, int accumulate, int residual, int64_t rows, int H) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H) return;
  int64_t row = i / H;
  int c = (int)(i - row * H);
  float d = dout[i], zv = z[i], t = cand[i], b = base[row * ldb + c];
  dcpre[row * lddc + c] = d * zv * (1.f - t * t);
  dz[i] = d * (t - b);
  float v = d * (1.f - zv) + (residual ? d : 0.f);
  float* o = dbase + row * lddb + c;
  *o = accumulate ? *o + v : v;
}

// g[t] = agg over u in [max(0,t-w), min(T,t+w+1)) of p[u]; p is [T, inner] (inner = rows*H).
// A thread walks t for one float4 column so the re-reads of the window hit L1/L2.
__global__ void window_fwd_kernel(const float* __restrict__ p, float* __restrict__ out, int T, int64_t inner4,
                                  int w, int agg) {
  int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= inner4) return;
  const float4* p4 = reinterpret_cast<const float4*>(p);
  float4* o4 = reinterpret_cast<float4*>(out);
  for (int t = 0; t < T; ++t) {
    const int u0 = max(0, t - w), u1 = min(T, t + w + 1);
    float4 a = p4[(int64_t)u0 * inner4 + x];
    for (int u = u0 + 1; u < u1; ++u) {
      float4 v = p4[(int64_t)u * inner4 + x];
      if (agg == 1) { a.x = fmaxf(a.x, v.x); a.y = fmaxf(a.y, v.y); a.z = fmaxf(a.z, v.z); a.w = fmaxf(a.w, v.w); }
      else { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
    }
    if (agg == 0) {
      const float inv = 1.f / (float)(u1 - u0);   // torch.mean = sum / count
      a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
    }
    o4[(int64_t)t * inner4 + x] = a;
  }
}

__device__ __forceinline__ float max_route(float pv, float av, bool earlier_hit) { return (pv == av && !earlier_hit) ? 1.f : 0.f; }

// dp[u] = sum over t with u in window(t) of dg[t] * weight(t,u); scalar version (handles max routing).
__global__ void window_bwd_kernel(const float* __restrict__ dg, const float* __restrict__ p,
                                  const float* __restrict__ aggv, float* __restrict__ dp, int T, int64_t inner, int w,
                                  int agg) {
  int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= inner) return;
  for (int u = 0; u < T; ++u) {
    float s = 0.f;
    const int t0 = max(0, u - w), t1 = min(T, u + w + 1);
    for (int t = t0; t < t1; ++t) {
      const float d = dg[(int64_t)t * inner + x];
      if (agg == 0) {
        const int cnt = min(T, t + w + 1) - max(0, t - w);
        s += d / (float)cnt;
      } else if (agg == 2) {
        s += d;
      } else {
        // max: gradient goes to the first window element equal to the maximum
        const float av = aggv[(int64_t)t * inner + x];
        const float pv = p[(int64_t)u * inner + x];
        if (pv == av) {
          bool earlier = false;
          for (int u2 = max(0, t - w); u2 < u; ++u2)
            if (p[(int64_t)u2 * inner + x] == av) { earlier = true; break; }
          if (!earlier) s += d;
        }
      }
    }
    dp[(int64_t)u * inner + x] = s;
  }
}

}  // namespace

TAGAN_API int tagan_gates_fwd(const float* g, int64_t ldg, const float* second, int64_t lds, float* r, float* z,
                              float* rs, int64_t ldrs, int64_t rows, int32_t H, tagan_stream_t stream) {
  if (!g || !second || !r || !z || !rs || rows < 0 || H <= 0 || ldg < 2 * H) return TAGAN_E_INVALID;
  if (rows == 0) return 0;
  gates_fwd_kernel<<<ceil_div_i64(rows * H, 256), 256, 0, as_stream(stream)>>>(g, ldg, second, lds, r, z, rs, ldrs, rows, H);
  return tagan_launch_status();
}

TAGAN_API int tagan_gates_bwd(const float* drs, int64_t lddrs, const float* dz, const float* r, const float* z,
                              const float* second, int64_t lds, float* dg, int64_t lddg, float* dsecond,
                              int64_t ldds, int32_t accumulate, int64_t rows, int32_t H, tagan_stream_t stream) {
  if (!drs || !dz || !r || !z || !second || !dg || !dsecond || rows < 0 || H <= 0 || lddg < 2 * H) return TAGAN_E_INVALID;
  if (rows == 0) return 0;
  gates_bwd_kernel<<<ceil_div_i64(rows * H, 256), 256, 0, as_stream(stream)>>>(drs, lddrs, dz, r, z, second, lds, dg,
                                                                               lddg, dsecond, ldds, accumulate, rows, H);
  return tagan_launch_status();
}

TAGAN_API int tagan_blend_fwd(const float* cand_pre, int64_t ldc, const float* z, const float* base, int64_t ldb,
                              float* cand, float* out, int32_t residual, int64_t rows, int32_t H,
                              tagan_stream_t stream) {
  if (!cand_pre || !z || !base || !cand || !out || rows < 0 || H <= 0 || ldc < H) return TAGAN_E_INVALID;
  if (rows == 0) return 0;
  blend_fwd_kernel<<<ceil_div_i64(rows * H, 256), 256, 0, as_stream(stream)>>>(cand_pre, ldc, z, base, ldb, cand, out,
                                                                               residual, rows, H);
  return tagan_launch_status();
}

TAGAN_API int tagan_blend_bwd(const float* dout, const float* z, const float* cand, const float* base, int64_t ldb,
                              float* dcand_pre, int64_t lddc, float* dz, float* dbase, int64_t lddb,
                              int32_t accumulate, int32_t residual, int64_t rows, int32_t H, tagan_stream_t stream) {
  if (!dout || !z || !cand || !base || !dcand_pre || !dz || !dbase || rows < 0 || H <= 0 || lddc < H) return TAGAN_E_INVALID;
  if (rows == 0) return 0;
  blend_bwd_kernel<<<ceil_div_i64(rows * H, 256), 256, 0, as_stream(stream)>>>(dout, z, cand, base, ldb, dcand_pre, lddc,
                                                                               dz, dbase, lddb, accumulate, residual, rows, H);
  return tagan_launch_status();
}

TAGAN_API int tagan_skip_window_fwd(const float* p, float* out, int32_t T, int64_t inner, int32_t window, int32_t agg,
                                    tagan_stream_t stream) {
  if (!p || !out || T < 0 || inner < 0 || window < 0 || agg < 0 || agg > 2) return TAGAN_E_INVALID;
  if (inner % 4) return TAGAN_E_UNSUPPORTED;
  if (T == 0 || inner == 0) return 0;
  window_fwd_kernel<<<ceil_div_i64(inner / 4, 256), 256, 0, as_stream(stream)>>>(p, out, T, inner / 4, window, agg);
  return tagan_launch_status();
}

TAGAN_API int tagan_skip_window_bwd(const float* dg, const float* p, const float* agg_out, float* dp, int32_t T,
                                    int64_t inner, int32_t window, int32_t agg, tagan_stream_t stream) {
  if (!dg || !dp || T < 0 || inner < 0 || window < 0 || agg < 0 || agg > 2) return TAGAN_E_INVALID;
  if (agg == 1 && (!p || !agg_out)) return TAGAN_E_INVALID;
  if (T == 0 || inner == 0) return 0;
  window_bwd_kernel<<<ceil_div_i64(inner, 256), 256, 0, as_stream(stream)>>>(dg, p, agg_out, dp, T, inner, window, agg);
  return tagan_launch_status();
}
