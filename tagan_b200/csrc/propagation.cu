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

// ---- float4 forms of the four kernels above (H, every leading dimension and every base pointer 16-byte aligned):
// same per-element arithmetic, one 128-bit access per tensor per thread.
#define V4_OP(dst, expr)                                                       \
  { float4 _o; { const int q = 0; _o.x = (expr); } { const int q = 1; _o.y = (expr); } \
    { const int q = 2; _o.z = (expr); } { const int q = 3; _o.w = (expr); } dst = _o; }
__device__ __forceinline__ float f4(const float4& v, int q) { return q == 0 ? v.x : q == 1 ? v.y : q == 2 ? v.z : v.w; }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__global__ void gates_fwd_vec_kernel(const float* __restrict__ g, int64_t ldg, const float* __restrict__ second, int64_t lds,
                                     float* __restrict__ r, float* __restrict__ z, float* __restrict__ rs, int64_t ldrs,
                                     int64_t rows, int H4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H4) return;
  const int64_t row = i / H4;
  const int c = (int)(i - row * H4) * 4, H = H4 * 4;
  const float4 gr = ld4(g + row * ldg + c), gz = ld4(g + row * ldg + H + c), sv = ld4(second + row * lds + c);
  float4 rv, zv, o;
  V4_OP(rv, sigmoidf_(f4(gr, q)));
  V4_OP(zv, sigmoidf_(f4(gz, q)));
  V4_OP(o, f4(rv, q) * f4(sv, q));
  st4(r + row * H + c, rv);
  st4(z + row * H + c, zv);
  st4(rs + row * ldrs + c, o);
}

__global__ void gates_bwd_vec_kernel(const float* __restrict__ drs, int64_t lddrs, const float* __restrict__ dz,
                                     const float* __restrict__ r, const float* __restrict__ z,
                                     const float* __restrict__ second, int64_t lds, float* __restrict__ dg, int64_t lddg,
                                     float* __restrict__ dsecond, int64_t ldds, int accumulate, int64_t rows, int H4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H4) return;
  const int64_t row = i / H4;
  const int c = (int)(i - row * H4) * 4, H = H4 * 4;
  const float4 rv = ld4(r + row * H + c), zv = ld4(z + row * H + c), d = ld4(drs + row * lddrs + c);
  const float4 sv = ld4(second + row * lds + c), dzv = ld4(dz + row * H + c);
  float4 a, b, v;
  V4_OP(a, f4(d, q) * f4(sv, q) * f4(rv, q) * (1.f - f4(rv, q)));
  V4_OP(b, f4(dzv, q) * f4(zv, q) * (1.f - f4(zv, q)));
  st4(dg + row * lddg + c, a);
  st4(dg + row * lddg + H + c, b);
  float* o = dsecond + row * ldds + c;
  if (accumulate) { const float4 old = ld4(o); V4_OP(v, f4(old, q) + f4(d, q) * f4(rv, q)); }
  else V4_OP(v, f4(d, q) * f4(rv, q));
  st4(o, v);
}

__global__ void blend_fwd_vec_kernel(const float* __restrict__ cpre, int64_t ldc, const float* __restrict__ z,
                                     const float* __restrict__ base, int64_t ldb, float* __restrict__ cand,
                                     float* __restrict__ out, int residual, int64_t rows, int H4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H4) return;
  const int64_t row = i / H4;
  const int c = (int)(i - row * H4) * 4, H = H4 * 4;
  const float4 cp = ld4(cpre + row * ldc + c), zv = ld4(z + row * H + c), b = ld4(base + row * ldb + c);
  float4 t, o;
  V4_OP(t, tanhf(f4(cp, q)));
  if (residual) V4_OP(o, ((1.f - f4(zv, q)) * f4(b, q) + f4(zv, q) * f4(t, q)) + f4(b, q))
  else V4_OP(o, (1.f - f4(zv, q)) * f4(b, q) + f4(zv, q) * f4(t, q))
  st4(cand + row * H + c, t);
  st4(out + row * H + c, o);
}

__global__ void blend_bwd_vec_kernel(const float* __restrict__ dout, const float* __restrict__ z,
                                     const float* __restrict__ cand, const float* __restrict__ base, int64_t ldb,
                                     float* __restrict__ dcpre, int64_t lddc, float* __restrict__ dz, float* __restrict__ dbase,
                                     int64_t lddb, int accumulate, int residual, int64_t rows, int H4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H4) return;
  const int64_t row = i / H4;
  const int c = (int)(i - row * H4) * 4, H = H4 * 4;
  const float4 d = ld4(dout + row * H + c), zv = ld4(z + row * H + c), t = ld4(cand + row * H + c), b = ld4(base + row * ldb + c);
  float4 a, e, v;
  V4_OP(a, f4(d, q) * f4(zv, q) * (1.f - f4(t, q) * f4(t, q)));
  V4_OP(e, f4(d, q) * (f4(t, q) - f4(b, q)));
  st4(dcpre + row * lddc + c, a);
  st4(dz + row * H + c, e);
  float* o = dbase + row * lddb + c;
  if (accumulate) { const float4 old = ld4(o); V4_OP(v, f4(old, q) + (f4(d, q) * (1.f - f4(zv, q)) + (residual ? f4(d, q) : 0.f))); }
  else V4_OP(v, f4(d, q) * (1.f - f4(zv, q)) + (residual ? f4(d, q) : 0.f));
  st4(o, v);
}

__host__ inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

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

// mean / sum aggregation, float4 columns: a thread walks u keeping the 2W+1 scaled gradients of its window in
// registers, so every dg element is loaded once (the scalar kernel re-reads each 2W+1 times through L2).  The sum
// runs over ascending t from 0.f exactly like the scalar kernel (out-of-range slots hold 0.f), so results are
// bit-identical.
template <int W>
__global__ void window_bwd_vec_kernel(const float* __restrict__ dg, float* __restrict__ dp, int T, int64_t inner4, int agg) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= inner4) return;
  const float4* g4 = reinterpret_cast<const float4*>(dg);
  float4* o4 = reinterpret_cast<float4*>(dp);
  auto fetch = [&](int t) -> float4 {
    if (t < 0 || t >= T) return make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v = g4[(int64_t)t * inner4 + x];
    if (agg == 0) {
      const float c = (float)(min(T, t + W + 1) - max(0, t - W));
      v.x = v.x / c; v.y = v.y / c; v.z = v.z / c; v.w = v.w / c;
    }
    return v;
  };
  float4 win[2 * W + 1];
#pragma unroll
  for (int k = 0; k <= 2 * W; ++k) win[k] = fetch(k - W);
  for (int u = 0; u < T; ++u) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k <= 2 * W; ++k) { s.x += win[k].x; s.y += win[k].y; s.z += win[k].z; s.w += win[k].w; }
    o4[(int64_t)u * inner4 + x] = s;
#pragma unroll
    for (int k = 0; k < 2 * W; ++k) win[k] = win[k + 1];
    win[2 * W] = fetch(u + 1 + W);
  }
}

}  // namespace

TAGAN_API int tagan_gates_fwd(const float* g, int64_t ldg, const float* second, int64_t lds, float* r, float* z,
                              float* rs, int64_t ldrs, int64_t rows, int32_t H, tagan_stream_t stream) {
  if (!g || !second || !r || !z || !rs || rows < 0 || H <= 0 || ldg < 2 * H) return TAGAN_E_INVALID;
  if (rows == 0) return 0;
  if (H % 4 == 0 && ldg % 4 == 0 && lds % 4 == 0 && ldrs % 4 == 0 && al16(g) && al16(second) && al16(r) && al16(z) && al16(rs))
    gates_fwd_vec_kernel<<<ceil_div_i64(rows * (H / 4), 256), 256, 0, as_stream(stream)>>>(g, ldg, second, lds, r, z, rs, ldrs, rows, H / 4);
  else
    gates_fwd_kernel<<<ceil_div_i64(rows * H, 256), 256, 0, as_stream(stream)>>>(g, ldg, second, lds, r, z, rs, ldrs, rows, H);
  return tagan_launch_status();
}

TAGAN_API int tagan_gates_bwd(const float* drs, int64_t lddrs, const float* dz, const float* r, const float* z,
                              const float* second, int64_t lds, float* dg, int64_t lddg, float* dsecond,
                              int64_t ldds, int32_t accumulate, int64_t rows, int32_t H, tagan_stream_t stream) {
  if (!drs || !dz || !r || !z || !second || !dg || !dsecond || rows < 0 || H <= 0 || lddg < 2 * H) return TAGAN_E_INVALID;
  if (rows == 0) return 0;
  if (H % 4 == 0 && lddrs % 4 == 0 && lds % 4 == 0 && lddg % 4 == 0 && ldds % 4 == 0 && al16(drs) && al16(dz) && al16(r) &&
      al16(z) && al16(second) && al16(dg) && al16(dsecond))
    gates_bwd_vec_kernel<<<ceil_div_i64(rows * (H / 4), 256), 256, 0, as_stream(stream)>>>(drs, lddrs, dz, r, z, second, lds, dg,
                                                                                         lddg, dsecond, ldds, accumulate, rows, H / 4);
  else
    gates_bwd_kernel<<<ceil_div_i64(rows * H, 256), 256, 0, as_stream(stream)>>>(drs, lddrs, dz, r, z, second, lds, dg,
                                                                                 lddg, dsecond, ldds, accumulate, rows, H);
  return tagan_launch_status();
}

TAGAN_API int tagan_blend_fwd(const float* cand_pre, int64_t ldc, const float* z, const float* base, int64_t ldb,
                              float* cand, float* out, int32_t residual, int64_t rows, int32_t H,
                              tagan_stream_t stream) {
  if (!cand_pre || !z || !base || !cand || !out || rows < 0 || H <= 0 || ldc < H) return TAGAN_E_INVALID;
  if (rows == 0) return 0;
  if (H % 4 == 0 && ldc % 4 == 0 && ldb % 4 == 0 && al16(cand_pre) && al16(z) && al16(base) && al16(cand) && al16(out))
    blend_fwd_vec_kernel<<<ceil_div_i64(rows * (H / 4), 256), 256, 0, as_stream(stream)>>>(cand_pre, ldc, z, base, ldb, cand, out,
                                                                                         residual, rows, H / 4);
  else
    blend_fwd_kernel<<<ceil_div_i64(rows * H, 256), 256, 0, as_stream(stream)>>>(cand_pre, ldc, z, base, ldb, cand, out,
                                                                                 residual, rows, H);
  return tagan_launch_status();
}

TAGAN_API int tagan_blend_bwd(const float* dout, const float* z, const float* cand, const float* base, int64_t ldb,
                              float* dcand_pre, int64_t lddc, float* dz, float* dbase, int64_t lddb,
                              int32_t accumulate, int32_t residual, int64_t rows, int32_t H, tagan_stream_t stream) {
  if (!dout || !z || !cand || !base || !dcand_pre || !dz || !dbase || rows < 0 || H <= 0 || lddc < H) return TAGAN_E_INVALID;
  if (rows == 0) return 0;
  if (H % 4 == 0 && ldb % 4 == 0 && lddc % 4 == 0 && lddb % 4 == 0 && al16(dout) && al16(z) && al16(cand) && al16(base) &&
      al16(dcand_pre) && al16(dz) && al16(dbase))
    blend_bwd_vec_kernel<<<ceil_div_i64(rows * (H / 4), 256), 256, 0, as_stream(stream)>>>(dout, z, cand, base, ldb, dcand_pre, lddc,
                                                                                         dz, dbase, lddb, accumulate, residual, rows, H / 4);
  else
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
  if (agg != 1 && inner % 4 == 0 && al16(dg) && al16(dp) && window >= 1 && window <= 4) {
    const unsigned grid = ceil_div_i64(inner / 4, 256);
    cudaStream_t st = as_stream(stream);
    switch (window) {
      case 1: window_bwd_vec_kernel<1><<<grid, 256, 0, st>>>(dg, dp, T, inner / 4, agg); break;
      case 2: window_bwd_vec_kernel<2><<<grid, 256, 0, st>>>(dg, dp, T, inner / 4, agg); break;
      case 3: window_bwd_vec_kernel<3><<<grid, 256, 0, st>>>(dg, dp, T, inner / 4, agg); break;
      default: window_bwd_vec_kernel<4><<<grid, 256, 0, st>>>(dg, dp, T, inner / 4, agg); break;
    }
    return tagan_launch_status();
  }
  window_bwd_kernel<<<ceil_div_i64(inner, 256), 256, 0, as_stream(stream)>>>(dg, p, agg_out, dp, T, inner, window, agg);
  return tagan_launch_status();
}
