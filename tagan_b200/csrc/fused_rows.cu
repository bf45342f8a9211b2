// Fused node-stream passes of the propagation core: each kernel replaces a chain of element-wise / LayerNorm
// launches of the reference by ONE pass over its rows (every tensor is read once and written once).
//
//   tagan_ln_pair_fwd/bwd      s_t = LN_out(hn_t);  h^_{t+1} = LN_h(s_t) * exp(-clamp(dt,0,10))
//                              (TemporalGRUCell.forward, src/tagan/layers/temporal_propagation.py:545-546 of step t and
//                              :505-514 of step t+1 -- the state of the GRU scan leaves and re-enters the cell
//                              through two LayerNorms with nothing in between)
//   tagan_gelu_ln_fwd/bwd      p = LN1(GELU(a))                      (TemporalSkipConnection.forward :866-877)
//   tagan_window_gelu_fwd/bwd  gg_t = GELU(mean/sum_{|u-t|<=w} p_u)  (:880-894 + the activation of :929-933)
//   tagan_mse_fwd/bwd          loss = mean(x^2) and its gradient (the benchmark's / trainer's scalar objective)
//
// Layout: one warp per row, the row held in registers (float4 lanes when cols % 128 == 0, scalar lanes otherwise,
// cols <= 512); LayerNorm statistics exactly as tagan_layernorm_fwd (mean, then the centred second moment).  The
// affine-parameter gradients are per-block partials reduced in a fixed order => deterministic.
#include "common.cuh"

namespace {

constexpr int RW = 8;                 // warps (rows in flight) per block
constexpr float EPS = 1e-5f;

// ---- a row in registers ---------------------------------------------------------------------------------------
// VEC: element e of a lane is column (e/4)*128 + lane*4 + e%4 (cols == 128*NV exactly);
// scalar: element e is column e*32 + lane, valid while < cols (cols <= 32*NV).
template <int NV, bool VEC>
struct Row {
  static constexpr int E = VEC ? 4 * NV : NV;
  static __device__ __forceinline__ int col(int e, int lane) { return VEC ? (e >> 2) * 128 + lane * 4 + (e & 3) : e * 32 + lane; }
  static __device__ __forceinline__ void load(float (&v)[E], const float* p, int lane, int cols) {
    if (VEC) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 t = *reinterpret_cast<const float4*>(p + j * 128 + lane * 4);
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) { const int c = e * 32 + lane; v[e] = c < cols ? p[c] : 0.f; }
    }
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[E], int lane, int cols) {
    if (VEC) {
#pragma unroll
      for (int j = 0; j < NV; ++j)
        *reinterpret_cast<float4*>(p + j * 128 + lane * 4) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) { const int c = e * 32 + lane; if (c < cols) p[c] = v[e]; }
    }
  }
  static __device__ __forceinline__ bool valid(int e, int lane, int cols) { return VEC ? true : (e * 32 + lane) < cols; }
};

template <int E>
__device__ __forceinline__ float lane_sum(const float (&v)[E]) {
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) s += v[e];
  return s;
}

// mean and rstd of a row held in registers (invalid elements are 0 and excluded from the centred moment)
template <int NV, bool VEC>
__device__ __forceinline__ void row_stats(const float (&v)[Row<NV, VEC>::E], int lane, int cols, float& mean, float& rstd) {
  using R = Row<NV, VEC>;
  mean = warp_sum(lane_sum<R::E>(v)) / (float)cols;
  float q = 0.f;
#pragma unroll
  for (int e = 0; e < R::E; ++e)
    if (R::valid(e, lane, cols)) { const float d = v[e] - mean; q = fmaf(d, d, q); }
  rstd = 1.f / sqrtf(warp_sum(q) / (float)cols + EPS);
}

__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad(float v) {
  const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
  return cdf + v * pdf;
}

// per-block partial sums of NACC per-column accumulators: acc[k][e] of every warp -> partial[block][k][cols]
template <int NV, bool VEC, int NACC>
__device__ __forceinline__ void block_partials(float (&acc)[NACC][Row<NV, VEC>::E], float* red /*[RW][NACC][cols]*/,
                                               float* partial, int cols) {
  using R = Row<NV, VEC>;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NACC; ++k)
#pragma unroll
    for (int e = 0; e < R::E; ++e)
      if (R::valid(e, lane, cols)) red[((size_t)w * NACC + k) * cols + R::col(e, lane)] = acc[k][e];
  __syncthreads();
  for (int i = threadIdx.x; i < NACC * cols; i += RW * 32) {
    float s = red[i];
#pragma unroll
    for (int ww = 1; ww < RW; ++ww) s += red[(size_t)ww * NACC * cols + i];
    partial[(size_t)blockIdx.x * NACC * cols + i] = s;
  }
}

// out[i] (+)= sum over parts of partial[p][i], ascending p: 8 interleaved slices, then the slices in order
__global__ void __launch_bounds__(256)
reduce_cols_kernel(const float* __restrict__ partial, int parts, int width, float* __restrict__ out, int accumulate) {
  __shared__ float red[8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), slice = threadIdx.x >> 5;
  float s = 0.f;
  if (c < width)
    for (int p = slice; p < parts; p += 8) s += partial[(size_t)p * width + c];
  red[slice][threadIdx.x & 31] = s;
  __syncthreads();
  if (slice == 0 && c < width) {
    float t = red[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += red[k][threadIdx.x];
    out[c] = accumulate ? out[c] + t : t;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// LN pair
// ---------------------------------------------------------------------------------------------------------------
template <int NV, bool VEC>
__global__ void __launch_bounds__(RW * 32)
ln_pair_fwd_kernel(const float* __restrict__ hn, int64_t ldhn, const float* __restrict__ bz, const float* __restrict__ bbase,
                   int64_t ldbase, float* __restrict__ cand_out, const float* __restrict__ go, const float* __restrict__ bo,
                   const float* __restrict__ gh, const float* __restrict__ bh, const float* __restrict__ ts, int64_t ldts,
                   int t_hi, float* __restrict__ s, int64_t lds, float* __restrict__ hh, int64_t ldhh,
                   float* __restrict__ mean_o, float* __restrict__ rstd_o, float* __restrict__ mean_h,
                   float* __restrict__ rstd_h, float* __restrict__ decay, int64_t rows, int cols) {
  using R = Row<NV, VEC>;
  const int64_t row = (int64_t)blockIdx.x * RW + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[R::E], g[R::E], b[R::E];
  R::load(v, hn + row * ldhn, lane, cols);
  if (bz != nullptr) {
    // blend prologue: the row handed in is the candidate PRE-activation; cand = tanh(.), hn = (1-z) h^ + z cand (:538-542)
    R::load(g, bz + row * (int64_t)cols, lane, cols);
    R::load(b, bbase + row * ldbase, lane, cols);
#pragma unroll
    for (int e = 0; e < R::E; ++e) {
      const float t = tanhf(v[e]);
      v[e] = R::valid(e, lane, cols) ? (1.f - g[e]) * b[e] + g[e] * t : 0.f;
      g[e] = t;
    }
    R::store(cand_out + row * (int64_t)cols, g, lane, cols);
  }
  R::load(g, go, lane, cols);
  R::load(b, bo, lane, cols);
  float mu, rs;
  row_stats<NV, VEC>(v, lane, cols, mu, rs);
#pragma unroll
  for (int e = 0; e < R::E; ++e) v[e] = R::valid(e, lane, cols) ? (v[e] - mu) * rs * g[e] + b[e] : 0.f;
  R::store(s + row * lds, v, lane, cols);
  if (lane == 0) { mean_o[row] = mu; rstd_o[row] = rs; }
  if (hh == nullptr) return;
  float dec = 1.f;
  if (ts != nullptr) {                                      // exp(-clamp(ts[:,t_hi] - ts[:,t_hi-1], 0, 10)) (:509-514)
    float d = ts[row * ldts + t_hi] - ts[row * ldts + t_hi - 1];
    d = fminf(fmaxf(d, 0.f), 10.f);
    dec = expf(-d);
  }
  R::load(g, gh, lane, cols);
  R::load(b, bh, lane, cols);
  row_stats<NV, VEC>(v, lane, cols, mu, rs);
#pragma unroll
  for (int e = 0; e < R::E; ++e) v[e] = ((v[e] - mu) * rs * g[e] + b[e]) * dec;
  R::store(hh + row * ldhh, v, lane, cols);
  if (lane == 0) { mean_h[row] = mu; rstd_h[row] = rs; if (decay) decay[row] = dec; }
}

// dhn = LN_out'( ds_ext + LN_h'(dhh * decay) );  partial[block][4][cols] = d gamma_o, d beta_o, d gamma_h, d beta_h
template <int NV, bool VEC>
__global__ void __launch_bounds__(RW * 32)
ln_pair_bwd_kernel(const float* __restrict__ ds_ext, int64_t ldds, const float* __restrict__ dhh, int64_t lddhh,
                   const float* __restrict__ hn, int64_t ldhn, const float* __restrict__ bz, const float* __restrict__ bbase,
                   int64_t ldbase, float* __restrict__ dgz, float* __restrict__ dgc, int64_t lddg, float* __restrict__ dbase,
                   int64_t lddbase, const float* __restrict__ go, const float* __restrict__ bo,
                   const float* __restrict__ gh, const float* __restrict__ mean_o, const float* __restrict__ rstd_o,
                   const float* __restrict__ mean_h, const float* __restrict__ rstd_h, const float* __restrict__ decay,
                   float* __restrict__ dhn, int64_t lddhn, float* __restrict__ partial, int64_t rows, int cols,
                   int64_t rows_per_block) {
  using R = Row<NV, VEC>;
  extern __shared__ float red[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[4][R::E];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int e = 0; e < R::E; ++e) acc[k][e] = 0.f;
  float g_o[R::E], b_o[R::E], g_h[R::E];
  R::load(g_o, go, lane, cols);
  R::load(b_o, bo, lane, cols);
  if (dhh != nullptr) R::load(g_h, gh, lane, cols);
  const float inv = 1.f / (float)cols;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  for (int64_t row = r0 + w; row < r1; row += RW) {
    float xo[R::E], ds[R::E];
    R::load(xo, hn + row * ldhn, lane, cols);
    float zv[R::E], bv[R::E], cv[R::E];
    if (bz != nullptr) {                       // blend fused: the row handed in is cand; hn is rebuilt, never stored
      R::load(zv, bz + row * (int64_t)cols, lane, cols);
      R::load(bv, bbase + row * ldbase, lane, cols);
#pragma unroll
      for (int e = 0; e < R::E; ++e) { cv[e] = xo[e]; xo[e] = (1.f - zv[e]) * bv[e] + zv[e] * cv[e]; }
    }
    if (ds_ext != nullptr) R::load(ds, ds_ext + row * ldds, lane, cols);
    else {
#pragma unroll
      for (int e = 0; e < R::E; ++e) ds[e] = 0.f;
    }
    const float muo = mean_o[row], rso = rstd_o[row];
#pragma unroll
    for (int e = 0; e < R::E; ++e) xo[e] = R::valid(e, lane, cols) ? (xo[e] - muo) * rso : 0.f;
    if (dhh != nullptr) {
      float d[R::E];
      R::load(d, dhh + row * lddhh, lane, cols);
      const float dec = decay ? decay[row] : 1.f;
      const float muh = mean_h[row], rsh = rstd_h[row];
      float s1 = 0.f, s2 = 0.f;
      float xh[R::E];
#pragma unroll
      for (int e = 0; e < R::E; ++e) {
        const bool ok = R::valid(e, lane, cols);
        d[e] *= dec;
        xh[e] = ok ? ((xo[e] * g_o[e] + b_o[e]) - muh) * rsh : 0.f;
        const float gg = ok ? d[e] * g_h[e] : 0.f;
        s1 += gg;
        s2 = fmaf(gg, xh[e], s2);
        acc[2][e] = fmaf(d[e], xh[e], acc[2][e]);
        acc[3][e] += d[e];
      }
      s1 = warp_sum(s1) * inv;
      s2 = warp_sum(s2) * inv;
#pragma unroll
      for (int e = 0; e < R::E; ++e)
        if (R::valid(e, lane, cols)) ds[e] += rsh * (d[e] * g_h[e] - s1 - xh[e] * s2);
    }
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int e = 0; e < R::E; ++e) {
      const float gg = R::valid(e, lane, cols) ? ds[e] * g_o[e] : 0.f;
      t1 += gg;
      t2 = fmaf(gg, xo[e], t2);
      acc[0][e] = fmaf(ds[e], xo[e], acc[0][e]);
      acc[1][e] += ds[e];
    }
    t1 = warp_sum(t1) * inv;
    t2 = warp_sum(t2) * inv;
    float o[R::E];
#pragma unroll
    for (int e = 0; e < R::E; ++e) o[e] = rso * (ds[e] * g_o[e] - t1 - xo[e] * t2);
    if (bz == nullptr) {
      R::store(dhn + row * lddhn, o, lane, cols);
    } else {
      // autograd of :538-542 on dhn = o: d z_pre = dhn (cand - h^) z(1-z); d cand_pre = dhn z (1 - cand^2); d h^ = dhn (1-z)
      float a[R::E];
#pragma unroll
      for (int e = 0; e < R::E; ++e) a[e] = o[e] * (cv[e] - bv[e]) * zv[e] * (1.f - zv[e]);
      R::store(dgz + row * lddg, a, lane, cols);
#pragma unroll
      for (int e = 0; e < R::E; ++e) a[e] = o[e] * zv[e] * (1.f - cv[e] * cv[e]);
      R::store(dgc + row * lddg, a, lane, cols);
#pragma unroll
      for (int e = 0; e < R::E; ++e) a[e] = o[e] * (1.f - zv[e]);
      R::store(dbase + row * lddbase, a, lane, cols);
    }
  }
  block_partials<NV, VEC, 4>(acc, red, partial, cols);
}

// ---------------------------------------------------------------------------------------------------------------
// GELU + LayerNorm
// ---------------------------------------------------------------------------------------------------------------
template <int NV, bool VEC>
__global__ void __launch_bounds__(RW * 32)
gelu_ln_fwd_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ y, int64_t ldy, float* __restrict__ mean, float* __restrict__ rstd, int64_t rows,
                   int cols) {
  using R = Row<NV, VEC>;
  const int64_t row = (int64_t)blockIdx.x * RW + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float v[R::E], g[R::E], b[R::E];
  R::load(v, a + row * lda, lane, cols);
  R::load(g, gamma, lane, cols);
  R::load(b, beta, lane, cols);
#pragma unroll
  for (int e = 0; e < R::E; ++e) v[e] = R::valid(e, lane, cols) ? gelu_f(v[e]) : 0.f;
  float mu, rs;
  row_stats<NV, VEC>(v, lane, cols, mu, rs);
#pragma unroll
  for (int e = 0; e < R::E; ++e) v[e] = (v[e] - mu) * rs * g[e] + b[e];
  R::store(y + row * ldy, v, lane, cols);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
}

template <int NV, bool VEC>
__global__ void __launch_bounds__(RW * 32)
gelu_ln_bwd_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ a, int64_t lda,
                   const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                   float* __restrict__ da, int64_t ldda, float* __restrict__ partial, int64_t rows, int cols,
                   int64_t rows_per_block) {
  using R = Row<NV, VEC>;
  extern __shared__ float red[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[2][R::E];
#pragma unroll
  for (int e = 0; e < R::E; ++e) { acc[0][e] = 0.f; acc[1][e] = 0.f; }
  float g[R::E];
  R::load(g, gamma, lane, cols);
  const float inv = 1.f / (float)cols;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  for (int64_t row = r0 + w; row < r1; row += RW) {
    float av[R::E], d[R::E], xh[R::E];
    R::load(av, a + row * lda, lane, cols);
    R::load(d, dy + row * lddy, lane, cols);
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < R::E; ++e) {
      const bool ok = R::valid(e, lane, cols);
      xh[e] = ok ? (gelu_f(av[e]) - mu) * rs : 0.f;
      const float gg = ok ? d[e] * g[e] : 0.f;
      s1 += gg;
      s2 = fmaf(gg, xh[e], s2);
      acc[0][e] = fmaf(d[e], xh[e], acc[0][e]);
      acc[1][e] += d[e];
    }
    s1 = warp_sum(s1) * inv;
    s2 = warp_sum(s2) * inv;
    float o[R::E];
#pragma unroll
    for (int e = 0; e < R::E; ++e) o[e] = rs * (d[e] * g[e] - s1 - xh[e] * s2) * gelu_grad(av[e]);
    R::store(da + row * ldda, o, lane, cols);
  }
  block_partials<NV, VEC, 2>(acc, red, partial, cols);
}

// ---------------------------------------------------------------------------------------------------------------
// sliding window (mean / sum over |u-t| <= W) + GELU, float4 columns; a thread walks t with its window in registers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

template <int W>
__global__ void window_gelu_fwd_kernel(const float* __restrict__ p, float* __restrict__ out, int T, int64_t inner4, int agg) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= inner4) return;
  const float4* p4 = reinterpret_cast<const float4*>(p);
  float4* o4 = reinterpret_cast<float4*>(out);
  float4 win[2 * W + 1];
#pragma unroll
  for (int k = 0; k <= 2 * W; ++k) win[k] = (k - W >= 0 && k - W < T) ? p4[(int64_t)(k - W) * inner4 + x] : f4zero();
  for (int t = 0; t < T; ++t) {
    // ascending u from max(0,t-W), as torch.stack(...).mean(0) sums; out-of-range slots hold 0
    float4 s = f4zero();
#pragma unroll
    for (int k = 0; k <= 2 * W; ++k) s = f4add(s, win[k]);
    if (agg == 0) {
      const float c = (float)(min(T, t + W + 1) - max(0, t - W));
      s.x /= c; s.y /= c; s.z /= c; s.w /= c;
    }
    o4[(int64_t)t * inner4 + x] = make_float4(gelu_f(s.x), gelu_f(s.y), gelu_f(s.z), gelu_f(s.w));
#pragma unroll
    for (int k = 0; k < 2 * W; ++k) win[k] = win[k + 1];
    win[2 * W] = (t + 1 + W < T) ? p4[(int64_t)(t + 1 + W) * inner4 + x] : f4zero();
  }
}

// dp[u] = sum_{|t-u|<=W} dgg[t] * gelu'(agg[t]) * weight(t); agg[t] is rebuilt from p (never stored)
template <int W>
__global__ void window_gelu_bwd_kernel(const float* __restrict__ dgg, const float* __restrict__ p, float* __restrict__ dp,
                                       int T, int64_t inner4, int agg) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= inner4) return;
  const float4* p4 = reinterpret_cast<const float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(dgg);
  float4* o4 = reinterpret_cast<float4*>(dp);
  float4 pw[2 * W + 1];                      // p[t-W .. t+W] for the step t being differentiated
  float4 ew[2 * W + 1];                      // e[u-W .. u+W] for the output row u = t - W
#pragma unroll
  for (int k = 0; k <= 2 * W; ++k) { pw[k] = (k - W >= 0 && k - W < T) ? p4[(int64_t)(k - W) * inner4 + x] : f4zero(); ew[k] = f4zero(); }
  for (int t = 0; t < T + W; ++t) {
    float4 e = f4zero();
    if (t < T) {
      float4 s = f4zero();
#pragma unroll
      for (int k = 0; k <= 2 * W; ++k) s = f4add(s, pw[k]);
      float c = 1.f;
      if (agg == 0) { c = (float)(min(T, t + W + 1) - max(0, t - W)); s.x /= c; s.y /= c; s.z /= c; s.w /= c; }
      const float4 d = g4[(int64_t)t * inner4 + x];
      e = make_float4(d.x * gelu_grad(s.x) / c, d.y * gelu_grad(s.y) / c, d.z * gelu_grad(s.z) / c, d.w * gelu_grad(s.w) / c);
#pragma unroll
      for (int k = 0; k < 2 * W; ++k) pw[k] = pw[k + 1];
      pw[2 * W] = (t + 1 + W < T) ? p4[(int64_t)(t + 1 + W) * inner4 + x] : f4zero();
    }
#pragma unroll
    for (int k = 0; k < 2 * W; ++k) ew[k] = ew[k + 1];
    ew[2 * W] = e;                           // ew now holds e[t-2W .. t]
    const int u = t - W;
    if (u >= 0) {
      float4 s = f4zero();
#pragma unroll
      for (int k = 0; k <= 2 * W; ++k) s = f4add(s, ew[k]);
      o4[(int64_t)u * inner4 + x] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// mean of squares
// ---------------------------------------------------------------------------------------------------------------
constexpr int MSE_BLOCKS = 1184;             // 8 per SM
__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const float* __restrict__ x, int64_t n4, int64_t n, float* __restrict__ partial) {
  __shared__ float red[8];
  float s = 0.f;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
    const float4 v = x4[i];
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  if (blockIdx.x == 0)
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += 256) s = fmaf(x[i], x[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = red[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += red[k];
    partial[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256)
sumsq_final_kernel(const float* __restrict__ partial, int parts, float scale, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < parts; i += 256) s += partial[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = red[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += red[k];
    *out = t * scale;
  }
}
__global__ void scale_by_dev_scalar_kernel(const float* __restrict__ x, const float* __restrict__ g, float coeff,
                                           float* __restrict__ out, int64_t n4, int64_t n) {
  const float c = (g ? *g : 1.f) * coeff;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<float4*>(out)[i] = make_float4(v.x * c, v.y * c, v.z * c, v.w * c);
  }
  if (i == 0)
    for (int64_t j = n4 * 4; j < n; ++j) out[j] = x[j] * c;
}

// reset-gate backward of the GRU step: dg_r = d(rs) h^ r(1-r); dhh += d(rs) r        (autograd of :531-538)
__global__ void gru_reset_bwd_kernel(const float* __restrict__ drs, const float* __restrict__ r, const float* __restrict__ hh,
                                     int64_t ldhh, float* __restrict__ dgr, int64_t lddg, float* __restrict__ dhh, int64_t lddhh,
                                     int64_t rows, int H4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H4) return;
  const int64_t row = i / H4;
  const int c = (int)(i - row * H4) * 4, H = H4 * 4;
  const float4 d = *reinterpret_cast<const float4*>(drs + row * H + c), rv = *reinterpret_cast<const float4*>(r + row * H + c);
  const float4 b = *reinterpret_cast<const float4*>(hh + row * ldhh + c);
  float4* op = reinterpret_cast<float4*>(dhh + row * lddhh + c);
  const float4 o = *op;
  *reinterpret_cast<float4*>(dgr + row * lddg + c) =
      make_float4(d.x * b.x * rv.x * (1.f - rv.x), d.y * b.y * rv.y * (1.f - rv.y), d.z * b.z * rv.z * (1.f - rv.z),
                  d.w * b.w * rv.w * (1.f - rv.w));
  *op = make_float4(o.x + d.x * rv.x, o.y + d.y * rv.y, o.z + d.z * rv.z, o.w + d.w * rv.w);
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
int parts_for(int64_t rows) {
  int64_t p = (rows + 63) / 64;
  if (p > 592) p = 592;
  if (p < 1) p = 1;
  return (int)p;
}

}  // namespace

// dispatch on the row layout: float4 lanes for cols in {128, 256} with aligned operands, scalar lanes (cols <= 512) otherwise
#define ROW_DISPATCH(vec_ok, cols, CALL)                                     \
  do {                                                                       \
    if ((vec_ok) && (cols) == 128) { CALL(1, true); }                        \
    else if ((vec_ok) && (cols) == 256) { CALL(2, true); }                   \
    else if ((cols) <= 64) { CALL(2, false); }                               \
    else { CALL(16, false); }                                                \
  } while (0)

static int ln_pair_fwd_impl(const float* hn, int64_t ldhn, const float* bz, const float* bbase, int64_t ldbase, float* cand,
                           const float* gamma_o, const float* beta_o, const float* gamma_h, const float* beta_h,
                           const float* ts, int64_t ldts, int32_t t_hi, float* s, int64_t lds, float* hhat, int64_t ldhh,
                           float* mean_o, float* rstd_o, float* mean_h, float* rstd_h, float* decay, int64_t rows, int32_t cols,
                           tagan_stream_t stream) {
  if (!hn || !gamma_o || !beta_o || !s || !mean_o || !rstd_o || rows < 0 || cols <= 0) return TAGAN_E_INVALID;
  if (hhat && (!gamma_h || !beta_h || !mean_h || !rstd_h)) return TAGAN_E_INVALID;
  if (bz && (!bbase || !cand)) return TAGAN_E_INVALID;
  if (ts && t_hi < 1) return TAGAN_E_INVALID;
  if (cols > 512) return TAGAN_E_UNSUPPORTED;
  if (rows == 0) return 0;
  const bool vec = ldhn % 4 == 0 && lds % 4 == 0 && (!hhat || ldhh % 4 == 0) && al16(hn) && al16(s) && al16(hhat) &&
                   al16(gamma_o) && al16(beta_o) && al16(gamma_h) && al16(beta_h) &&
                   (!bz || (al16(bz) && al16(bbase) && al16(cand) && ldbase % 4 == 0));
  cudaStream_t st = as_stream(stream);
  const unsigned grid = ceil_div_i64(rows, RW);
#define CALL(NV, VEC) ln_pair_fwd_kernel<NV, VEC><<<grid, RW * 32, 0, st>>>(hn, ldhn, bz, bbase, ldbase, cand, gamma_o, beta_o, gamma_h, \
      beta_h, ts, ldts, t_hi, s, lds, hhat, ldhh, mean_o, rstd_o, mean_h, rstd_h, decay, rows, cols)
  ROW_DISPATCH(vec, cols, CALL);
#undef CALL
  return tagan_launch_status();
}

TAGAN_API int tagan_ln_pair_fwd(const float* hn, int64_t ldhn, const float* gamma_o, const float* beta_o,
                                const float* gamma_h, const float* beta_h, const float* ts, int64_t ldts, int32_t t_hi,
                                float* s, int64_t lds, float* hhat, int64_t ldhh, float* mean_o, float* rstd_o,
                                float* mean_h, float* rstd_h, float* decay, int64_t rows, int32_t cols,
                                tagan_stream_t stream) {
  return ln_pair_fwd_impl(hn, ldhn, nullptr, nullptr, 0, nullptr, gamma_o, beta_o, gamma_h, beta_h, ts, ldts, t_hi, s, lds, hhat,
                          ldhh, mean_o, rstd_o, mean_h, rstd_h, decay, rows, cols, stream);
}

TAGAN_API int tagan_gru_blend_ln_fwd(const float* cand_pre, int64_t ldc, const float* z, const float* hhat_cur, int64_t ldcur,
                                     float* cand, const float* gamma_o, const float* beta_o, const float* gamma_h,
                                     const float* beta_h, const float* ts, int64_t ldts, int32_t t_hi, float* s, int64_t lds,
                                     float* hhat_next, int64_t ldhh, float* mean_o, float* rstd_o, float* mean_h, float* rstd_h,
                                     float* decay, int64_t rows, int32_t cols, tagan_stream_t stream) {
  if (!z || !hhat_cur || !cand) return TAGAN_E_INVALID;
  return ln_pair_fwd_impl(cand_pre, ldc, z, hhat_cur, ldcur, cand, gamma_o, beta_o, gamma_h, beta_h, ts, ldts, t_hi, s, lds,
                          hhat_next, ldhh, mean_o, rstd_o, mean_h, rstd_h, decay, rows, cols, stream);
}

TAGAN_API size_t tagan_ln_pair_bwd_workspace_bytes(int64_t rows, int32_t cols) {
  return (size_t)parts_for(rows) * 4 * (size_t)cols * sizeof(float);
}

static int ln_pair_bwd_impl(const float* ds_ext, int64_t ldds, const float* dhh, int64_t lddhh, const float* hn, int64_t ldhn,
                           const float* bz, const float* bbase, int64_t ldbase, float* dgz, float* dgc, int64_t lddg,
                           float* dbase, int64_t lddbase, const float* gamma_o, const float* beta_o, const float* gamma_h,
                           const float* mean_o, const float* rstd_o, const float* mean_h, const float* rstd_h,
                           const float* decay, float* dhn, int64_t lddhn, float* daffine, int32_t accumulate, void* workspace,
                           size_t workspace_bytes, int64_t rows, int32_t cols, tagan_stream_t stream) {
  if (!hn || !gamma_o || !beta_o || !mean_o || !rstd_o || !daffine || rows < 0 || cols <= 0) return TAGAN_E_INVALID;
  if (!bz && !dhn) return TAGAN_E_INVALID;
  if (bz && (!bbase || !dgz || !dgc || !dbase)) return TAGAN_E_INVALID;
  if (dhh && (!gamma_h || !mean_h || !rstd_h)) return TAGAN_E_INVALID;
  if (cols > 512) return TAGAN_E_UNSUPPORTED;
  if (!workspace || workspace_bytes < tagan_ln_pair_bwd_workspace_bytes(rows, cols)) return TAGAN_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (rows == 0) { if (!accumulate) cudaMemsetAsync(daffine, 0, sizeof(float) * 4 * cols, st); return 0; }
  const int parts = parts_for(rows);
  const int64_t rpb = (rows + parts - 1) / parts;
  float* part = static_cast<float*>(workspace);
  const bool vec = (!ds_ext || (ldds % 4 == 0 && al16(ds_ext))) && (!dhh || (lddhh % 4 == 0 && al16(dhh))) && ldhn % 4 == 0 &&
                   (!dhn || (lddhn % 4 == 0 && al16(dhn))) && al16(hn) && al16(gamma_o) && al16(beta_o) && al16(gamma_h) &&
                   (!bz || (al16(bz) && al16(bbase) && al16(dgz) && al16(dgc) && al16(dbase) && ldbase % 4 == 0 && lddg % 4 == 0 &&
                            lddbase % 4 == 0));
  const size_t smem = (size_t)RW * 4 * cols * sizeof(float);
#define CALL(NV, VEC) ln_pair_bwd_kernel<NV, VEC><<<parts, RW * 32, smem, st>>>(ds_ext, ldds, dhh, lddhh, hn, ldhn, bz, bbase, ldbase, \
      dgz, dgc, lddg, dbase, lddbase, gamma_o, beta_o, gamma_h, mean_o, rstd_o, mean_h, rstd_h, decay, dhn, lddhn, part, rows, cols, rpb)
  if (smem > 48 * 1024) {          // cols > 384 with four accumulators: opt in once per launch (cheap, idempotent)
    cudaFuncSetAttribute(ln_pair_bwd_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  ROW_DISPATCH(vec, cols, CALL);
#undef CALL
  reduce_cols_kernel<<<(4 * cols + 31) / 32, 256, 0, st>>>(part, parts, 4 * cols, daffine, accumulate);
  return tagan_launch_status();
}

TAGAN_API int tagan_ln_pair_bwd(const float* ds_ext, int64_t ldds, const float* dhh, int64_t lddhh, const float* hn,
                                int64_t ldhn, const float* gamma_o, const float* beta_o, const float* gamma_h,
                                const float* mean_o, const float* rstd_o, const float* mean_h, const float* rstd_h,
                                const float* decay, float* dhn, int64_t lddhn, float* daffine /*[4][cols]*/,
                                int32_t accumulate, void* workspace, size_t workspace_bytes, int64_t rows, int32_t cols,
                                tagan_stream_t stream) {
  if (!dhn) return TAGAN_E_INVALID;
  return ln_pair_bwd_impl(ds_ext, ldds, dhh, lddhh, hn, ldhn, nullptr, nullptr, 0, nullptr, nullptr, 0, nullptr, 0, gamma_o, beta_o,
                          gamma_h, mean_o, rstd_o, mean_h, rstd_h, decay, dhn, lddhn, daffine, accumulate, workspace,
                          workspace_bytes, rows, cols, stream);
}

TAGAN_API int tagan_gru_blend_ln_bwd(const float* ds_ext, int64_t ldds, const float* dhh_next, int64_t lddhh, const float* cand,
                                     const float* z, const float* hhat_cur, int64_t ldcur, float* dgz, float* dgc, int64_t lddg,
                                     float* dhh_cur, int64_t lddcur, const float* gamma_o, const float* beta_o,
                                     const float* gamma_h, const float* mean_o, const float* rstd_o, const float* mean_h,
                                     const float* rstd_h, const float* decay, float* daffine, int32_t accumulate, void* workspace,
                                     size_t workspace_bytes, int64_t rows, int32_t cols, tagan_stream_t stream) {
  if (!cand || !z || !hhat_cur) return TAGAN_E_INVALID;
  return ln_pair_bwd_impl(ds_ext, ldds, dhh_next, lddhh, cand, cols, z, hhat_cur, ldcur, dgz, dgc, lddg, dhh_cur, lddcur, gamma_o,
                          beta_o, gamma_h, mean_o, rstd_o, mean_h, rstd_h, decay, nullptr, 0, daffine, accumulate, workspace,
                          workspace_bytes, rows, cols, stream);
}

TAGAN_API int tagan_gru_reset_bwd(const float* drs, const float* r, const float* hhat, int64_t ldhh, float* dgr, int64_t lddg,
                                  float* dhh, int64_t lddhh, int64_t rows, int32_t H, tagan_stream_t stream) {
  if (!drs || !r || !hhat || !dgr || !dhh || rows < 0 || H <= 0) return TAGAN_E_INVALID;
  if (H % 4 || ldhh % 4 || lddg % 4 || lddhh % 4 || !al16(drs) || !al16(r) || !al16(hhat) || !al16(dgr) || !al16(dhh))
    return TAGAN_E_UNSUPPORTED;
  if (rows == 0) return 0;
  gru_reset_bwd_kernel<<<ceil_div_i64(rows * (H / 4), 256), 256, 0, as_stream(stream)>>>(drs, r, hhat, ldhh, dgr, lddg, dhh, lddhh,
                                                                                       rows, H / 4);
  return tagan_launch_status();
}

TAGAN_API int tagan_gelu_ln_fwd(const float* a, int64_t lda, const float* gamma, const float* beta, float* y, int64_t ldy,
                                float* mean, float* rstd, int64_t rows, int32_t cols, tagan_stream_t stream) {
  if (!a || !gamma || !beta || !y || !mean || !rstd || rows < 0 || cols <= 0) return TAGAN_E_INVALID;
  if (cols > 512) return TAGAN_E_UNSUPPORTED;
  if (rows == 0) return 0;
  const bool vec = lda % 4 == 0 && ldy % 4 == 0 && al16(a) && al16(y) && al16(gamma) && al16(beta);
  cudaStream_t st = as_stream(stream);
  const unsigned grid = ceil_div_i64(rows, RW);
#define CALL(NV, VEC) gelu_ln_fwd_kernel<NV, VEC><<<grid, RW * 32, 0, st>>>(a, lda, gamma, beta, y, ldy, mean, rstd, rows, cols)
  ROW_DISPATCH(vec, cols, CALL);
#undef CALL
  return tagan_launch_status();
}

TAGAN_API size_t tagan_gelu_ln_bwd_workspace_bytes(int64_t rows, int32_t cols) {
  return (size_t)parts_for(rows) * 2 * (size_t)cols * sizeof(float);
}

TAGAN_API int tagan_gelu_ln_bwd(const float* dy, int64_t lddy, const float* a, int64_t lda, const float* gamma,
                                const float* mean, const float* rstd, float* da, int64_t ldda, float* daffine /*[2][cols]*/,
                                void* workspace, size_t workspace_bytes, int64_t rows, int32_t cols, tagan_stream_t stream) {
  if (!dy || !a || !gamma || !mean || !rstd || !da || !daffine || rows < 0 || cols <= 0) return TAGAN_E_INVALID;
  if (cols > 512) return TAGAN_E_UNSUPPORTED;
  if (!workspace || workspace_bytes < tagan_gelu_ln_bwd_workspace_bytes(rows, cols)) return TAGAN_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (rows == 0) { cudaMemsetAsync(daffine, 0, sizeof(float) * 2 * cols, st); return 0; }
  const int parts = parts_for(rows);
  const int64_t rpb = (rows + parts - 1) / parts;
  float* part = static_cast<float*>(workspace);
  const bool vec = lddy % 4 == 0 && lda % 4 == 0 && ldda % 4 == 0 && al16(dy) && al16(a) && al16(da) && al16(gamma);
  const size_t smem = (size_t)RW * 2 * cols * sizeof(float);
#define CALL(NV, VEC) gelu_ln_bwd_kernel<NV, VEC><<<parts, RW * 32, smem, st>>>(dy, lddy, a, lda, gamma, mean, rstd, da, ldda, part, \
      rows, cols, rpb)
  ROW_DISPATCH(vec, cols, CALL);
#undef CALL
  reduce_cols_kernel<<<(2 * cols + 31) / 32, 256, 0, st>>>(part, parts, 2 * cols, daffine, 0);
  return tagan_launch_status();
}

TAGAN_API int tagan_window_gelu_fwd(const float* p, float* out, int32_t T, int64_t inner, int32_t window, int32_t agg,
                                    tagan_stream_t stream) {
  if (!p || !out || T < 0 || inner < 0 || (agg != 0 && agg != 2)) return TAGAN_E_INVALID;
  if (inner % 4 || window < 1 || window > 4 || !al16(p) || !al16(out)) return TAGAN_E_UNSUPPORTED;
  if (T == 0 || inner == 0) return 0;
  const unsigned grid = ceil_div_i64(inner / 4, 256);
  cudaStream_t st = as_stream(stream);
  switch (window) {
    case 1: window_gelu_fwd_kernel<1><<<grid, 256, 0, st>>>(p, out, T, inner / 4, agg); break;
    case 2: window_gelu_fwd_kernel<2><<<grid, 256, 0, st>>>(p, out, T, inner / 4, agg); break;
    case 3: window_gelu_fwd_kernel<3><<<grid, 256, 0, st>>>(p, out, T, inner / 4, agg); break;
    default: window_gelu_fwd_kernel<4><<<grid, 256, 0, st>>>(p, out, T, inner / 4, agg); break;
  }
  return tagan_launch_status();
}

TAGAN_API int tagan_window_gelu_bwd(const float* dgg, const float* p, float* dp, int32_t T, int64_t inner, int32_t window,
                                    int32_t agg, tagan_stream_t stream) {
  if (!dgg || !p || !dp || T < 0 || inner < 0 || (agg != 0 && agg != 2)) return TAGAN_E_INVALID;
  if (inner % 4 || window < 1 || window > 4 || !al16(p) || !al16(dgg) || !al16(dp)) return TAGAN_E_UNSUPPORTED;
  if (T == 0 || inner == 0) return 0;
  const unsigned grid = ceil_div_i64(inner / 4, 256);
  cudaStream_t st = as_stream(stream);
  switch (window) {
    case 1: window_gelu_bwd_kernel<1><<<grid, 256, 0, st>>>(dgg, p, dp, T, inner / 4, agg); break;
    case 2: window_gelu_bwd_kernel<2><<<grid, 256, 0, st>>>(dgg, p, dp, T, inner / 4, agg); break;
    case 3: window_gelu_bwd_kernel<3><<<grid, 256, 0, st>>>(dgg, p, dp, T, inner / 4, agg); break;
    default: window_gelu_bwd_kernel<4><<<grid, 256, 0, st>>>(dgg, p, dp, T, inner / 4, agg); break;
  }
  return tagan_launch_status();
}

TAGAN_API size_t tagan_mse_workspace_bytes(void) { return (size_t)MSE_BLOCKS * sizeof(float); }

TAGAN_API int tagan_mse_fwd(const float* x, int64_t n, float* loss, void* workspace, size_t workspace_bytes,
                            tagan_stream_t stream) {
  if (!x || !loss || n <= 0) return TAGAN_E_INVALID;
  if (!al16(x)) return TAGAN_E_UNSUPPORTED;
  if (!workspace || workspace_bytes < tagan_mse_workspace_bytes()) return TAGAN_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > MSE_BLOCKS) blocks = MSE_BLOCKS;
  if (blocks < 1) blocks = 1;
  sumsq_partial_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n / 4, n, static_cast<float*>(workspace));
  sumsq_final_kernel<<<1, 256, 0, st>>>(static_cast<const float*>(workspace), (int)blocks, 1.f / (float)n, loss);
  return tagan_launch_status();
}

TAGAN_API int tagan_mse_bwd(const float* x, int64_t n, const float* dloss, float* dx, tagan_stream_t stream) {
  if (!x || !dx || n <= 0) return TAGAN_E_INVALID;
  if (!al16(x) || !al16(dx)) return TAGAN_E_UNSUPPORTED;
  const int64_t n4 = n / 4;
  scale_by_dev_scalar_kernel<<<ceil_div_i64(n4 > 0 ? n4 : 1, 256), 256, 0, as_stream(stream)>>>(x, dloss, 2.f / (float)n, dx, n4, n);
  return tagan_launch_status();
}
