// Row-wise building blocks: LayerNorm (eps 1e-5, biased variance, as nn.LayerNorm) forward /
// backward with fused residual add and per-row scale, deterministic column sums (bias and
// LayerNorm-affine gradients), and small fused element-wise ops.  All HBM-bound streams:
// one warp per row; rows are re-read from L1 rather than held in registers so any width works.
#include "common.cuh"

namespace {

constexpr int LN_WARPS = 8;
constexpr int MAX_LN_COLS = 512;
constexpr float LN_EPS = 1e-5f;

// y = ((x+res) - mean) * rstd * gamma + beta, optionally * rowscale[row].
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ res, int64_t ldres,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ rowscale, float* __restrict__ y, int64_t ldy,
                     float* __restrict__ sum_out, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                     int64_t rows, int cols) {
  const int64_t row = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * ldx;
  const float* rr = res ? res + row * ldres : nullptr;
  float* yr = y + row * ldy;
  float* sr = sum_out ? sum_out + row * (int64_t)cols : nullptr;
  const float rs = rowscale ? rowscale[row] : 1.f;
  if (gamma == nullptr) {   // use_layer_norm = False: y = (x + res) * rowscale
    for (int c = lane; c < cols; c += 32) {
      float v = rr ? xr[c] + rr[c] : xr[c];
      if (sr) sr[c] = v;
      yr[c] = v * rs;
    }
    return;
  }
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) {
    float v = rr ? xr[c] + rr[c] : xr[c];
    if (sr) sr[c] = v;
    s += v;
  }
  const float mean = warp_sum(s) / (float)cols;
  float q = 0.f;
  for (int c = lane; c < cols; c += 32) {
    float v = (rr ? xr[c] + rr[c] : xr[c]) - mean;
    q = fmaf(v, v, q);
  }
  const float rstd = 1.f / sqrtf(warp_sum(q) / (float)cols + LN_EPS);
  for (int c = lane; c < cols; c += 32) {
    float v = (rr ? xr[c] + rr[c] : xr[c]);
    yr[c] = ((v - mean) * rstd * gamma[c] + beta[c]) * rs;
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g*xhat)), g = dy*rowscale*gamma; per-block partial
// dgamma/dbeta accumulated in a fixed order (rows ascending inside a warp, warps summed in order).
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ xsum, int64_t ldx,
                     const float* __restrict__ gamma, const float* __restrict__ rowscale,
                     const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dx,
                     int64_t lddx, int accumulate, float* __restrict__ partial /*[parts][2][cols]*/,
                     int64_t rows, int cols, int64_t rows_per_block) {
  __shared__ float acc[LN_WARPS][2][MAX_LN_COLS];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = lane; c < cols; c += 32) { acc[w][0][c] = 0.f; acc[w][1][c] = 0.f; }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  for (int64_t row = r0 + w; row < r1; row += LN_WARPS) {
    const float* dyr = dy + row * lddy;
    float* dxr = dx + row * lddx;
    const float rs = rowscale ? rowscale[row] : 1.f;
    if (gamma == nullptr) {
      for (int c = lane; c < cols; c += 32) {
        float g = dyr[c] * rs;
        dxr[c] = accumulate ? dxr[c] + g : g;
      }
      continue;
    }
    const float* xr = xsum + row * ldx;
    const float mu = mean[row], rsd = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < cols; c += 32) {
      float d = dyr[c] * rs;
      float xh = (xr[c] - mu) * rsd;
      float g = d * gamma[c];
      s1 += g;
      s2 = fmaf(g, xh, s2);
      acc[w][0][c] = fmaf(d, xh, acc[w][0][c]);
      acc[w][1][c] += d;
    }
    s1 = warp_sum(s1) / (float)cols;
    s2 = warp_sum(s2) / (float)cols;
    for (int c = lane; c < cols; c += 32) {
      float xh = (xr[c] - mu) * rsd;
      float g = dyr[c] * rs * gamma[c];
      float v = rsd * (g - s1 - xh * s2);
      dxr[c] = accumulate ? dxr[c] + v : v;
    }
  }
  __syncthreads();
  if (partial != nullptr && gamma != nullptr) {
    for (int c = threadIdx.x; c < cols; c += LN_WARPS * 32) {
      float a = 0.f, b = 0.f;
#pragma unroll
      for (int k = 0; k < LN_WARPS; ++k) { a += acc[k][0][c]; b += acc[k][1][c]; }
      partial[((int64_t)blockIdx.x * 2 + 0) * cols + c] = a;
      partial[((int64_t)blockIdx.x * 2 + 1) * cols + c] = b;
    }
  }
}

__global__ void reduce_partials2(const float* __restrict__ partial, int parts, int cols, float* __restrict__ out0,
                                 float* __restrict__ out1) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float a = 0.f, b = 0.f;
  for (int p = 0; p < parts; ++p) {
    a += partial[((int64_t)p * 2 + 0) * cols + c];
    b += partial[((int64_t)p * 2 + 1) * cols + c];
  }
  if (out0) out0[c] = a;
  if (out1) out1[c] = b;
}

__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ partial, int64_t rows, int cols,
                      int64_t rows_per_block) {
  // thread t owns columns t, t+256, ...; rows walked in order -> deterministic
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  for (int c = threadIdx.x; c < cols; c += 256) {
    float s = 0.f;
    for (int64_t r = r0; r < r1; ++r) s += x[r * ldx + c];
    partial[(int64_t)blockIdx.x * cols + c] = s;
  }
}
__global__ void reduce_partials1(const float* __restrict__ partial, int parts, int cols, float* __restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float a = 0.f;
  for (int p = 0; p < parts; ++p) a += partial[(int64_t)p * cols + c];
  out[c] = a;
}


// ---- vectorised fast paths: cols = 128*NV, 16-byte aligned rows; a lane keeps its NV float4 of the row in
// registers, so every element is read once and the two reductions are pure shuffles.
template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_vec_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ res, int64_t ldres,
                         const float* __restrict__ gamma, const float* __restrict__ beta,
                         const float* __restrict__ rowscale, float* __restrict__ y, int64_t ldy,
                         float* __restrict__ sum_out, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                         int64_t rows) {
  constexpr int COLS = 128 * NV;
  const int64_t row = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j] = __ldg(reinterpret_cast<const float4*>(x + row * ldx + j * 128 + lane * 4));
    if (res) {
      float4 r4 = __ldg(reinterpret_cast<const float4*>(res + row * ldres + j * 128 + lane * 4));
      v[j].x += r4.x; v[j].y += r4.y; v[j].z += r4.z; v[j].w += r4.w;
    }
    if (sum_out) *reinterpret_cast<float4*>(sum_out + row * (int64_t)COLS + j * 128 + lane * 4) = v[j];
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float rs = rowscale ? rowscale[row] : 1.f;
  if (gamma == nullptr) {
#pragma unroll
    for (int j = 0; j < NV; ++j)
      *reinterpret_cast<float4*>(y + row * ldy + j * 128 + lane * 4) =
          make_float4(v[j].x * rs, v[j].y * rs, v[j].z * rs, v[j].w * rs);
    return;
  }
  const float mean = warp_sum(s) / (float)COLS;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
    q = fmaf(a, a, q); q = fmaf(b, b, q); q = fmaf(c, c, q); q = fmaf(d, d, q);
  }
  const float rstd = 1.f / sqrtf(warp_sum(q) / (float)COLS + LN_EPS);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + j * 128 + lane * 4));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + j * 128 + lane * 4));
    float4 o;
    o.x = ((v[j].x - mean) * rstd * g4.x + b4.x) * rs;
    o.y = ((v[j].y - mean) * rstd * g4.y + b4.y) * rs;
    o.z = ((v[j].z - mean) * rstd * g4.z + b4.z) * rs;
    o.w = ((v[j].w - mean) * rstd * g4.w + b4.w) * rs;
    *reinterpret_cast<float4*>(y + row * ldy + j * 128 + lane * 4) = o;
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

template <int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_vec_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ xsum, int64_t ldx,
                         const float* __restrict__ gamma, const float* __restrict__ rowscale,
                         const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dx,
                         int64_t lddx, int accumulate, float* __restrict__ partial, int64_t rows,
                         int64_t rows_per_block) {
  constexpr int COLS = 128 * NV;
  __shared__ float4 red[LN_WARPS][2][NV][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 ag[NV], ab[NV], g4[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    ag[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    g4[j] = __ldg(reinterpret_cast<const float4*>(gamma + j * 128 + lane * 4));
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  for (int64_t row = r0 + w; row < r1; row += LN_WARPS) {
    const float rs = rowscale ? rowscale[row] : 1.f;
    const float mu = mean[row], rsd = rstd[row];
    float4 d[NV], xh[NV];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      d[j] = __ldg(reinterpret_cast<const float4*>(dy + row * lddy + j * 128 + lane * 4));
      const float4 xv = __ldg(reinterpret_cast<const float4*>(xsum + row * ldx + j * 128 + lane * 4));
      d[j].x *= rs; d[j].y *= rs; d[j].z *= rs; d[j].w *= rs;
      xh[j] = make_float4((xv.x - mu) * rsd, (xv.y - mu) * rsd, (xv.z - mu) * rsd, (xv.w - mu) * rsd);
      const float gx = d[j].x * g4[j].x, gy = d[j].y * g4[j].y, gz = d[j].z * g4[j].z, gw = d[j].w * g4[j].w;
      s1 += (gx + gy) + (gz + gw);
      s2 = fmaf(gx, xh[j].x, s2); s2 = fmaf(gy, xh[j].y, s2); s2 = fmaf(gz, xh[j].z, s2); s2 = fmaf(gw, xh[j].w, s2);
      ag[j].x = fmaf(d[j].x, xh[j].x, ag[j].x); ag[j].y = fmaf(d[j].y, xh[j].y, ag[j].y);
      ag[j].z = fmaf(d[j].z, xh[j].z, ag[j].z); ag[j].w = fmaf(d[j].w, xh[j].w, ag[j].w);
      ab[j].x += d[j].x; ab[j].y += d[j].y; ab[j].z += d[j].z; ab[j].w += d[j].w;
    }
    s1 = warp_sum(s1) / (float)COLS;
    s2 = warp_sum(s2) / (float)COLS;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float4 o;
      o.x = rsd * (d[j].x * g4[j].x - s1 - xh[j].x * s2);
      o.y = rsd * (d[j].y * g4[j].y - s1 - xh[j].y * s2);
      o.z = rsd * (d[j].z * g4[j].z - s1 - xh[j].z * s2);
      o.w = rsd * (d[j].w * g4[j].w - s1 - xh[j].w * s2);
      float4* op = reinterpret_cast<float4*>(dx + row * lddx + j * 128 + lane * 4);
      if (accumulate) { const float4 p4 = *op; o.x += p4.x; o.y += p4.y; o.z += p4.z; o.w += p4.w; }
      *op = o;
    }
  }
  if (partial == nullptr) return;
#pragma unroll
  for (int j = 0; j < NV; ++j) { red[w][0][j][lane] = ag[j]; red[w][1][j][lane] = ab[j]; }
  __syncthreads();
  // fixed-order sum over the 8 warps; thread t handles float4 slot t of the [2][NV][32] table
  for (int t = threadIdx.x; t < 2 * NV * 32; t += LN_WARPS * 32) {
    const int which = t / (NV * 32), rem = t - which * NV * 32;
    float4 a = red[0][which][rem / 32][rem % 32];
#pragma unroll
    for (int k = 1; k < LN_WARPS; ++k) {
      const float4 b = red[k][which][rem / 32][rem % 32];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    *reinterpret_cast<float4*>(partial + ((int64_t)blockIdx.x * 2 + which) * COLS + (rem / 32) * 128 + (rem % 32) * 4) = a;
  }
}

// column sums, vectorised: warp w of a block walks rows r0+w, r0+w+8, ...; lanes own float4 columns
template <int NV>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ partial, int64_t rows,
                  int64_t rows_per_block) {
  constexpr int COLS = 128 * NV;
  __shared__ float4 red[8][NV][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 a[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  for (int64_t row = r0 + w; row < r1; row += 8) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + row * ldx + j * 128 + lane * 4));
      a[j].x += v.x; a[j].y += v.y; a[j].z += v.z; a[j].w += v.w;
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) red[w][j][lane] = a[j];
  __syncthreads();
  for (int t = threadIdx.x; t < NV * 32; t += 256) {
    float4 s = red[0][t / 32][t % 32];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 b = red[k][t / 32][t % 32];
      s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
    }
    *reinterpret_cast<float4*>(partial + (int64_t)blockIdx.x * COLS + (t / 32) * 128 + (t % 32) * 4) = s;
  }
}

// the two affine gradients of a LayerNorm in one launch: partial rows are [dgamma cols | dbeta cols]
__global__ void __launch_bounds__(256)
reduce_parts2_kernel(const float* __restrict__ partial, int parts, int64_t stride, int cols, float* __restrict__ out_a,
                     float* __restrict__ out_b) {
  __shared__ float red[8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), slice = threadIdx.x >> 5;
  float s = 0.f;
  if (c < 2 * cols)
    for (int p = slice; p < parts; p += 8) s += partial[(int64_t)p * stride + c];
  red[slice][threadIdx.x & 31] = s;
  __syncthreads();
  if (slice == 0 && c < 2 * cols) {
    float t = red[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += red[k][threadIdx.x];
    if (c < cols) out_a[c] = t; else out_b[c - cols] = t;
  }
}

// out[c] = sum over parts of partial[p][c] for `nvec` stacked vectors of `cols` (fixed order: 8 interleaved
// slices, then the slices in order) -- 32 columns per block
__global__ void __launch_bounds__(256)
reduce_parts_kernel(const float* __restrict__ partial, int parts, int64_t stride, int cols, float* __restrict__ out) {
  __shared__ float red[8][32];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), slice = threadIdx.x >> 5;
  float s = 0.f;
  if (c < cols)
    for (int p = slice; p < parts; p += 8) s += partial[(int64_t)p * stride + c];
  red[slice][threadIdx.x & 31] = s;
  __syncthreads();
  if (slice == 0 && c < cols) {
    float t = red[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += red[k][threadIdx.x];
    out[c] = t;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int parts_for(int64_t rows) {
  int64_t p = (rows + 63) / 64;
  if (p > 592) p = 592;
  if (p < 1) p = 1;
  return (int)p;
}

__global__ void axpby_kernel(const float* __restrict__ a, float alpha, const float* __restrict__ b, float beta,
                             float* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = alpha * a[i];
  if (b) v = fmaf(beta, b[i], v);
  out[i] = v;
}

__global__ void gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i];
  y[i] = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
}
__global__ void gelu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dx, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i];
  float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752440f));
  float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
  dx[i] = dy[i] * (cdf + v * pdf);
}

__global__ void scale_rows_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ rowscale,
                                  float* __restrict__ y, int64_t ldy, int64_t rows, int cols, int accumulate) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  int64_t r = i / cols;
  int c = (int)(i - r * cols);
  float v = x[r * ldx + c] * (rowscale ? rowscale[r] : 1.f);
  float* o = y + r * ldy + c;
  *o = accumulate ? *o + v : v;
}

__global__ void decay_scale_kernel(const float* __restrict__ ts, int64_t ldts, int t, float* __restrict__ out, int64_t rows) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  float d = ts[i * ldts + t] - ts[i * ldts + t - 1];
  d = fminf(fmaxf(d, 0.f), 10.f);
  out[i] = expf(-d);
}

}  // namespace

TAGAN_API int tagan_layernorm_fwd(const float* x, int64_t ldx, const float* res, int64_t ldres, const float* gamma,
                                  const float* beta, const float* rowscale, float* y, int64_t ldy, float* sum_out,
                                  float* mean, float* rstd, int64_t rows, int32_t cols, tagan_stream_t stream) {
  if (!x || !y || rows < 0 || cols <= 0 || (gamma && !beta)) return TAGAN_E_INVALID;
  if (rows == 0) return 0;
  unsigned grid = ceil_div_i64(rows, LN_WARPS);
  cudaStream_t st = as_stream(stream);
  const bool vec = cols % 128 == 0 && cols <= 512 && ldx % 4 == 0 && ldy % 4 == 0 && (!res || ldres % 4 == 0) &&
                   aligned16(x) && aligned16(y) && (!res || aligned16(res)) && (!sum_out || aligned16(sum_out)) &&
                   (!gamma || (aligned16(gamma) && aligned16(beta)));
  if (vec) {
    switch (cols / 128) {
      case 1: layernorm_fwd_vec_kernel<1><<<grid, LN_WARPS * 32, 0, st>>>(x, ldx, res, ldres, gamma, beta, rowscale, y, ldy, sum_out, mean, rstd, rows); break;
      case 2: layernorm_fwd_vec_kernel<2><<<grid, LN_WARPS * 32, 0, st>>>(x, ldx, res, ldres, gamma, beta, rowscale, y, ldy, sum_out, mean, rstd, rows); break;
      case 3: layernorm_fwd_vec_kernel<3><<<grid, LN_WARPS * 32, 0, st>>>(x, ldx, res, ldres, gamma, beta, rowscale, y, ldy, sum_out, mean, rstd, rows); break;
      default: layernorm_fwd_vec_kernel<4><<<grid, LN_WARPS * 32, 0, st>>>(x, ldx, res, ldres, gamma, beta, rowscale, y, ldy, sum_out, mean, rstd, rows); break;
    }
    return tagan_launch_status();
  }
  layernorm_fwd_kernel<<<grid, LN_WARPS * 32, 0, st>>>(x, ldx, res, ldres, gamma, beta, rowscale, y, ldy,
                                                       sum_out, mean, rstd, rows, cols);
  return tagan_launch_status();
}

TAGAN_API size_t tagan_layernorm_bwd_workspace_bytes(int64_t rows, int32_t cols) {
  return (size_t)parts_for(rows) * 2 * (size_t)cols * sizeof(float);
}

TAGAN_API int tagan_layernorm_bwd(const float* dy, int64_t lddy, const float* xsum, int64_t ldx, const float* gamma,
                                  const float* rowscale, const float* mean, const float* rstd, float* dx, int64_t lddx,
                                  int32_t dx_accumulate, float* dgamma, float* dbeta, void* workspace,
                                  size_t workspace_bytes, int64_t rows, int32_t cols, tagan_stream_t stream) {
  if (!dy || !dx || rows < 0 || cols <= 0) return TAGAN_E_INVALID;
  if (gamma && (!xsum || !mean || !rstd)) return TAGAN_E_INVALID;
  if (cols > MAX_LN_COLS) return TAGAN_E_UNSUPPORTED;
  const bool want_affine = gamma && (dgamma || dbeta);
  if (want_affine && (!workspace || workspace_bytes < tagan_layernorm_bwd_workspace_bytes(rows, cols)))
    return TAGAN_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (rows == 0) {
    if (dgamma) cudaMemsetAsync(dgamma, 0, sizeof(float) * cols, st);
    if (dbeta) cudaMemsetAsync(dbeta, 0, sizeof(float) * cols, st);
    return 0;
  }
  const int parts = parts_for(rows);
  const int64_t rpb = (rows + parts - 1) / parts;
  float* part = want_affine ? (float*)workspace : nullptr;
  const bool vec = gamma && cols % 128 == 0 && cols <= 512 && lddy % 4 == 0 && ldx % 4 == 0 && lddx % 4 == 0 &&
                   aligned16(dy) && aligned16(xsum) && aligned16(dx) && aligned16(gamma);
  if (vec) {
    switch (cols / 128) {
      case 1: layernorm_bwd_vec_kernel<1><<<parts, LN_WARPS * 32, 0, st>>>(dy, lddy, xsum, ldx, gamma, rowscale, mean, rstd, dx, lddx, dx_accumulate, part, rows, rpb); break;
      case 2: layernorm_bwd_vec_kernel<2><<<parts, LN_WARPS * 32, 0, st>>>(dy, lddy, xsum, ldx, gamma, rowscale, mean, rstd, dx, lddx, dx_accumulate, part, rows, rpb); break;
      case 3: layernorm_bwd_vec_kernel<3><<<parts, LN_WARPS * 32, 0, st>>>(dy, lddy, xsum, ldx, gamma, rowscale, mean, rstd, dx, lddx, dx_accumulate, part, rows, rpb); break;
      default: layernorm_bwd_vec_kernel<4><<<parts, LN_WARPS * 32, 0, st>>>(dy, lddy, xsum, ldx, gamma, rowscale, mean, rstd, dx, lddx, dx_accumulate, part, rows, rpb); break;
    }
  } else {
    layernorm_bwd_kernel<<<parts, LN_WARPS * 32, 0, st>>>(dy, lddy, xsum, ldx, gamma, rowscale, mean, rstd, dx, lddx,
                                                          dx_accumulate, part, rows, cols, rpb);
  }
  if (want_affine) {
    // partial layout [parts][2][cols]: dgamma rows at offset 0, dbeta rows at offset cols
    if (dgamma && dbeta)          // one launch over the 2*cols stacked columns, same summation order per column
      reduce_parts2_kernel<<<(2 * cols + 31) / 32, 256, 0, st>>>(part, parts, 2 * (int64_t)cols, cols, dgamma, dbeta);
    else if (dgamma) reduce_parts_kernel<<<(cols + 31) / 32, 256, 0, st>>>(part, parts, 2 * (int64_t)cols, cols, dgamma);
    else if (dbeta) reduce_parts_kernel<<<(cols + 31) / 32, 256, 0, st>>>(part + cols, parts, 2 * (int64_t)cols, cols, dbeta);
  }
  return tagan_launch_status();
}

TAGAN_API size_t tagan_colsum_workspace_bytes(int64_t rows, int32_t cols) {
  return (size_t)parts_for(rows) * (size_t)cols * sizeof(float);
}

TAGAN_API int tagan_colsum(const float* x, int64_t ldx, float* out, void* workspace, size_t workspace_bytes,
                           int64_t rows, int32_t cols, tagan_stream_t stream) {
  if (!x || !out || rows < 0 || cols <= 0) return TAGAN_E_INVALID;
  if (!workspace || workspace_bytes < tagan_colsum_workspace_bytes(rows, cols)) return TAGAN_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  if (rows == 0) { cudaMemsetAsync(out, 0, sizeof(float) * cols, st); return 0; }
  const int parts = parts_for(rows);
  const int64_t rpb = (rows + parts - 1) / parts;
  if (cols % 128 == 0 && cols <= 768 && ldx % 4 == 0 && aligned16(x)) {
    switch (cols / 128) {
      case 1: colsum_vec_kernel<1><<<parts, 256, 0, st>>>(x, ldx, (float*)workspace, rows, rpb); break;
      case 2: colsum_vec_kernel<2><<<parts, 256, 0, st>>>(x, ldx, (float*)workspace, rows, rpb); break;
      case 3: colsum_vec_kernel<3><<<parts, 256, 0, st>>>(x, ldx, (float*)workspace, rows, rpb); break;
      case 4: colsum_vec_kernel<4><<<parts, 256, 0, st>>>(x, ldx, (float*)workspace, rows, rpb); break;
      case 5: colsum_vec_kernel<5><<<parts, 256, 0, st>>>(x, ldx, (float*)workspace, rows, rpb); break;
      default: colsum_vec_kernel<6><<<parts, 256, 0, st>>>(x, ldx, (float*)workspace, rows, rpb); break;
    }
  } else {
    colsum_partial_kernel<<<parts, 256, 0, st>>>(x, ldx, (float*)workspace, rows, cols, rpb);
  }
  reduce_parts_kernel<<<(cols + 31) / 32, 256, 0, st>>>((const float*)workspace, parts, cols, cols, out);
  return tagan_launch_status();
}

TAGAN_API int tagan_axpby(const float* a, float alpha, const float* b, float beta, float* out, int64_t n,
                          tagan_stream_t stream) {
  if (!a || !out || n < 0) return TAGAN_E_INVALID;
  if (n == 0) return 0;
  axpby_kernel<<<ceil_div_i64(n, 256), 256, 0, as_stream(stream)>>>(a, alpha, b, beta, out, n);
  return tagan_launch_status();
}

TAGAN_API int tagan_gelu_fwd(const float* x, float* y, int64_t n, tagan_stream_t stream) {
  if (!x || !y || n < 0) return TAGAN_E_INVALID;
  if (n == 0) return 0;
  gelu_fwd_kernel<<<ceil_div_i64(n, 256), 256, 0, as_stream(stream)>>>(x, y, n);
  return tagan_launch_status();
}
TAGAN_API int tagan_gelu_bwd(const float* dy, const float* x, float* dx, int64_t n, tagan_stream_t stream) {
  if (!dy || !x || !dx || n < 0) return TAGAN_E_INVALID;
  if (n == 0) return 0;
  gelu_bwd_kernel<<<ceil_div_i64(n, 256), 256, 0, as_stream(stream)>>>(dy, x, dx, n);
  return tagan_launch_status();
}

TAGAN_API int tagan_scale_rows(const float* x, int64_t ldx, const float* rowscale, float* y, int64_t ldy,
                               int64_t rows, int32_t cols, int32_t accumulate, tagan_stream_t stream) {
  if (!x || !y || rows < 0 || cols <= 0) return TAGAN_E_INVALID;
  if (rows == 0) return 0;
  scale_rows_kernel<<<ceil_div_i64(rows * cols, 256), 256, 0, as_stream(stream)>>>(x, ldx, rowscale, y, ldy, rows, cols, accumulate);
  return tagan_launch_status();
}

TAGAN_API int tagan_decay_scale(const float* ts, int64_t ldts, int32_t t, float* rowscale, int64_t rows,
                                tagan_stream_t stream) {
  if (!ts || !rowscale || rows < 0 || t < 1) return TAGAN_E_INVALID;
  if (rows == 0) return 0;
  decay_scale_kernel<<<ceil_div_i64(rows, 256), 256, 0, as_stream(stream)>>>(ts, ldts, t, rowscale, rows);
  return tagan_launch_status();
}
