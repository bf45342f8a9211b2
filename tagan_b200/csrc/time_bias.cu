// (b2) RBF time bias for PER-NODE timestamps, computed on device chunk by chunk -- never the [B,T,T,H] tensor the
// reference materialises (AsymmetricTemporalAttention._compute_time_based_attention, src/tagan/layers/
// temporal_attention.py:792-871 with TimeEncoding._get_basis_encoding :122-220), and never a [B,h,T,T] tensor for the
// whole batch either: the host loops over node chunks whose bias / dBias tiles stay L2-resident.
//
//   dt[b,i,j]   = ts[b,i] - ts[b,j]
//   tn          = (dt - tmin) / (tmax - tmin), tmin/tmax over the WHOLE batch (:142-152); 0 if the range is degenerate.
//                 max over (b,i,j) of dt is max_b (max_t ts - min_t ts) =: R and min is -R, so one range reduction suffices
//   phi_k       = exp(clamp(-(tn - mu_k)^2 / (2 sigma_k^2), -88, 88))                      (:166-189)
//   bias[b,h,i,j] = pos[h,i,j] + sum_k wc[h,k] phi_k + bc[h],  wc = time_q_proj.W @ basis_proj.W  (:848; time_k_proj unused)
//
// Backward (dBias of a chunk -> parameter gradients) puts one basis function on each lane, so the per-pair work is one
// exp and 2*heads FMAs per lane with NO cross-lane traffic; per-block partials are reduced in a fixed order.
#include "common.cuh"

namespace {

constexpr int MAX_HEADS = 16;
constexpr int TB_WARPS = 8;

__global__ void ts_range_kernel(const float* __restrict__ ts, int64_t B, int T, int* __restrict__ range_bits) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float lo = INFINITY, hi = -INFINITY;
  for (int t = 0; t < T; ++t) { const float v = ts[b * T + t]; lo = fminf(lo, v); hi = fmaxf(hi, v); }
  const float r = hi - lo;                               // >= 0: its bit pattern orders like the float
  if (r >= 0.f) atomicMax(range_bits, __float_as_int(r));
}
__global__ void set_int_kernel(int* p, int v) { *p = v; }

__device__ __forceinline__ float norm_dt(float dt, float R) {
  const float tmin = -R, rng = R - tmin;                // = 2R exactly
  return (R > tmin && rng > 1e-7f) ? (dt - tmin) / rng : 0.f;
}

// one thread per (node, i, j) of the chunk
__global__ void __launch_bounds__(256)
time_bias_fwd_kernel(const float* __restrict__ ts, int64_t b0, int64_t Bc, int T, int heads, int nb, const int* __restrict__ range_bits,
                     const float* __restrict__ mu, const float* __restrict__ sigma, const float* __restrict__ wc,
                     const float* __restrict__ bc, const float* __restrict__ pos, float* __restrict__ bias, float* __restrict__ bias_t) {
  extern __shared__ float sm[];                         // mu[nb], inv2s2[nb], wc[heads*nb]
  float* s_mu = sm;
  float* s_i2 = sm + nb;
  float* s_wc = sm + 2 * nb;
  for (int k = threadIdx.x; k < nb; k += blockDim.x) { s_mu[k] = mu[k]; const float sg = sigma[k]; s_i2[k] = 2.f * sg * sg; }
  for (int k = threadIdx.x; k < heads * nb; k += blockDim.x) s_wc[k] = wc[k];
  __syncthreads();
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= Bc * T * T) return;
  const int64_t b = x / (T * T);
  const int ij = (int)(x - b * T * T), i = ij / T, j = ij - i * T;
  const float R = __int_as_float(*range_bits);
  const float tn = norm_dt(ts[(b0 + b) * T + i] - ts[(b0 + b) * T + j], R);
  float acc[MAX_HEADS];
#pragma unroll
  for (int h = 0; h < MAX_HEADS; ++h) acc[h] = 0.f;
  for (int k = 0; k < nb; ++k) {
    const float d = tn - s_mu[k];
    const float e = fminf(fmaxf(-(d * d) / s_i2[k], -88.f), 88.f);
    const float phi = expf(e);
#pragma unroll
    for (int h = 0; h < MAX_HEADS; ++h)
      if (h < heads) acc[h] = fmaf(phi, s_wc[h * nb + k], acc[h]);
  }
#pragma unroll
  for (int h = 0; h < MAX_HEADS; ++h) {
    if (h < heads) {
      const float v = acc[h] + bc[h] + (pos ? pos[((int64_t)h * T + i) * T + j] : 0.f);
      bias[((b * heads + h) * T + i) * (int64_t)T + j] = v;
      bias_t[((b * heads + h) * T + j) * (int64_t)T + i] = v;
    }
  }
}

// lane <-> basis function k (k = lane, lane+32, ...: KPL per lane); a warp walks pairs p = warp, warp+W, ... of its block's
// slice.  partial[block][ (heads*nb) dwc | heads dbc | nb dmu | nb dsigma ]
template <int KPL>
__global__ void __launch_bounds__(TB_WARPS * 32)
time_bias_bwd_kernel(const float* __restrict__ ts, int64_t b0, int64_t Bc, int T, int heads, int nb, const int* __restrict__ range_bits,
                     const float* __restrict__ mu, const float* __restrict__ sigma, const float* __restrict__ wc,
                     const float* __restrict__ dbias /*[Bc,heads,T,T]*/, float* __restrict__ partial, int64_t pairs_per_block) {
  extern __shared__ float red[];                         // [TB_WARPS][width]
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int width = heads * nb + heads + 2 * nb;
  float k_mu[KPL], k_sg[KPL], k_wc[KPL][MAX_HEADS], a_wc[KPL][MAX_HEADS], a_mu[KPL], a_sg[KPL];
#pragma unroll
  for (int q = 0; q < KPL; ++q) {
    const int k = lane + 32 * q;
    k_mu[q] = k < nb ? mu[k] : 0.f;
    k_sg[q] = k < nb ? sigma[k] : 1.f;
    a_mu[q] = a_sg[q] = 0.f;
#pragma unroll
    for (int h = 0; h < MAX_HEADS; ++h) { k_wc[q][h] = (k < nb && h < heads) ? wc[h * nb + k] : 0.f; a_wc[q][h] = 0.f; }
  }
  float a_bc = 0.f;                                     // lane h accumulates dbc[h]
  const float R = __int_as_float(*range_bits);
  const int64_t TT = (int64_t)T * T, total = Bc * TT;
  const int64_t p0 = (int64_t)blockIdx.x * pairs_per_block, p1 = min(total, p0 + pairs_per_block);
  for (int64_t p = p0 + w; p < p1; p += TB_WARPS) {
    const int64_t b = p / TT;
    const int ij = (int)(p - b * TT), i = ij / T, j = ij - i * T;
    const float tn = norm_dt(ts[(b0 + b) * T + i] - ts[(b0 + b) * T + j], R);
    float ds[MAX_HEADS];
#pragma unroll
    for (int h = 0; h < MAX_HEADS; ++h) ds[h] = h < heads ? __ldg(dbias + ((b * heads + h) * T + i) * (int64_t)T + j) : 0.f;
#pragma unroll
    for (int h = 0; h < MAX_HEADS; ++h)
      if (h == lane) a_bc += ds[h];
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
      const float d = tn - k_mu[q];
      const float s2 = k_sg[q] * k_sg[q];
      const float e = -(d * d) / (2.f * s2);
      const bool live = e >= -88.f && e <= 88.f;        // clamp: no gradient outside
      const float phi = expf(fminf(fmaxf(e, -88.f), 88.f));
      float g = 0.f;
#pragma unroll
      for (int h = 0; h < MAX_HEADS; ++h) {
        a_wc[q][h] = fmaf(ds[h], phi, a_wc[q][h]);
        g = fmaf(ds[h], k_wc[q][h], g);
      }
      if (live) {
        const float gp = g * phi;
        a_mu[q] = fmaf(gp, d / s2, a_mu[q]);
        a_sg[q] = fmaf(gp, d * d / (s2 * k_sg[q]), a_sg[q]);
      }
    }
  }
  // warp-private rows of the shared table, then a fixed-order sum over the warps
  float* row = red + (size_t)w * width;
  for (int x = lane; x < width; x += 32) row[x] = 0.f;
  __syncwarp();
#pragma unroll
  for (int q = 0; q < KPL; ++q) {
    const int k = lane + 32 * q;
    if (k < nb) {
#pragma unroll
      for (int h = 0; h < MAX_HEADS; ++h)
        if (h < heads) row[h * nb + k] = a_wc[q][h];
      row[heads * nb + heads + k] = a_mu[q];
      row[heads * nb + heads + nb + k] = a_sg[q];
    }
  }
  if (lane < heads) row[heads * nb + lane] = a_bc;
  __syncthreads();
  for (int x = threadIdx.x; x < width; x += TB_WARPS * 32) {
    float s = red[x];
    for (int ww = 1; ww < TB_WARPS; ++ww) s += red[(size_t)ww * width + x];
    partial[(size_t)blockIdx.x * width + x] = s;
  }
}

__global__ void __launch_bounds__(256)
tb_reduce_kernel(const float* __restrict__ partial, int parts, int width, float* __restrict__ out, int accumulate) {
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

// dpos[h,i,j] (+)= sum over the chunk's nodes of dbias[b,h,i,j] (ascending b: deterministic)
__global__ void dpos_kernel(const float* __restrict__ dbias, int64_t Bc, int64_t hTT, float* __restrict__ dpos, int accumulate) {
  const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= hTT) return;
  float s = 0.f;
  for (int64_t b = 0; b < Bc; ++b) s += dbias[b * hTT + x];
  dpos[x] = accumulate ? dpos[x] + s : s;
}

constexpr int TB_BLOCKS = 592;

}  // namespace

TAGAN_API int tagan_ts_range(const float* ts, int64_t B, int32_t T, float* range, tagan_stream_t stream) {
  if (!ts || !range || B < 0 || T <= 0) return TAGAN_E_INVALID;
  cudaStream_t st = as_stream(stream);
  set_int_kernel<<<1, 1, 0, st>>>(reinterpret_cast<int*>(range), 0);
  if (B > 0) ts_range_kernel<<<ceil_div_i64(B, 256), 256, 0, st>>>(ts, B, T, reinterpret_cast<int*>(range));
  return tagan_launch_status();
}

TAGAN_API int tagan_time_bias_fwd(const float* ts, int64_t node_begin, int64_t nodes, int32_t T, int32_t heads, int32_t num_bases,
                                  const float* range, const float* mu, const float* sigma, const float* wc, const float* bc,
                                  const float* pos_bias, float* bias, float* bias_t, tagan_stream_t stream) {
  if (!ts || !range || !mu || !sigma || !wc || !bc || !bias || !bias_t || node_begin < 0 || nodes < 0 || T <= 0 || heads <= 0 ||
      num_bases <= 0)
    return TAGAN_E_INVALID;
  if (heads > MAX_HEADS || num_bases > 128) return TAGAN_E_UNSUPPORTED;
  if (nodes == 0) return 0;
  const size_t smem = (size_t)(2 * num_bases + heads * num_bases) * sizeof(float);
  time_bias_fwd_kernel<<<ceil_div_i64(nodes * T * T, 256), 256, smem, as_stream(stream)>>>(
      ts, node_begin, nodes, T, heads, num_bases, reinterpret_cast<const int*>(range), mu, sigma, wc, bc, pos_bias, bias, bias_t);
  return tagan_launch_status();
}

TAGAN_API size_t tagan_time_bias_bwd_workspace_bytes(int32_t heads, int32_t num_bases) {
  if (heads <= 0 || num_bases <= 0) return 0;
  return (size_t)TB_BLOCKS * (size_t)(heads * num_bases + heads + 2 * num_bases) * sizeof(float);
}

TAGAN_API int tagan_time_bias_bwd(const float* ts, int64_t node_begin, int64_t nodes, int32_t T, int32_t heads, int32_t num_bases,
                                  const float* range, const float* mu, const float* sigma, const float* wc, const float* dbias,
                                  float* dparams /* [heads*nb dwc | heads dbc | nb dmu | nb dsigma] */, float* dpos /* [heads,T,T] or NULL */,
                                  int32_t accumulate, void* workspace, size_t workspace_bytes, tagan_stream_t stream) {
  if (!ts || !range || !mu || !sigma || !wc || !dbias || !dparams || node_begin < 0 || nodes < 0 || T <= 0 || heads <= 0 ||
      num_bases <= 0)
    return TAGAN_E_INVALID;
  if (heads > MAX_HEADS || num_bases > 128) return TAGAN_E_UNSUPPORTED;
  if (!workspace || workspace_bytes < tagan_time_bias_bwd_workspace_bytes(heads, num_bases)) return TAGAN_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  const int width = heads * num_bases + heads + 2 * num_bases;
  const int64_t hTT = (int64_t)heads * T * T;
  if (nodes == 0) {
    if (!accumulate) {
      cudaMemsetAsync(dparams, 0, sizeof(float) * width, st);
      if (dpos) cudaMemsetAsync(dpos, 0, sizeof(float) * hTT, st);
    }
    return 0;
  }
  const int64_t total = nodes * T * T;
  int blocks = (int)((total + 255) / 256);
  if (blocks > TB_BLOCKS) blocks = TB_BLOCKS;
  const int64_t ppb = (total + blocks - 1) / blocks;
  float* part = static_cast<float*>(workspace);
  const size_t smem = (size_t)TB_WARPS * width * sizeof(float);
  const int kpl = (num_bases + 31) / 32;
#define TB_LAUNCH(K)                                                                                                        \
  {                                                                                                                         \
    if (smem > 48 * 1024) cudaFuncSetAttribute(time_bias_bwd_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    time_bias_bwd_kernel<K><<<blocks, TB_WARPS * 32, smem, st>>>(ts, node_begin, nodes, T, heads, num_bases,                 \
        reinterpret_cast<const int*>(range), mu, sigma, wc, dbias, part, ppb);                                              \
  }
  switch (kpl) {
    case 1: TB_LAUNCH(1) break;
    case 2: TB_LAUNCH(2) break;
    case 3: TB_LAUNCH(3) break;
    default: TB_LAUNCH(4) break;
  }
#undef TB_LAUNCH
  tb_reduce_kernel<<<(width + 31) / 32, 256, 0, st>>>(part, blocks, width, dparams, accumulate);
  if (dpos) dpos_kernel<<<ceil_div_i64(hTT, 256), 256, 0, st>>>(dbias, nodes, hTT, dpos, accumulate);
  return tagan_launch_status();
}
