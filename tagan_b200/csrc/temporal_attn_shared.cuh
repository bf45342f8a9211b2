// Pieces shared by temporal_attn.cu (generic kernels + C ABI) and temporal_attn_fast.cu (T <= 32 fast path).
#pragma once
#include "common.cuh"

namespace tagan_tattn {

constexpr int MAX_WARPS = 8;
constexpr int GEN_MAX_THREADS = 512;   // generic kernels: up to 4 warps per (node, head) slot for T > 32
constexpr int SMEM_LIMIT = 200 * 1024;

struct MaskSpec {
  const float* ts;            // [B,T] or null
  int flags;                  // bit0 causal, bit1 band, bit2 allones=>causal
  float band;
  const int* allones_flag;    // device
  const uint8_t* mask;        // [mask_b, mask_h, T, T] keep-mask or null
  int mask_b, mask_h;
};

__device__ __forceinline__ bool key_valid(const MaskSpec& ms, bool causal, const float* ts_s, const uint8_t* mrow_base,
                                          int T, int i, int j) {
  bool v = true;
  if (causal) v = j <= i;
  if ((ms.flags & 2) && ts_s) v = v && (fabsf(ts_s[i] - ts_s[j]) <= ms.band);
  if (mrow_base) v = v && (mrow_base[(int64_t)i * T + j] != 0);
  return v;
}

// Conservative index window [lo, hi] outside of which key_valid(., .) is false for row / key `x` (the generic kernels then
// walk the warp's union of windows instead of all T positions; every position inside is still tested with key_valid, so a
// loose bound only costs time).  Band masks bound the window only when the node's timestamps are non-decreasing
// (`ts_sorted`: checked per node, NaNs fail the check); an explicit mask tensor gives no bound.
// upper_side: x is a query row and the window is over keys (causal: j <= x); else x is a key and the window is over rows.
__device__ __forceinline__ void valid_window(const MaskSpec& ms, bool causal, const float* ts_s, bool ts_sorted, int T, int x,
                                             bool upper_side, int* lo_out, int* hi_out) {
  int lo = 0, hi = T - 1;
  if (causal) { if (upper_side) hi = x; else lo = x; }
  if ((ms.flags & 2) && ts_s && ts_sorted && x < T) {
    const float tx = ts_s[x];
    // key_valid tests fl(|ts_i - ts_j|) <= band; the thresholds are widened by a few ulps of the larger operand (and the
    // window by one position) so that rounding can never exclude a position that test accepts
    const float slack = 4.f * 1.1920929e-7f * fmaxf(fabsf(tx), fabsf(ms.band));
    const float tlo = tx - ms.band - slack, thi = tx + ms.band + slack;
    int a = 0, b = T;                                      // first index with ts >= tlo
    while (a < b) { const int m = (a + b) >> 1; if (ts_s[m] < tlo) a = m + 1; else b = m; }
    lo = max(lo, a - 1);
    a = 0; b = T;                                          // first index with ts > thi
    while (a < b) { const int m = (a + b) >> 1; if (ts_s[m] <= thi) a = m + 1; else b = m; }
    hi = min(hi, a);
  }
  *lo_out = max(lo, 0);
  *hi_out = min(hi, T - 1);
}
// all lanes of the warp: are ts_s[0..T) non-decreasing (false if any NaN)?
__device__ __forceinline__ bool ts_sorted_warp(const float* ts_s, int T, int lane) {
  bool ok = true;
  for (int t = lane + 1; t < T; t += 32) ok = ok && (ts_s[t] >= ts_s[t - 1]);
  if (T > 0 && lane == 0) ok = ok && (ts_s[0] == ts_s[0]);
  return __all_sync(0xffffffffu, ok);
}

// floats of shared memory per (head) slot of the fast backward kernel: Q,K,V,dO tiles, P and dS (pitch TP+1), ts
__host__ __device__ inline int tattn_bwd_fast_slot_floats(int T, int D, int TP) {
  return 4 * T * D + 2 * TP * (TP + 1) + ((T + 3) & ~3) + (((2 * TP * (TP + 1)) & 3) ? 4 - ((2 * TP * (TP + 1)) & 3) : 0);
}

template <int D>
__device__ __forceinline__ float dot_smem(const float* q, const float* ks) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < D; c += 4) {
    float4 k4 = *reinterpret_cast<const float4*>(ks + c);
    s = fmaf(q[c], k4.x, s); s = fmaf(q[c + 1], k4.y, s); s = fmaf(q[c + 2], k4.z, s); s = fmaf(q[c + 3], k4.w, s);
  }
  return s;
}

// cooperative copy of a [T,D] head slice (row stride ld) into smem by `nl` lanes starting at lane id `li`
template <int D>
__device__ __forceinline__ void load_tile(float* dst, const float* src, int64_t ld, int T, int li, int nl) {
  constexpr int C4 = D / 4;
  for (int idx = li; idx < T * C4; idx += nl) {
    int t = idx / C4, c = idx - t * C4;
    float4 v = __ldg(reinterpret_cast<const float4*>(src + (int64_t)t * ld + c * 4));
    *reinterpret_cast<float4*>(dst + t * D + c * 4) = v;
  }
}


}  // namespace tagan_tattn
using namespace tagan_tattn;

// fast-path launchers (temporal_attn_fast.cu); return false when (D, TP) is not instantiated
bool tagan_tattn_fwd_fast_launch(int D, int TP, int grid, int threads, size_t smem, cudaStream_t st, const float* Q,
                                 const float* K, const float* V, int64_t ld, int64_t B, int T, int heads, int64_t rsb,
                                 int64_t rst, const float* bias, MaskSpec ms, float* ctx, float* lse, float* attn);
bool tagan_tattn_bwd_fast_launch(int D, int TP, int grid, int threads, size_t smem, cudaStream_t st, const float* Q,
                                 const float* K, const float* V, int64_t ld, int64_t B, int T, int heads, int64_t rsb,
                                 int64_t rst, const float* bias, MaskSpec ms, const float* ctx, const float* lse,
                                 const float* dctx, float* dQ, float* dK, float* dV, int64_t ldd, float* dbias_partial);

// tensor-core kernels for 8 < T <= 16 (temporal_attn_fast.cu); return false when the shape is not covered
bool tagan_tattn_fwd_mma_launch(int D, int grid, cudaStream_t st, const float* Q, const float* K, const float* V, int64_t ld,
                                int64_t B, int T, int heads, int64_t rsb, int64_t rst, const float* bias, MaskSpec ms,
                                float* ctx, float* lse, float* attn);
