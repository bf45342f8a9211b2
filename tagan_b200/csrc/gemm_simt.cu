// Dense projections, fp32 FFMA path (precision 0): exact-fp32 tiled SIMT GEMM with register
// prefetch.  This is the bring-up / parity-reference path for nn.Linear forward (NT), dX (NN)
// and dW (TN, split-K with a fixed-order reduction => deterministic).  The tensor-core path
// (tcgen05, 3xTF32) lives in gemm_tc.cu and is selected with precision 1/2.
#include "common.cuh"

namespace {

constexpr int BK = 8;
constexpr int THREADS = 256;

// Loads a (ROWS x BK) tile of a K-contiguous matrix (element (r,k) at p[r*ld+k]) into
// smem[k][r], zero-filling out-of-range elements.
template <int ROWS>
__device__ __forceinline__ void fetch_kc(const float* __restrict__ p, int64_t ld, int64_t r0, int64_t rmax, int64_t k0,
                                         int64_t kmax, bool vec_ok, float (&reg)[ROWS * BK / THREADS]) {
  constexpr int PER = ROWS * BK / THREADS;   // 4 for ROWS=128, 2 for ROWS=64
  // thread t covers row t / (BK/PER), k offset (t % (BK/PER)) * PER
  constexpr int TPR = BK / PER;
  const int r = threadIdx.x / TPR, kq = (threadIdx.x % TPR) * PER;
  const int64_t gr = r0 + r, gk = k0 + kq;
  if (PER == 4 && vec_ok && gr < rmax && gk + 3 < kmax) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p + gr * ld + gk));
    reg[0] = v.x; reg[1] = v.y; reg[2] = v.z; reg[3] = v.w;
  } else if (PER == 2 && vec_ok && gr < rmax && gk + 1 < kmax) {
    float2 v = __ldg(reinterpret_cast<const float2*>(p + gr * ld + gk));
    reg[0] = v.x; reg[1] = v.y;
  } else {
#pragma unroll
    for (int i = 0; i < PER; ++i) reg[i] = (gr < rmax && gk + i < kmax) ? __ldg(p + gr * ld + gk + i) : 0.f;
  }
}
template <int ROWS>
__device__ __forceinline__ void stash_kc(float (*sm)[ROWS], const float (&reg)[ROWS * BK / THREADS]) {
  constexpr int PER = ROWS * BK / THREADS;
  constexpr int TPR = BK / PER;
  const int r = threadIdx.x / TPR, kq = (threadIdx.x % TPR) * PER;
#pragma unroll
  for (int i = 0; i < PER; ++i) sm[kq + i][r] = reg[i];
}

// Loads a (BK x COLS) tile of an MN-contiguous matrix (element (k,c) at p[k*ld+c]) into smem[k][c].
template <int COLS>
__device__ __forceinline__ void fetch_mc(const float* __restrict__ p, int64_t ld, int64_t c0, int64_t cmax, int64_t k0,
                                         int64_t kmax, bool vec_ok, float (&reg)[COLS * BK / THREADS]) {
  constexpr int PER = COLS * BK / THREADS;
  constexpr int TPK = COLS / PER;            // threads per k row
  const int k = threadIdx.x / TPK, cq = (threadIdx.x % TPK) * PER;
  const int64_t gk = k0 + k, gc = c0 + cq;
  if (PER == 4 && vec_ok && gk < kmax && gc + 3 < cmax) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p + gk * ld + gc));
    reg[0] = v.x; reg[1] = v.y; reg[2] = v.z; reg[3] = v.w;
  } else if (PER == 2 && vec_ok && gk < kmax && gc + 1 < cmax) {
    float2 v = __ldg(reinterpret_cast<const float2*>(p + gk * ld + gc));
    reg[0] = v.x; reg[1] = v.y;
  } else {
#pragma unroll
    for (int i = 0; i < PER; ++i) reg[i] = (gk < kmax && gc + i < cmax) ? __ldg(p + gk * ld + gc + i) : 0.f;
  }
}
template <int COLS>
__device__ __forceinline__ void stash_mc(float (*sm)[COLS], const float (&reg)[COLS * BK / THREADS]) {
  constexpr int PER = COLS * BK / THREADS;
  constexpr int TPK = COLS / PER;
  const int k = threadIdx.x / TPK, cq = (threadIdx.x % TPK) * PER;
#pragma unroll
  for (int i = 0; i < PER; ++i) sm[k][cq + i] = reg[i];
}

// C[M,N] (+)= op(A) . op(B).  A_KC: A stored [M,K] (K contiguous) else [K,M].  B_KC: B stored
// [N,K] else [K,N].  gridDim.z = split-K parts; with parts > 1 results go to `partial`.
template <int BM, int BN, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(THREADS)
gemm_kernel(int64_t M, int64_t N, int64_t K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
            int64_t ldb, const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int accumulate,
            float* __restrict__ partial, int64_t k_per_split, int a_vec, int b_vec) {
  constexpr int TM = BM / 16, TN = BN / 16;
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_per_split;
  const int64_t kend = min(K, kbeg + k_per_split);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float ra[BM * BK / THREADS], rb[BN * BK / THREADS];
  auto fetch = [&](int64_t k0) {
    if (A_KC) fetch_kc<BM>(A, lda, m0, M, k0, kend, a_vec, ra);
    else fetch_mc<BM>(A, lda, m0, M, k0, kend, a_vec, ra);
    if (B_KC) fetch_kc<BN>(B, ldb, n0, N, k0, kend, b_vec, rb);
    else fetch_mc<BN>(B, ldb, n0, N, k0, kend, b_vec, rb);
  };
  auto stash = [&](int buf) {
    if (A_KC) stash_kc<BM>(As[buf], ra); else stash_mc<BM>(As[buf], ra);
    if (B_KC) stash_kc<BN>(Bs[buf], rb); else stash_mc<BN>(Bs[buf], rb);
  };
  int buf = 0;
  if (kbeg < kend) { fetch(kbeg); stash(0); }
  __syncthreads();
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = k0 + BK < kend;
    if (more) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4 + i * 16]);   // rows ty*4+{0..3} + 64*(i/4)
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4 + j * 16]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) { stash(buf ^ 1); __syncthreads(); buf ^= 1; }
  }
  // thread owns rows m0 + ty*4 + (i%4) + 64*(i/4), cols n0 + tx*4 + (j%4) + 64*(j/4)
  float* out = partial ? partial + (int64_t)blockIdx.z * M * N : C;
  const int64_t ldo = partial ? N : ldc;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t m = m0 + ty * 4 + (i & 3) + 64 * (i >> 2);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + tx * 4 + (j & 3) + 64 * (j >> 2);
      if (n >= N) continue;
      float v = acc[i][j];
      if (!partial) {
        if (bias) v += bias[n];
        if (accumulate) v += out[m * ldo + n];
      }
      out[m * ldo + n] = v;
    }
  }
}

__global__ void splitk_reduce(const float* __restrict__ partial, int parts, int64_t M, int64_t N,
                              const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int accumulate) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  float s = 0.f;
  for (int p = 0; p < parts; ++p) s += partial[(int64_t)p * M * N + i];
  const int64_t m = i / N, n = i - m * N;
  if (bias) s += bias[n];
  float* o = C + m * ldc + n;
  *o = accumulate ? *o + s : s;
}

struct Plan { int bm, bn; int splits; int64_t k_per_split; };
Plan make_plan(int op, int64_t M, int64_t N, int64_t K) {
  Plan p;
  const bool big = (M >= 128 && N >= 128) && ((M + 127) / 128) * ((N + 127) / 128) >= 148;
  p.bm = big ? 128 : 64;
  p.bn = big ? 128 : 64;
  int64_t tiles = ((M + p.bm - 1) / p.bm) * ((N + p.bn - 1) / p.bn);
  p.splits = 1;
  if (op == 2 && tiles < 296 && K >= 2048) {
    int64_t want = (296 + tiles - 1) / tiles;
    int64_t maxs = K / 512;
    p.splits = (int)(want < maxs ? want : maxs);
    if (p.splits < 1) p.splits = 1;
  }
  int64_t kps = (K + p.splits - 1) / p.splits;
  kps = (kps + BK - 1) / BK * BK;
  p.k_per_split = kps;
  p.splits = (int)((K + kps - 1) / (kps > 0 ? kps : 1));
  if (p.splits < 1) p.splits = 1;
  return p;
}

template <bool A_KC, bool B_KC>
void launch(const Plan& p, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb,
            const float* bias, float* C, int64_t ldc, int accumulate, float* partial, cudaStream_t st) {
  const int a_vec = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (lda % 4 == 0);
  const int b_vec = ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (ldb % 4 == 0);
  dim3 grid((unsigned)((N + p.bn - 1) / p.bn), (unsigned)((M + p.bm - 1) / p.bm), (unsigned)p.splits);
  if (p.bm == 128)
    gemm_kernel<128, 128, A_KC, B_KC><<<grid, THREADS, 0, st>>>(M, N, K, A, lda, B, ldb, bias, C, ldc, accumulate, partial, p.k_per_split, a_vec, b_vec);
  else
    gemm_kernel<64, 64, A_KC, B_KC><<<grid, THREADS, 0, st>>>(M, N, K, A, lda, B, ldb, bias, C, ldc, accumulate, partial, p.k_per_split, a_vec, b_vec);
}

}  // namespace

size_t tagan_gemm_simt_workspace_bytes(int32_t op, int64_t m, int64_t n, int64_t k) {
  Plan p = make_plan(op, m, n, k);
  return p.splits > 1 ? (size_t)p.splits * (size_t)m * (size_t)n * sizeof(float) : 0;
}

int tagan_gemm_simt(int32_t op, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                    int64_t ldb, const float* bias, float* C, int64_t ldc, int32_t accumulate, void* workspace,
                    size_t workspace_bytes, cudaStream_t st) {
  Plan p = make_plan(op, M, N, K);
  float* partial = nullptr;
  if (p.splits > 1) {
    if (!workspace || workspace_bytes < (size_t)p.splits * M * N * sizeof(float)) return TAGAN_E_WORKSPACE;
    partial = static_cast<float*>(workspace);
  }
  if (op == 0) launch<true, true>(p, M, N, K, A, lda, B, ldb, bias, C, ldc, accumulate, partial, st);
  else if (op == 1) launch<true, false>(p, M, N, K, A, lda, B, ldb, bias, C, ldc, accumulate, partial, st);
  else launch<false, false>(p, M, N, K, A, lda, B, ldb, bias, C, ldc, accumulate, partial, st);
  if (partial)
    splitk_reduce<<<ceil_div_i64(M * N, 256), 256, 0, st>>>(partial, p.splits, M, N, bias, C, ldc, accumulate);
  return tagan_launch_status();
}
