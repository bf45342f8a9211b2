// tagan_gemm: precision dispatch between the fp32 FFMA path (gemm_simt.cu) and the tcgen05
// tensor-core path (gemm_tc.cu).
#include "common.cuh"

size_t tagan_gemm_simt_workspace_bytes(int32_t op, int64_t m, int64_t n, int64_t k);
int tagan_gemm_simt(int32_t op, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                    int64_t ldb, const float* bias, float* C, int64_t ldc, int32_t accumulate, void* workspace,
                    size_t workspace_bytes, cudaStream_t st);

size_t tagan_gemm_tc_workspace_bytes(int32_t op, int64_t M, int64_t N, int64_t K);
int tagan_gemm_tc(int32_t op, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                  int64_t ldb, const float* bias, float* C, int64_t ldc, int32_t accumulate, int32_t passes,
                  void* workspace, size_t workspace_bytes, cudaStream_t st);

bool tagan_gemm_tma_supported(int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B, int64_t ldb);
size_t tagan_gemm_tma_workspace_bytes(int32_t op, int64_t M, int64_t N, int64_t K);
int tagan_gemm_tma(int32_t op, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                   int64_t ldb, const float* bias, float* C, int64_t ldc, int32_t accumulate, int32_t passes,
                   void* workspace, size_t workspace_bytes, cudaStream_t st);

TAGAN_API size_t tagan_gemm_workspace_bytes(int32_t op, int64_t m, int64_t n, int64_t k) {
  if (op < 0 || op > 2 || m < 0 || n < 0 || k < 0) return 0;
  size_t a = tagan_gemm_simt_workspace_bytes(op, m, n, k), b = tagan_gemm_tc_workspace_bytes(op, m, n, k);
  size_t c = tagan_gemm_tma_workspace_bytes(op, m, n, k);
  a = a > b ? a : b;
  return a > c ? a : c;
}

TAGAN_API int tagan_gemm(int32_t op, int64_t m, int64_t n, int64_t k, const float* A, int64_t lda, const float* B,
                         int64_t ldb, const float* bias, float* C, int64_t ldc, int32_t accumulate, int32_t precision,
                         void* workspace, size_t workspace_bytes, tagan_stream_t stream) {
  if (op < 0 || op > 2 || m < 0 || n < 0 || k < 0 || !C || (k > 0 && (!A || !B))) return TAGAN_E_INVALID;
  if (precision < 0 || precision > 7) return TAGAN_E_INVALID;
  if (m == 0 || n == 0) return 0;
  // tensor-core path: tiles are 128x128, so tiny problems (toy configs) stay on the FFMA kernel
  if (precision > 0 && k > 0 && m * n >= 64 * 64) {
    const int passes = (precision & 3) == 1 ? 3 : ((precision & 3) == 3 ? 4 : 1);
    // precision bit 2 (value 4) forces the LDG-fed kernel (used by the tests to cover both tensor-core paths)
    if (!(precision & 4) && tagan_gemm_tma_supported(m, n, k, A, lda, B, ldb))
      return tagan_gemm_tma(op, m, n, k, A, lda, B, ldb, bias, C, ldc, accumulate, passes, workspace, workspace_bytes,
                            as_stream(stream));
    return tagan_gemm_tc(op, m, n, k, A, lda, B, ldb, bias, C, ldc, accumulate, passes, workspace, workspace_bytes,
                         as_stream(stream));
  }
  return tagan_gemm_simt(op, m, n, k, A, lda, B, ldb, bias, C, ldc, accumulate, workspace, workspace_bytes,
                         as_stream(stream));
}

TAGAN_API int tagan_abi_version(void) { return 1; }
