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
size_t tagan_gemm_tma_colsum_bytes(int64_t M, int64_t N, int64_t K);
int tagan_gemm_tma(int32_t op, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                   int64_t ldb, const float* bias, float* C, int64_t ldc, int32_t accumulate, int32_t passes,
                   void* workspace, size_t workspace_bytes, cudaStream_t st, float* colsum_a,
                   const float* A2 = nullptr, int64_t lda2 = 0, int64_t K1 = 0, const tagan_epilogue* epi = nullptr,
                   int32_t epi_fast = 0);

void tagan_gemm_tma_set_tuning(int key, int value);
/* tuning knobs of the tcgen05 GEMM for A/B measurements inside one process (tools/profile_gemm_shapes.py) */
TAGAN_API void tagan_gemm_set_tuning(int32_t key, int32_t value) { tagan_gemm_tma_set_tuning(key, value); }
void tagan_gemm_tma_set_trace(void* buf);
/* debug: device buffer of 16 x 512 int64 that CTA 0 of every following tcgen05 GEMM launch fills with clock64() stamps
 * (per k-block: TMA issue, bytes landed, split done, MMA start, MMA issued; per tile: accumulator free, accumulator full,
 * epilogue done); null switches tracing off */
TAGAN_API void tagan_gemm_set_trace(void* buf) { tagan_gemm_tma_set_trace(buf); }

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
                            as_stream(stream), nullptr);
    return tagan_gemm_tc(op, m, n, k, A, lda, B, ldb, bias, C, ldc, accumulate, passes, workspace, workspace_bytes,
                         as_stream(stream));
  }
  return tagan_gemm_simt(op, m, n, k, A, lda, B, ldb, bias, C, ldc, accumulate, workspace, workspace_bytes,
                         as_stream(stream));
}

// dW = A^T . B together with colsum(A) (Linear backward: A = dY [rows, out], B = X [rows, in]; colsum(A) = db).
// On the TMA-fed tensor-core path the column sums are accumulated by the warps that split A anyway, so the bias
// gradient costs no second pass over dY; otherwise it is tagan_gemm followed by tagan_colsum.
TAGAN_API size_t tagan_gemm_tn_colsum_workspace_bytes(int64_t m, int64_t n, int64_t k) {
  if (m < 0 || n < 0 || k < 0) return 0;
  size_t a = tagan_gemm_workspace_bytes(2, m, n, k) + tagan_gemm_tma_colsum_bytes(m, n, k);
  size_t b = tagan_colsum_workspace_bytes(k, (int32_t)m);
  return a > b ? a : b;
}

TAGAN_API int tagan_gemm_tn_colsum(int64_t m, int64_t n, int64_t k, const float* A, int64_t lda, const float* B, int64_t ldb,
                                   float* C, int64_t ldc, float* colsum_a, int32_t precision, void* workspace,
                                   size_t workspace_bytes, tagan_stream_t stream) {
  if (m < 0 || n < 0 || k < 0 || !C || !colsum_a || (k > 0 && (!A || !B))) return TAGAN_E_INVALID;
  if (precision < 0 || precision > 7 || m >= (1LL << 31)) return TAGAN_E_INVALID;
  if (m == 0) return 0;
  if (n > 0 && precision > 0 && k > 0 && m * n >= 64 * 64 && !(precision & 4) && tagan_gemm_tma_supported(m, n, k, A, lda, B, ldb)) {
    const int passes = (precision & 3) == 1 ? 3 : ((precision & 3) == 3 ? 4 : 1);
    return tagan_gemm_tma(2, m, n, k, A, lda, B, ldb, nullptr, C, ldc, 0, passes, workspace, workspace_bytes,
                          as_stream(stream), colsum_a);
  }
  if (n > 0) {
    int rc = tagan_gemm(2, m, n, k, A, lda, B, ldb, nullptr, C, ldc, 0, precision, workspace, workspace_bytes, stream);
    if (rc) return rc;
  }
  return tagan_colsum(A, lda, colsum_a, workspace, workspace_bytes, k, (int32_t)m, stream);
}

TAGAN_API int tagan_abi_version(void) { return 1; }


// ---- fused-epilogue projections (tensor-core path only; the caller composes the stand-alone kernels otherwise) ----
TAGAN_API size_t tagan_gemm_fused_workspace_bytes(int32_t op, int64_t m, int64_t n, int64_t k) {
  if (op < 0 || op > 1 || m < 0 || n < 0 || k < 0) return 0;
  return tagan_gemm_tma_workspace_bytes(op, m, n, k);
}

static inline bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

TAGAN_API int tagan_gemm_fused(int32_t op, int64_t m, int64_t n, int64_t k, const float* A, int64_t lda, const float* A2,
                               int64_t lda2, int64_t k1, const float* B, int64_t ldb, const float* bias,
                               const struct tagan_epilogue* epi, int32_t precision, void* workspace, size_t workspace_bytes,
                               tagan_stream_t stream) {
  if (op < 0 || op > 1 || m < 0 || n <= 0 || k <= 0 || !A || !B || !epi || !epi->out0) return TAGAN_E_INVALID;
  const int32_t epi_fast = precision & 8;                  // +8: MUFU-based sigmoid / tanh in the epilogue
  precision &= 7;
  if (precision < 1 || precision > 3) return TAGAN_E_INVALID;
  if (m == 0) return 0;
  const tagan_epilogue& e = *epi;
  if (e.mode < TAGAN_EPI_STORE || e.mode > TAGAN_EPI_STORE_BF16) return TAGAN_E_INVALID;
  if (n % 4 || !al16p(bias) || e.ld_out0 % 4) return TAGAN_E_UNSUPPORTED;
  if (e.mode == TAGAN_EPI_STORE_BF16 ? (reinterpret_cast<uintptr_t>(e.out0) & 7) != 0 : !al16p(e.out0)) return TAGAN_E_UNSUPPORTED;
  if ((e.in0 && (!al16p(e.in0) || e.ld_in0 % 4)) || (e.in1 && (!al16p(e.in1) || e.ld_in1 % 4)) ||
      (e.out1 && (!al16p(e.out1) || e.ld_out1 % 4)) || (e.out2 && (!al16p(e.out2) || e.ld_out2 % 4)))
    return TAGAN_E_UNSUPPORTED;
  switch (e.mode) {
    case TAGAN_EPI_RES_LN:
      if (e.gamma && (!e.beta || !al16p(e.gamma) || !al16p(e.beta) || n > 128)) return e.beta ? TAGAN_E_UNSUPPORTED : TAGAN_E_INVALID;
      break;
    case TAGAN_EPI_GATES:
      if (!e.in0 || !e.out1 || !e.out2) return TAGAN_E_INVALID;
      if (e.split <= 0 || e.split >= n || e.split % 4) return TAGAN_E_UNSUPPORTED;
      break;
    case TAGAN_EPI_BLEND:
      if (!e.in0 || !e.in1 || !e.out1) return TAGAN_E_INVALID;
      break;
    case TAGAN_EPI_GATES_BWD:
      if (!e.in0 || !e.in1 || !e.out1) return TAGAN_E_INVALID;
      break;
    default: break;
  }
  if (!tagan_gemm_tma_supported(m, n, k, A, lda, B, ldb)) return TAGAN_E_UNSUPPORTED;
  if (A2 && ((reinterpret_cast<uintptr_t>(A2) & 15) || lda2 % 4)) return TAGAN_E_UNSUPPORTED;
  const int passes = (precision & 3) == 1 ? 3 : ((precision & 3) == 3 ? 4 : 1);
  return tagan_gemm_tma(op, m, n, k, A, lda, B, ldb, bias, e.out0, e.ld_out0, 0, passes, workspace, workspace_bytes,
                        as_stream(stream), nullptr, A2, lda2, k1, epi, epi_fast);
}
