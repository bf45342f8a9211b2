/*
 * tagan_b200.h -- C ABI of libtagan_b200.so: the sm_100a kernels behind TAGAN's
 * per-snapshot attention-and-propagation core.
 *
 * The reference (MaLoskins/Temporal-Asymmetric-Graph-Attention-Network) is pure eager
 * PyTorch and has no FFI of its own; its boundary for this path is the nn.Module surface
 * (SURVEY.md section 8b).  Each entry point below names the reference code it replaces
 * (paths relative to the reference checkout).  The Python host side (tagan_b200/*.py)
 * mirrors the reference's modules and binds these symbols with ctypes; INTEGRATION.md shows
 * the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous memory unless a leading dimension is
 *     given; fp32 unless the name says otherwise; indices are int32 except edge_index (int64,
 *     the torch.long the reference is handed);
 *   - the library allocates nothing and keeps no pointer after return: the caller owns all
 *     memory, including workspaces whose size the *_workspace_bytes functions report;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t): no host sync, no
 *     allocation, CUDA-graph capturable, re-entrant;
 *   - return value: 0 = ok, <0 = TAGAN_E_* (invalid argument / unsupported shape),
 *     >0 = a cudaError_t from the launch.  Nothing throws across the boundary.
 */
#ifndef TAGAN_B200_H
#define TAGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tagan_stream_t; /* cudaStream_t */

#define TAGAN_E_INVALID (-1)     /* null pointer, negative size, ... */
#define TAGAN_E_UNSUPPORTED (-2) /* shape outside what the kernels are built for */
#define TAGAN_E_WORKSPACE (-3)   /* workspace too small */

/* DistanceMetric ids (src/tagan/layers/geometric_attention.py:196-225). */
enum tagan_metric {
  TAGAN_METRIC_SCALED_DOT = 0,
  TAGAN_METRIC_DOT = 1,
  TAGAN_METRIC_COSINE_SIM = 2,
  TAGAN_METRIC_EUCLIDEAN = 3,
  TAGAN_METRIC_SQ_EUCLIDEAN = 4,
  TAGAN_METRIC_MANHATTAN = 5,
  TAGAN_METRIC_COSINE_DIST = 6,
  TAGAN_METRIC_GAUSSIAN = 7,
  TAGAN_METRIC_RBF = 8
};

int tagan_abi_version(void);

/* ---------------------------------------------------------------------------------------
 * (a1) edge_index -> adjacency-with-self-loops as CSR, built on device.
 * Replaces the dense mask of TAGANGraphAttention.forward: `adj[ei[0],ei[1]] = 1; adj += eye`
 * (src/tagan/layers/graph_attention.py:98-102).  Entry set = unique(edges) U {(i,i)}, row =
 * edge_index[0] (query node), col = edge_index[1] (key/value node); negative indices wrap
 * like torch indexing; entries row-major sorted (== torch.unique(row*N+col) order), bit-exact.
 *   rowptr[N+1], col[E+N], row[E+N] (row of every CSR entry)       -- outputs, capacity E+N
 *   rowptr_t[N+1], row_t[E+N], perm_t[E+N]  -- transposed CSR (sorted by (col,row)); perm_t[k]
 *                                              = CSR position of transposed entry k.  All three
 *                                              may be NULL to skip the transpose.
 *   status[1] (int32): set to 1 if any index was outside [-N, N) (the reference raises
 *                      IndexError there; such edges are dropped here).  nnz = rowptr[N].
 * ------------------------------------------------------------------------------------- */
size_t tagan_csr_workspace_bytes(int64_t num_edges, int32_t num_nodes);
int tagan_csr_build(const int64_t* edge_index, int64_t num_edges, int32_t num_nodes,
                    int32_t* rowptr, int32_t* col, int32_t* row,
                    int32_t* rowptr_t, int32_t* row_t, int32_t* perm_t,
                    int32_t* status, void* workspace, size_t workspace_bytes,
                    tagan_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * (a2-a4) fused geometric attention over the CSR: per-entry per-head score from the
 * DistanceMetric, segment softmax over each row, weighted aggregation of V.  One warp per
 * destination row, online softmax, no atomics.
 * Replaces GeometricAttention._get_attention_weights + `attn @ v`
 * (src/tagan/layers/geometric_attention.py:332-516, :579).
 *   Q,K,V: [N,H] with row stride ld (floats) -- column slices of a fused [N,3H] projection
 *   metric_param: [heads] sigma (gaussian) / gamma (rbf) when learnable, else NULL (=1.0)
 *   ctx[N,H] (dense), lse[N,heads] = log-sum-exp of each row's scores (saved for backward)
 *   attn: [nnz,heads] per-entry softmax weights, or NULL
 * ------------------------------------------------------------------------------------- */
int tagan_geo_attn_fwd(const float* Q, const float* K, const float* V, int64_t ld,
                       const int32_t* rowptr, const int32_t* col,
                       int32_t num_nodes, int32_t hidden, int32_t heads, int32_t metric,
                       const float* metric_param, float* ctx, float* lse, float* attn,
                       tagan_stream_t stream);

/* Deterministic backward: row pass over the CSR (dQ, delta) then column pass over the
 * transposed CSR (dK, dV); scores are recomputed, nothing is accumulated with atomics.
 *   dQ,dK,dV: [N,H] with row stride ldd.  delta_ws[N,heads] scratch.
 *   dparam_ws[N,heads] scratch and dparam[heads] output, both NULL unless metric_param given. */
int tagan_geo_attn_bwd(const float* Q, const float* K, const float* V, int64_t ld,
                       const int32_t* rowptr, const int32_t* col,
                       const int32_t* rowptr_t, const int32_t* row_t,
                       int32_t num_nodes, int32_t hidden, int32_t heads, int32_t metric,
                       const float* metric_param, const float* ctx, const float* lse,
                       const float* dctx, float* dQ, float* dK, float* dV, int64_t ldd,
                       float* delta_ws, float* dparam_ws, float* dparam,
                       tagan_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Row-wise building blocks shared by all layers (the reference uses nn.LayerNorm eps=1e-5,
 * residual adds, GELU(erf), sigmoid/tanh gates as separate eager ops).
 * ------------------------------------------------------------------------------------- */
/* y = LN(x [+ res]) * gamma + beta; optionally y *= rowscale[row]; saves mean,rstd [rows].
 * `sum_out` (nullable) receives x+res (pre-norm value, needed by backward).
 * gamma/beta NULL => plain copy of (x+res) (use_layer_norm=False). */
int tagan_layernorm_fwd(const float* x, int64_t ldx, const float* res, int64_t ldres,
                        const float* gamma, const float* beta, const float* rowscale,
                        float* y, int64_t ldy, float* sum_out, float* mean, float* rstd,
                        int64_t rows, int32_t cols, tagan_stream_t stream);
/* dx = LN backward of dy (dy already multiplied by rowscale if given);
 * dgamma/dbeta [cols] accumulated deterministically through partial_ws[2*parts*cols].
 * `dx_accumulate` != 0 adds into dx instead of overwriting. */
size_t tagan_layernorm_bwd_workspace_bytes(int64_t rows, int32_t cols);
int tagan_layernorm_bwd(const float* dy, int64_t lddy, const float* xsum, int64_t ldx,
                        const float* gamma, const float* rowscale, const float* mean,
                        const float* rstd, float* dx, int64_t lddx, int32_t dx_accumulate,
                        float* dgamma, float* dbeta, void* workspace, size_t workspace_bytes,
                        int64_t rows, int32_t cols, tagan_stream_t stream);
/* out[cols] = sum over rows of x (bias gradients), deterministic two-stage reduction. */
size_t tagan_colsum_workspace_bytes(int64_t rows, int32_t cols);
int tagan_colsum(const float* x, int64_t ldx, float* out, void* workspace, size_t workspace_bytes,
                 int64_t rows, int32_t cols, tagan_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Dense projections (nn.Linear: q/k/v/output_proj and the GRU / gating / skip Linears).
 *   NT: C[M,N] = A[M,K] . B[N,K]^T (+ bias[N]) (+ C if accumulate)        forward Linear
 *   NN: C[M,N] = A[M,K] . B[K,N]            (+ C if accumulate)           dX = dY . W
 *   TN: C[M,N] = A[Kd,M]^T . B[Kd,N]        (+ C if accumulate)           dW = dY^T . X
 * fp32 in/out.  precision: 0 = fp32 FFMA, 1 = 3xTF32 on tcgen05 (fp32-accurate), 2 = 1xTF32.
 * ------------------------------------------------------------------------------------- */
size_t tagan_gemm_workspace_bytes(int32_t op, int64_t m, int64_t n, int64_t k);
int tagan_gemm(int32_t op /*0=NT,1=NN,2=TN*/, int64_t m, int64_t n, int64_t k,
               const float* A, int64_t lda, const float* B, int64_t ldb,
               const float* bias, float* C, int64_t ldc, int32_t accumulate, int32_t precision,
               void* workspace, size_t workspace_bytes, tagan_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Small fused element-wise helpers.
 * ------------------------------------------------------------------------------------- */
/* out = alpha*a + beta*b (b may be NULL) */
int tagan_axpby(const float* a, float alpha, const float* b, float beta, float* out, int64_t n,
                tagan_stream_t stream);
/* y = gelu(x) (erf form, nn.GELU default) and dx = dy * gelu'(x) */
int tagan_gelu_fwd(const float* x, float* y, int64_t n, tagan_stream_t stream);
int tagan_gelu_bwd(const float* dy, const float* x, float* dx, int64_t n, tagan_stream_t stream);
/* y[r,:] (+)= x[r,:] * rowscale[r] */
int tagan_scale_rows(const float* x, int64_t ldx, const float* rowscale, float* y, int64_t ldy,
                     int64_t rows, int32_t cols, int32_t accumulate, tagan_stream_t stream);
/* rowscale[i] = exp(-clamp(ts[i,t]-ts[i,t-1], 0, 10))  (TemporalGRUCell.forward,
 * src/tagan/layers/temporal_propagation.py:509-514) */
int tagan_decay_scale(const float* ts, int64_t ldts, int32_t t, float* rowscale, int64_t rows,
                      tagan_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TAGAN_B200_H */
