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

/* Batched form: T snapshots (T <= 128) as ONE block-diagonal graph in one launch set -- snapshot t's nodes are rows
 * [off_t, off_t + node_counts[t]) with off_t = sum of the previous counts, its column ids are offset the same way, so the
 * result equals the per-snapshot CSRs concatenated (bit-exact) and kernel (a) runs ONCE over the stacked [sum N_t, 3H]
 * projection.  src / dst / edge_counts / node_counts are HOST arrays (T device pointers to the int64 source / destination
 * ids of each snapshot; negative ids wrap within their snapshot); they are read during the call only.
 * Outputs have capacity sum(E_t) + sum(N_t); workspace = tagan_csr_workspace_bytes(sum E_t, sum N_t). */
int tagan_csr_build_batched(const int64_t* const* src, const int64_t* const* dst, const int64_t* edge_counts,
                            const int32_t* node_counts, int32_t T, int32_t* rowptr, int32_t* col, int32_t* row,
                            int32_t* rowptr_t, int32_t* row_t, int32_t* perm_t, int32_t* status, void* workspace,
                            size_t workspace_bytes, tagan_stream_t stream);

/* Node-partitioned variant (single large graph): this rank owns the query rows
 * [row_begin, row_begin+num_rows); edges whose row lies elsewhere are skipped.  rowptr[num_rows+1], row[]
 * hold LOCAL row ids, col[] GLOBAL node ids (self loop of local row i = row_begin+i); the transposed CSR is
 * indexed by global source node (rowptr_t[num_nodes+1]) with local row ids in row_t. */
int tagan_csr_build_part(const int64_t* edge_index, int64_t num_edges, int32_t num_nodes,
                         int32_t row_begin, int32_t num_rows,
                         int32_t* rowptr, int32_t* col, int32_t* row,
                         int32_t* rowptr_t, int32_t* row_t, int32_t* perm_t,
                         int32_t* status, void* workspace, size_t workspace_bytes,
                         tagan_stream_t stream);

/* Profiling knob (synchronous; default 1): kernel (a) loads the gathered K / V / Q / dctx rows with an L2 evict_last
 * policy and streams its own rows, results and gradients with evict_first, so a snapshot's K|V stay L2-resident while the
 * kernel walks it; 0 switches the hints off.  The `_bf16` variant sets the bf16-storage build. */
int tagan_geo_attn_set_l2_policy(int32_t mode);
int tagan_geo_attn_set_l2_policy_bf16(int32_t mode);

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

/* Rectangular forms for the node-partitioned graph: Q/ctx/dQ have n_rows LOCAL rows (stride ldq/lddq),
 * K/V and dK/dV have n_src GLOBAL rows (stride ldkv/lddkv) -- K|V is the all-gathered projection, dK|dV
 * the partial sums that are reduce-scattered afterwards. */
int tagan_geo_attn_fwd_part(const float* Q, int64_t ldq, const float* K, const float* V, int64_t ldkv,
                            const int32_t* rowptr, const int32_t* col, int32_t n_rows,
                            int32_t hidden, int32_t heads, int32_t metric, const float* metric_param,
                            float* ctx, float* lse, float* attn, tagan_stream_t stream);
int tagan_geo_attn_bwd_part(const float* Q, int64_t ldq, const float* K, const float* V, int64_t ldkv,
                            const int32_t* rowptr, const int32_t* col,
                            const int32_t* rowptr_t, const int32_t* row_t,
                            int32_t n_rows, int32_t n_src, int32_t hidden, int32_t heads, int32_t metric,
                            const float* metric_param, const float* ctx, const float* lse,
                            const float* dctx, float* dQ, int64_t lddq, float* dK, float* dV, int64_t lddkv,
                            float* delta_ws, float* dparam_ws, float* dparam, tagan_stream_t stream);

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
 * fp32 in/out.  precision: 0 = fp32 FFMA, 1 = TF32 hi/lo split on tcgen05 (3 MMAs, fp32-accurate),
 * 2 = plain TF32, 3 = hi/lo split with the lo.lo term as well (4 MMAs);
 * +4 forces the LDG-fed tensor-core kernel instead of the TMA-fed one (TMA needs 16-byte aligned operands).
 * ------------------------------------------------------------------------------------- */
size_t tagan_gemm_workspace_bytes(int32_t op, int64_t m, int64_t n, int64_t k);
int tagan_gemm(int32_t op /*0=NT,1=NN,2=TN*/, int64_t m, int64_t n, int64_t k,
               const float* A, int64_t lda, const float* B, int64_t ldb,
               const float* bias, float* C, int64_t ldc, int32_t accumulate, int32_t precision,
               void* workspace, size_t workspace_bytes, tagan_stream_t stream);
/* nn.Linear backward in one call (autograd of the reference's `nn.Linear`s): C[M,N] = A[Kd,M]^T . B[Kd,N] (= dW for
 * A = dY, B = X) and colsum_a[M] = sum over the Kd rows of A (= db).  On the tensor-core path the column sums are
 * accumulated while A is split, so the bias gradient needs no second pass over dY. */
size_t tagan_gemm_tn_colsum_workspace_bytes(int64_t m, int64_t n, int64_t k);
int tagan_gemm_tn_colsum(int64_t m, int64_t n, int64_t k, const float* A, int64_t lda, const float* B, int64_t ldb,
                         float* C, int64_t ldc, float* colsum_a, int32_t precision,
                         void* workspace, size_t workspace_bytes, tagan_stream_t stream);

/* bf16-STORAGE variants of kernel (a) (north_star: "bf16 tolerances stated separately"): Q, K, V rows are bf16 (raw 16-bit
 * words, leading dimension in elements), which halves the gather traffic; scores, softmax, aggregation and every gradient
 * are computed and returned in fp32.  Same arguments otherwise.  Parity: against the reference arithmetic with q, k, v rounded
 * to bf16 (round-to-nearest-even) at the fp32 tolerance; against the fp32 reference itself at rtol 2e-2 / atol 2e-2. */
int tagan_geo_attn_fwd_bf16(const uint16_t* Q, const uint16_t* K, const uint16_t* V, int64_t ld, const int32_t* rowptr,
                            const int32_t* col, int32_t N, int32_t H, int32_t heads, int32_t metric,
                            const float* metric_param, float* ctx, float* lse, float* attn, tagan_stream_t stream);
int tagan_geo_attn_bwd_bf16(const uint16_t* Q, const uint16_t* K, const uint16_t* V, int64_t ld, const int32_t* rowptr,
                            const int32_t* col, const int32_t* rowptr_t, const int32_t* row_t, int32_t N, int32_t H,
                            int32_t heads, int32_t metric, const float* metric_param, const float* ctx, const float* lse,
                            const float* dctx, float* dQ, float* dK, float* dV, int64_t ldd, float* delta_ws,
                            float* dparam_ws, float* dparam, tagan_stream_t stream);
int tagan_geo_attn_fwd_part_bf16(const uint16_t* Q, int64_t ldq, const uint16_t* K, const uint16_t* V, int64_t ldkv,
                                 const int32_t* rowptr, const int32_t* col, int32_t n_rows, int32_t H, int32_t heads,
                                 int32_t metric, const float* metric_param, float* ctx, float* lse, float* attn,
                                 tagan_stream_t stream);
int tagan_geo_attn_bwd_part_bf16(const uint16_t* Q, int64_t ldq, const uint16_t* K, const uint16_t* V, int64_t ldkv,
                                 const int32_t* rowptr, const int32_t* col, const int32_t* rowptr_t, const int32_t* row_t,
                                 int32_t n_rows, int32_t n_src, int32_t H, int32_t heads, int32_t metric,
                                 const float* metric_param, const float* ctx, const float* lse, const float* dctx,
                                 float* dQ, int64_t lddq, float* dK, float* dV, int64_t lddkv, float* delta_ws,
                                 float* dparam_ws, float* dparam, tagan_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Projections with the reference's surrounding element-wise / LayerNorm code fused into the GEMM
 * (SURVEY.md section 8b "tagan_gemm_ln_qkv / tagan_gemm_out_res_ln / tagan_gru_epilogue_*"): the
 * accumulator tile is consumed in registers, the intermediate never makes an HBM round trip.
 *
 *   acc[M,N] = [A | A2][M, K] . op(B) (+ bias[N])            op 0 = NT (B[N,K]), 1 = NN (B[K,N])
 *   A2 != NULL: the A operand is the column concatenation of A[M,k1] and A2[M,K-k1] (k1 % 32 == 0, NT only) --
 *   the reference's `torch.cat([x, h], dim=-1)` feeding reset/update/candidate Linears
 *   (src/tagan/layers/temporal_propagation.py:531-538) without materialising the concatenation.
 *
 *   TAGAN_EPI_STORE      out0 = acc
 *   TAGAN_EPI_RES_LN     v = acc + in0 (in0 NULL: none); out1 = v if out1 != NULL (saved for backward);
 *                        out0 = LayerNorm(v; gamma, beta, eps 1e-5), mean/rstd [M] optional; gamma NULL: out0 = v.
 *                        `output_proj -> dropout(p=0) -> + identity -> layer_norm2`
 *                        (geometric_attention.py:586-596, temporal_attention.py:1186-1200,
 *                        temporal_propagation.py:738-753, :929-944, :1487-1500).  Needs N <= 128.
 *   TAGAN_EPI_GATES      s = sigmoid(acc); columns < split: out0 = s (reset gate r), out1 = s * in0 (r * h^);
 *                        columns >= split: out2[:, col-split] = s (update gate z)   (temporal_propagation.py:531-535)
 *   TAGAN_EPI_BLEND      t = tanh(acc); out0 = t (candidate); out1 = (1-in0)*in1 + in0*t, in0 = z, in1 = h^  (:538-542)
 *   TAGAN_EPI_GATES_BWD  d = acc (= d(r*h^)); out0 = d * in1 * r(1-r) with r = in0, h^ = in1; out1 += d * r
 *                        (autograd of :531-538; out0 is the gradient of the reset pre-activation)
 * All pointers 16-byte aligned, every leading dimension and N a multiple of 4.  Returns TAGAN_E_UNSUPPORTED when
 * the shape cannot take the fused path (the caller then composes tagan_gemm with the stand-alone kernels).
 * ------------------------------------------------------------------------------------- */
enum tagan_epi_mode {
  TAGAN_EPI_STORE = 0,
  TAGAN_EPI_RES_LN = 1,
  TAGAN_EPI_GATES = 2,
  TAGAN_EPI_BLEND = 3,
  TAGAN_EPI_GATES_BWD = 4,
  TAGAN_EPI_STORE_BF16 = 5 /* out0 is a bf16 matrix (raw 16-bit words, ld_out0 in elements): acc + bias, round-to-nearest-even */
};
struct tagan_epilogue {
  int32_t mode;
  int32_t split;
  const float* in0; int64_t ld_in0;
  const float* in1; int64_t ld_in1;
  float* out0; int64_t ld_out0;
  float* out1; int64_t ld_out1;
  float* out2; int64_t ld_out2;
  const float* gamma;
  const float* beta;
  float* mean;
  float* rstd;
};
size_t tagan_gemm_fused_workspace_bytes(int32_t op, int64_t m, int64_t n, int64_t k);
/* Tuning knobs of the tcgen05 GEMM, for A/B measurements inside one process (boxes of the pool differ by ~15 %):
 *   key 0  resident weights (default 1): NT / NN projections with K <= 128 keep the CTA's pre-split weight panel (B_hi, B_lo:
 *          K x 128 x 8 bytes) in shared memory for the whole persistent kernel; only the activation tiles stream
 *   key 1  L2 prefetch distance of the TMA producer in 32-wide k-blocks (default 0 = off: measured slower at every distance)
 *   key 2  epilogue issues the TMEM load of the next 32-column chunk before storing the current one (default 0)
 *   key 3  resident mode: activation smem slots are released by the split warps instead of the MMA commit (default 0)
 *   key 4  suspend-time hint (ns) of the mbarrier waits, 0 = plain polling (default 10 000 000, as CUTLASS)
 *   key 5  256-bit global stores (STG.256) in the plain epilogue of interior tiles when C is 32-byte aligned (default 1)
 *   key 6  lean kernel instantiations (default 3): bit 0 pre-split K-major weights without split-K / accumulate / knobs,
 *          bit 1 the dW products (TN, in-kernel split of both operands, CTA-private split-K partial tiles)
 * Keys 1-3 made no difference or a small loss in same-process A/B runs (profiles/r02_SUMMARY.md); they stay for re-measurement. */
void tagan_gemm_set_tuning(int32_t key, int32_t value);
/* Debug aid (tools/trace_gemm.py): a device buffer of 16 x 512 int64 that CTA 0 of every following tcgen05 GEMM launch
 * fills with clock64() stamps of its pipeline (per k-block: TMA issue, bytes landed, split done, MMA start, MMA issued; per
 * tile: accumulator free, accumulator full, epilogue done).  NULL (default) switches it off. */
void tagan_gemm_set_trace(void* buf);
int tagan_gemm_fused(int32_t op /*0=NT,1=NN*/, int64_t m, int64_t n, int64_t k,
                     const float* A, int64_t lda, const float* A2, int64_t lda2, int64_t k1,
                     const float* B, int64_t ldb, const float* bias, const struct tagan_epilogue* epi,
                     int32_t precision, void* workspace, size_t workspace_bytes, tagan_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Fused node-stream passes of the propagation core (csrc/fused_rows.cu): one pass over the rows instead of a chain
 * of element-wise / LayerNorm launches.  cols <= 512; affine gradients are deterministic (fixed-order partials).
 *
 * tagan_ln_pair_fwd: s = LN_out(hn) and, if hhat != NULL, hhat = LN_h(s) * exp(-clamp(ts[:,t_hi]-ts[:,t_hi-1],0,10))
 *   -- the state of the GRU scan between two steps (TemporalGRUCell.forward, temporal_propagation.py:545-546 then
 *   :505-514 of the next step).  ts NULL: no decay.  decay[rows] (optional) receives the row scale for backward.
 * tagan_ln_pair_bwd: dhn = dLN_out(ds_ext + dLN_h(dhh * decay)); daffine[4][cols] (+)= d gamma_o, d beta_o,
 *   d gamma_h, d beta_h.  ds_ext / dhh may be NULL (no external gradient / last step).
 * tagan_gelu_ln_fwd/bwd: y = LN(GELU(a)) (TemporalSkipConnection.forward :866-877); daffine[2][cols] = d gamma, d beta.
 * tagan_window_gelu_fwd/bwd: out[t] = GELU(agg_{|u-t|<=window} p[u]) over the leading axis of p [T, inner], agg 0 = mean,
 *   2 = sum (:880-894 followed by the activation of :929-933); backward rebuilds the aggregate from p.
 * tagan_mse_fwd/bwd: *loss = mean(x^2); dx = (*dloss) * 2 x / n (dloss NULL = 1).
 * ------------------------------------------------------------------------------------- */
int tagan_ln_pair_fwd(const float* hn, int64_t ldhn, const float* gamma_o, const float* beta_o,
                      const float* gamma_h, const float* beta_h, const float* ts, int64_t ldts, int32_t t_hi,
                      float* s, int64_t lds, float* hhat, int64_t ldhh, float* mean_o, float* rstd_o,
                      float* mean_h, float* rstd_h, float* decay, int64_t rows, int32_t cols, tagan_stream_t stream);
size_t tagan_ln_pair_bwd_workspace_bytes(int64_t rows, int32_t cols);
int tagan_ln_pair_bwd(const float* ds_ext, int64_t ldds, const float* dhh, int64_t lddhh, const float* hn, int64_t ldhn,
                      const float* gamma_o, const float* beta_o, const float* gamma_h,
                      const float* mean_o, const float* rstd_o, const float* mean_h, const float* rstd_h,
                      const float* decay, float* dhn, int64_t lddhn, float* daffine, int32_t accumulate,
                      void* workspace, size_t workspace_bytes, int64_t rows, int32_t cols, tagan_stream_t stream);
int tagan_gelu_ln_fwd(const float* a, int64_t lda, const float* gamma, const float* beta, float* y, int64_t ldy,
                      float* mean, float* rstd, int64_t rows, int32_t cols, tagan_stream_t stream);
size_t tagan_gelu_ln_bwd_workspace_bytes(int64_t rows, int32_t cols);
int tagan_gelu_ln_bwd(const float* dy, int64_t lddy, const float* a, int64_t lda, const float* gamma,
                      const float* mean, const float* rstd, float* da, int64_t ldda, float* daffine,
                      void* workspace, size_t workspace_bytes, int64_t rows, int32_t cols, tagan_stream_t stream);
int tagan_window_gelu_fwd(const float* p, float* out, int32_t T, int64_t inner, int32_t window, int32_t agg,
                          tagan_stream_t stream);
int tagan_window_gelu_bwd(const float* dgg, const float* p, float* dp, int32_t T, int64_t inner, int32_t window,
                          int32_t agg, tagan_stream_t stream);
/* GRU step with the blend fused into the LayerNorm pair (one pass per step and direction):
 *   forward:  cand = tanh(cand_pre); hn = (1-z) hhat_cur + z cand (temporal_propagation.py:538-542), then as
 *             tagan_ln_pair_fwd on hn (hn itself is never stored).
 *   backward: rebuilds hn from (cand, z, hhat_cur), runs tagan_ln_pair_bwd and the autograd of the blend:
 *             dgz = dhn (cand-hhat) z(1-z), dgc = dhn z (1-cand^2) (column slices of the step's [rows,3H] gradient
 *             tile, leading dimension lddg), dhh_cur = dhn (1-z).
 * tagan_gru_reset_bwd: dgr = drs * hhat * r(1-r); dhh += drs * r   (autograd of :531-538, drs = d(r*hhat)). */
int tagan_gru_blend_ln_fwd(const float* cand_pre, int64_t ldc, const float* z, const float* hhat_cur, int64_t ldcur,
                           float* cand, const float* gamma_o, const float* beta_o, const float* gamma_h,
                           const float* beta_h, const float* ts, int64_t ldts, int32_t t_hi, float* s, int64_t lds,
                           float* hhat_next, int64_t ldhh, float* mean_o, float* rstd_o, float* mean_h, float* rstd_h,
                           float* decay, int64_t rows, int32_t cols, tagan_stream_t stream);
int tagan_gru_blend_ln_bwd(const float* ds_ext, int64_t ldds, const float* dhh_next, int64_t lddhh, const float* cand,
                           const float* z, const float* hhat_cur, int64_t ldcur, float* dgz, float* dgc, int64_t lddg,
                           float* dhh_cur, int64_t lddcur, const float* gamma_o, const float* beta_o,
                           const float* gamma_h, const float* mean_o, const float* rstd_o, const float* mean_h,
                           const float* rstd_h, const float* decay, float* daffine, int32_t accumulate, void* workspace,
                           size_t workspace_bytes, int64_t rows, int32_t cols, tagan_stream_t stream);
int tagan_gru_reset_bwd(const float* drs, const float* r, const float* hhat, int64_t ldhh, float* dgr, int64_t lddg,
                        float* dhh, int64_t lddhh, int64_t rows, int32_t H, tagan_stream_t stream);
size_t tagan_mse_workspace_bytes(void);
int tagan_mse_fwd(const float* x, int64_t n, float* loss, void* workspace, size_t workspace_bytes, tagan_stream_t stream);
int tagan_mse_bwd(const float* x, int64_t n, const float* dloss, float* dx, tagan_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * (b1,b2) per-node temporal attention over the snapshot axis.
 * Replaces the score/bias/mask/softmax/`attn @ v` core of AsymmetricTemporalAttention.forward
 * (src/tagan/layers/temporal_attention.py:1008-1183).  One CTA per node, (node,head) pairs on
 * (sub-)warps, one lane per query row, K/V staged in shared memory, online softmax.
 *   Q,K,V: rows of a fused [B*T,3H] projection (row stride ld).  Row of (node b, step t) is
 *          b*T+t (time_major = 0, the [B,T,H] layout) or t*B+b (time_major = 1, the stacked
 *          list-of-snapshots layout [T,B,H] -- avoids the permute of :972-976).
 *   bias / bias_t: additive score bias [heads,T,T] in (i,j) and (j,i) order, shared by all nodes
 *          (bias_bstride = 0) or per node (bias_bstride = heads*T*T); NULL = none.  It carries
 *          relative_pos_table (:1011-1021), asymmetric_kernel (:1024-1027) and the RBF time bias
 *          (:792-871); entries may be -inf (shared masks folded in).
 *   ts [B,T] per-node timestamps (needed for mask bit1), band = 10.0 in the reference (:873-903)
 *   mask_flags: bit0 causal (j<=i, :1073-1076); bit1 time band |ts_i-ts_j| <= band;
 *          bit2 "mask is all ones => make it causal" (:1142-1148), decided on device from
 *          *allones_flag (see tagan_tattn_mask_allones) -- no host sync.
 *   mask: explicit uint8 keep-mask [mask_b, mask_h, T, T], mask_b in {1,B}, mask_h in {1,heads};
 *          0 = masked, 1 = value was exactly 1.0, 2 = other non-zero.  NULL = none.
 *   ctx: [rows,H] same row order as Q; lse [B,heads,T]; attn [B,heads,T,T] or NULL.
 * ------------------------------------------------------------------------------------- */
int tagan_tattn_fwd(const float* Q, const float* K, const float* V, int64_t ld,
                    int64_t batch, int32_t T, int32_t hidden, int32_t heads, int32_t time_major,
                    const float* bias, const float* bias_t, int64_t bias_bstride, const float* ts,
                    int32_t mask_flags, float band, const int32_t* allones_flag,
                    const uint8_t* mask, int32_t mask_b, int32_t mask_h,
                    float* ctx, float* lse, float* attn, tagan_stream_t stream);
/* Deterministic backward.  dQ,dK,dV rows like Q (row stride ldd).  dbias (nullable): gradient of
 * the bias in (i,j) order -- [heads,T,T] reduced over nodes through `workspace` (per-CTA partial
 * tables summed in a fixed order) when the bias is shared, [B,heads,T,T] when it is per node. */
size_t tagan_tattn_bwd_workspace_bytes(int64_t batch, int32_t T, int32_t heads);
int tagan_tattn_bwd(const float* Q, const float* K, const float* V, int64_t ld,
                    int64_t batch, int32_t T, int32_t hidden, int32_t heads, int32_t time_major,
                    const float* bias, const float* bias_t, int64_t bias_bstride, const float* ts,
                    int32_t mask_flags, float band, const int32_t* allones_flag,
                    const uint8_t* mask, int32_t mask_b, int32_t mask_h,
                    const float* ctx, const float* lse, const float* dctx,
                    float* dQ, float* dK, float* dV, int64_t ldd,
                    float* dbias, void* workspace, size_t workspace_bytes, tagan_stream_t stream);
/* allones_flag[0] = 1 iff every entry of the effective mask is exactly 1: all band tests
 * |ts_i-ts_j| <= band pass (ts nullable) and every explicit mask byte == 1 (mask nullable).
 * The data-dependent test of temporal_attention.py:1144, evaluated on device. */
int tagan_tattn_mask_allones(const float* ts, int64_t batch, int32_t T, float band,
                             const uint8_t* mask, int64_t mask_elems, int32_t* allones_flag,
                             tagan_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * (b3-b6) element-wise stages of temporal propagation; the Linears go through tagan_gemm and
 * the LayerNorms (with the exp(-dt) decay as `rowscale`) through tagan_layernorm_*.
 *   GRU cell  (TemporalGRUCell.forward, src/tagan/layers/temporal_propagation.py:531-539):
 *       r,z = sigmoid(W[x^,h^]);  rh = r*h^;  h~ = tanh(Wc[x^,rh]);  hn = (1-z)*h^ + z*h~
 *   Gating    (TemporalGatingUnit.forward :1043-1060):
 *       u,r = sigmoid(W[c,p]);  rp = r*p;  cand = tanh(Wo[c,rp]);  out = (1-u)*c + u*cand (+c)
 * `second` is the gated half of the concatenation (h^ / p), `base` the blended one (h^ / c).
 * ------------------------------------------------------------------------------------- */
/* g[rows, >=2H] (row stride ldg) = [reset | update] pre-activations -> r, z = sigmoid; rs = r * second.
 * The strides let g / cand_pre be column slices of one fused [rows,3H] pre-activation buffer. */
int tagan_gates_fwd(const float* g, int64_t ldg, const float* second, int64_t lds, float* r, float* z,
                    float* rs, int64_t ldrs, int64_t rows, int32_t H, tagan_stream_t stream);
int tagan_gates_bwd(const float* drs, int64_t lddrs, const float* dz, const float* r, const float* z,
                    const float* second, int64_t lds, float* dg, int64_t lddg, float* dsecond, int64_t ldds,
                    int32_t accumulate, int64_t rows, int32_t H, tagan_stream_t stream);
/* cand = tanh(cand_pre); out = (1-z)*base + z*cand (+ base if residual) */
int tagan_blend_fwd(const float* cand_pre, int64_t ldc, const float* z, const float* base, int64_t ldb,
                    float* cand, float* out, int32_t residual, int64_t rows, int32_t H,
                    tagan_stream_t stream);
int tagan_blend_bwd(const float* dout, const float* z, const float* cand, const float* base, int64_t ldb,
                    float* dcand_pre, int64_t lddc, float* dz, float* dbase, int64_t lddb, int32_t accumulate,
                    int32_t residual, int64_t rows, int32_t H, tagan_stream_t stream);
/* sliding window over the snapshot axis of p[T, inner] (TemporalSkipConnection.forward :880-926):
 * out[t] = agg_{u in [max(0,t-w), min(T,t+w+1))} p[u];  agg: 0 mean, 1 max, 2 sum. */
int tagan_skip_window_fwd(const float* p, float* out, int32_t T, int64_t inner, int32_t window,
                          int32_t agg, tagan_stream_t stream);
int tagan_skip_window_bwd(const float* dg, const float* p, const float* agg_out, float* dp,
                          int32_t T, int64_t inner, int32_t window, int32_t agg,
                          tagan_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * (c1-c3) node memory bank on dense device tables keyed by slot (= node id).
 * Replaces NodeMemoryBank's python dicts (src/tagan/utils/memory_bank.py:14-360).
 *   table[cap,H] fp32 states; valid,has_seen u8[cap]; last_seen,inactivity,frequency i32[cap].
 *   status[1] int32 is set to 1 if an id was outside [0, cap).
 * ------------------------------------------------------------------------------------- */
/* get_states (:187-211): out[i] = table[ids[i]] or zeros; unknown ids are inserted with a zero
 * state, inactivity 0 and no last_seen. */
int tagan_bank_gather(float* table, uint8_t* valid, int32_t* inactivity, const int32_t* ids,
                      float* out, int64_t M, int32_t H, int32_t capacity, int32_t* status,
                      tagan_stream_t stream);
/* update (:65-173): inactivity++ on every known id; frequency++ per occurrence; the LAST occurrence
 * of a listed id writes its state -- blended with the stored one (w*prev + (1-w)*cur, :121-129)
 * only if the id is listed once and reappears after a gap (last_seen < timestep-1); every stored id
 * NOT listed is multiplied by decay^inactivity (:149-153); ids with inactivity > max_inactivity
 * are pruned (:156-166); size_out[0] = surviving ids (:169).  Only min(num_ids, num_states) ids
 * have a state row (bounds check :95); the rest still count as listed.
 *   w23 (HOST): {w(gap 2), 1-w(gap 2), w(gap>=3), 1-w(gap>=3)} rounded to fp32 from the python
 *   doubles max(0.4, decay**min(gap,3)); decay_pow (DEVICE) [decay_pow_len]: decay**k rounded from
 *   double, `decay` (double) used beyond the table.  mark_ws: int32[2*capacity] scratch.
 * States and all integer bookkeeping are bit-exact with the reference. */
int tagan_bank_update(float* table, uint8_t* valid, uint8_t* has_seen, int32_t* last_seen,
                      int32_t* inactivity, int32_t* frequency, const int32_t* ids, int64_t num_ids,
                      const float* states, int64_t lds, int64_t num_states, int32_t H,
                      int32_t capacity, int32_t timestep, const float* w23, const float* decay_pow,
                      int32_t decay_pow_len, double decay, int32_t max_inactivity, int32_t* mark_ws,
                      int32_t* size_out, int32_t* status, tagan_stream_t stream);
/* decay_all (:222-225): every stored state *= decay */
int tagan_bank_decay_all(float* table, const uint8_t* valid, float decay, int32_t H, int32_t capacity,
                         tagan_stream_t stream);

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

/* ---------------------------------------------------------------------------------------
 * (b2) RBF time bias for PER-NODE timestamps (AsymmetricTemporalAttention._compute_time_based_attention,
 * temporal_attention.py:792-871, with TimeEncoding._get_basis_encoding :122-220), on device and chunk by chunk:
 *   tagan_ts_range: *range = max over nodes of (max_t ts - min_t ts); the reference's global min / max of all pairwise
 *     differences (:142-152) are -range and +range.
 *   tagan_time_bias_fwd: for nodes [node_begin, node_begin+nodes): bias[b,h,i,j] = pos_bias[h,i,j] + bc[h] +
 *     sum_k wc[h,k] exp(clamp(-(tn-mu_k)^2/(2 sigma_k^2), -88, 88)), tn = (ts_i - ts_j + range) / (2 range) (0 if the
 *     range is degenerate); bias_t is the (j,i) transpose.  wc = time_q_proj.weight @ basis_proj.weight (time_k_proj is
 *     unused, :848).  The tiles feed tagan_tattn_fwd/bwd_strided with bias_bstride = heads*T*T.
 *   tagan_time_bias_bwd: dbias tile of the same nodes -> dparams (+)= [heads*nb d wc | heads d bc | nb d mu | nb d sigma],
 *     dpos (+)= sum_b dbias (optional).  Deterministic (fixed-order partials).
 * tagan_tattn_fwd/bwd_strided: as tagan_tattn_fwd/bwd with the row strides of the node and snapshot axes given explicitly
 *   (row of (node b, step t) = b*row_stride_b + t*row_stride_t), so a chunk of nodes of a time-major tensor can be addressed.
 * ------------------------------------------------------------------------------------- */
int tagan_ts_range(const float* ts, int64_t B, int32_t T, float* range, tagan_stream_t stream);
int tagan_time_bias_fwd(const float* ts, int64_t node_begin, int64_t nodes, int32_t T, int32_t heads, int32_t num_bases,
                        const float* range, const float* mu, const float* sigma, const float* wc, const float* bc,
                        const float* pos_bias, float* bias, float* bias_t, tagan_stream_t stream);
size_t tagan_time_bias_bwd_workspace_bytes(int32_t heads, int32_t num_bases);
int tagan_time_bias_bwd(const float* ts, int64_t node_begin, int64_t nodes, int32_t T, int32_t heads, int32_t num_bases,
                        const float* range, const float* mu, const float* sigma, const float* wc, const float* dbias,
                        float* dparams, float* dpos, int32_t accumulate, void* workspace, size_t workspace_bytes,
                        tagan_stream_t stream);
int tagan_tattn_fwd_strided(const float* Q, const float* K, const float* V, int64_t ld, int64_t B, int32_t T,
                            int32_t H, int32_t heads, int64_t row_stride_b, int64_t row_stride_t, const float* bias,
                            const float* bias_t, int64_t bias_bstride, const float* ts, int32_t mask_flags, float band,
                            const int32_t* allones_flag, const uint8_t* mask, int32_t mask_b, int32_t mask_h,
                            float* ctx, float* lse, float* attn, tagan_stream_t stream);
int tagan_tattn_bwd_strided(const float* Q, const float* K, const float* V, int64_t ld, int64_t B, int32_t T,
                            int32_t H, int32_t heads, int64_t row_stride_b, int64_t row_stride_t, const float* bias,
                            const float* bias_t, int64_t bias_bstride, const float* ts, int32_t mask_flags, float band,
                            const int32_t* allones_flag, const uint8_t* mask, int32_t mask_b, int32_t mask_h,
                            const float* ctx, const float* lse, const float* dctx, float* dQ, float* dK, float* dV,
                            int64_t ldd, float* dbias, void* workspace, size_t workspace_bytes, tagan_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * After the hot path: pooling, classification head + loss, optimizer step -- one launch
 * each, no host sync, so TAGAN.forward + backward + the trainer step capture into one CUDA graph.
 *
 * tagan_pack_padded_fwd/bwd: packed rows [sum N_t, H] + offsets[T+1] (int32) <-> zero-padded [T, maxn, H]; the pad + stack
 *   of AsymmetricTemporalAttention.forward for ragged snapshots (temporal_attention.py:928-976).
 * tagan_pool_blocks_fwd/bwd: out[t] = mean of logical rows [t*B, (t+1)*B) of x viewed as the row-major [B*T, H] matrix of
 *   x[B,T,H] (time_major = 1: the storage is [T,B,H]) -- TAGAN.forward's node pooling (model.py:377-427; both of its branches
 *   reduce to this block mean, including the `view(T,-1,H)` scrambling of node and time).
 * tagan_head_fwd/bwd: TemporalClassificationHead with attention pooling (classification.py:743-975): Linear+Tanh+Linear(no
 *   bias) scores -> softmax over T -> weighted sum -> Linear -> LayerNorm (ln_weight NULL: none) -> ReLU -> Linear, then
 *   loss_type 0: binary_cross_entropy_with_logits mean over labels [label_rows, O] (label_rows == Bsz, or Bsz == 1 broadcast:
 *   TemporalLossFunction :420-456), 1: cross entropy over class_index[Bsz] (model.py:436-438).  gf [Bsz,T,H].
 *   Forward saves u [Bsz,T,H], alpha [Bsz,T], pooled/h1/hn [Bsz,H], stats [Bsz,2], logits [Bsz,O]; loss may be NULL.
 *   Backward: dloss (device scalar, NULL = no loss term) and/or dlogits [Bsz,O]; dw holds the gradient buffers.
 * tagan_adam_clip_step: *step += 1; clip = min(1, max_grad_norm / (sqrt(*grad_mean_sq * n) + 1e-6)) (max_grad_norm <= 0: off;
 *   torch.nn.utils.clip_grad_norm_), then torch.optim.Adam's update (L2 weight decay, bias correction) on flat buffers
 *   (trainer.py:295-311).  grad_mean_sq = tagan_mse_fwd of the flat gradient.
 * ------------------------------------------------------------------------------------- */
struct tagan_head_weights {
  const float* attn0_weight; /* [H,H] */
  const float* attn0_bias;   /* [H]   */
  const float* attn2_weight; /* [H] (Linear(H,1,bias=False)) */
  const float* fc0_weight;   /* [H,H] */
  const float* fc0_bias;     /* [H]   */
  const float* ln_weight;    /* [H] or NULL */
  const float* ln_bias;      /* [H] or NULL */
  const float* fc1_weight;   /* [O,H] */
  const float* fc1_bias;     /* [O]   */
};
int tagan_pack_padded_fwd(const float* packed, const int32_t* offsets, float* padded, int32_t T, int32_t maxn, int32_t H,
                          tagan_stream_t stream);
int tagan_pack_padded_bwd(const float* dpadded, const int32_t* offsets, float* dpacked, int32_t T, int32_t maxn, int32_t H,
                          tagan_stream_t stream);
size_t tagan_pool_blocks_workspace_bytes(int32_t T, int32_t H);
int tagan_pool_blocks_fwd(const float* x, int64_t B, int32_t T, int32_t H, int32_t time_major, float* out,
                          void* workspace, size_t workspace_bytes, tagan_stream_t stream);
int tagan_pool_blocks_bwd(const float* dout, int64_t B, int32_t T, int32_t H, int32_t time_major, float* dx,
                          tagan_stream_t stream);
int tagan_head_fwd(const struct tagan_head_weights* w, const float* gf, int32_t Bsz, int32_t T, int32_t H, int32_t O,
                   int32_t loss_type, const float* labels, int32_t label_rows, const int64_t* class_index,
                   float* u, float* alpha, float* pooled, float* h1, float* hn, float* stats, float* logits,
                   float* loss, tagan_stream_t stream);
size_t tagan_head_bwd_workspace_bytes(int32_t Bsz, int32_t T, int32_t H, int32_t O);
int tagan_head_bwd(const struct tagan_head_weights* w, const float* gf, int32_t Bsz, int32_t T, int32_t H, int32_t O,
                   int32_t loss_type, const float* labels, int32_t label_rows, const int64_t* class_index,
                   const float* u, const float* alpha, const float* pooled, const float* h1, const float* hn,
                   const float* stats, const float* logits, const float* dloss, const float* dlogits,
                   float* dgf, struct tagan_head_weights* dw, void* workspace, size_t workspace_bytes,
                   tagan_stream_t stream);
int tagan_adam_clip_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                         float beta1, float beta2, float eps, float weight_decay, float max_grad_norm,
                         const float* grad_mean_sq, int32_t* step, tagan_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TAGAN_B200_H */
