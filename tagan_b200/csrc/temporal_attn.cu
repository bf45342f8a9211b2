// (b1,b2) Per-node temporal attention over the snapshot axis.
// Replaces the score / bias / mask / softmax / `attn @ v` core of the reference's
// AsymmetricTemporalAttention.forward (src/tagan/layers/temporal_attention.py:1008-1183), which
// materialises [N,h,T,T] score tensors (and [N,T,T,H] for the RBF time bias).
//
// Work unit = one (node, head) pair: K_h,V_h [T,D] staged in shared memory, one lane per query
// row (T <= 32: several pairs share a warp as sub-warps of TP = pow2 >= T lanes; T > 32: a lane
// walks rows lane, lane+32, ...).  Scores, additive bias table, masks and an online softmax are
// evaluated in registers; keys are read from shared memory as broadcasts.  One CTA processes
// all heads of one node at a time (grid-stride over nodes), so a node's [T,3H] block of the fused
// QKV projection is read from HBM once.
//
// Additive bias: bias[h,i,j] (+ its transpose bias_t[h,j,i], so both passes read it coalesced),
// shared by all nodes (bias_bstride = 0) or per node.  It carries the relative-position table
// (:1011-1021), the asymmetric window kernel (:1024-1027) and -- when timestamps are shared -- the
// RBF time bias (:792-871) and any shared mask folded in as -inf.  Per-node masks: causal flag,
// time band |ts_i - ts_j| <= band (:873-903), explicit uint8 keep-mask, and the data-dependent
// "all-ones mask => causal" rule (:1142-1148) read from a device flag (no host sync).
//
// Backward is deterministic: phase 1 (lane = query row) gives dQ and dBias, phase 2 (lane = key
// row) recomputes the probabilities and gives dK, dV.  dBias for a shared table is reduced over
// nodes through per-CTA partial tables summed in a fixed order.
#include "common.cuh"
#include "temporal_attn_shared.cuh"

namespace {

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
// `rbw` warps share one (node, head) slot when T > 32: they load its K/V tile together and each takes the query-row blocks
// rb = r, r + rbw, ... (T = 128: 4 warps per head instead of one warp walking 4 row blocks -- the kernel is bound by
// instruction issue at low occupancy, not by memory).  `warps` counts SLOT groups; the CTA has warps * rbw warps.
template <int D, bool WIDE>
__global__ void __launch_bounds__(WIDE ? GEN_MAX_THREADS : MAX_WARPS * 32)
tattn_fwd_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, int64_t ld,
                 int64_t B, int T, int heads, int64_t rsb, int64_t rst, const float* __restrict__ bias_t,
                 int64_t bias_bstride, MaskSpec ms, float* __restrict__ ctx, float* __restrict__ lse,
                 float* __restrict__ attn, int TP, int warps, int rbw) {
  extern __shared__ __align__(16) float smem[];
  const int H = heads * D;
  const int wfull = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = wfull / rbw, rw = wfull - w * rbw;   // slot group, row-block worker inside it
  const int PPW = 32 / TP;                       // pairs per warp (1 when T >= 17)
  const int sub = lane / TP, li = lane - sub * TP;
  const int slot_floats = 2 * T * D + ((T + 3) & ~3);
  float* slot = smem + (size_t)(w * PPW + sub) * slot_floats;
  float* Ks = slot;
  float* Vs = slot + T * D;
  float* ts_s = slot + 2 * T * D;
  const float scale = 1.f / sqrtf((float)D);
  const bool causal = (ms.flags & 1) || ((ms.flags & 4) && ms.allones_flag && *ms.allones_flag != 0);
  const int RB = (T + 31) / 32;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    for (int hb = 0; hb < heads; hb += warps * PPW) {
      const int hd = hb + w * PPW + sub;
      const bool pv = hd < heads;
      if (rbw > 1) __syncthreads(); else __syncwarp();
      if (pv) {
        const int64_t base = b * rsb * ld + (int64_t)hd * D;
        load_tile<D>(Ks, K + base, rst * ld, T, li + rw * TP, TP * rbw);
        load_tile<D>(Vs, V + base, rst * ld, T, li + rw * TP, TP * rbw);
        if (ms.ts) for (int t = li + rw * TP; t < T; t += TP * rbw) ts_s[t] = ms.ts[b * T + t];
      }
      if (rbw > 1) __syncthreads(); else __syncwarp();
      const uint8_t* mbase = nullptr;
      if (ms.mask && pv)
        mbase = ms.mask + ((int64_t)(ms.mask_b > 1 ? b : 0) * ms.mask_h + (ms.mask_h > 1 ? hd : 0)) * T * T;
      const float* bt = (bias_t && pv) ? bias_t + b * bias_bstride + (int64_t)hd * T * T : nullptr;
      const bool windowed = mbase == nullptr && TP == 32;           // one slot per warp, no explicit mask tensor
      const bool sorted = windowed && (ms.flags & 2) && ms.ts && pv && ts_sorted_warp(ts_s, T, lane);
      for (int rb = rw; rb < RB; rb += rbw) {
        const int i = rb * 32 + li;
        const bool rv = pv && i < T;
        int jlo = 0, jhi = T - 1;
        if (windowed) {                                              // the warp's union of key windows
          int lo = T, hi = -1;
          if (rv) valid_window(ms, causal, ms.ts ? ts_s : nullptr, sorted, T, i, true, &lo, &hi);
          jlo = __reduce_min_sync(FULL_MASK, lo);
          jhi = __reduce_max_sync(FULL_MASK, hi);
        }
        float q[D], acc[D];
        float m = -INFINITY, l = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = 0.f;
        if (rv) {
          const float* qp = Q + (b * rsb + i * rst) * ld + (int64_t)hd * D;
#pragma unroll
          for (int c = 0; c < D; c += 4) {
            float4 t4 = __ldg(reinterpret_cast<const float4*>(qp + c));
            q[c] = t4.x; q[c + 1] = t4.y; q[c + 2] = t4.z; q[c + 3] = t4.w;
          }
        } else {
#pragma unroll
          for (int c = 0; c < D; ++c) q[c] = 0.f;
        }
        for (int j = jlo; j <= jhi; ++j) {
          const bool kv = rv && key_valid(ms, causal, ms.ts ? ts_s : nullptr, mbase, T, i, j);
          if (!__any_sync(FULL_MASK, kv)) continue;
          float s = -INFINITY;
          if (kv) {
            s = dot_smem<D>(q, Ks + j * D) * scale;
            if (bt) s += bt[(int64_t)j * T + i];
          }
          if (s > -INFINITY) {
            const float mn = fmaxf(m, s);
            const float sc = expf(m - mn);
            const float p = expf(s - mn);
            l = fmaf(l, sc, p);
            const float* vr = Vs + j * D;
#pragma unroll
            for (int c = 0; c < D; c += 4) {
              float4 v4 = *reinterpret_cast<const float4*>(vr + c);
              acc[c] = fmaf(acc[c], sc, p * v4.x); acc[c + 1] = fmaf(acc[c + 1], sc, p * v4.y);
              acc[c + 2] = fmaf(acc[c + 2], sc, p * v4.z); acc[c + 3] = fmaf(acc[c + 3], sc, p * v4.w);
            }
            m = mn;
          }
        }
        const float ls = m + logf(l);
        if (rv) {
          const float inv = 1.f / l;
          float* op = ctx + (b * rsb + i * rst) * (int64_t)H + (int64_t)hd * D;
#pragma unroll
          for (int c = 0; c < D; c += 4)
            *reinterpret_cast<float4*>(op + c) = make_float4(acc[c] * inv, acc[c + 1] * inv, acc[c + 2] * inv, acc[c + 3] * inv);
          lse[(b * heads + hd) * T + i] = ls;
        }
        if (attn != nullptr) {
          for (int j = 0; j < T; ++j) {
            const bool kv = rv && key_valid(ms, causal, ms.ts ? ts_s : nullptr, mbase, T, i, j);
            float p = 0.f;
            if (kv) {
              float s = dot_smem<D>(q, Ks + j * D) * scale;
              if (bt) s += bt[(int64_t)j * T + i];
              p = expf(s - ls);
            }
            if (rv) attn[((b * heads + hd) * T + i) * (int64_t)T + j] = p;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------
template <int D, bool WIDE>
__global__ void __launch_bounds__(WIDE ? GEN_MAX_THREADS : MAX_WARPS * 32)
tattn_bwd_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V, int64_t ld,
                 int64_t B, int T, int heads, int64_t rsb, int64_t rst, const float* __restrict__ bias,
                 const float* __restrict__ bias_t, int64_t bias_bstride, MaskSpec ms, const float* __restrict__ ctx, const float* __restrict__ lse,
                 const float* __restrict__ dctx, float* __restrict__ dQ, float* __restrict__ dK, float* __restrict__ dV,
                 int64_t ldd, float* __restrict__ dbias_out /* per-node [B,h,T,T] or per-CTA partial [grid,h,T,T] */,
                 int dbias_per_node, int TP, int warps, int rbw) {
  extern __shared__ __align__(16) float smem[];
  const int H = heads * D;
  const int wfull = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = wfull / rbw, rw = wfull - w * rbw;   // slot group, row-block worker inside it (see tattn_fwd_kernel)
  const int PPW = 32 / TP;
  const int sub = lane / TP, li = lane - sub * TP;
  const int tpad = (T + 3) & ~3;
  const int slot_floats = 4 * T * D + 3 * tpad;
  float* slot = smem + (size_t)(w * PPW + sub) * slot_floats;
  float* Qs = slot;
  float* Ks = Qs + T * D;
  float* Vs = Ks + T * D;
  float* Gs = Vs + T * D;
  float* lse_s = Gs + T * D;
  float* del_s = lse_s + tpad;
  float* ts_s = del_s + tpad;
  const float scale = 1.f / sqrtf((float)D);
  const bool causal = (ms.flags & 1) || ((ms.flags & 4) && ms.allones_flag && *ms.allones_flag != 0);
  const int RB = (T + 31) / 32;
  const int64_t hTT = (int64_t)heads * T * T;
  const bool want_pn = dbias_out != nullptr && dbias_per_node != 0;
  // zero this CTA's partial table once (shared-bias mode)
  if (dbias_out && !dbias_per_node) {
    float* part = dbias_out + (int64_t)blockIdx.x * hTT;
    for (int64_t x = threadIdx.x; x < hTT; x += blockDim.x) part[x] = 0.f;
    __syncthreads();
  }
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    for (int hb = 0; hb < heads; hb += warps * PPW) {
      const int hd = hb + w * PPW + sub;
      const bool pv = hd < heads;
      const int lw = li + rw * TP, nlw = TP * rbw;           // this lane among the lanes that share the slot
      if (rbw > 1) __syncthreads(); else __syncwarp();
      if (pv) {
        const int64_t base = b * rsb * ld + (int64_t)hd * D;
        load_tile<D>(Qs, Q + base, rst * ld, T, lw, nlw);
        load_tile<D>(Ks, K + base, rst * ld, T, lw, nlw);
        load_tile<D>(Vs, V + base, rst * ld, T, lw, nlw);
        load_tile<D>(Gs, dctx + b * rsb * (int64_t)H + (int64_t)hd * D, rst * H, T, lw, nlw);
        if (ms.ts) for (int t = lw; t < T; t += nlw) ts_s[t] = ms.ts[b * T + t];
      }
      if (rbw > 1) __syncthreads(); else __syncwarp();
      if (pv) {
        for (int t = lw; t < T; t += nlw) {
          const float* cp = ctx + (b * rsb + t * rst) * (int64_t)H + (int64_t)hd * D;
          float dl = 0.f;
#pragma unroll
          for (int c = 0; c < D; c += 4) {
            float4 c4 = __ldg(reinterpret_cast<const float4*>(cp + c));
            const float* g = Gs + t * D + c;
            dl = fmaf(g[0], c4.x, dl); dl = fmaf(g[1], c4.y, dl); dl = fmaf(g[2], c4.z, dl); dl = fmaf(g[3], c4.w, dl);
          }
          del_s[t] = dl;
          lse_s[t] = lse[(b * heads + hd) * T + t];
        }
      }
      if (rbw > 1) __syncthreads(); else __syncwarp();
      const uint8_t* mbase = nullptr;
      if (ms.mask && pv)
        mbase = ms.mask + ((int64_t)(ms.mask_b > 1 ? b : 0) * ms.mask_h + (ms.mask_h > 1 ? hd : 0)) * T * T;
      const float* bij = (bias && pv) ? bias + b * bias_bstride + (int64_t)hd * T * T : nullptr;
      const float* bji = (bias_t && pv) ? bias_t + b * bias_bstride + (int64_t)hd * T * T : nullptr;
      float* db = nullptr;
      if (dbias_out && pv)
        db = dbias_out + (dbias_per_node ? b * hTT : (int64_t)blockIdx.x * hTT) + (int64_t)hd * T * T;
      // per-node bias gradients are written for every (i, j), so only the shared-bias mode may skip positions
      const bool windowed2 = mbase == nullptr && TP == 32;         // phase 2 (dK, dV) may always skip
      const bool windowed = windowed2 && !want_pn;
      const bool sorted = windowed2 && (ms.flags & 2) && ms.ts && pv && ts_sorted_warp(ts_s, T, lane);
      // ---- phase 1: lane owns query row i -> dQ_i, dBias[i,:]
      for (int rb = rw; rb < RB; rb += rbw) {
        const int i = rb * 32 + li;
        const bool rv = pv && i < T;
        int jlo = 0, jhi = T - 1;
        if (windowed) {
          int lo = T, hi = -1;
          if (rv) valid_window(ms, causal, ms.ts ? ts_s : nullptr, sorted, T, i, true, &lo, &hi);
          jlo = __reduce_min_sync(FULL_MASK, lo);
          jhi = __reduce_max_sync(FULL_MASK, hi);
        }
        float q[D], g[D], dq[D];
#pragma unroll
        for (int c = 0; c < D; ++c) { q[c] = rv ? Qs[i * D + c] : 0.f; g[c] = rv ? Gs[i * D + c] : 0.f; dq[c] = 0.f; }
        const float ls = rv ? lse_s[i] : 0.f, dl = rv ? del_s[i] : 0.f;
        for (int j = jlo; j <= jhi; ++j) {
          const bool kv = rv && key_valid(ms, causal, ms.ts ? ts_s : nullptr, mbase, T, i, j);
          if (!__any_sync(FULL_MASK, kv) && !want_pn) continue;
          float ds = 0.f;
          if (kv) {
            float s = dot_smem<D>(q, Ks + j * D) * scale;
            if (bji) s += bji[(int64_t)j * T + i];
            if (s > -INFINITY) {
              const float p = expf(s - ls);
              const float dp = dot_smem<D>(g, Vs + j * D);
              ds = p * (dp - dl);
              const float dss = ds * scale;
              const float* kr = Ks + j * D;
#pragma unroll
              for (int c = 0; c < D; ++c) dq[c] = fmaf(dss, kr[c], dq[c]);
            }
          }
          if (db && rv) {
            // (cta, head, i, j) is owned by exactly this lane.  The CTA-private partial table (shared-bias mode) is kept
            // TRANSPOSED ([j][i]): lanes are consecutive i, so the read-modify-write is one 128-byte line per warp instead of 32
            // sectors a stride of T apart (T = 128: this was most of the kernel's time); the reduction transposes it back
            if (dbias_per_node) db[(int64_t)i * T + j] = ds;
            else if (kv) { float* o = db + (int64_t)j * T + i; *o += ds; }   // the partial table starts at zero: masked pairs add nothing
          }
        }
        if (rv) {
          float* op = dQ + (b * rsb + i * rst) * ldd + (int64_t)hd * D;
#pragma unroll
          for (int c = 0; c < D; c += 4) *reinterpret_cast<float4*>(op + c) = make_float4(dq[c], dq[c + 1], dq[c + 2], dq[c + 3]);
        }
      }
      // ---- phase 2: lane owns key row j -> dK_j, dV_j
      for (int rb = rw; rb < RB; rb += rbw) {
        const int j = rb * 32 + li;
        const bool rv = pv && j < T;
        float k[D], v[D], dk[D], dv[D];
#pragma unroll
        for (int c = 0; c < D; ++c) { k[c] = rv ? Ks[j * D + c] : 0.f; v[c] = rv ? Vs[j * D + c] : 0.f; dk[c] = 0.f; dv[c] = 0.f; }
        int ilo = 0, ihi = T - 1;
        if (windowed2) {                                              // the warp's union of row windows (dBias is phase 1's)
          int lo = T, hi = -1;
          if (rv) valid_window(ms, causal, ms.ts ? ts_s : nullptr, sorted, T, j, false, &lo, &hi);
          ilo = __reduce_min_sync(FULL_MASK, lo);
          ihi = __reduce_max_sync(FULL_MASK, hi);
        }
        for (int i = ilo; i <= ihi; ++i) {
          const bool kv = rv && key_valid(ms, causal, ms.ts ? ts_s : nullptr, mbase, T, i, j);
          if (!__any_sync(FULL_MASK, kv)) continue;
          if (kv) {
            float s = dot_smem<D>(k, Qs + i * D) * scale;
            if (bij) s += bij[(int64_t)i * T + j];
            if (s > -INFINITY) {
              const float p = expf(s - lse_s[i]);
              const float* gr = Gs + i * D;
              const float dp = dot_smem<D>(v, gr);
              const float dss = p * (dp - del_s[i]) * scale;
              const float* qr = Qs + i * D;
#pragma unroll
              for (int c = 0; c < D; ++c) { dv[c] = fmaf(p, gr[c], dv[c]); dk[c] = fmaf(dss, qr[c], dk[c]); }
            }
          }
        }
        if (rv) {
          float* okp = dK + (b * rsb + j * rst) * ldd + (int64_t)hd * D;
          float* ovp = dV + (b * rsb + j * rst) * ldd + (int64_t)hd * D;
#pragma unroll
          for (int c = 0; c < D; c += 4) {
            *reinterpret_cast<float4*>(okp + c) = make_float4(dk[c], dk[c + 1], dk[c + 2], dk[c + 3]);
            *reinterpret_cast<float4*>(ovp + c) = make_float4(dv[c], dv[c + 1], dv[c + 2], dv[c + 3]);
          }
        }
      }
    }
  }
}


__global__ void reduce_parts(const float* __restrict__ partial, int parts, int64_t n, float* __restrict__ out) {
  int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  float s = 0.f;
  for (int p = 0; p < parts; ++p) s += partial[(int64_t)p * n + x];
  out[x] = s;
}
// same, for partial tables stored [head][j][i] (the generic backward kernel): out[head][i][j]
__global__ void reduce_parts_transposed(const float* __restrict__ partial, int parts, int heads, int T, float* __restrict__ out) {
  const int64_t n = (int64_t)heads * T * T;
  int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;        // index into the [head][j][i] layout (coalesced reads)
  if (x >= n) return;
  float s = 0.f;
  for (int p = 0; p < parts; ++p) s += partial[(int64_t)p * n + x];
  const int64_t h = x / ((int64_t)T * T);
  const int r = (int)(x - h * T * T), j = r / T, i = r - j * T;
  out[(h * T + i) * T + j] = s;
}

// allones[0] = 1 iff all band tests pass for every node and every explicit mask byte is non-zero
__global__ void mask_allones_kernel(const float* __restrict__ ts, int64_t B, int T, float band,
                                    const uint8_t* __restrict__ mask, int64_t mask_elems, int* __restrict__ flag) {
  int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool bad = false;
  if (ts != nullptr) {
    // a node's band mask is all ones iff every pairwise |ts_i - ts_j| <= band (checked pairwise so
    // that NaN timestamps behave like the reference's comparison)
    for (int64_t p = x; p < B * T; p += (int64_t)gridDim.x * blockDim.x) {
      const int64_t b = p / T;
      const float ti = ts[p];
      for (int j = 0; j < T; ++j)
        if (!(fabsf(ti - ts[b * T + j]) <= band)) bad = true;
    }
  }
  if (mask != nullptr)
    for (int64_t p = x; p < mask_elems; p += (int64_t)gridDim.x * blockDim.x)
      if (mask[p] != 1) bad = true;   // 1 = exactly 1.0, 2 = other non-zero, 0 = masked
  if (bad) *flag = 0;
}
__global__ void set_flag(int* flag, int v) { *flag = v; }

struct Cfg { int TP, PPW, warps, rbw; size_t smem; };
bool make_cfg(int T, int D, int heads, int slot_floats, Cfg* c) {
  int TP = 32;
  if (T <= 16) { TP = 1; while (TP < T) TP <<= 1; if (TP < 4) TP = 4; }
  c->TP = TP;
  c->PPW = 32 / TP;
  size_t per_warp = (size_t)c->PPW * slot_floats * sizeof(float);
  int want = (heads + c->PPW - 1) / c->PPW;
  int fit = (int)(SMEM_LIMIT / per_warp);
  if (fit < 1) return false;
  int warps = want < fit ? want : fit;
  if (warps > MAX_WARPS) warps = MAX_WARPS;
  // long sequences (more than one block of 32 query rows): up to 4 warps share a (node, head) slot, GEN_MAX_THREADS per CTA
  int rbw = 1;
  if (T > 32 && D <= 32) {                 // (D = 64 keeps one warp per slot: its register footprint needs the 256-thread bound)
    const int rb = (T + 31) / 32;
    rbw = rb < 4 ? rb : 4;
    while (warps * rbw * 32 > GEN_MAX_THREADS && warps > 1) --warps;
    while (warps * rbw * 32 > GEN_MAX_THREADS && rbw > 1) --rbw;
  }
  c->warps = warps;
  c->rbw = rbw;
  c->smem = per_warp * warps;
  return true;
}

bool check_shape(int T, int H, int heads, int* D) {
  if (T <= 0 || H <= 0 || heads <= 0 || H % heads) return false;
  *D = H / heads;
  return *D == 4 || *D == 8 || *D == 16 || *D == 32 || *D == 64;
}

#define DISPATCH_D(FN, WIDE, ...)                                  \
  if (WIDE) {                                                      \
    switch (D) {                                                   \
      case 4: FN<4, true> __VA_ARGS__; break;                      \
      case 8: FN<8, true> __VA_ARGS__; break;                      \
      case 16: FN<16, true> __VA_ARGS__; break;                    \
      default: FN<32, true> __VA_ARGS__; break;                    \
    }                                                              \
  } else {                                                         \
    switch (D) {                                                   \
      case 4: FN<4, false> __VA_ARGS__; break;                     \
      case 8: FN<8, false> __VA_ARGS__; break;                     \
      case 16: FN<16, false> __VA_ARGS__; break;                   \
      case 32: FN<32, false> __VA_ARGS__; break;                   \
      default: FN<64, false> __VA_ARGS__; break;                   \
    }                                                              \
  }

template <typename KernelT>
int set_smem(KernelT kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

int bwd_grid(int64_t B) { return (int)(B < 148 * 4 ? B : 148 * 4); }

}  // namespace

static int tattn_fwd_impl(const float* Q, const float* K, const float* V, int64_t ld, int64_t B, int32_t T,
                          int32_t H, int32_t heads, int64_t rsb, int64_t rst, const float* bias, const float* bias_t,
                          int64_t bias_bstride, const float* ts,
                          int32_t mask_flags, float band, const int32_t* allones_flag, const uint8_t* mask,
                          int32_t mask_b, int32_t mask_h, float* ctx, float* lse, float* attn,
                          tagan_stream_t stream) {
  if (!Q || !K || !V || !ctx || !lse || B < 0 || ld < H) return TAGAN_E_INVALID;
  if ((mask_flags & 2) && !ts) return TAGAN_E_INVALID;
  if ((mask_flags & 4) && !allones_flag) return TAGAN_E_INVALID;
  int D;
  if (!check_shape(T, H, heads, &D) || (ld % 4)) return TAGAN_E_UNSUPPORTED;
  if (B == 0) return 0;
  Cfg c;
  if (!make_cfg(T, D, heads, 2 * T * D + ((T + 3) & ~3), &c)) return TAGAN_E_UNSUPPORTED;
  MaskSpec ms{ts, mask_flags, band, allones_flag, mask, mask_b, mask_h};
  const int grid = (int)(B < 148 * 8 ? B : 148 * 8);
  cudaStream_t st = as_stream(stream);
  int rc = 0;
  if (c.rbw > 1) {
    switch (D) {
      case 4: rc = set_smem(tattn_fwd_kernel<4, true>, c.smem); break;
      case 8: rc = set_smem(tattn_fwd_kernel<8, true>, c.smem); break;
      case 16: rc = set_smem(tattn_fwd_kernel<16, true>, c.smem); break;
      default: rc = set_smem(tattn_fwd_kernel<32, true>, c.smem); break;
    }
  } else {
    switch (D) {
      case 4: rc = set_smem(tattn_fwd_kernel<4, false>, c.smem); break;
      case 8: rc = set_smem(tattn_fwd_kernel<8, false>, c.smem); break;
      case 16: rc = set_smem(tattn_fwd_kernel<16, false>, c.smem); break;
      case 32: rc = set_smem(tattn_fwd_kernel<32, false>, c.smem); break;
      default: rc = set_smem(tattn_fwd_kernel<64, false>, c.smem); break;
    }
  }
  if (rc) return rc;
  const bool shared_bias = bias_bstride == 0 && (bias == nullptr) == (bias_t == nullptr);
  if (shared_bias && tagan_tattn_fwd_mma_launch(D, (int)(B < 148 * 4 ? B : 148 * 4), st, Q, K, V, ld, B, T, heads, rsb, rst,
                                                bias, ms, ctx, lse, attn))
    return tagan_launch_status();
  const bool fast = T <= 32 && c.warps * c.PPW >= heads && shared_bias && c.smem <= 48 * 1024;
  if (fast && tagan_tattn_fwd_fast_launch(D, c.TP, grid, c.warps * 32, c.smem, st, Q, K, V, ld, B, T, heads, rsb, rst, bias, ms,
                                          ctx, lse, attn))
    return tagan_launch_status();
  DISPATCH_D(tattn_fwd_kernel, c.rbw > 1, <<<grid, c.warps * c.rbw * 32, c.smem, st>>>(Q, K, V, ld, B, T, heads, rsb, rst, bias_t, bias_bstride, ms, ctx, lse, attn, c.TP, c.warps, c.rbw))
  return tagan_launch_status();
}

TAGAN_API int tagan_tattn_fwd(const float* Q, const float* K, const float* V, int64_t ld, int64_t B, int32_t T,
                              int32_t H, int32_t heads, int32_t time_major, const float* bias, const float* bias_t,
                              int64_t bias_bstride, const float* ts,
                              int32_t mask_flags, float band, const int32_t* allones_flag, const uint8_t* mask,
                              int32_t mask_b, int32_t mask_h, float* ctx, float* lse, float* attn,
                              tagan_stream_t stream) {
  return tattn_fwd_impl(Q, K, V, ld, B, T, H, heads, time_major ? 1 : T, time_major ? B : 1, bias, bias_t, bias_bstride, ts,
                        mask_flags, band, allones_flag, mask, mask_b, mask_h, ctx, lse, attn, stream);
}

TAGAN_API int tagan_tattn_fwd_strided(const float* Q, const float* K, const float* V, int64_t ld, int64_t B, int32_t T,
                                      int32_t H, int32_t heads, int64_t row_stride_b, int64_t row_stride_t, const float* bias,
                                      const float* bias_t, int64_t bias_bstride, const float* ts, int32_t mask_flags, float band,
                                      const int32_t* allones_flag, const uint8_t* mask, int32_t mask_b, int32_t mask_h,
                                      float* ctx, float* lse, float* attn, tagan_stream_t stream) {
  if (row_stride_b <= 0 || row_stride_t <= 0) return TAGAN_E_INVALID;
  return tattn_fwd_impl(Q, K, V, ld, B, T, H, heads, row_stride_b, row_stride_t, bias, bias_t, bias_bstride, ts, mask_flags,
                        band, allones_flag, mask, mask_b, mask_h, ctx, lse, attn, stream);
}

TAGAN_API size_t tagan_tattn_bwd_workspace_bytes(int64_t B, int32_t T, int32_t heads) {
  if (B < 0 || T <= 0 || heads <= 0) return 0;
  return (size_t)bwd_grid(B) * heads * (size_t)T * T * sizeof(float);
}

static int tattn_bwd_impl(const float* Q, const float* K, const float* V, int64_t ld, int64_t B, int32_t T,
                          int32_t H, int32_t heads, int64_t rsb, int64_t rst, const float* bias, const float* bias_t, int64_t bias_bstride,
                              const float* ts, int32_t mask_flags, float band, const int32_t* allones_flag,
                              const uint8_t* mask, int32_t mask_b, int32_t mask_h, const float* ctx, const float* lse,
                              const float* dctx, float* dQ, float* dK, float* dV, int64_t ldd, float* dbias,
                              void* workspace, size_t workspace_bytes, tagan_stream_t stream) {
  if (!Q || !K || !V || !ctx || !lse || !dctx || !dQ || !dK || !dV || B < 0 || ld < H || ldd < H) return TAGAN_E_INVALID;
  if ((bias == nullptr) != (bias_t == nullptr)) return TAGAN_E_INVALID;
  if ((mask_flags & 2) && !ts) return TAGAN_E_INVALID;
  if ((mask_flags & 4) && !allones_flag) return TAGAN_E_INVALID;
  int D;
  if (!check_shape(T, H, heads, &D) || (ld % 4) || (ldd % 4)) return TAGAN_E_UNSUPPORTED;
  if (B == 0) {
    if (dbias && bias_bstride == 0) cudaMemsetAsync(dbias, 0, sizeof(float) * heads * T * T, as_stream(stream));
    return 0;
  }
  Cfg c;
  if (!make_cfg(T, D, heads, 4 * T * D + 3 * ((T + 3) & ~3), &c)) return TAGAN_E_UNSUPPORTED;
  const int grid = bwd_grid(B);
  const bool per_node = bias_bstride != 0;
  float* db_target = nullptr;
  if (dbias) {
    if (per_node) db_target = dbias;
    else {
      if (!workspace || workspace_bytes < tagan_tattn_bwd_workspace_bytes(B, T, heads)) return TAGAN_E_WORKSPACE;
      db_target = static_cast<float*>(workspace);
    }
  }
  MaskSpec ms{ts, mask_flags, band, allones_flag, mask, mask_b, mask_h};
  cudaStream_t st = as_stream(stream);
  int rc = 0;
  if (c.rbw > 1) {
    switch (D) {
      case 4: rc = set_smem(tattn_bwd_kernel<4, true>, c.smem); break;
      case 8: rc = set_smem(tattn_bwd_kernel<8, true>, c.smem); break;
      case 16: rc = set_smem(tattn_bwd_kernel<16, true>, c.smem); break;
      default: rc = set_smem(tattn_bwd_kernel<32, true>, c.smem); break;
    }
  } else {
    switch (D) {
      case 4: rc = set_smem(tattn_bwd_kernel<4, false>, c.smem); break;
      case 8: rc = set_smem(tattn_bwd_kernel<8, false>, c.smem); break;
      case 16: rc = set_smem(tattn_bwd_kernel<16, false>, c.smem); break;
      case 32: rc = set_smem(tattn_bwd_kernel<32, false>, c.smem); break;
      default: rc = set_smem(tattn_bwd_kernel<64, false>, c.smem); break;
    }
  }
  if (rc) return rc;
  Cfg cf;                                                // the fast kernel parks P and dS in smem: larger slots
  const bool fast = T <= 32 && !per_node && make_cfg(T, D, heads, tattn_bwd_fast_slot_floats(T, D, c.TP), &cf) &&
                    cf.warps * cf.PPW >= heads && cf.smem <= 100 * 1024;
  bool transposed_parts = false;
  if (fast && tagan_tattn_bwd_fast_launch(D, cf.TP, grid, cf.warps * 32, cf.smem, st, Q, K, V, ld, B, T, heads, rsb, rst, bias, ms,
                                          ctx, lse, dctx, dQ, dK, dV, ldd, db_target)) {
  } else {
    transposed_parts = true;
    DISPATCH_D(tattn_bwd_kernel, c.rbw > 1, <<<grid, c.warps * c.rbw * 32, c.smem, st>>>(Q, K, V, ld, B, T, heads, rsb, rst, bias, bias_t, bias_bstride, ms, ctx, lse, dctx, dQ, dK, dV, ldd, db_target, per_node ? 1 : 0, c.TP, c.warps, c.rbw))
  }
  if (dbias && !per_node) {
    const int64_t n = (int64_t)heads * T * T;
    if (transposed_parts) reduce_parts_transposed<<<ceil_div_i64(n, 256), 256, 0, st>>>(db_target, grid, heads, T, dbias);
    else reduce_parts<<<ceil_div_i64(n, 256), 256, 0, st>>>(db_target, grid, n, dbias);
  }
  return tagan_launch_status();
}

TAGAN_API int tagan_tattn_bwd(const float* Q, const float* K, const float* V, int64_t ld, int64_t B, int32_t T,
                              int32_t H, int32_t heads, int32_t time_major, const float* bias, const float* bias_t, int64_t bias_bstride,
                              const float* ts, int32_t mask_flags, float band, const int32_t* allones_flag,
                              const uint8_t* mask, int32_t mask_b, int32_t mask_h, const float* ctx, const float* lse,
                              const float* dctx, float* dQ, float* dK, float* dV, int64_t ldd, float* dbias,
                              void* workspace, size_t workspace_bytes, tagan_stream_t stream) {
  return tattn_bwd_impl(Q, K, V, ld, B, T, H, heads, time_major ? 1 : T, time_major ? B : 1, bias, bias_t, bias_bstride, ts,
                        mask_flags, band, allones_flag, mask, mask_b, mask_h, ctx, lse, dctx, dQ, dK, dV, ldd, dbias, workspace,
                        workspace_bytes, stream);
}

TAGAN_API int tagan_tattn_bwd_strided(const float* Q, const float* K, const float* V, int64_t ld, int64_t B, int32_t T,
                                      int32_t H, int32_t heads, int64_t row_stride_b, int64_t row_stride_t, const float* bias,
                                      const float* bias_t, int64_t bias_bstride, const float* ts, int32_t mask_flags, float band,
                                      const int32_t* allones_flag, const uint8_t* mask, int32_t mask_b, int32_t mask_h,
                                      const float* ctx, const float* lse, const float* dctx, float* dQ, float* dK, float* dV,
                                      int64_t ldd, float* dbias, void* workspace, size_t workspace_bytes, tagan_stream_t stream) {
  if (row_stride_b <= 0 || row_stride_t <= 0) return TAGAN_E_INVALID;
  return tattn_bwd_impl(Q, K, V, ld, B, T, H, heads, row_stride_b, row_stride_t, bias, bias_t, bias_bstride, ts, mask_flags, band,
                        allones_flag, mask, mask_b, mask_h, ctx, lse, dctx, dQ, dK, dV, ldd, dbias, workspace, workspace_bytes, stream);
}

TAGAN_API int tagan_tattn_mask_allones(const float* ts, int64_t B, int32_t T, float band, const uint8_t* mask,
                                       int64_t mask_elems, int32_t* allones_flag, tagan_stream_t stream) {
  if (!allones_flag || B < 0 || T < 0 || mask_elems < 0) return TAGAN_E_INVALID;
  cudaStream_t st = as_stream(stream);
  set_flag<<<1, 1, 0, st>>>(allones_flag, 1);
  int64_t work = (ts ? B * T : 0) > mask_elems ? (ts ? B * T : 0) : mask_elems;
  if (work > 0) {
    int grid = ceil_div_i64(work, 256);
    if (grid > 148 * 8) grid = 148 * 8;
    mask_allones_kernel<<<grid, 256, 0, st>>>(ts, B, T, band, mask, mask ? mask_elems : 0, allones_flag);
  }
  return tagan_launch_status();
}
