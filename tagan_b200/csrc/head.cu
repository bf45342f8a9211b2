// What follows the hot path in TAGAN.forward / the trainer step: everything here
// works on tiny tensors ([1,T,H] pooled features, ~0.5-2 M parameters), i.e. it is launch latency, not bandwidth --
// so each stage is ONE launch and the whole step stays CUDA-graph capturable (no host sync, no allocation).
//
//   tagan_pack_padded_fwd/bwd   packed rows [sum N_t, H] + offsets -> zero-padded [T, maxN, H]
//                               (AsymmetricTemporalAttention.forward's pad + stack, temporal_attention.py:928-976)
//   tagan_pool_blocks_fwd/bwd   graph_features[t] = mean of the t-th block of maxN consecutive rows of the [maxN*T, H]
//                               matrix the temporal attention returns (TAGAN.forward model.py:377-427: both of its
//                               branches -- `x[t].mean(0)` when maxN == T and `x.view(T,-1,H)[t].mean(0)` otherwise --
//                               are this block mean; the "view, not permute" scrambling of SURVEY.md section 3.1 is kept)
//   tagan_head_fwd/bwd          TemporalClassificationHead (classification.py:743-975; attention pooling over T, Linear
//                               -> LayerNorm -> ReLU -> Linear) + TemporalLossFunction 'classification' (BCE with logits,
//                               mean, :453-456) or nn.CrossEntropyLoss (model.py:436-438), single CTA
//   tagan_adam_clip_step        clip_grad_norm_ + Adam (trainer.py:295-311) on flat parameter / gradient buffers, step
//                               counter and gradient norm on device
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------------
// pack / pad
// ---------------------------------------------------------------------------------------------------------------
__global__ void pack_padded_kernel(const float* __restrict__ packed, const int32_t* __restrict__ off, float* __restrict__ out,
                                   int maxn, int H4, int to_padded) {
  const int t = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)maxn * H4) return;
  const int row = (int)(i / H4);
  const int c = (int)(i - (int64_t)row * H4);
  const int n_t = off[t + 1] - off[t];
  float4* o4 = reinterpret_cast<float4*>(out);
  const float4* p4 = reinterpret_cast<const float4*>(packed);
  if (to_padded) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < n_t) v = p4[(int64_t)(off[t] + row) * H4 + c];
    o4[((int64_t)t * maxn + row) * H4 + c] = v;
  } else if (row < n_t) {        // backward: `out` is d packed, `packed` is d padded
    o4[(int64_t)(off[t] + row) * H4 + c] = p4[((int64_t)t * maxn + row) * H4 + c];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// block-mean pooling
// ---------------------------------------------------------------------------------------------------------------
constexpr int POOL_PARTS = 64;
// logical row r = b*T + s of x[B,T,H]; physical row = s*B + b when the storage is time-major [T,B,H]
__device__ __forceinline__ int64_t phys_row(int64_t r, int64_t B, int T, int time_major) {
  if (!time_major) return r;
  const int64_t b = r / T;
  const int s = (int)(r - b * T);
  return (int64_t)s * B + b;
}
__global__ void __launch_bounds__(256)
pool_partial_kernel(const float* __restrict__ x, int64_t B, int T, int H4, int time_major, float* __restrict__ partial) {
  __shared__ float4 red[256];
  const int t = blockIdx.x, part = blockIdx.y;
  const int slots = 256 / H4;                           // row slots per iteration (H4 <= 256)
  const int slot = threadIdx.x / H4, c = threadIdx.x - slot * H4;
  const int64_t per = (B + POOL_PARTS - 1) / POOL_PARTS;
  const int64_t r0 = (int64_t)t * B + part * per, r1 = min((int64_t)t * B + B, r0 + per);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (slot < slots)
    for (int64_t r = r0 + slot; r < r1; r += slots) {
      const float4 v = reinterpret_cast<const float4*>(x)[phys_row(r, B, T, time_major) * H4 + c];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  red[threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x < H4) {
    float4 s = red[threadIdx.x];
    for (int k = 1; k < slots; ++k) {
      const float4 b4 = red[k * H4 + threadIdx.x];
      s.x += b4.x; s.y += b4.y; s.z += b4.z; s.w += b4.w;
    }
    reinterpret_cast<float4*>(partial)[((int64_t)t * POOL_PARTS + part) * H4 + threadIdx.x] = s;
  }
}
__global__ void pool_final_kernel(const float* __restrict__ partial, int T, int H, float inv, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * H) return;
  const int t = i / H, c = i - t * H;
  float s = 0.f;
  for (int p = 0; p < POOL_PARTS; ++p) s += partial[((int64_t)t * POOL_PARTS + p) * H + c];
  out[i] = s * inv;
}
__global__ void pool_bwd_kernel(const float* __restrict__ dgf, int64_t B, int T, int H4, int time_major, float inv,
                                float* __restrict__ dx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * T * H4) return;
  const int64_t r = i / H4;                              // logical row
  const int c = (int)(i - r * H4);
  const int t = (int)(r / B);
  float4 g = reinterpret_cast<const float4*>(dgf)[(int64_t)t * H4 + c];
  g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
  reinterpret_cast<float4*>(dx)[phys_row(r, B, T, time_major) * H4 + c] = g;
}

// ---------------------------------------------------------------------------------------------------------------
// classification head, one CTA
// ---------------------------------------------------------------------------------------------------------------
constexpr int HT = 512;                                  // threads of the head kernels

struct HeadParams {
  const float *wa1, *ba1, *wa2, *w1, *b1, *lng, *lnb, *w2, *b2;
  int Bsz, T, H, O, use_ln, loss_type, label_rows;      // loss_type 0 = BCE with logits (mean), 1 = cross entropy (mean)
};

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int k = 0; k < HT / 32; ++k) s += red[k];
  __syncthreads();
  return s;
}

__global__ void __launch_bounds__(HT)
head_fwd_kernel(HeadParams p, const float* __restrict__ gf, const float* __restrict__ labels, const int64_t* __restrict__ cls,
                float* __restrict__ u, float* __restrict__ alpha, float* __restrict__ pooled, float* __restrict__ h1,
                float* __restrict__ hn, float* __restrict__ stats, float* __restrict__ logits, float* __restrict__ loss) {
  __shared__ float red[HT / 32];
  const int T = p.T, H = p.H, O = p.O;
  float loss_acc = 0.f;
  for (int b = 0; b < p.Bsz; ++b) {
    const float* g = gf + (int64_t)b * T * H;
    float* ub = u + (int64_t)b * T * H;
    for (int i = threadIdx.x; i < T * H; i += HT) {      // attention[0] + Tanh
      const int t = i / H, j = i - t * H;
      float s = p.ba1[j];
      for (int c = 0; c < H; ++c) s = fmaf(p.wa1[(int64_t)j * H + c], g[t * H + c], s);
      ub[i] = tanhf(s);
    }
    __syncthreads();
    float* ab = alpha + b * T;
    for (int t = threadIdx.x; t < T; t += HT) {          // attention[2] (no bias)
      float s = 0.f;
      for (int j = 0; j < H; ++j) s = fmaf(p.wa2[j], ub[t * H + j], s);
      ab[t] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {                               // softmax over the T snapshots
      float m = -INFINITY, l = 0.f;
      for (int t = 0; t < T; ++t) m = fmaxf(m, ab[t]);
      for (int t = 0; t < T; ++t) { ab[t] = expf(ab[t] - m); l += ab[t]; }
      for (int t = 0; t < T; ++t) ab[t] /= l;
    }
    __syncthreads();
    float* pb = pooled + b * H;
    for (int c = threadIdx.x; c < H; c += HT) {
      float s = 0.f;
      for (int t = 0; t < T; ++t) s = fmaf(ab[t], g[t * H + c], s);
      pb[c] = s;
    }
    __syncthreads();
    float* h1b = h1 + b * H;
    for (int j = threadIdx.x; j < H; j += HT) {          // classifier[0]
      float s = p.b1[j];
      for (int c = 0; c < H; ++c) s = fmaf(p.w1[(int64_t)j * H + c], pb[c], s);
      h1b[j] = s;
    }
    __syncthreads();
    float mean = 0.f, rstd = 1.f;
    if (p.use_ln) {
      float s = 0.f;
      for (int j = threadIdx.x; j < H; j += HT) s += h1b[j];
      mean = block_sum(s, red) / (float)H;
      float q = 0.f;
      for (int j = threadIdx.x; j < H; j += HT) { const float d = h1b[j] - mean; q = fmaf(d, d, q); }
      rstd = 1.f / sqrtf(block_sum(q, red) / (float)H + 1e-5f);
    }
    if (threadIdx.x == 0) { stats[2 * b] = mean; stats[2 * b + 1] = rstd; }
    float* hnb = hn + b * H;
    for (int j = threadIdx.x; j < H; j += HT)
      hnb[j] = p.use_ln ? (h1b[j] - mean) * rstd * p.lng[j] + p.lnb[j] : h1b[j];
    __syncthreads();
    for (int o = threadIdx.x; o < O; o += HT) {          // ReLU + classifier[-1]
      float s = p.b2[o];
      for (int j = 0; j < H; ++j) s = fmaf(p.w2[(int64_t)o * H + j], fmaxf(hnb[j], 0.f), s);
      logits[b * O + o] = s;
    }
    __syncthreads();
  }
  if (loss == nullptr) return;
  if (threadIdx.x == 0) {
    if (p.loss_type == 0) {                               // F.binary_cross_entropy_with_logits(..., reduction='none').mean()
      const int rows = p.label_rows;
      for (int r = 0; r < rows; ++r)
        for (int o = 0; o < O; ++o) {
          const float x = logits[(p.Bsz == rows ? r : 0) * O + o], y = labels[r * O + o];
          loss_acc += fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));
        }
      *loss = loss_acc / (float)(rows * O);
    } else {                                              // nn.CrossEntropyLoss()(logits, class indices)
      for (int b = 0; b < p.Bsz; ++b) {
        float m = -INFINITY, l = 0.f;
        for (int o = 0; o < O; ++o) m = fmaxf(m, logits[b * O + o]);
        for (int o = 0; o < O; ++o) l += expf(logits[b * O + o] - m);
        loss_acc += m + logf(l) - logits[b * O + (int)cls[b]];
      }
      *loss = loss_acc / (float)p.Bsz;
    }
  }
}

// gradients of every head parameter and of gf; dlogits_ext (optional) adds an external gradient on the logits
__global__ void __launch_bounds__(HT)
head_bwd_kernel(HeadParams p, const float* __restrict__ gf, const float* __restrict__ labels, const int64_t* __restrict__ cls,
                const float* __restrict__ u, const float* __restrict__ alpha, const float* __restrict__ pooled,
                const float* __restrict__ h1, const float* __restrict__ hn, const float* __restrict__ stats,
                const float* __restrict__ logits, const float* __restrict__ dloss, const float* __restrict__ dlogits_ext,
                float* __restrict__ scratch /* Bsz*(O + 3H + 2T) + Bsz*T*H */, float* __restrict__ dgf, float* __restrict__ dwa1,
                float* __restrict__ dba1, float* __restrict__ dwa2, float* __restrict__ dw1, float* __restrict__ db1,
                float* __restrict__ dlng, float* __restrict__ dlnb, float* __restrict__ dw2, float* __restrict__ db2) {
  __shared__ float red[HT / 32];
  const int T = p.T, H = p.H, O = p.O, Bsz = p.Bsz;
  float* dlog = scratch;                                  // [Bsz,O]
  float* dh1 = dlog + Bsz * O;                            // [Bsz,H]
  float* dhn = dh1 + Bsz * H;                             // [Bsz,H]
  float* dpool = dhn + Bsz * H;                           // [Bsz,H]
  float* dsc = dpool + Bsz * H;                           // [Bsz,T]
  float* dapre = dsc + Bsz * T + Bsz * T;                 // [Bsz,T,H]
  const float gl = dloss ? *dloss : 0.f;
  for (int i = threadIdx.x; i < Bsz * O; i += HT) {
    const int b = i / O, o = i - b * O;
    float d = dlogits_ext ? dlogits_ext[i] : 0.f;
    if (dloss) {
      if (p.loss_type == 0) {
        const int rows = p.label_rows;
        if (Bsz == rows) d += gl * (1.f / (1.f + expf(-logits[i])) - labels[i]) / (float)(rows * O);
        else if (b == 0)
          for (int r = 0; r < rows; ++r) d += gl * (1.f / (1.f + expf(-logits[o])) - labels[r * O + o]) / (float)(rows * O);
      } else {
        float m = -INFINITY, l = 0.f;
        for (int q = 0; q < O; ++q) m = fmaxf(m, logits[b * O + q]);
        for (int q = 0; q < O; ++q) l += expf(logits[b * O + q] - m);
        d += gl * (expf(logits[i] - m) / l - ((int)cls[b] == o ? 1.f : 0.f)) / (float)Bsz;
      }
    }
    dlog[i] = d;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < O * H; i += HT) {        // classifier[-1]
    const int o = i / H, j = i - o * H;
    float s = 0.f;
    for (int b = 0; b < Bsz; ++b) s = fmaf(dlog[b * O + o], fmaxf(hn[b * H + j], 0.f), s);
    dw2[i] = s;
  }
  for (int o = threadIdx.x; o < O; o += HT) {
    float s = 0.f;
    for (int b = 0; b < Bsz; ++b) s += dlog[b * O + o];
    db2[o] = s;
  }
  for (int i = threadIdx.x; i < Bsz * H; i += HT) {      // through ReLU
    const int b = i / H, j = i - b * H;
    float s = 0.f;
    for (int o = 0; o < O; ++o) s = fmaf(dlog[b * O + o], p.w2[(int64_t)o * H + j], s);
    dhn[i] = hn[i] > 0.f ? s : 0.f;
  }
  __syncthreads();
  for (int b = 0; b < Bsz; ++b) {                         // LayerNorm backward (per row)
    const float mean = stats[2 * b], rstd = stats[2 * b + 1];
    if (p.use_ln) {
      float s1 = 0.f, s2 = 0.f;
      for (int j = threadIdx.x; j < H; j += HT) {
        const float g = dhn[b * H + j] * p.lng[j], xh = (h1[b * H + j] - mean) * rstd;
        s1 += g;
        s2 = fmaf(g, xh, s2);
      }
      s1 = block_sum(s1, red) / (float)H;
      s2 = block_sum(s2, red) / (float)H;
      for (int j = threadIdx.x; j < H; j += HT) {
        const float xh = (h1[b * H + j] - mean) * rstd;
        dh1[b * H + j] = rstd * (dhn[b * H + j] * p.lng[j] - s1 - xh * s2);
      }
    } else {
      for (int j = threadIdx.x; j < H; j += HT) dh1[b * H + j] = dhn[b * H + j];
    }
  }
  __syncthreads();
  if (p.use_ln)
    for (int j = threadIdx.x; j < H; j += HT) {
      float a = 0.f, c = 0.f;
      for (int b = 0; b < Bsz; ++b) {
        a = fmaf(dhn[b * H + j], (h1[b * H + j] - stats[2 * b]) * stats[2 * b + 1], a);
        c += dhn[b * H + j];
      }
      dlng[j] = a;
      dlnb[j] = c;
    }
  for (int i = threadIdx.x; i < H * H; i += HT) {        // classifier[0]
    const int j = i / H, c = i - j * H;
    float s = 0.f;
    for (int b = 0; b < Bsz; ++b) s = fmaf(dh1[b * H + j], pooled[b * H + c], s);
    dw1[i] = s;
  }
  for (int j = threadIdx.x; j < H; j += HT) {
    float s = 0.f;
    for (int b = 0; b < Bsz; ++b) s += dh1[b * H + j];
    db1[j] = s;
  }
  for (int i = threadIdx.x; i < Bsz * H; i += HT) {
    const int b = i / H, c = i - b * H;
    float s = 0.f;
    for (int j = 0; j < H; ++j) s = fmaf(dh1[b * H + j], p.w1[(int64_t)j * H + c], s);
    dpool[i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Bsz * T; i += HT) {      // d alpha[t] = dpooled . gf[t]
    const int b = i / T, t = i - b * T;
    float s = 0.f;
    for (int c = 0; c < H; ++c) s = fmaf(dpool[b * H + c], gf[((int64_t)b * T + t) * H + c], s);
    dsc[Bsz * T + i] = s;                                 // second half of the buffer: d alpha
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Bsz * T; i += HT) {      // softmax backward
    const int b = i / T;
    float dot = 0.f;
    for (int t = 0; t < T; ++t) dot = fmaf(alpha[b * T + t], dsc[Bsz * T + b * T + t], dot);
    dsc[i] = alpha[i] * (dsc[Bsz * T + i] - dot);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < H; j += HT) {            // attention[2]
    float s = 0.f;
    for (int i = 0; i < Bsz * T; ++i) s = fmaf(dsc[i], u[(int64_t)i * H + j], s);
    dwa2[j] = s;
  }
  for (int64_t i = threadIdx.x; i < (int64_t)Bsz * T * H; i += HT) {   // through Tanh
    const int j = (int)(i % H);
    const float uv = u[i];
    dapre[i] = dsc[i / H] * p.wa2[j] * (1.f - uv * uv);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * H; i += HT) {        // attention[0]
    const int j = i / H, c = i - j * H;
    float s = 0.f;
    for (int r = 0; r < Bsz * T; ++r) s = fmaf(dapre[(int64_t)r * H + j], gf[(int64_t)r * H + c], s);
    dwa1[i] = s;
  }
  for (int j = threadIdx.x; j < H; j += HT) {
    float s = 0.f;
    for (int r = 0; r < Bsz * T; ++r) s += dapre[(int64_t)r * H + j];
    dba1[j] = s;
  }
  for (int64_t i = threadIdx.x; i < (int64_t)Bsz * T * H; i += HT) {   // d gf = alpha * dpooled + Wa1^T d a_pre
    const int64_t r = i / H;
    const int c = (int)(i - r * H), b = (int)(r / T);
    float s = alpha[r] * dpool[b * H + c];
    for (int j = 0; j < H; ++j) s = fmaf(dapre[r * H + j], p.wa1[(int64_t)j * H + c], s);
    dgf[i] = s;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// clip_grad_norm_ + Adam
// ---------------------------------------------------------------------------------------------------------------
__global__ void step_inc_kernel(int32_t* step) { *step += 1; }

__global__ void adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                 int64_t n, float lr, float b1, float b2, float eps, float wd, float max_norm,
                                 const float* __restrict__ grad_msq, const int32_t* __restrict__ step) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float clip = 1.f;
  if (max_norm > 0.f && grad_msq != nullptr) {            // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
    const float norm = sqrtf(*grad_msq * (float)n);
    clip = fminf(1.f, max_norm / (norm + 1e-6f));
  }
  const float t = (float)(*step);
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  float gi = g[i] * clip;
  const float pi = p[i];
  if (wd != 0.f) gi = fmaf(wd, pi, gi);                    // torch.optim.Adam: L2 penalty added to the gradient
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
  p[i] = pi - (lr / bc1) * (mi / denom);
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

TAGAN_API int tagan_pack_padded_fwd(const float* packed, const int32_t* offsets, float* padded, int32_t T, int32_t maxn,
                                    int32_t H, tagan_stream_t stream) {
  if (!packed || !offsets || !padded || T < 0 || maxn < 0 || H <= 0) return TAGAN_E_INVALID;
  if (H % 4 || !al16(packed) || !al16(padded)) return TAGAN_E_UNSUPPORTED;
  if (T == 0 || maxn == 0) return 0;
  dim3 grid(ceil_div_i64((int64_t)maxn * (H / 4), 256), T);
  pack_padded_kernel<<<grid, 256, 0, as_stream(stream)>>>(packed, offsets, padded, maxn, H / 4, 1);
  return tagan_launch_status();
}

TAGAN_API int tagan_pack_padded_bwd(const float* dpadded, const int32_t* offsets, float* dpacked, int32_t T, int32_t maxn,
                                    int32_t H, tagan_stream_t stream) {
  if (!dpadded || !offsets || !dpacked || T < 0 || maxn < 0 || H <= 0) return TAGAN_E_INVALID;
  if (H % 4 || !al16(dpadded) || !al16(dpacked)) return TAGAN_E_UNSUPPORTED;
  if (T == 0 || maxn == 0) return 0;
  dim3 grid(ceil_div_i64((int64_t)maxn * (H / 4), 256), T);
  pack_padded_kernel<<<grid, 256, 0, as_stream(stream)>>>(dpadded, offsets, dpacked, maxn, H / 4, 0);
  return tagan_launch_status();
}

TAGAN_API size_t tagan_pool_blocks_workspace_bytes(int32_t T, int32_t H) {
  return (size_t)(T > 0 ? T : 0) * POOL_PARTS * (size_t)(H > 0 ? H : 0) * sizeof(float);
}

TAGAN_API int tagan_pool_blocks_fwd(const float* x, int64_t B, int32_t T, int32_t H, int32_t time_major, float* out,
                                    void* workspace, size_t workspace_bytes, tagan_stream_t stream) {
  if (!x || !out || B <= 0 || T <= 0 || H <= 0) return TAGAN_E_INVALID;
  if (H % 4 || H > 1024 || !al16(x)) return TAGAN_E_UNSUPPORTED;
  if (!workspace || workspace_bytes < tagan_pool_blocks_workspace_bytes(T, H)) return TAGAN_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  pool_partial_kernel<<<dim3(T, POOL_PARTS), 256, 0, st>>>(x, B, T, H / 4, time_major, static_cast<float*>(workspace));
  pool_final_kernel<<<ceil_div_i64((int64_t)T * H, 256), 256, 0, st>>>(static_cast<const float*>(workspace), T, H, 1.f / (float)B, out);
  return tagan_launch_status();
}

TAGAN_API int tagan_pool_blocks_bwd(const float* dout, int64_t B, int32_t T, int32_t H, int32_t time_major, float* dx,
                                    tagan_stream_t stream) {
  if (!dout || !dx || B <= 0 || T <= 0 || H <= 0) return TAGAN_E_INVALID;
  if (H % 4 || !al16(dout) || !al16(dx)) return TAGAN_E_UNSUPPORTED;
  pool_bwd_kernel<<<ceil_div_i64(B * T * (H / 4), 256), 256, 0, as_stream(stream)>>>(dout, B, T, H / 4, time_major, 1.f / (float)B, dx);
  return tagan_launch_status();
}

static HeadParams head_params(const struct tagan_head_weights* w, int32_t Bsz, int32_t T, int32_t H, int32_t O, int32_t loss_type,
                              int32_t label_rows) {
  HeadParams p;
  p.wa1 = w->attn0_weight; p.ba1 = w->attn0_bias; p.wa2 = w->attn2_weight; p.w1 = w->fc0_weight; p.b1 = w->fc0_bias;
  p.lng = w->ln_weight; p.lnb = w->ln_bias; p.w2 = w->fc1_weight; p.b2 = w->fc1_bias;
  p.Bsz = Bsz; p.T = T; p.H = H; p.O = O; p.use_ln = w->ln_weight != nullptr; p.loss_type = loss_type; p.label_rows = label_rows;
  return p;
}

TAGAN_API int tagan_head_fwd(const struct tagan_head_weights* w, const float* gf, int32_t Bsz, int32_t T, int32_t H, int32_t O,
                             int32_t loss_type, const float* labels, int32_t label_rows, const int64_t* class_index,
                             float* u, float* alpha, float* pooled, float* h1, float* hn, float* stats, float* logits,
                             float* loss, tagan_stream_t stream) {
  if (!w || !gf || !u || !alpha || !pooled || !h1 || !hn || !stats || !logits || Bsz <= 0 || T <= 0 || H <= 0 || O <= 0)
    return TAGAN_E_INVALID;
  if (!w->attn0_weight || !w->attn0_bias || !w->attn2_weight || !w->fc0_weight || !w->fc0_bias || !w->fc1_weight || !w->fc1_bias)
    return TAGAN_E_INVALID;
  if (loss && ((loss_type == 0 && (!labels || (label_rows != Bsz && Bsz != 1))) || (loss_type == 1 && !class_index) ||
               loss_type < 0 || loss_type > 1))
    return TAGAN_E_INVALID;
  head_fwd_kernel<<<1, HT, 0, as_stream(stream)>>>(head_params(w, Bsz, T, H, O, loss_type, label_rows), gf, labels, class_index, u,
                                                   alpha, pooled, h1, hn, stats, logits, loss);
  return tagan_launch_status();
}

TAGAN_API size_t tagan_head_bwd_workspace_bytes(int32_t Bsz, int32_t T, int32_t H, int32_t O) {
  if (Bsz <= 0 || T <= 0 || H <= 0 || O <= 0) return 0;
  return ((size_t)Bsz * (O + 3 * (size_t)H + 2 * (size_t)T) + (size_t)Bsz * T * H) * sizeof(float);
}

TAGAN_API int tagan_head_bwd(const struct tagan_head_weights* w, const float* gf, int32_t Bsz, int32_t T, int32_t H, int32_t O,
                             int32_t loss_type, const float* labels, int32_t label_rows, const int64_t* class_index,
                             const float* u, const float* alpha, const float* pooled, const float* h1, const float* hn,
                             const float* stats, const float* logits, const float* dloss, const float* dlogits,
                             float* dgf, struct tagan_head_weights* dw, void* workspace, size_t workspace_bytes,
                             tagan_stream_t stream) {
  if (!w || !dw || !gf || !u || !alpha || !pooled || !h1 || !hn || !stats || !logits || !dgf || Bsz <= 0 || T <= 0 || H <= 0 || O <= 0)
    return TAGAN_E_INVALID;
  if (!workspace || workspace_bytes < tagan_head_bwd_workspace_bytes(Bsz, T, H, O)) return TAGAN_E_WORKSPACE;
  head_bwd_kernel<<<1, HT, 0, as_stream(stream)>>>(
      head_params(w, Bsz, T, H, O, loss_type, label_rows), gf, labels, class_index, u, alpha, pooled, h1, hn, stats, logits, dloss,
      dlogits, static_cast<float*>(workspace), dgf, const_cast<float*>(dw->attn0_weight), const_cast<float*>(dw->attn0_bias),
      const_cast<float*>(dw->attn2_weight), const_cast<float*>(dw->fc0_weight), const_cast<float*>(dw->fc0_bias),
      const_cast<float*>(dw->ln_weight), const_cast<float*>(dw->ln_bias), const_cast<float*>(dw->fc1_weight),
      const_cast<float*>(dw->fc1_bias));
  return tagan_launch_status();
}

TAGAN_API int tagan_adam_clip_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, float max_grad_norm,
                                   const float* grad_mean_sq, int32_t* step, tagan_stream_t stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step || n < 0) return TAGAN_E_INVALID;
  if (max_grad_norm > 0.f && !grad_mean_sq) return TAGAN_E_INVALID;
  if (n == 0) return 0;
  cudaStream_t st = as_stream(stream);
  step_inc_kernel<<<1, 1, 0, st>>>(step);
  adam_clip_kernel<<<ceil_div_i64(n, 256), 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                       max_grad_norm, grad_mean_sq, step);
  return tagan_launch_status();
}
