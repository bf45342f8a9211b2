// (a1) Device CSR build: edge_index[2,E] (int64) -> unique(edges) U self-loops as CSR, plus the
// transposed CSR used by the deterministic backward.  Replaces the dense N x N mask of
// TAGANGraphAttention.forward (reference src/tagan/layers/graph_attention.py:98-102).
//
// Integer, HBM/atomic-bound work; every output is a pure function of the input *set* (atomics
// only decide the order inside a bucket, and each bucket is then sorted), so results are
// bit-exact and deterministic:
//   count rows (atomicAdd) -> scan -> scatter into row buckets -> per-row rank sort + dedup
//   -> scan unique counts = rowptr -> compact -> [count cols -> scan -> scatter -> per-col sort].
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  // exclusive scan of one value per thread across a SCAN_THREADS block
  __shared__ int warp_tot[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(FULL_MASK, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    int t = lane < SCAN_THREADS / 32 ? warp_tot[lane] : 0;
    int ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(FULL_MASK, ti, o);
      if (lane >= o) ti += u;
    }
    if (lane < SCAN_THREADS / 32) warp_tot[lane] = ti - t;  // exclusive warp offsets
    if (lane == 31) *total = ti;
  }
  __syncthreads();
  int r = warp_tot[w] + inc - v;
  __syncthreads();
  return r;
}

__global__ void scan_block_sums(const int* __restrict__ in, int* __restrict__ bsum, int64_t n) {
  __shared__ int total;
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) s += (base + i < n) ? in[base + i] : 0;
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

// out[i] = boff[block] + exclusive prefix inside the block; the last block also writes out[n].
__global__ void scan_apply(const int* __restrict__ in, const int* __restrict__ boff, int* __restrict__ out, int64_t n) {
  __shared__ int total;
  int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = (base + i < n) ? in[base + i] : 0; s += v[i]; }
  int off = block_exclusive_scan(s, &total) + (boff ? boff[blockIdx.x] : 0);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = off;
    off += v[i];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) out[n] = off;
}

int64_t scan_ws_ints(int64_t n) {
  int64_t tot = 0;
  while (n > SCAN_TILE) {
    int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    tot += 2 * (nb + 1);
    n = nb;
  }
  return tot + 2;
}

// exclusive scan of in[0..n) into out[0..n] (n+1 outputs).  in may alias out only if n <= SCAN_TILE... never aliased here.
void exclusive_scan(const int* in, int* out, int64_t n, int* ws, cudaStream_t st) {
  int64_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (nb <= 1) {
    scan_apply<<<1, SCAN_THREADS, 0, st>>>(in, nullptr, out, n);
    return;
  }
  int* bsum = ws;
  int* bscan = ws + (nb + 1);
  scan_block_sums<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, bsum, n);
  exclusive_scan(bsum, bscan, nb, ws + 2 * (nb + 1), st);
  scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(in, bscan, out, n);
}

__device__ __forceinline__ bool norm_index(int64_t v, int n, int* out) {
  if (v < 0) v += n;  // torch advanced indexing wraps negatives
  if (v < 0 || v >= n) return false;
  *out = (int)v;
  return true;
}

__global__ void fill_i32(int* p, int v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// rows outside [row_begin, row_begin + nrows) belong to another partition and are skipped silently
__global__ void count_rows(const int64_t* __restrict__ ei, int64_t E, int n, int row_begin, int nrows,
                           int* __restrict__ cnt, int* __restrict__ status) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int r, c;
  if (norm_index(ei[e], n, &r) && norm_index(ei[E + e], n, &c)) {
    r -= row_begin;
    if (r >= 0 && r < nrows) atomicAdd(&cnt[r], 1);
  } else {
    *status = 1;
  }
}

// bucket layout: position off[r] holds the self loop, the rest is filled through cursor[r] (starts at 1)
__global__ void place_self_loops(const int* __restrict__ off, int* __restrict__ bucket, int n, int row_begin) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) bucket[off[i]] = row_begin + i;
}

__global__ void scatter_edges(const int64_t* __restrict__ ei, int64_t E, int n, int row_begin, int nrows,
                              const int* __restrict__ off, int* __restrict__ cursor, int* __restrict__ bucket) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int r, c;
  if (norm_index(ei[e], n, &r) && norm_index(ei[E + e], n, &c)) {
    r -= row_begin;
    if (r >= 0 && r < nrows) {
      int p = atomicAdd(&cursor[r], 1);
      bucket[off[r] + p] = c;
    }
  }
}

// ---- batched (block-diagonal) form: T snapshots as ONE graph whose node ids are noff[t] + local id.  Only the two kernels
// that read edge_index differ; everything after the bucket scatter is the single-graph pipeline on sum(N_t) rows.
constexpr int MAX_BATCH = 128;
struct Batch {
  int T;
  const int64_t* src[MAX_BATCH];
  const int64_t* dst[MAX_BATCH];
  int64_t eoff[MAX_BATCH + 1];     // edge e of the concatenation belongs to snapshot t with eoff[t] <= e < eoff[t+1]
  int noff[MAX_BATCH + 1];         // first global row of snapshot t
};
__device__ __forceinline__ int batch_find(const Batch& b, int64_t e) {
  int lo = 0, hi = b.T;            // invariant: eoff[lo] <= e < eoff[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (b.eoff[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}
__global__ void count_rows_batched(const __grid_constant__ Batch b, int64_t E, int* __restrict__ cnt, int* __restrict__ status) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int t = batch_find(b, e);
  const int64_t le = e - b.eoff[t];
  const int n = b.noff[t + 1] - b.noff[t];
  int r, c;
  if (norm_index(b.src[t][le], n, &r) && norm_index(b.dst[t][le], n, &c)) atomicAdd(&cnt[b.noff[t] + r], 1);
  else *status = 1;
}
__global__ void scatter_edges_batched(const __grid_constant__ Batch b, int64_t E, const int* __restrict__ off,
                                      int* __restrict__ cursor, int* __restrict__ bucket) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int t = batch_find(b, e);
  const int64_t le = e - b.eoff[t];
  const int n = b.noff[t + 1] - b.noff[t];
  int r, c;
  if (norm_index(b.src[t][le], n, &r) && norm_index(b.dst[t][le], n, &c)) {
    r += b.noff[t];
    const int p = atomicAdd(&cursor[r], 1);
    bucket[off[r] + p] = b.noff[t] + c;
  }
}

// ---- per-segment rank sort.  UNIQUE: keep the first of each run of equal keys and pack them
// at the segment start (count -> ucount[seg]); otherwise a stable full sort with a payload.
constexpr int SORT_SMEM_KEYS = 8192;

// Segments of up to WARP_SEG entries: one warp, each lane holds two keys (positions lane and lane+32);
// ranks are counted with register shuffles only.
constexpr int WARP_SEG = 64;

template <bool UNIQUE, bool PAYLOAD>
__global__ void seg_sort_warp(const int* __restrict__ off, const int* __restrict__ kin, const int* __restrict__ pin,
                              int* __restrict__ kout, int* __restrict__ pout, int* __restrict__ ucount, int nseg) {
  int seg = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  if (seg >= nseg) return;
  int b = off[seg], len = off[seg + 1] - b;
  if (len > WARP_SEG) return;  // handled by seg_sort_block
  const int i0 = lane, i1 = lane + 32;
  const int key0 = i0 < len ? kin[b + i0] : 0x7fffffff;
  const int key1 = i1 < len ? kin[b + i1] : 0x7fffffff;
  int pay0 = 0, pay1 = 0;
  if (PAYLOAD) { pay0 = i0 < len ? pin[b + i0] : 0; pay1 = i1 < len ? pin[b + i1] : 0; }
  const int lo_n = len < 32 ? len : 32, hi_n = len - lo_n;   // valid entries in the two halves
  if (UNIQUE) {
    bool first0 = i0 < len, first1 = i1 < len;
    for (int j = 0; j < lo_n; ++j) {
      const int kj = __shfl_sync(FULL_MASK, key0, j);
      if (kj == key0 && j < i0) first0 = false;
      if (kj == key1) first1 = false;                          // every position of the low half precedes i1
    }
    for (int j = 0; j < hi_n; ++j) {
      const int kj = __shfl_sync(FULL_MASK, key1, j);
      if (kj == key1 && j + 32 < i1) first1 = false;
    }
    const unsigned fm0 = __ballot_sync(FULL_MASK, first0), fm1 = __ballot_sync(FULL_MASK, first1);
    int ur0 = 0, ur1 = 0;
    for (int j = 0; j < lo_n; ++j) {
      const int kj = __shfl_sync(FULL_MASK, key0, j);
      if ((fm0 >> j) & 1u) { ur0 += kj < key0; ur1 += kj < key1; }
    }
    for (int j = 0; j < hi_n; ++j) {
      const int kj = __shfl_sync(FULL_MASK, key1, j);
      if ((fm1 >> j) & 1u) { ur0 += kj < key0; ur1 += kj < key1; }
    }
    if (first0) kout[b + ur0] = key0;
    if (first1) kout[b + ur1] = key1;
    if (lane == 0) ucount[seg] = __popc(fm0) + __popc(fm1);
  } else {
    int r0 = 0, r1 = 0;
    for (int j = 0; j < lo_n; ++j) {
      const int kj = __shfl_sync(FULL_MASK, key0, j);
      r0 += (kj < key0) || (kj == key0 && j < i0);
      r1 += (kj <= key1);                                      // low-half positions precede i1: ties rank before
    }
    for (int j = 0; j < hi_n; ++j) {
      const int kj = __shfl_sync(FULL_MASK, key1, j);
      r0 += (kj < key0);                                       // high-half positions follow i0: ties rank after
      r1 += (kj < key1) || (kj == key1 && j + 32 < i1);
    }
    if (i0 < len) { kout[b + r0] = key0; if (PAYLOAD) pout[b + r0] = pay0; }
    if (i1 < len) { kout[b + r1] = key1; if (PAYLOAD) pout[b + r1] = pay1; }
  }
}

template <bool UNIQUE, bool PAYLOAD>
__global__ void seg_sort_block(const int* __restrict__ off, const int* __restrict__ kin, const int* __restrict__ pin,
                               int* __restrict__ kout, int* __restrict__ pout, int* __restrict__ ucount, int nseg) {
  __shared__ int skey[SORT_SMEM_KEYS];
  __shared__ int scount;
  __shared__ int long_list[256];
  __shared__ int long_count;
  // find the long segments 256 at a time (one thread per segment), then sort each with the whole block
  // segments are dealt to the blocks round-robin: long segments (hubs) tend to have neighbouring ids
  for (int64_t base = 0; base < nseg; base += (int64_t)gridDim.x * 256) {
   if (threadIdx.x == 0) long_count = 0;
   __syncthreads();
   {
     const int64_t sg = base + (int64_t)threadIdx.x * gridDim.x + blockIdx.x;
     if (sg < nseg && off[sg + 1] - off[sg] > WARP_SEG) long_list[atomicAdd(&long_count, 1)] = (int)sg;
   }
   __syncthreads();
   const int nlong = long_count;
   for (int li = 0; li < nlong; ++li) {
    const int seg = long_list[li];
    int b = off[seg], len = off[seg + 1] - b;
    const bool in_smem = len <= SORT_SMEM_KEYS;
    __syncthreads();
    if (in_smem)
      for (int i = threadIdx.x; i < len; i += blockDim.x) skey[i] = kin[b + i];
    if (threadIdx.x == 0) scount = 0;
    __syncthreads();
    const int* keys = in_smem ? skey : (kin + b);
    int local_first = 0;
    // long segments that fit in shared memory: bitonic sort (O(len log^2 len)) instead of the O(len^2) rank sort --
    // a 2000-entry hub of a power-law graph costs ~20 us instead of ~200 us.  Same result: ascending unique keys /
    // ascending keys with their payload (keys of a transposed segment are distinct, so no tie-breaking is needed).
    const bool bitonic = in_smem && (!PAYLOAD || len <= SORT_SMEM_KEYS / 2);
    if (bitonic) {
      int npad = 1;
      while (npad < len) npad <<= 1;
      int* spay = skey + SORT_SMEM_KEYS / 2;                       // payload half (PAYLOAD only; then npad <= KEYS/2)
      for (int i = len + threadIdx.x; i < npad; i += blockDim.x) skey[i] = 0x7fffffff;
      if (PAYLOAD)
        for (int i = threadIdx.x; i < npad; i += blockDim.x) spay[i] = i < len ? pin[b + i] : 0;
      __syncthreads();
      for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = threadIdx.x; i < npad; i += blockDim.x) {
            const int ixj = i ^ j;
            if (ixj > i) {
              const int a = skey[i], c = skey[ixj];
              if ((a > c) == ((i & k) == 0)) {
                skey[i] = c; skey[ixj] = a;
                if (PAYLOAD) { const int t = spay[i]; spay[i] = spay[ixj]; spay[ixj] = t; }
              }
            }
          }
          __syncthreads();
        }
      if (UNIQUE) {
        // contiguous chunk per thread: count first occurrences, exclusive scan over the 256 counts, then write
        __shared__ int chunk_cnt[256];
        const int per = (len + (int)blockDim.x - 1) / (int)blockDim.x;
        const int c0 = min(len, (int)threadIdx.x * per), c1 = min(len, c0 + per);
        int cnt = 0;
        for (int i = c0; i < c1; ++i) cnt += (i == 0 || skey[i] != skey[i - 1]);
        chunk_cnt[threadIdx.x] = cnt;
        __syncthreads();
        int offs = 0;
        for (int t = 0; t < (int)threadIdx.x; ++t) offs += chunk_cnt[t];
        for (int i = c0; i < c1; ++i)
          if (i == 0 || skey[i] != skey[i - 1]) kout[b + offs++] = skey[i];
        local_first = cnt;
      } else {
        for (int i = threadIdx.x; i < len; i += blockDim.x) {
          kout[b + i] = skey[i];
          if (PAYLOAD) pout[b + i] = spay[i];
        }
      }
    } else if (UNIQUE) {
      // phase 1: first-occurrence flags (scratch lives in pout); phase 2: rank among distinct keys
      for (int i = threadIdx.x; i < len; i += blockDim.x) {
        int key = keys[i];
        int first = 1;
        for (int j = 0; j < i; ++j)
          if (keys[j] == key) { first = 0; break; }
        pout[b + i] = first;
      }
      __syncthreads();
      for (int i = threadIdx.x; i < len; i += blockDim.x) {
        if (!pout[b + i]) continue;
        int key = keys[i];
        int urank = 0;
        for (int j = 0; j < len; ++j) urank += (keys[j] < key) & pout[b + j];
        kout[b + urank] = key;
        ++local_first;
      }
    } else {
      for (int i = threadIdx.x; i < len; i += blockDim.x) {
        int key = keys[i];
        int rank = 0;
        for (int j = 0; j < len; ++j) {
          int kj = keys[j];
          if (kj < key || (kj == key && j < i)) ++rank;
        }
        kout[b + rank] = key;
        if (PAYLOAD) pout[b + rank] = pin[b + i];
      }
    }
    if (UNIQUE) {
      atomicAdd(&scount, local_first);
      __syncthreads();
      if (threadIdx.x == 0) ucount[seg] = scount;
    }
   }
   __syncthreads();
  }
}

// copy the packed unique keys of each row to their final place; also emit the row of every entry
__global__ void compact_rows(const int* __restrict__ off, const int* __restrict__ rowptr, const int* __restrict__ staged,
                             int* __restrict__ col, int* __restrict__ row, int n) {
  int r = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  if (r >= n) return;
  int src = off[r], dst = rowptr[r], len = rowptr[r + 1] - dst;
  for (int i = lane; i < len; i += 32) {
    col[dst + i] = staged[src + i];
    row[dst + i] = r;
  }
}

__global__ void count_cols(const int* __restrict__ rowptr, const int* __restrict__ col, int n, int* __restrict__ cnt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rowptr[n]) atomicAdd(&cnt[col[i]], 1);
}

__global__ void scatter_cols(const int* __restrict__ rowptr, const int* __restrict__ col, const int* __restrict__ row, int n,
                             const int* __restrict__ off_t, int* __restrict__ cursor, int* __restrict__ brow, int* __restrict__ bperm) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rowptr[n]) return;
  int c = col[i];
  int p = off_t[c] + atomicAdd(&cursor[c], 1);
  brow[p] = row[i];
  bperm[p] = (int)i;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct WsLayout {
  size_t cnt, off, cursor, bucket, staged, flags, ucount, scan, total;
};
WsLayout ws_layout(int64_t E, int32_t N) {
  WsLayout L;
  size_t p = 0;
  int64_t tot = E + N;
  auto take = [&](int64_t ints) { size_t o = p; p = align_up(p + (size_t)ints * 4, 256); return o; };
  L.cnt = take(N + 1);
  L.off = take(N + 1);
  L.cursor = take(N + 1);
  L.bucket = take(tot + 1);
  L.staged = take(tot + 1);
  L.flags = take(tot + 1);
  L.ucount = take(N + 1);
  L.scan = take(scan_ws_ints(tot > N ? tot : N) + scan_ws_ints(N + 1) + 16);
  L.total = p;
  return L;
}

}  // namespace

TAGAN_API size_t tagan_csr_workspace_bytes(int64_t num_edges, int32_t num_nodes) {
  if (num_edges < 0 || num_nodes < 0) return 0;
  return ws_layout(num_edges, num_nodes).total;
}

static int csr_build_impl(const int64_t* edge_index, const Batch* batch, int64_t E, int32_t N, int32_t row_begin, int32_t R,
                          int32_t* rowptr, int32_t* col, int32_t* row, int32_t* rowptr_t, int32_t* row_t,
                          int32_t* perm_t, int32_t* status, void* workspace, size_t workspace_bytes,
                          tagan_stream_t stream) {
  if (E < 0 || N < 0 || R < 0 || row_begin < 0 || row_begin + (int64_t)R > N || !rowptr || !col || !row || !status ||
      (E > 0 && !edge_index && !batch))
    return TAGAN_E_INVALID;
  if (E + (int64_t)N >= 0x7fffffffLL) return TAGAN_E_UNSUPPORTED;
  const bool transpose = rowptr_t || row_t || perm_t;
  if (transpose && !(rowptr_t && row_t && perm_t)) return TAGAN_E_INVALID;
  WsLayout L = ws_layout(E, N);
  if (!workspace || workspace_bytes < L.total) return TAGAN_E_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  char* w = static_cast<char*>(workspace);
  int* cnt = (int*)(w + L.cnt);
  int* off = (int*)(w + L.off);
  int* cursor = (int*)(w + L.cursor);
  int* bucket = (int*)(w + L.bucket);
  int* staged = (int*)(w + L.staged);
  int* flags = (int*)(w + L.flags);
  int* ucount = (int*)(w + L.ucount);
  int* scanws = (int*)(w + L.scan);
  const int TB = 256;
  cudaMemsetAsync(status, 0, sizeof(int), st);
  if (transpose) cudaMemsetAsync(rowptr_t, 0, sizeof(int) * ((size_t)N + 1), st);
  if (R == 0) {
    cudaMemsetAsync(rowptr, 0, sizeof(int), st);
    if (E > 0 && N == 0) fill_i32<<<1, 1, 0, st>>>(status, 1, 1);
    return tagan_launch_status();
  }
  const unsigned gR = ceil_div_i64(R, TB), gE = ceil_div_i64(E > 0 ? E : 1, TB);
  const unsigned gW = ceil_div_i64((int64_t)R * 32, TB);
  const unsigned gHeavy = 148 * 4;

  fill_i32<<<gR, TB, 0, st>>>(cnt, 1, R);         // one self loop per local row
  fill_i32<<<gR, TB, 0, st>>>(cursor, 1, R);
  if (E > 0) {
    if (batch) count_rows_batched<<<gE, TB, 0, st>>>(*batch, E, cnt, status);
    else count_rows<<<gE, TB, 0, st>>>(edge_index, E, N, row_begin, R, cnt, status);
  }
  exclusive_scan(cnt, off, R, scanws, st);
  place_self_loops<<<gR, TB, 0, st>>>(off, bucket, R, row_begin);
  if (E > 0) {
    if (batch) scatter_edges_batched<<<gE, TB, 0, st>>>(*batch, E, off, cursor, bucket);
    else scatter_edges<<<gE, TB, 0, st>>>(edge_index, E, N, row_begin, R, off, cursor, bucket);
  }
  seg_sort_warp<true, false><<<gW, TB, 0, st>>>(off, bucket, nullptr, staged, nullptr, ucount, R);
  seg_sort_block<true, false><<<gHeavy, TB, 0, st>>>(off, bucket, nullptr, staged, flags, ucount, R);
  exclusive_scan(ucount, rowptr, R, scanws, st);
  compact_rows<<<gW, TB, 0, st>>>(off, rowptr, staged, col, row, R);

  if (transpose) {                                 // indexed by GLOBAL source node, rows are local ids
    const unsigned gN = ceil_div_i64(N, TB), gT = ceil_div_i64(E + R, TB);
    const unsigned gWN = ceil_div_i64((int64_t)N * 32, TB);
    cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)N, st);
    cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)N, st);
    (void)gN;
    count_cols<<<gT, TB, 0, st>>>(rowptr, col, R, cnt);
    exclusive_scan(cnt, rowptr_t, N, scanws, st);
    scatter_cols<<<gT, TB, 0, st>>>(rowptr, col, row, R, rowptr_t, cursor, bucket, staged);
    seg_sort_warp<false, true><<<gWN, TB, 0, st>>>(rowptr_t, bucket, staged, row_t, perm_t, nullptr, N);
    seg_sort_block<false, true><<<gHeavy, TB, 0, st>>>(rowptr_t, bucket, staged, row_t, perm_t, nullptr, N);
  }
  return tagan_launch_status();
}

TAGAN_API int tagan_csr_build_part(const int64_t* edge_index, int64_t E, int32_t N, int32_t row_begin, int32_t R,
                                   int32_t* rowptr, int32_t* col, int32_t* row, int32_t* rowptr_t, int32_t* row_t,
                                   int32_t* perm_t, int32_t* status, void* workspace, size_t workspace_bytes,
                                   tagan_stream_t stream) {
  return csr_build_impl(edge_index, nullptr, E, N, row_begin, R, rowptr, col, row, rowptr_t, row_t, perm_t, status, workspace,
                        workspace_bytes, stream);
}

TAGAN_API int tagan_csr_build_batched(const int64_t* const* src, const int64_t* const* dst, const int64_t* edge_counts,
                                      const int32_t* node_counts, int32_t T, int32_t* rowptr, int32_t* col, int32_t* row,
                                      int32_t* rowptr_t, int32_t* row_t, int32_t* perm_t, int32_t* status, void* workspace,
                                      size_t workspace_bytes, tagan_stream_t stream) {
  if (T <= 0 || T > MAX_BATCH || !edge_counts || !node_counts || !src || !dst) return T > MAX_BATCH ? TAGAN_E_UNSUPPORTED : TAGAN_E_INVALID;
  Batch b;
  b.T = T;
  b.eoff[0] = 0;
  b.noff[0] = 0;
  for (int t = 0; t < T; ++t) {
    if (edge_counts[t] < 0 || node_counts[t] < 0 || (edge_counts[t] > 0 && (!src[t] || !dst[t]))) return TAGAN_E_INVALID;
    b.src[t] = src[t];
    b.dst[t] = dst[t];
    b.eoff[t + 1] = b.eoff[t] + edge_counts[t];
    const int64_t nn = (int64_t)b.noff[t] + node_counts[t];
    if (nn >= 0x7fffffffLL) return TAGAN_E_UNSUPPORTED;
    b.noff[t + 1] = (int)nn;
  }
  for (int t = T; t < MAX_BATCH; ++t) { b.src[t] = nullptr; b.dst[t] = nullptr; b.eoff[t + 1] = b.eoff[T]; b.noff[t + 1] = b.noff[T]; }
  const int32_t N = b.noff[T];
  return csr_build_impl(nullptr, &b, b.eoff[T], N, 0, N, rowptr, col, row, rowptr_t, row_t, perm_t, status, workspace,
                        workspace_bytes, stream);
}

TAGAN_API int tagan_csr_build(const int64_t* edge_index, int64_t E, int32_t N, int32_t* rowptr, int32_t* col,
                              int32_t* row, int32_t* rowptr_t, int32_t* row_t, int32_t* perm_t, int32_t* status,
                              void* workspace, size_t workspace_bytes, tagan_stream_t stream) {
  return tagan_csr_build_part(edge_index, E, N, 0, N, rowptr, col, row, rowptr_t, row_t, perm_t, status, workspace,
                              workspace_bytes, stream);
}
