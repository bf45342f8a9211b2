// (c1-c3) Node memory bank on dense device tables keyed by slot (= node id).
// Replaces the python-dict NodeMemoryBank of the reference (src/tagan/utils/memory_bank.py:14-360):
//   table[cap,H] fp32 states; valid/has_seen u8; last_seen/inactivity/frequency i32.
// update() reproduces the sequential loop of :65-173 in four stream-ordered passes; integer
// bookkeeping is bit-exact, and states are bit-exact too because every product and sum is rounded
// separately (__fmul_rn/__fadd_rn: no FMA contraction), with the python-double scalars
// (blend weights, decay^k) rounded to fp32 once on the host exactly as torch does.
#include "common.cuh"
#include <limits.h>

namespace {

// pass A: every known id gets inactivity+1 (:89-90); reset the per-call occurrence marks
__global__ void bank_tick_kernel(const uint8_t* __restrict__ valid, int* __restrict__ inactivity,
                                 int* __restrict__ mark_min, int* __restrict__ mark_max, int cap) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  if (valid[i]) inactivity[i] += 1;
  mark_min[i] = INT_MAX;
  mark_max[i] = -1;
}

// pass B: first listed position (over all ids) and last updatable position (over ids that have a
// state row), frequency++ per updatable occurrence (:97)
__global__ void bank_mark_kernel(const int* __restrict__ ids, int64_t m_all, int64_t m_upd, int cap,
                                 int* __restrict__ mark_min, int* __restrict__ mark_max, int* __restrict__ frequency,
                                 int* __restrict__ status) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m_all) return;
  int id = ids[i];
  if (id < 0 || id >= cap) { *status = 1; return; }
  atomicMin(&mark_min[id], (int)i);
  if (i < m_upd) {
    atomicMax(&mark_max[id], (int)i);
    atomicAdd(&frequency[id], 1);
  }
}

// pass C: one warp per state row; only the LAST occurrence of an id writes (:93-141).  A blend
// with the stored state happens only for a reappearing id listed once (later duplicates see
// last_seen == timestep and overwrite, exactly like the sequential loop).
__global__ void bank_write_kernel(float* __restrict__ table, uint8_t* __restrict__ valid, uint8_t* __restrict__ has_seen,
                                  int* __restrict__ last_seen, int* __restrict__ inactivity,
                                  const int* __restrict__ ids, const float* __restrict__ states, int64_t lds,
                                  int64_t m_upd, int H, int cap, int timestep, float w2, float omw2, float w3,
                                  float omw3, const int* __restrict__ mark_min, const int* __restrict__ mark_max) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (i >= m_upd) return;
  int id = ids[i];
  if (id < 0 || id >= cap) return;
  if (mark_max[id] != (int)i) return;
  const bool once = mark_min[id] == (int)i;
  const bool reappearing = once && valid[id] && has_seen[id] && last_seen[id] < timestep - 1;   // :100-103
  float* row = table + (int64_t)id * H;
  const float* cur = states + i * lds;
  if (reappearing) {
    const int gap = timestep - last_seen[id];
    const float w = gap >= 3 ? w3 : w2, omw = gap >= 3 ? omw3 : omw2;                          // :124
    for (int c = lane; c < H; c += 32) row[c] = __fadd_rn(__fmul_rn(w, row[c]), __fmul_rn(omw, cur[c]));   // :127
  } else {
    for (int c = lane; c < H; c += 32) row[c] = cur[c];                                        // :135
  }
  __syncwarp();
  if (lane == 0) {
    valid[id] = 1;
    has_seen[id] = 1;
    inactivity[id] = 0;      // :138
    last_seen[id] = timestep;  // :141
  }
}

// pass D: decay every stored id that was not listed (:149-153), prune (:156-166), count (:169)
__global__ void bank_sweep_kernel(float* __restrict__ table, uint8_t* __restrict__ valid, uint8_t* __restrict__ has_seen,
                                  int* __restrict__ last_seen, int* __restrict__ inactivity,
                                  const int* __restrict__ mark_min, int H, int cap, const float* __restrict__ decay_pow,
                                  int pow_len, double decay, int max_inactivity, int* __restrict__ size_out) {
  int id = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  int keep = 0;
  if (id < cap && valid[id]) {
    const int ina = inactivity[id];
    if (mark_min[id] == INT_MAX) {
      const float d = ina < pow_len ? decay_pow[ina] : (float)pow(decay, (double)ina);
      float* row = table + (int64_t)id * H;
      for (int c = lane; c < H; c += 32) row[c] = __fmul_rn(row[c], d);
    }
    __syncwarp();
    if (ina > max_inactivity) {
      if (lane == 0) { valid[id] = 0; has_seen[id] = 0; last_seen[id] = 0; inactivity[id] = 0; }
      float* row = table + (int64_t)id * H;
      for (int c = lane; c < H; c += 32) row[c] = 0.f;
    } else {
      keep = 1;
    }
  }
  // integer count of surviving ids: warp leaders -> block -> one atomic per block (deterministic)
  __shared__ int block_count;
  if (threadIdx.x == 0) block_count = 0;
  __syncthreads();
  if (lane == 0 && keep) atomicAdd(&block_count, 1);
  __syncthreads();
  if (threadIdx.x == 0 && block_count) atomicAdd(size_out, block_count);
}

// get_states (:187-211)
__global__ void bank_gather_kernel(float* __restrict__ table, uint8_t* __restrict__ valid, int* __restrict__ inactivity,
                                   const int* __restrict__ ids, float* __restrict__ out, int64_t M, int H, int cap,
                                   int* __restrict__ status) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (i >= M) return;
  int id = ids[i];
  float* o = out + i * H;
  if (id < 0 || id >= cap) {
    if (lane == 0) *status = 1;
    for (int c = lane; c < H; c += 32) o[c] = 0.f;
    return;
  }
  float* row = table + (int64_t)id * H;
  if (valid[id]) {
    for (int c = lane; c < H; c += 32) o[c] = row[c];
  } else {
    // unknown id: zeros, and inserted with inactivity 0 and no last_seen (:203-209).  Duplicate
    // unknown ids in one call all write the same values.
    for (int c = lane; c < H; c += 32) { o[c] = 0.f; row[c] = 0.f; }
    __syncwarp();
    if (lane == 0) { inactivity[id] = 0; }
  }
}
// validity is flipped in a second pass so that duplicates of an unknown id inside one call all take
// the "unknown" branch above regardless of scheduling
__global__ void bank_gather_commit_kernel(uint8_t* __restrict__ valid, const int* __restrict__ ids, int64_t M, int cap) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  int id = ids[i];
  if (id >= 0 && id < cap) valid[id] = 1;
}

__global__ void bank_decay_all_kernel(float* __restrict__ table, const uint8_t* __restrict__ valid, float decay, int H,
                                      int cap) {
  int id = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  if (id >= cap || !valid[id]) return;
  float* row = table + (int64_t)id * H;
  for (int c = lane; c < H; c += 32) row[c] = __fmul_rn(row[c], decay);
}

}  // namespace

TAGAN_API int tagan_bank_gather(float* table, uint8_t* valid, int32_t* inactivity, const int32_t* ids, float* out,
                                int64_t M, int32_t H, int32_t capacity, int32_t* status, tagan_stream_t stream) {
  if (!table || !valid || !inactivity || !status || M < 0 || H <= 0 || capacity < 0 || (M > 0 && (!ids || !out)))
    return TAGAN_E_INVALID;
  if (M == 0) return 0;
  cudaStream_t st = as_stream(stream);
  bank_gather_kernel<<<ceil_div_i64(M * 32, 256), 256, 0, st>>>(table, valid, inactivity, ids, out, M, H, capacity, status);
  bank_gather_commit_kernel<<<ceil_div_i64(M, 256), 256, 0, st>>>(valid, ids, M, capacity);
  return tagan_launch_status();
}

TAGAN_API int tagan_bank_update(float* table, uint8_t* valid, uint8_t* has_seen, int32_t* last_seen,
                                int32_t* inactivity, int32_t* frequency, const int32_t* ids, int64_t num_ids,
                                const float* states, int64_t lds, int64_t num_states, int32_t H, int32_t capacity,
                                int32_t timestep, const float* w23, const float* decay_pow, int32_t decay_pow_len,
                                double decay, int32_t max_inactivity, int32_t* mark_ws, int32_t* size_out,
                                int32_t* status, tagan_stream_t stream) {
  if (!table || !valid || !has_seen || !last_seen || !inactivity || !frequency || !w23 || !decay_pow || !mark_ws ||
      !size_out || !status || num_ids < 0 || num_states < 0 || H <= 0 || capacity < 0 || decay_pow_len < 0)
    return TAGAN_E_INVALID;
  if (num_ids > 0 && !ids) return TAGAN_E_INVALID;
  const int64_t m_upd = num_ids < num_states ? num_ids : num_states;   // bounds check of :95
  if (m_upd > 0 && !states) return TAGAN_E_INVALID;
  cudaStream_t st = as_stream(stream);
  int* mark_min = mark_ws;
  int* mark_max = mark_ws + capacity;
  cudaMemsetAsync(size_out, 0, sizeof(int), st);
  if (capacity == 0) return 0;
  bank_tick_kernel<<<ceil_div_i64(capacity, 256), 256, 0, st>>>(valid, inactivity, mark_min, mark_max, capacity);
  if (num_ids > 0)
    bank_mark_kernel<<<ceil_div_i64(num_ids, 256), 256, 0, st>>>(ids, num_ids, m_upd, capacity, mark_min, mark_max,
                                                                 frequency, status);
  if (m_upd > 0)
    bank_write_kernel<<<ceil_div_i64(m_upd * 32, 256), 256, 0, st>>>(table, valid, has_seen, last_seen, inactivity, ids,
                                                                     states, lds, m_upd, H, capacity, timestep, w23[0],
                                                                     w23[1], w23[2], w23[3], mark_min, mark_max);
  bank_sweep_kernel<<<ceil_div_i64((int64_t)capacity * 32, 256), 256, 0, st>>>(table, valid, has_seen, last_seen,
                                                                               inactivity, mark_min, H, capacity,
                                                                               decay_pow, decay_pow_len, decay,
                                                                               max_inactivity, size_out);
  return tagan_launch_status();
}

TAGAN_API int tagan_bank_decay_all(float* table, const uint8_t* valid, float decay, int32_t H, int32_t capacity,
                                   tagan_stream_t stream) {
  if (!table || !valid || H <= 0 || capacity < 0) return TAGAN_E_INVALID;
  if (capacity == 0) return 0;
  bank_decay_all_kernel<<<ceil_div_i64((int64_t)capacity * 32, 256), 256, 0, as_stream(stream)>>>(table, valid, decay, H, capacity);
  return tagan_launch_status();
}
