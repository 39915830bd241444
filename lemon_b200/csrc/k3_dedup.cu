// Exact-duplicate handling for the kNN database.  Classification datasets have only C distinct text
// embeddings (run_lemon.py:117-119,140-143) and caption-noise injection duplicates captions
// (lib/datasets/noise_captioning.py:44-53), so thousands of DB rows can be bit-identical.  Identical rows
// are searched once: rows are hashed, grouped (host: sort), verified bit-for-bit, the search runs on one
// representative per group and the result is expanded back to the members by ascending DB index — the same
// list the full search returns under the documented total order (value best-first, then index ascending).
#include "lemon_common.cuh"

namespace lemon {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
hash_rows_kernel(const uint32_t* __restrict__ x, int64_t n, int d, int64_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
    const uint32_t* r = x + row * d;
    uint64_t h = 0x9e3779b97f4a7c15ull * uint64_t(lane + 1);
    for (int c = lane; c < d; c += 32) h = mix64(h ^ (uint64_t(r[c]) | (uint64_t(c) << 32)));
    // order-independent combine across lanes is fine: every lane's stream is position-tagged
    uint64_t acc = h;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += shfl_xor_u64(acc, o);
    if (lane == 0) out[row] = int64_t(mix64(acc) >> 1);   // non-negative: sorts the same signed or unsigned
  }
}

// flag[0] |= 1 if some row differs bit-wise from its group's representative row
__global__ void __launch_bounds__(256)
rows_equal_kernel(const uint32_t* __restrict__ x, const int64_t* __restrict__ rep_of_row, int64_t n, int d,
                  int32_t* __restrict__ flag) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
    const int64_t rep = rep_of_row[row];
    if (rep == row) continue;
    const uint32_t* a = x + row * d;
    const uint32_t* b = x + rep * d;
    bool diff = false;
    for (int c = lane; c < d; c += 32) diff |= a[c] != b[c];
    if (__any_sync(kFull, diff) && lane == 0) atomicOr(flag, 1);
  }
}

// top lists over the UNIQUE rows -> top lists over the original DB rows
template <int METRIC>
__global__ void __launch_bounds__(256)
expand_groups_kernel(const float* __restrict__ uval, const int32_t* __restrict__ uidx, const int64_t* __restrict__ offsets,
                     const int32_t* __restrict__ members, int64_t nq, int kp, float* __restrict__ top_val,
                     int32_t* __restrict__ top_idx) {
  const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= nq) return;
  const float* uv = uval + row * kp;
  const int32_t* ui = uidx + row * kp;
  float* ov = top_val + row * kp;
  int32_t* oi = top_idx + row * kp;
  int pos = 0;
  for (int r = 0; r < kp && pos < kp; ++r) {
    const int u = ui[r];
    if (u < 0) break;
    const float v = uv[r];
    const int64_t e = offsets[u + 1];
    for (int64_t j = offsets[u]; j < e && pos < kp; ++j) { ov[pos] = v; oi[pos] = members[j]; ++pos; }
  }
  for (; pos < kp; ++pos) { ov[pos] = METRIC == LEMON_METRIC_IP ? -CUDART_INF_F : CUDART_INF_F; oi[pos] = -1; }
}

}  // namespace lemon

extern "C" int lemon_hash_rows(lemon_ctx* ctx, const float* x, int64_t n, int d, int64_t* out, void* stream) {
  if (!ctx) return LEMON_ERR_INVALID;
  if (!x || !out || n < 0 || d <= 0) return lemon_set_error(ctx, LEMON_ERR_INVALID, "hash_rows: bad args");
  if (n == 0) return LEMON_OK;
  int64_t blocks = (n + 7) / 8;
  const int64_t cap = int64_t(ctx->num_sms) * 16;
  if (blocks > cap) blocks = cap;
  lemon::hash_rows_kernel<<<unsigned(blocks), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t*>(x), n, d, out);
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}

extern "C" int lemon_rows_equal(lemon_ctx* ctx, const float* x, const int64_t* rep_of_row, int64_t n, int d,
                                int32_t* flag, void* stream) {
  if (!ctx) return LEMON_ERR_INVALID;
  if (!x || !rep_of_row || !flag || n < 0 || d <= 0) return lemon_set_error(ctx, LEMON_ERR_INVALID, "rows_equal: bad args");
  if (n == 0) return LEMON_OK;
  int64_t blocks = (n + 7) / 8;
  const int64_t cap = int64_t(ctx->num_sms) * 16;
  if (blocks > cap) blocks = cap;
  lemon::rows_equal_kernel<<<unsigned(blocks), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t*>(x), rep_of_row,
                                                                             n, d, flag);
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}

extern "C" int lemon_expand_groups(lemon_ctx* ctx, const float* uval, const int32_t* uidx, const int64_t* offsets,
                                   const int32_t* members, int64_t nq, int kp, int metric, float* top_val,
                                   int32_t* top_idx, void* stream) {
  if (!ctx) return LEMON_ERR_INVALID;
  if (!uval || !uidx || !offsets || !members || !top_val || !top_idx || nq < 0 || kp < 1 || kp > LEMON_MAX_KP)
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "expand_groups: bad args");
  if (nq == 0) return LEMON_OK;
  const unsigned blocks = unsigned((nq + 255) / 256);
  if (metric == LEMON_METRIC_IP)
    lemon::expand_groups_kernel<LEMON_METRIC_IP><<<blocks, 256, 0, (cudaStream_t)stream>>>(uval, uidx, offsets, members, nq, kp, top_val, top_idx);
  else
    lemon::expand_groups_kernel<LEMON_METRIC_L2><<<blocks, 256, 0, (cudaStream_t)stream>>>(uval, uidx, offsets, members, nq, kp, top_val, top_idx);
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}
