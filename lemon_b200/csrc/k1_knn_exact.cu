// fp32 brute-force exact kNN on CUDA cores.
//
// Role: (1) GPU fallback for rows the tensor-core path could not certify, (2) general path for
// shapes the tensor-core kernel does not take (padded d > 768, databases under 2048 rows),
// (3) `knn_mode="exact"`.  Replaces faiss IndexFlatIP / IndexFlatL2 .search (run_lemon.py:235-236) with the
// documented total order (value best-first, then DB index ascending).
//
// Layout: one CTA = 32 query rows staged in shared memory (fp32) and 8 warps.  Warp w works on query rows
// 8*(w&3) .. +7 and on the DB row groups (4 rows each) of parity w>>2, so two warps share a query group and
// split the DB.  Per step a warp forms 8 x 4 = 32 pair values: every lane accumulates its column slice
// (float4 index lane, lane+32, ...) exactly as `warp_pair_value` does, then a 5-step VALUE-SPLITTING butterfly
// (xor 16, 8, 4, 2, 1; 31 shuffles instead of 160) leaves the fully reduced value #l in lane l with the same
// summation tree as `warp_sum` -- the reported values are bit-identical to every other kernel's.  Each lane
// compares its value with the query row's running threshold; the rare survivors go to a per-(warp,row)
// 128-slot key buffer that a warp-wide bitonic sort prunes to the best 64.
#include "lemon_common.cuh"

namespace lemon {

constexpr int kExWarps = 8;
constexpr int kExRows = 32;            // query rows per CTA
constexpr int kExR = 8;                // query rows per warp
constexpr int kExJ = 4;                // DB rows per step
constexpr int kExCap = 128;            // key slots per (warp, query row)

// 32 partial sums per lane -> lane l holds the warp-wide sum of value #l (same tree as warp_sum)
__device__ __forceinline__ float split_butterfly32(float (&a)[32], int lane) {
  float b16[16];
  {
    const bool hi = (lane & 16) != 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float send = hi ? a[k] : a[k + 16];
      const float recv = __shfl_xor_sync(kFull, send, 16);
      b16[k] = (hi ? a[k + 16] : a[k]) + recv;
    }
  }
  float b8[8];
  {
    const bool hi = (lane & 8) != 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float send = hi ? b16[k] : b16[k + 8];
      const float recv = __shfl_xor_sync(kFull, send, 8);
      b8[k] = (hi ? b16[k + 8] : b16[k]) + recv;
    }
  }
  float b4[4];
  {
    const bool hi = (lane & 4) != 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float send = hi ? b8[k] : b8[k + 4];
      const float recv = __shfl_xor_sync(kFull, send, 4);
      b4[k] = (hi ? b8[k + 4] : b8[k]) + recv;
    }
  }
  float b2[2];
  {
    const bool hi = (lane & 2) != 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float send = hi ? b4[k] : b4[k + 2];
      const float recv = __shfl_xor_sync(kFull, send, 2);
      b2[k] = (hi ? b4[k + 2] : b4[k]) + recv;
    }
  }
  const bool hi = (lane & 1) != 0;
  const float send = hi ? b2[0] : b2[1];
  const float recv = __shfl_xor_sync(kFull, send, 1);
  return (hi ? b2[1] : b2[0]) + recv;
}

// sorts one 128-slot buffer (cnt valid keys) descending through the 256-key network and keeps the best `keep`
__device__ __forceinline__ void sort_keep128(uint64_t* buf, int cnt, int keep, int lane, uint64_t (&key)[8]) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) { const int e = lane * 8 + i; key[i] = (e < cnt && e < kExCap) ? buf[e] : 0ull; }
  warp_sort256_desc(key, lane);
  __syncwarp();
  if (lane < keep / 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) buf[lane * 8 + i] = key[i];
  }
  __syncwarp();
}

template <int METRIC>
__global__ void __launch_bounds__(kExWarps * 32, 1)
knn_exact_kernel(const float* __restrict__ q, const float* __restrict__ db, const int32_t* __restrict__ rows,
                 const int32_t* __restrict__ n_rows_ptr, int64_t nq, int64_t m, int d, int kp,
                 float* __restrict__ top_val, int32_t* __restrict__ top_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bufs = reinterpret_cast<uint64_t*>(smem_raw);                                   // [8 warps][8 rows][128]
  float* qs = reinterpret_cast<float*>(smem_raw + size_t(kExWarps) * kExR * kExCap * 8);    // [32][d]
  __shared__ float theta_s[kExWarps][kExR];
  __shared__ int cnt_s[kExWarps][kExR];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qgrp = warp & 3, par = warp >> 2;
  const int64_t n_rows = rows ? int64_t(*n_rows_ptr) : nq;
  const int d4 = d >> 2;
  const float4* db4 = reinterpret_cast<const float4*>(db);
  uint64_t* wbuf = bufs + size_t(warp) * kExR * kExCap;
  const int my_r = lane >> 2, my_j = lane & 3;          // after the butterfly lane l holds pair (query my_r, DB row my_j)

  for (int64_t g = blockIdx.x; g * kExRows < n_rows; g += gridDim.x) {
    __syncthreads();
    for (int r = 0; r < kExRows; ++r) {                  // stage the 32 query rows
      const int64_t li = g * kExRows + r;
      const int64_t row = li < n_rows ? (rows ? int64_t(rows[li]) : li) : -1;
      for (int c = threadIdx.x; c < d; c += blockDim.x) qs[r * d + c] = row >= 0 ? q[row * d + c] : 0.f;
    }
    if (lane < kExR) { theta_s[warp][lane] = -CUDART_INF_F; cnt_s[warp][lane] = 0; }
    __syncthreads();
    const float4* q4 = reinterpret_cast<const float4*>(qs + size_t(qgrp) * kExR * d);

    const int64_t ngroups = (m + kExJ - 1) / kExJ;
    for (int64_t jg = par; jg < ngroups; jg += 2) {
      const int64_t j0 = jg * kExJ;
      const float4* bp[kExJ];
#pragma unroll
      for (int j = 0; j < kExJ; ++j) bp[j] = db4 + min(j0 + j, m - 1) * d4;
      float acc[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) acc[k] = 0.f;
      for (int c = lane; c < d4; c += 32) {
        float4 y[kExJ];
#pragma unroll
        for (int j = 0; j < kExJ; ++j) y[j] = __ldg(bp[j] + c);
#pragma unroll
        for (int r = 0; r < kExR; ++r) {
          const float4 x = q4[r * d4 + c];
#pragma unroll
          for (int j = 0; j < kExJ; ++j) {
            float& a = acc[r * kExJ + j];
            if (METRIC == LEMON_METRIC_IP) {
              a = fmaf(x.x, y[j].x, a); a = fmaf(x.y, y[j].y, a); a = fmaf(x.z, y[j].z, a); a = fmaf(x.w, y[j].w, a);
            } else {
              float t;
              t = x.x - y[j].x; a = fmaf(t, t, a); t = x.y - y[j].y; a = fmaf(t, t, a);
              t = x.z - y[j].z; a = fmaf(t, t, a); t = x.w - y[j].w; a = fmaf(t, t, a);
            }
          }
        }
      }
      float v = split_butterfly32(acc, lane);            // value of (query my_r, DB row j0 + my_j)
      if (METRIC == LEMON_METRIC_L2) v = -v;             // larger == better
      const bool hit = (j0 + my_j) < m && v > theta_s[warp][my_r];
      unsigned mask = __ballot_sync(kFull, hit);
      while (mask) {                                     // rare; lanes are in (row, index) ascending order
        const int L = __ffs(mask) - 1;
        mask &= mask - 1;
        const int r = L >> 2;
        const float vL = __shfl_sync(kFull, v, L);
        const int64_t idxL = j0 + (L & 3);
        int c = cnt_s[warp][r];
        if (vL > theta_s[warp][r]) {                     // the threshold may have risen inside this loop
          if (lane == 0) { wbuf[r * kExCap + c] = make_key(vL, uint32_t(idxL)); cnt_s[warp][r] = c + 1; }
          ++c;
          if (c == kExCap) {
            uint64_t key[8];
            sort_keep128(wbuf + r * kExCap, c, kKeep, lane, key);
            const float th = key_val(shfl_u64(key[7], kKeep / 8 - 1));
            if (lane == 0) { theta_s[warp][r] = th; cnt_s[warp][r] = kKeep; }
          }
          __syncwarp();
        }
      }
    }
    // ---- both warps of a query group sort their lists; the parity-0 warp merges and emits the best kp
    uint64_t key[8];
    for (int r = 0; r < kExR; ++r) sort_keep128(wbuf + r * kExCap, cnt_s[warp][r], kKeep, lane, key);
    __syncthreads();
    if (par == 0) {
      const uint64_t* obuf = bufs + size_t(warp + 4) * kExR * kExCap;
      for (int r = 0; r < kExR; ++r) {
        const int c0 = min(cnt_s[warp][r], kKeep), c1 = min(cnt_s[warp + 4][r], kKeep);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int e = lane * 8 + i;            // slots 0..63: own list, 64..127: the other warp's
          key[i] = e < kKeep ? (e < c0 ? wbuf[r * kExCap + e] : 0ull)
                             : (e < 2 * kKeep && (e - kKeep) < c1 ? obuf[r * kExCap + e - kKeep] : 0ull);
        }
        warp_sort256_desc(key, lane);
        const int64_t li = g * kExRows + qgrp * kExR + r;
        const int64_t row = li < n_rows ? (rows ? int64_t(rows[li]) : li) : -1;
        if (row >= 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int e = lane * 8 + i;
            if (e < kp) {
              const bool ok = key[i] != 0ull;
              float v = ok ? key_val(key[i]) : -CUDART_INF_F;
              if (METRIC == LEMON_METRIC_L2) v = -v;
              top_val[row * kp + e] = v;
              top_idx[row * kp + e] = ok ? key_idx(key[i]) : -1;
            }
          }
        }
      }
    }
  }
}

}  // namespace lemon

extern "C" int lemon_knn_exact(lemon_ctx* ctx, const float* q, const float* db, const int32_t* rows,
                               const int32_t* n_rows, int64_t max_rows, int64_t nq, int64_t m, int d, int kp,
                               int metric, float* top_val, int32_t* top_idx, void* stream) {
  using namespace lemon;
  if (!ctx) return LEMON_ERR_INVALID;
  if (!q || !db || !top_val || !top_idx || nq < 0 || m < 0 || d <= 0 || (d & 3) || kp < 1 || kp > LEMON_MAX_KP ||
      (rows && !n_rows) || m >= (int64_t(1) << 31))
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "knn_exact: bad args (d %% 4 == 0, 1 <= kp <= %d required)", LEMON_MAX_KP);
  if (!rows) max_rows = nq;
  if (max_rows <= 0) return LEMON_OK;
  if (m == 0) return lemon_set_error(ctx, LEMON_ERR_INVALID, "knn_exact: empty database");
  const size_t smem = size_t(kExWarps) * kExR * kExCap * 8 + size_t(kExRows) * d * 4;
  if (smem > 227 * 1024) return lemon_set_error(ctx, LEMON_ERR_INVALID, "knn_exact: d=%d too large for shared staging", d);
  auto kern = metric == LEMON_METRIC_IP ? knn_exact_kernel<LEMON_METRIC_IP> : knn_exact_kernel<LEMON_METRIC_L2>;
  LEMON_CUDA_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int64_t groups = (max_rows + kExRows - 1) / kExRows;
  int64_t blocks = groups < int64_t(ctx->num_sms) * 4 ? groups : int64_t(ctx->num_sms) * 4;
  kern<<<unsigned(blocks), kExWarps * 32, smem, (cudaStream_t)stream>>>(q, db, rows, n_rows, nq, m, d, kp, top_val, top_idx);
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}
