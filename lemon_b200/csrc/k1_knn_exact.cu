// fp32 brute-force exact kNN on CUDA cores.
//
// Role: (1) GPU fallback for rows the tensor-core path could not certify, (2) general path for
// shapes / metrics the tensor-core kernel does not take, (3) the first-correct path of round 1.
// Replaces faiss IndexFlatIP / IndexFlatL2 .search (run_lemon.py:235-236) with the documented
// total order (value best-first, then DB index ascending).
//
// Layout: one CTA = 8 warps x 4 query rows = 32 query rows held in shared memory (fp32); every
// warp walks the whole DB two rows at a time (coalesced float4 loads; the 8 warps hit the same
// lines in L1), keeps one running threshold per query row and appends the rare survivors to a
// 256-entry shared buffer that a warp-wide bitonic sort compacts to the best 64.
#include "lemon_common.cuh"

namespace lemon {

constexpr int kExWarps = 8;
constexpr int kExRowsPerWarp = 4;
constexpr int kExRows = kExWarps * kExRowsPerWarp;

template <int METRIC>
__global__ void __launch_bounds__(kExWarps * 32, 1)
knn_exact_kernel(const float* __restrict__ q, const float* __restrict__ db, const int32_t* __restrict__ rows,
                 const int32_t* __restrict__ n_rows_ptr, int64_t nq, int64_t m, int d, int kp,
                 float* __restrict__ top_val, int32_t* __restrict__ top_idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* buf = reinterpret_cast<uint64_t*>(smem_raw);                       // [8][4][256]
  float* qs = reinterpret_cast<float*>(smem_raw + size_t(kExRows) * kCap * 8); // [32][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_rows = rows ? int64_t(*n_rows_ptr) : nq;
  const int d4 = d >> 2;
  const float4* db4 = reinterpret_cast<const float4*>(db);

  for (int64_t g = blockIdx.x; g * kExRows < n_rows; g += gridDim.x) {
    __syncthreads();
    // stage the 32 query rows
    int64_t my_row[kExRowsPerWarp];
    for (int r = 0; r < kExRows; ++r) {
      const int64_t li = g * kExRows + r;
      const int64_t row = li < n_rows ? (rows ? int64_t(rows[li]) : li) : -1;
      for (int c = threadIdx.x; c < d; c += blockDim.x) qs[r * d + c] = row >= 0 ? q[row * d + c] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < kExRowsPerWarp; ++r) {
      const int64_t li = g * kExRows + warp * kExRowsPerWarp + r;
      my_row[r] = li < n_rows ? (rows ? int64_t(rows[li]) : li) : -1;
    }
    __syncthreads();

    float theta[kExRowsPerWarp];
    int cnt[kExRowsPerWarp];
#pragma unroll
    for (int r = 0; r < kExRowsPerWarp; ++r) { theta[r] = -CUDART_INF_F; cnt[r] = 0; }
    const float4* q4 = reinterpret_cast<const float4*>(qs + size_t(warp) * kExRowsPerWarp * d);
    uint64_t* wbuf = buf + size_t(warp) * kExRowsPerWarp * kCap;

    for (int64_t j = 0; j < m; j += 2) {
      const bool has2 = (j + 1) < m;
      const float4* b0 = db4 + j * d4;
      const float4* b1 = db4 + (has2 ? j + 1 : j) * d4;
      float acc0[kExRowsPerWarp], acc1[kExRowsPerWarp];
#pragma unroll
      for (int r = 0; r < kExRowsPerWarp; ++r) { acc0[r] = 0.f; acc1[r] = 0.f; }
      for (int c = lane; c < d4; c += 32) {
        const float4 y0 = __ldg(b0 + c), y1 = __ldg(b1 + c);
#pragma unroll
        for (int r = 0; r < kExRowsPerWarp; ++r) {
          const float4 x = q4[r * d4 + c];
          if (METRIC == LEMON_METRIC_IP) {
            acc0[r] = fmaf(x.x, y0.x, acc0[r]); acc0[r] = fmaf(x.y, y0.y, acc0[r]);
            acc0[r] = fmaf(x.z, y0.z, acc0[r]); acc0[r] = fmaf(x.w, y0.w, acc0[r]);
            acc1[r] = fmaf(x.x, y1.x, acc1[r]); acc1[r] = fmaf(x.y, y1.y, acc1[r]);
            acc1[r] = fmaf(x.z, y1.z, acc1[r]); acc1[r] = fmaf(x.w, y1.w, acc1[r]);
          } else {
            float t;
            t = x.x - y0.x; acc0[r] = fmaf(t, t, acc0[r]); t = x.y - y0.y; acc0[r] = fmaf(t, t, acc0[r]);
            t = x.z - y0.z; acc0[r] = fmaf(t, t, acc0[r]); t = x.w - y0.w; acc0[r] = fmaf(t, t, acc0[r]);
            t = x.x - y1.x; acc1[r] = fmaf(t, t, acc1[r]); t = x.y - y1.y; acc1[r] = fmaf(t, t, acc1[r]);
            t = x.z - y1.z; acc1[r] = fmaf(t, t, acc1[r]); t = x.w - y1.w; acc1[r] = fmaf(t, t, acc1[r]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < kExRowsPerWarp; ++r) {
        float v0 = warp_sum(acc0[r]);
        float v1 = warp_sum(acc1[r]);
        if (METRIC == LEMON_METRIC_L2) { v0 = -v0; v1 = -v1; }   // larger == better
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float v = h ? v1 : v0;
          if (h && !has2) break;
          if (v > theta[r]) {                                    // warp-uniform
            if (lane == 0) wbuf[r * kCap + cnt[r]] = make_key(v, uint32_t(j + h));
            cnt[r]++;
            if (cnt[r] == kCap) {
              __syncwarp();
              uint64_t key[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) key[i] = wbuf[r * kCap + lane * 8 + i];
              warp_sort256_desc(key, lane);
              __syncwarp();
              if (lane < kKeep / 8) {
#pragma unroll
                for (int i = 0; i < 8; ++i) wbuf[r * kCap + lane * 8 + i] = key[i];
              }
              theta[r] = key_val(shfl_u64(key[7], kKeep / 8 - 1));
              cnt[r] = kKeep;
              __syncwarp();
            }
          }
        }
      }
    }
    // final: sort what is left, emit the best kp
#pragma unroll
    for (int r = 0; r < kExRowsPerWarp; ++r) {
      __syncwarp();
      uint64_t key[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int e = lane * 8 + i;
        key[i] = e < cnt[r] ? wbuf[r * kCap + e] : 0ull;
      }
      warp_sort256_desc(key, lane);
      if (my_row[r] >= 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int e = lane * 8 + i;
          if (e < kp) {
            const bool ok = key[i] != 0ull;
            float v = ok ? key_val(key[i]) : -CUDART_INF_F;
            if (METRIC == LEMON_METRIC_L2) v = -v;
            top_val[my_row[r] * kp + e] = v;
            top_idx[my_row[r] * kp + e] = ok ? key_idx(key[i]) : -1;
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
  const size_t smem = size_t(kExRows) * kCap * 8 + size_t(kExRows) * d * 4;
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
