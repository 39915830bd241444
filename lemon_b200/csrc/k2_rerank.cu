// K2a: fp32 exact re-rank of the tensor-core candidates + per-row exactness certificate.
// One warp per query row.  HBM-bound gather: ncand * d * 4 B per row.
// (North star: "low-precision candidates get an fp32 exact re-rank of a margin-widened set".)
#include "lemon_common.cuh"

namespace lemon {

constexpr int kRrWarps = 8;

template <int METRIC>
__global__ void __launch_bounds__(kRrWarps * 32)
rerank_kernel(const float* __restrict__ q, const float* __restrict__ db, const float* __restrict__ cand_val,
              const int32_t* __restrict__ cand_idx, const float* __restrict__ q_row_stats,
              const float* __restrict__ db_stats_max, float acc_eps, int64_t nq, int64_t m, int d, int ncand,
              int nseg, int kp, float* __restrict__ top_val, int32_t* __restrict__ top_idx,
              int32_t* __restrict__ uncert_rows, int32_t* __restrict__ n_uncert) {
  __shared__ uint64_t sbuf[kRrWarps][kCap];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* buf = sbuf[warp];
  const int64_t warps = int64_t(gridDim.x) * kRrWarps;
  for (int64_t row = int64_t(blockIdx.x) * kRrWarps + warp; row < nq; row += warps) {
    const float* qr = q + row * d;
    const int32_t* ci = cand_idx + row * ncand;
    const float* cv = cand_val + row * ncand;
    int cnt = 0;
    __syncwarp();
    // rounding-error bound of this row (see include/lemon_b200.h) and the candidate cut-off: the kp best
    // approximate values a_(1..kp) certify kp elements with exact value >= a_(kp) - eps, so a candidate whose
    // approximate value is below a_(kp) - 2 eps cannot be in the exact top-kp and is not gathered.
    float eps = 0.f, qsq = 1.f, dbdev = 0.f;
    if (q_row_stats) {
      const float4 st = reinterpret_cast<const float4*>(q_row_stats)[row];   // {||q||, ||q16||, ||q-q16||, ||q||^2}
      eps = st.z * db_stats_max[1] + st.x * db_stats_max[2] + acc_eps;
      qsq = st.w;
      dbdev = db_stats_max[3];
    }
    float akp = -CUDART_INF_F;
    for (int s = lane; s < nseg; s += 32) {
      const int pos = s * kKeep + kp - 1;
      if (ci[pos] >= 0) akp = fmaxf(akp, cv[pos]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) akp = fmaxf(akp, __shfl_xor_sync(kFull, akp, o));
    const float cut = akp - 2.f * eps - (METRIC == LEMON_METRIC_L2 ? dbdev : 0.f);
    for (int c0 = 0; c0 < ncand; c0 += 32) {
      // lanes fetch 32 candidate ids at once, then the warp evaluates them one by one
      int my = (c0 + lane) < ncand ? ci[c0 + lane] : -1;
      if (my >= 0 && cv[c0 + lane] < cut) my = -1;
      const int nc = min(32, ncand - c0);
      for (int t = 0; t < nc; ++t) {
        const int idx = __shfl_sync(kFull, my, t);
        if (idx < 0 || int64_t(idx) >= m) continue;               // padding (warp-uniform)
        float v = warp_pair_value<METRIC>(qr, db + int64_t(idx) * d, d, lane);
        if (METRIC == LEMON_METRIC_L2) v = -v;
        if (lane == 0) buf[cnt] = make_key(v, uint32_t(idx));
        cnt++;
        if (cnt == kCap) {
          __syncwarp();
          uint64_t key[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) key[i] = buf[lane * 8 + i];
          warp_sort256_desc(key, lane);
          __syncwarp();
          if (lane < kKeep / 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) buf[lane * 8 + i] = key[i];
          }
          cnt = kKeep;
          __syncwarp();
        }
      }
    }
    __syncwarp();
    uint64_t key[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int e = lane * 8 + i; key[i] = e < cnt ? buf[e] : 0ull; }
    warp_sort256_desc(key, lane);
    // emit the exact top list
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
    // certificate: every non-candidate has approx ip <= B (B = max over segments of the segment's
    // last kept value, -inf when the segment kept everything it saw)
    float B = -CUDART_INF_F;
    for (int s = lane; s < nseg; s += 32) {
      const int last = s * kKeep + kKeep - 1;
      if (ci[last] >= 0) B = fmaxf(B, cv[last]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) B = fmaxf(B, __shfl_xor_sync(kFull, B, o));
    uint64_t kth_sel = 0ull;
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i == ((kp - 1) & 7)) kth_sel = key[i];
    const uint64_t kth_key = shfl_u64(kth_sel, (kp - 1) >> 3);
    if (lane == 0 && B > -CUDART_INF_F) {
      const float dbmin = 1.f - dbdev;
      float T = B + eps;
      if (METRIC == LEMON_METRIC_L2) T = 2.f * T - qsq - dbmin;
      const bool certified = kth_key != 0ull && key_val(kth_key) > T;
      if (!certified) {
        const int pos = atomicAdd(n_uncert, 1);
        uncert_rows[pos] = int32_t(row);
      }
    }
  }
}

}  // namespace lemon

extern "C" int lemon_rerank(lemon_ctx* ctx, const float* q, const float* db, const float* cand_val,
                            const int32_t* cand_idx, const float* q_row_stats, const float* db_stats_max,
                            float acc_eps, int64_t nq, int64_t m, int d, int ncand, int nseg, int kp,
                            int metric, float* top_val, int32_t* top_idx, int32_t* uncert_rows,
                            int32_t* n_uncert, void* stream) {
  using namespace lemon;
  if (!ctx) return LEMON_ERR_INVALID;
  if (!q || !db || !cand_val || !cand_idx || !top_val || !top_idx || !uncert_rows || !n_uncert || nq < 0 || d <= 0 ||
      kp < 1 || kp > LEMON_MAX_KP || nseg < 1 || ncand != nseg * LEMON_KPRIME || (q_row_stats && !db_stats_max))
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "rerank: bad args");
  if (nq == 0) return LEMON_OK;
  int64_t blocks = (nq + kRrWarps - 1) / kRrWarps;
  const int64_t cap = int64_t(ctx->num_sms) * 8;
  if (blocks > cap) blocks = cap;
  if (metric == LEMON_METRIC_IP)
    rerank_kernel<LEMON_METRIC_IP><<<unsigned(blocks), kRrWarps * 32, 0, (cudaStream_t)stream>>>(
        q, db, cand_val, cand_idx, q_row_stats, db_stats_max, acc_eps, nq, m, d, ncand, nseg, kp, top_val, top_idx,
        uncert_rows, n_uncert);
  else
    rerank_kernel<LEMON_METRIC_L2><<<unsigned(blocks), kRrWarps * 32, 0, (cudaStream_t)stream>>>(
        q, db, cand_val, cand_idx, q_row_stats, db_stats_max, acc_eps, nq, m, d, ncand, nseg, kp, top_val, top_idx,
        uncert_rows, n_uncert);
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}
