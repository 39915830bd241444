// K2a: selection over the tensor-core candidate lists, fp32 exact re-rank, per-row exactness certificate.
// One warp per query row; thousands of warps in flight hide the selection latency that would otherwise sit
// on K1's critical path.  HBM-bound gather: (gathered candidates) * d * 4 B per row.
// (North star: "low-precision candidates get an fp32 exact re-rank of a margin-widened set".)
#include "lemon_common.cuh"

namespace lemon {

constexpr int kRrWarps = 8;

// sorts the warp's 256-slot shared buffer (cnt valid keys) descending and keeps the best `keep`
__device__ __forceinline__ void sort_keep(uint64_t* buf, int& cnt, int keep, int lane, uint64_t (&key)[8]) {
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 8; ++i) { const int e = lane * 8 + i; key[i] = e < cnt ? buf[e] : 0ull; }
  warp_sort256_desc(key, lane);
  __syncwarp();
  if (lane < keep / 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) buf[lane * 8 + i] = key[i];
  }
  cnt = min(cnt, keep);
  __syncwarp();
}

template <int METRIC>
__global__ void __launch_bounds__(kRrWarps * 32)
rerank_kernel(const float* __restrict__ q, const float* __restrict__ db, const uint64_t* __restrict__ cand_keys,
              const int32_t* __restrict__ cand_cnt, const float* __restrict__ cand_theta,
              const float* __restrict__ q_row_stats, const float* __restrict__ db_stats_max, float acc_eps, int64_t nq,
              int64_t m, int d, int nlist, int kp, float* __restrict__ top_val, int32_t* __restrict__ top_idx,
              int32_t* __restrict__ uncert_rows, int32_t* __restrict__ n_uncert) {
  __shared__ uint64_t sbuf[kRrWarps][kCap];
  __shared__ uint64_t sbuf2[kRrWarps][kKeep];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* buf = sbuf[warp];
  uint64_t* ebuf = sbuf2[warp];
  const int64_t warps = int64_t(gridDim.x) * kRrWarps;
  for (int64_t row = int64_t(blockIdx.x) * kRrWarps + warp; row < nq; row += warps) {
    const float* qr = q + row * d;
    uint64_t key[8];
    // ---- 1. the row's 64 best approximate candidates over the union of its lists, and the bound B on every
    //         column that is in no list (max of the lists' thresholds)
    int cnt = 0, total = 0;
    float B = -CUDART_INF_F;
    for (int l = 0; l < nlist; ++l) {
      const int c = min(cand_cnt[row * nlist + l], kCap);
      total += c;
      B = fmaxf(B, cand_theta[row * nlist + l]);
      const uint64_t* src = cand_keys + (row * nlist + l) * kCap;
      for (int off = 0; off < c; off += 32) {
        if (cnt > kCap - 32) sort_keep(buf, cnt, kKeep, lane, key);
        const int n = min(32, c - off);
        if (lane < n) buf[cnt + lane] = __ldg(src + off + lane);
        cnt += n;
      }
    }
    sort_keep(buf, cnt, kKeep, lane, key);      // buf[0 .. cnt) = best approximate candidates, descending
    if (total > kKeep) B = fmaxf(B, key_val(buf[kKeep - 1]));   // candidates dropped here are non-candidates too
    // ---- 2. rounding-error bound of this row (include/lemon_b200.h) and the candidate cut-off: the kp best
    //         approximate values certify kp elements with exact value >= a_(kp) - eps, so a candidate whose
    //         approximate value is below a_(kp) - 2 eps cannot be in the exact top-kp and is not gathered
    float eps = 0.f, qsq = 1.f, dbdev = 0.f;
    if (q_row_stats) {
      const float4 st = reinterpret_cast<const float4*>(q_row_stats)[row];   // {||q||, ||q16||, ||q-q16||, ||q||^2}
      eps = st.z * db_stats_max[1] + st.x * db_stats_max[2] + acc_eps;
      qsq = st.w;
      dbdev = db_stats_max[3];
    }
    const float akp = cnt >= kp ? key_val(buf[kp - 1]) : -CUDART_INF_F;
    const float cut = akp - 2.f * eps - (METRIC == LEMON_METRIC_L2 ? dbdev : 0.f);
    // ---- 3. exact fp32 values of the surviving candidates
    int ecnt = 0;
    for (int t = 0; t < cnt; ++t) {
      const uint64_t ck = buf[t];                                   // warp-uniform (shared memory broadcast)
      if (key_val(ck) < cut) break;                                 // sorted: everything after is below the cut too
      const int idx = key_idx(ck);
      if (idx < 0 || int64_t(idx) >= m) continue;
      float v = warp_pair_value<METRIC>(qr, db + int64_t(idx) * d, d, lane);
      if (METRIC == LEMON_METRIC_L2) v = -v;
      if (lane == 0) ebuf[ecnt] = make_key(v, uint32_t(idx));
      ecnt++;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int e = lane * 8 + i; key[i] = e < ecnt ? ebuf[e] : 0ull; }
    warp_sort256_desc(key, lane);
    // ---- 4. emit the exact top list
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
    // ---- 5. certificate: every column outside the lists has approximate ip <= B, hence exact ip <= B + eps
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
    __syncwarp();
  }
}

}  // namespace lemon

extern "C" int lemon_rerank(lemon_ctx* ctx, const float* q, const float* db, const uint64_t* cand_keys,
                            const int32_t* cand_cnt, const float* cand_theta, const float* q_row_stats,
                            const float* db_stats_max, float acc_eps, int64_t nq, int64_t m, int d, int nlist, int kp,
                            int metric, float* top_val, int32_t* top_idx, int32_t* uncert_rows, int32_t* n_uncert,
                            void* stream) {
  using namespace lemon;
  if (!ctx) return LEMON_ERR_INVALID;
  if (!q || !db || !cand_keys || !cand_cnt || !cand_theta || !top_val || !top_idx || !uncert_rows || !n_uncert || nq < 0 ||
      d <= 0 || kp < 1 || kp > LEMON_MAX_KP || nlist < 1 || (q_row_stats && !db_stats_max))
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "rerank: bad args");
  if (nq == 0) return LEMON_OK;
  int64_t blocks = (nq + kRrWarps - 1) / kRrWarps;
  const int64_t cap = int64_t(ctx->num_sms) * 8;
  if (blocks > cap) blocks = cap;
  if (metric == LEMON_METRIC_IP)
    rerank_kernel<LEMON_METRIC_IP><<<unsigned(blocks), kRrWarps * 32, 0, (cudaStream_t)stream>>>(
        q, db, cand_keys, cand_cnt, cand_theta, q_row_stats, db_stats_max, acc_eps, nq, m, d, nlist, kp, top_val, top_idx,
        uncert_rows, n_uncert);
  else
    rerank_kernel<LEMON_METRIC_L2><<<unsigned(blocks), kRrWarps * 32, 0, (cudaStream_t)stream>>>(
        q, db, cand_keys, cand_cnt, cand_theta, q_row_stats, db_stats_max, acc_eps, nq, m, d, nlist, kp, top_val, top_idx,
        uncert_rows, n_uncert);
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}
