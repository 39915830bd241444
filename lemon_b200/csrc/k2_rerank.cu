// K2a: selection over the tensor-core candidate lists, fp32 exact re-rank, per-row exactness certificate.
// One warp per query row; thousands of warps in flight hide the selection latency that would otherwise sit
// on K1's critical path.  HBM-bound gather: (gathered candidates) * d * 4 B per row.
// (North star: "low-precision candidates get an fp32 exact re-rank of a margin-widened set".)
#include "lemon_common.cuh"

namespace lemon {

constexpr int kRrWarps = 8;
constexpr int kSelCap = 512;             // selection buffer (keys) per warp; a chunk adds at most kRrChunk = 256
                                         // (6 KB of shared memory per warp with the exact-key buffer: 32 warps per SM)

// loads one chunk = 256 slots of a candidate list (8 keys per lane); invalid slots become 0
constexpr int kRrChunk = 256;
constexpr int kRrPerLane = kRrChunk / 32;
constexpr int kRrChunksPerList = kListCap / kRrChunk;
static_assert(kListCap % kRrChunk == 0, "candidate lists are read in 256-key chunks");
// cnt_lo / cnt_hi: the lengths of the row's lists, list l in lane l & 31 (loaded once per row: the key loads of a chunk
// then do not wait for a dependent length load first)
__device__ __forceinline__ void load_chunk(const uint64_t* __restrict__ cand_keys, int cnt_lo, int cnt_hi,
                                           int64_t row, int nlist, int chunk, int lane, uint64_t (&k)[kRrPerLane], int& tot) {
  const int l = chunk / kRrChunksPerList, part = chunk % kRrChunksPerList;
  const int len = __shfl_sync(kFull, l < 32 ? cnt_lo : cnt_hi, l & 31);
  const int c = max(0, min(min(len, kListCap) - part * kRrChunk, kRrChunk));
  tot = c;
  if (c == 0) return;
  const uint64_t* src = cand_keys + (row * nlist + l) * kListCap + part * kRrChunk;
#pragma unroll
  for (int i = 0; i < kRrPerLane; ++i) {
    const int e = i * 32 + lane;                      // coalesced: consecutive lanes read consecutive keys
    // K1 stores the RAW float bits in the high word (its append runs for all 32 lanes of a warp); everything below
    // compares keys, so the order-preserving pattern is applied here
    const uint64_t raw = e < c ? __ldg(src + e) : 0ull;
    k[i] = raw == 0ull ? 0ull : ((uint64_t(f2ord(__uint_as_float(uint32_t(raw >> 32)))) << 32) | (raw & 0xffffffffull));
  }
}

template <int METRIC>
__global__ void __launch_bounds__(kRrWarps * 32, 4)
rerank_kernel(const float* __restrict__ q, const float* __restrict__ db, const uint64_t* __restrict__ cand_keys,
              const int32_t* __restrict__ cand_cnt, const float* __restrict__ cand_theta,
              const float* __restrict__ q_row_stats, const float* __restrict__ db_stats_max, float acc_eps, int64_t nq,
              int64_t m, int d, int nlist, int kp, const int32_t* __restrict__ out_rows, float* __restrict__ top_val,
              int32_t* __restrict__ top_idx, int32_t* __restrict__ uncert_rows, int32_t* __restrict__ n_uncert) {
  // per warp: selection buffer (kSelCap keys: all candidates above the bound that may still be in the top-kp) and
  // the exact keys of the gathered candidates (kCap)
  extern __shared__ __align__(16) unsigned char rr_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t* cbuf = reinterpret_cast<uint64_t*>(rr_smem) + size_t(warp) * (kSelCap + kCap);
  uint64_t* ebuf = cbuf + kSelCap;
  const int64_t warps = int64_t(gridDim.x) * kRrWarps;
  const int nchunk = nlist * kRrChunksPerList;
  for (int64_t row = int64_t(blockIdx.x) * kRrWarps + warp; row < nq; row += warps) {
    const float* qr = q + row * d;
    // rounding-error bound of this row (include/lemon_b200.h)
    float eps = 0.f, qsq = 1.f, dbdev = 0.f;
    if (q_row_stats) {
      const float4 st = reinterpret_cast<const float4*>(q_row_stats)[row];   // {||q||, ||q16||, ||q-q16||, ||q||^2}
      // rounding of the two operands (Cauchy-Schwarz) + accumulation, the latter relative to the norms (acc_eps is
      // a coefficient: d-dependent, see lemon_b200.h) so that un-normalised rows stay rigorously bounded
      eps = st.z * db_stats_max[1] + st.x * db_stats_max[2] +
            acc_eps * fmaxf(st.x, st.y) * fmaxf(db_stats_max[0], db_stats_max[1]);
      qsq = st.w;
      dbdev = db_stats_max[3];
    }
    // ---- 1. B = bound on every column that is in no list
    float B = -CUDART_INF_F;
    for (int l = lane; l < nlist; l += 32) B = fmaxf(B, cand_theta[row * nlist + l]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) B = fmaxf(B, __shfl_xor_sync(kFull, B, o));
    // ---- 2. streaming selection over the UNION of the row's lists.  A certified top-kp has exact values > B + eps,
    //         hence approximate values > B: only keys above B are collected.  Whenever the buffer passes half its
    //         capacity, `sel` = (a lower bound of) the kp-th best key collected so far is found by bisection on the
    //         ordered bit pattern and everything below it is dropped; at the end `sel` is a lower bound of the kp-th
    //         best approximate value of the whole row however many lists (segments) there are.
    //         Keys within `margin` (twice the rounding bound) below `sel` stay: their exact value may still be in the top-kp.
    const float margin = 2.f * eps + (METRIC == LEMON_METRIC_L2 ? dbdev : 0.f);
    const uint32_t above_B = f2ord(B) + 1u;           // keys must be strictly above B (f2ord(-inf) + 1 is still tiny)
    uint32_t sel = 0u, keep_bits = above_B;
    int n = 0;
    bool overflow = false;
    auto compact_ge = [&](uint32_t lo_bits, float minval) {      // keep keys with ordered value >= lo_bits and value >= minval
      int w = 0;
      for (int s0 = 0; s0 < n; s0 += 32) {
        const int e = s0 + lane;
        const uint64_t key = e < n ? cbuf[e] : 0ull;
        const bool keep = key != 0ull && uint32_t(key >> 32) >= lo_bits && key_val(key) >= minval;
        const unsigned mask = __ballot_sync(kFull, keep);
        __syncwarp();
        if (keep) cbuf[w + __popc(mask & ((1u << lane) - 1u))] = key;     // w + rank <= e: in-place is safe
        w += __popc(mask);
      }
      __syncwarp();
      n = w;
    };
    auto kth_lower_bound = [&]() -> uint32_t {                  // a value with at least kp collected keys >= it (n >= kp)
      uint32_t lo = 0xffffffffu, hi = 0u;
      for (int e = lane; e < n; e += 32) { const uint32_t o = uint32_t(cbuf[e] >> 32); lo = min(lo, o); hi = max(hi, o); }
      lo = __reduce_min_sync(kFull, lo);
      hi = __reduce_max_sync(kFull, hi);              // count(>= lo) = n >= kp ; count(>= hi + 1) = 0
      for (int it = 0; it < 14 && hi > lo; ++it) {
        const uint32_t mid = lo + ((hi - lo + 1u) >> 1);
        int cge = 0;
        for (int e = lane; e < n; e += 32) cge += uint32_t(cbuf[e] >> 32) >= mid;
        cge = __reduce_add_sync(kFull, cge);
        if (cge >= kp) lo = mid; else hi = mid - 1u;
      }
      return lo;
    };
    uint64_t k[kRrPerLane];
    int tot;
    const int cnt_lo = lane < nlist ? cand_cnt[row * nlist + lane] : 0;
    const int cnt_hi = lane + 32 < nlist ? cand_cnt[row * nlist + lane + 32] : 0;
    for (int c = 0; c < nchunk; ++c) {
      load_chunk(cand_keys, cnt_lo, cnt_hi, row, nlist, c, lane, k, tot);
      if (tot == 0) continue;
#pragma unroll
      for (int i = 0; i < kRrPerLane; ++i) {
        const bool pred = k[i] != 0ull && uint32_t(k[i] >> 32) >= keep_bits && uint32_t(key_idx(k[i])) < uint32_t(m);
        const unsigned mask = __ballot_sync(kFull, pred);
        const int pos = n + __popc(mask & ((1u << lane) - 1u));
        if (pred && pos < kSelCap) cbuf[pos] = k[i];
        n += __popc(mask);
      }
      if (n > kSelCap) { overflow = true; n = kSelCap; }
      __syncwarp();
      if (n > kSelCap / 2 && n >= kp) {               // room for the next chunk (<= kRrChunk keys)
        sel = max(sel, kth_lower_bound());
        keep_bits = max(above_B, f2ord(ord2f(sel) - margin));
        compact_ge(keep_bits, -CUDART_INF_F);
        if (n > kSelCap / 2) overflow = true;         // mass ties around the kp-th value: the exact kernel takes the row
      }
    }
    __syncwarp();
    const uint32_t lb = n >= kp ? max(sel, kth_lower_bound()) : 0u;
    // the kp candidates above lb certify kp elements with exact value >= lb - eps, so a candidate whose
    // approximate value is below lb - 2 eps cannot be in the exact top-kp and is not gathered
    const float cut = lb ? ord2f(lb) - margin : -CUDART_INF_F;
    compact_ge(0u, cut);
    int ccnt = n;
    if (ccnt > kCap) { overflow = true; ccnt = kCap; }
    __syncwarp();
    const int ecnt = ccnt;
    for (int t = 0; t < ccnt; t += 4) {
      const float* bp[4];
      int idx4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        idx4[j] = key_idx(cbuf[min(t + j, ccnt - 1)]);          // warp-uniform (shared memory broadcast)
        bp[j] = db + int64_t(idx4[j]) * d;
      }
      float v4[4];
      warp_pair_value4<METRIC>(qr, bp[0], bp[1], bp[2], bp[3], d, lane, v4);
      if (lane < 4 && t + lane < ccnt) {
        float v = lane == 0 ? v4[0] : (lane == 1 ? v4[1] : (lane == 2 ? v4[2] : v4[3]));
        const int id = lane == 0 ? idx4[0] : (lane == 1 ? idx4[1] : (lane == 2 ? idx4[2] : idx4[3]));
        if (METRIC == LEMON_METRIC_L2) v = -v;
        ebuf[t + lane] = make_key(v, uint32_t(id));
      }
    }
    __syncwarp();
    uint64_t key[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int e = lane * 8 + i; key[i] = e < ecnt ? ebuf[e] : 0ull; }
    warp_sort256_desc(key, lane);
    // ---- 3. emit the exact top list (second-pass calls: query row r is row out_rows[r] of the caller's lists)
    const int64_t orow = out_rows ? int64_t(out_rows[row]) : row;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = lane * 8 + i;
      if (e < kp) {
        const bool ok = key[i] != 0ull;
        float v = ok ? key_val(key[i]) : -CUDART_INF_F;
        if (METRIC == LEMON_METRIC_L2) v = -v;
        top_val[orow * kp + e] = v;
        top_idx[orow * kp + e] = ok ? key_idx(key[i]) : -1;
      }
    }
    // ---- 4. certificate: every column outside the lists has approximate ip <= B, hence exact ip <= B + eps
    uint64_t kth_sel = 0ull;
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i == ((kp - 1) & 7)) kth_sel = key[i];
    const uint64_t kth_key = shfl_u64(kth_sel, (kp - 1) >> 3);
    if (lane == 0 && (B > -CUDART_INF_F || overflow)) {
      const float dbmin = 1.f - dbdev;
      float T = B + eps;
      if (METRIC == LEMON_METRIC_L2) T = 2.f * T - qsq - dbmin;
      const bool certified = !overflow && kth_key != 0ull && key_val(kth_key) > T;
      if (!certified) {
        const int pos = atomicAdd(n_uncert, 1);
        uncert_rows[pos] = int32_t(orow);
      }
    }
    __syncwarp();
  }
}

}  // namespace lemon

extern "C" int lemon_rerank(lemon_ctx* ctx, const float* q, const float* db, const uint64_t* cand_keys,
                            const int32_t* cand_cnt, const float* cand_theta, const float* q_row_stats,
                            const float* db_stats_max, float acc_eps, int64_t nq, int64_t m, int d, int nlist, int kp,
                            int metric, const int32_t* out_rows, float* top_val, int32_t* top_idx, int32_t* uncert_rows,
                            int32_t* n_uncert, void* stream) {
  using namespace lemon;
  if (!ctx) return LEMON_ERR_INVALID;
  if (!q || !db || !cand_keys || !cand_cnt || !cand_theta || !top_val || !top_idx || !uncert_rows || !n_uncert || nq < 0 ||
      d <= 0 || kp < 1 || kp > LEMON_MAX_KP || nlist < 1 || nlist > 64 || (q_row_stats && !db_stats_max))
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "rerank: bad args");
  LEMON_CUDA_CHECK(ctx, cudaMemsetAsync(n_uncert, 0, sizeof(int32_t), (cudaStream_t)stream));
  if (nq == 0) return LEMON_OK;
  int64_t blocks = (nq + kRrWarps - 1) / kRrWarps;
  const int64_t cap = int64_t(ctx->num_sms) * 8;      // two waves of four resident blocks
  if (blocks > cap) blocks = cap;
  const size_t smem = size_t(kRrWarps) * (kSelCap + kCap) * sizeof(uint64_t);     // 48 KB: four blocks per SM
  LEMON_CUDA_CHECK(ctx, cudaFuncSetAttribute(rerank_kernel<LEMON_METRIC_IP>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  LEMON_CUDA_CHECK(ctx, cudaFuncSetAttribute(rerank_kernel<LEMON_METRIC_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  if (metric == LEMON_METRIC_IP)
    rerank_kernel<LEMON_METRIC_IP><<<unsigned(blocks), kRrWarps * 32, smem, (cudaStream_t)stream>>>(
        q, db, cand_keys, cand_cnt, cand_theta, q_row_stats, db_stats_max, acc_eps, nq, m, d, nlist, kp, out_rows, top_val, top_idx,
        uncert_rows, n_uncert);
  else
    rerank_kernel<LEMON_METRIC_L2><<<unsigned(blocks), kRrWarps * 32, smem, (cudaStream_t)stream>>>(
        q, db, cand_keys, cand_cnt, cand_theta, q_row_stats, db_stats_max, acc_eps, nq, m, d, nlist, kp, out_rows, top_val, top_idx,
        uncert_rows, n_uncert);
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}
