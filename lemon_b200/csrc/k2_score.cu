// K2b: per-sample records + score.  One warp per query row; gathers the k neighbours'
// cross-modal rows (coalesced float4, 2*k*d*4 B per row: HBM-bound), applies the reference's
// self-exclusion rule, the exp-decay weighting and the final reduction.
// Restates run_lemon.py:250-307 and lib/metrics/utils.py:63-77.
#include "lemon_common.cuh"

namespace lemon {

constexpr int kScWarps = 8;

struct ScoreHp { double beta, gamma, t1n, t2n, t1m, t2m; int has; };

template <int METRIC>
__global__ void __launch_bounds__(kScWarps * 32, 3)
score_kernel(const float* __restrict__ xq, const float* __restrict__ yq, const float* __restrict__ xdb,
             const float* __restrict__ ydb, const float* __restrict__ dists_tr, const float* __restrict__ topn_val,
             const int32_t* __restrict__ topn_idx, const float* __restrict__ topm_val,
             const int32_t* __restrict__ topm_idx, const int64_t* __restrict__ query_in_db,
             const int32_t* __restrict__ label_q, const int32_t* __restrict__ label_db,
             const float* __restrict__ class_emb, const int32_t* __restrict__ noisy_label, int n_class, int64_t nq,
             int64_t m, int d, int k, int kp, ScoreHp hp, float* __restrict__ d1, float* __restrict__ Dn, float* __restrict__ dists_n,
             float* __restrict__ dists_tr_n, float* __restrict__ Dm, float* __restrict__ dists_m,
             float* __restrict__ dists_tr_m, int64_t* __restrict__ In, int64_t* __restrict__ Im, int idx32, int sides,
             double* __restrict__ sn, double* __restrict__ sm, double* __restrict__ score) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t warps = int64_t(gridDim.x) * kScWarps;
  const bool discrete = label_q != nullptr;
  const float nanv = __int_as_float(0x7fc00000);
  for (int64_t row = int64_t(blockIdx.x) * kScWarps + warp; row < nq; row += warps) {
    const float* xr = xq + row * d;
    const float* yr = yq + row * d;
    // d_1 (run_lemon.py:250-253)
    float d1v;
    if (class_emb == nullptr) {
      const float pv = warp_pair_value<METRIC>(xr, yr, d, lane);
      d1v = (METRIC == LEMON_METRIC_IP) ? 1.0f - pv : pv;
    } else {
      // --normalize_d1 (run_lemon.py:244-248): softmax over the class prompts of the image-to-prompt
      // distances, read at the sample's noisy label (max-subtracted like scipy.special.softmax)
      const int lab = noisy_label[row];
      float mxv = -CUDART_INF_F, dl = 0.f;
      for (int c = 0; c < n_class; ++c) {
        const float pv = warp_pair_value<METRIC>(xr, class_emb + int64_t(c) * d, d, lane);
        const float dc = (METRIC == LEMON_METRIC_IP) ? 1.0f - pv : pv;
        mxv = fmaxf(mxv, dc);
        if (c == lab) dl = dc;
      }
      float den = 0.f;
      for (int c = 0; c < n_class; ++c) {
        const float pv = warp_pair_value<METRIC>(xr, class_emb + int64_t(c) * d, d, lane);
        const float dc = (METRIC == LEMON_METRIC_IP) ? 1.0f - pv : pv;
        den += expf(dc - mxv);
      }
      d1v = (lab >= 0 && lab < n_class) ? expf(dl - mxv) / den : __int_as_float(0x7fc00000);
    }
    // self-exclusion (run_lemon.py:257-263, 277-283): kp == k+1 -> drop rank 0 if in DB else the last
    const int off = (query_in_db != nullptr && query_in_db[row] >= 0) ? 1 : 0;
    double acc_n = 0.0, acc_m = 0.0;
    // lane j (and j+32) keeps neighbour j's record
    for (int side = 0; side < 2; ++side) {
      if (!((sides >> side) & 1)) continue;     // image-neighbour side = bit 0, text-neighbour side = bit 1
      const float* tv = (side == 0 ? topn_val : topm_val) + row * kp + off;
      const int32_t* ti = (side == 0 ? topn_idx : topm_idx) + row * kp + off;
      const float* other_q = side == 0 ? yr : xr;        // image neighbours -> compare TEXT rows; text neighbours -> IMAGE rows
      const float* other_db = side == 0 ? ydb : xdb;
      float rD[2], rdist[2], rdtr[2];
      int ridx[2];
      for (int j0 = 0; j0 < k; j0 += 4) {
        int idx4[4];
        const float* bp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = min(j0 + u, k - 1);
          idx4[u] = ti[j];
          const bool ok = idx4[u] >= 0 && int64_t(idx4[u]) < m;
          bp[u] = other_db + int64_t(ok ? idx4[u] : 0) * d;
        }
        float v4[4] = {0.f, 0.f, 0.f, 0.f};
        if (!(side == 0 && discrete)) warp_pair_value4<METRIC>(other_q, bp[0], bp[1], bp[2], bp[3], d, lane, v4);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + u;
          if (j >= k) break;
          const int idx = idx4[u];
          float dist, Dv, dtr;
          if (idx < 0 || int64_t(idx) >= m) {
            dist = nanv; Dv = nanv; dtr = nanv;
          } else {
            if (side == 0 && discrete) {
              dist = 1.0f - float(label_db[idx] == label_q[row]);       // run_lemon.py:266-267
            } else {
              dist = (METRIC == LEMON_METRIC_IP) ? 1.0f - v4[u] : v4[u];  // :271,273,287,289
            }
            Dv = tv[j];
            // cosine: D = -<a,b> (:270,286); the negation is skipped for image neighbours under the
            // discrete text metric because it sits in the else-branch (:266-270)
            if (METRIC == LEMON_METRIC_IP && !(side == 0 && discrete)) Dv = -Dv;
            dtr = dists_tr[idx];
          }
          if (lane == (j & 31)) {
            if (j < 32) { rD[0] = Dv; rdist[0] = dist; rdtr[0] = dtr; ridx[0] = idx; }
            else        { rD[1] = Dv; rdist[1] = dist; rdtr[1] = dtr; ridx[1] = idx; }
          }
        }
      }
      double part = 0.0;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int j = lane + 32 * s;
        if (j < k) {
          const int64_t o = row * k + j;
          if (side == 0) {
            if (Dn) Dn[o] = rD[s]; if (dists_n) dists_n[o] = rdist[s]; if (dists_tr_n) dists_tr_n[o] = rdtr[s];
            if (In) { if (idx32) reinterpret_cast<int32_t*>(In)[o] = ridx[s]; else In[o] = ridx[s]; }
          } else {
            if (Dm) Dm[o] = rD[s]; if (dists_m) dists_m[o] = rdist[s]; if (dists_tr_m) dists_tr_m[o] = rdtr[s];
            if (Im) { if (idx32) reinterpret_cast<int32_t*>(Im)[o] = ridx[s]; else Im[o] = ridx[s]; }
          }
          if (hp.has) {   // utils.py:71-75
            const double t1 = side == 0 ? hp.t1n : hp.t1m, t2 = side == 0 ? hp.t2n : hp.t2m;
            part += exp(-t1 * double(rD[s])) * exp(-t2 * double(rdtr[s])) * double(rdist[s]);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(kFull, part, o);
      if (side == 0) acc_n = part / double(k); else acc_m = part / double(k);
    }
    if (lane == 0) {
      // a call that only does the image-neighbour side (sides == 1) leaves d_1 and the score to the call that does the
      // text side; that call reads s_n back (the two calls are stream-ordered)
      if (d1 && (sides & 2)) d1[row] = d1v;
      if (hp.has) {
        if (sn && (sides & 1)) sn[row] = acc_n;
        if (sm && (sides & 2)) sm[row] = acc_m;
        if (score && (sides & 2)) {
          const double s_n = (sides & 1) ? acc_n : sn[row];
          score[row] = double(d1v) + hp.beta * s_n + hp.gamma * acc_m;   // utils.py:77
        }
      }
    }
  }
}

__global__ void combine_scores_kernel(const float* __restrict__ Dn, const float* __restrict__ dtn,
                                      const float* __restrict__ dn, const float* __restrict__ Dm,
                                      const float* __restrict__ dtm, const float* __restrict__ dm,
                                      const double* __restrict__ d1, int64_t n, int k, ScoreHp hp,
                                      double* __restrict__ sn, double* __restrict__ sm, double* __restrict__ score) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
    double pn = 0.0, pm = 0.0;
    for (int j = lane; j < k; j += 32) {
      const int64_t o = row * k + j;
      pn += exp(-hp.t1n * double(Dn[o])) * exp(-hp.t2n * double(dtn[o])) * double(dn[o]);
      pm += exp(-hp.t1m * double(Dm[o])) * exp(-hp.t2m * double(dtm[o])) * double(dm[o]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      pn += __shfl_xor_sync(kFull, pn, o);
      pm += __shfl_xor_sync(kFull, pm, o);
    }
    if (lane == 0) {
      pn /= double(k); pm /= double(k);
      if (sn) sn[row] = pn;
      if (sm) sm[row] = pm;
      score[row] = d1[row] + hp.beta * pn + hp.gamma * pm;
    }
  }
}

static ScoreHp make_hp(const double* hp) {
  ScoreHp h{};
  if (hp) { h.beta = hp[0]; h.gamma = hp[1]; h.t1n = hp[2]; h.t2n = hp[3]; h.t1m = hp[4]; h.t2m = hp[5]; h.has = 1; }
  return h;
}

}  // namespace lemon

extern "C" int lemon_score(lemon_ctx* ctx, const float* xq, const float* yq, const float* xdb, const float* ydb,
                           const float* dists_tr, const float* topn_val, const int32_t* topn_idx,
                           const float* topm_val, const int32_t* topm_idx, const int64_t* query_in_db,
                           const int32_t* label_q, const int32_t* label_db, const float* class_emb,
                           const int32_t* noisy_label, int n_class, int64_t nq, int64_t m, int d, int k,
                           int kp, int metric, const double* hp, float* d1, float* Dn, float* dists_n,
                           float* dists_tr_n, float* Dm, float* dists_m, float* dists_tr_m, void* In,
                           void* Im, int index_bits, int sides, double* sn, double* sm, double* score, void* stream) {
  using namespace lemon;
  if (!ctx) return LEMON_ERR_INVALID;
  if (!xq || !yq || !xdb || !ydb || !dists_tr || ((sides & 1) && (!topn_val || !topn_idx)) ||
      ((sides & 2) && (!topm_val || !topm_idx)) || nq < 0 || d <= 0 ||
      k < 1 || k > 64 || (kp != k && kp != k + 1) || (query_in_db && kp != k + 1) || (!query_in_db && kp != k) ||
      ((label_q == nullptr) != (label_db == nullptr)) || ((class_emb == nullptr) != (noisy_label == nullptr)) ||
      (class_emb && n_class < 1) || (index_bits != 64 && index_bits != 32) || sides < 1 || sides > 3 ||
      (sides == 2 && hp && !sn))
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "score: bad args (kp must be k+1 with query_in_db, k without; index_bits 32 or 64)");
  if (nq == 0) return LEMON_OK;
  int64_t blocks = (nq + kScWarps - 1) / kScWarps;
  const int64_t cap = int64_t(ctx->num_sms) * 6;   // 3 resident blocks per SM (<= 80 registers), two waves
  if (blocks > cap) blocks = cap;
  const ScoreHp h = make_hp(hp);
  if (metric == LEMON_METRIC_IP)
    score_kernel<LEMON_METRIC_IP><<<unsigned(blocks), kScWarps * 32, 0, (cudaStream_t)stream>>>(
        xq, yq, xdb, ydb, dists_tr, topn_val, topn_idx, topm_val, topm_idx, query_in_db, label_q, label_db, class_emb, noisy_label, n_class,
        nq, m, d, k, kp, h, d1, Dn, dists_n, dists_tr_n, Dm, dists_m, dists_tr_m, (int64_t*)In, (int64_t*)Im, index_bits == 32, sides, sn, sm, score);
  else
    score_kernel<LEMON_METRIC_L2><<<unsigned(blocks), kScWarps * 32, 0, (cudaStream_t)stream>>>(
        xq, yq, xdb, ydb, dists_tr, topn_val, topn_idx, topm_val, topm_idx, query_in_db, label_q, label_db, class_emb, noisy_label, n_class,
        nq, m, d, k, kp, h, d1, Dn, dists_n, dists_tr_n, Dm, dists_m, dists_tr_m, (int64_t*)In, (int64_t*)Im, index_bits == 32, sides, sn, sm, score);
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}

extern "C" int lemon_combine_scores(lemon_ctx* ctx, const float* Dn, const float* dists_tr_n, const float* dists_n,
                                    const float* Dm, const float* dists_tr_m, const float* dists_m, const double* d1,
                                    int64_t n, int k, const double* hp, double* sn, double* sm, double* score,
                                    void* stream) {
  using namespace lemon;
  if (!ctx) return LEMON_ERR_INVALID;
  if (!Dn || !dists_tr_n || !dists_n || !Dm || !dists_tr_m || !dists_m || !d1 || !hp || !score || n < 0 || k < 1)
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "combine_scores: bad args");
  if (n == 0) return LEMON_OK;
  int64_t blocks = (n + 7) / 8;
  const int64_t cap = int64_t(ctx->num_sms) * 16;
  if (blocks > cap) blocks = cap;
  combine_scores_kernel<<<unsigned(blocks), 256, 0, (cudaStream_t)stream>>>(Dn, dists_tr_n, dists_n, Dm, dists_tr_m,
                                                                          dists_m, d1, n, k, make_hp(hp), sn, sm, score);
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}
