// Discrepancy / diversity baseline scores on the kNN kernels' output.
// Restates lib/baselines/discrepancy_baseline.py:213-230 for given neighbour lists:
//   dis_y / dis_x : second_nns = [l for j in I_m for l in cache[j]]  (cache[j] = kNN of DB row j without j itself)
//                   score = sum(1 - emb[second_nns] @ e_i) / len(second_nns)
//   div_y / div_x : U = 1 - emb[I_m] @ emb[I_m].T ;  score = U.sum() / k**2
// using  sum_l (1 - <e_l, e_i>) = L - <e_i, sum_l e_l>  and  sum_ab (1 - <e_a, e_b>) = n^2 - ||sum_a e_a||^2,
// so one warp per query only accumulates a sum vector over gathered rows (HBM-bound gather).
#include "lemon_common.cuh"

namespace lemon {

constexpr int kDiscMaxV4 = 8;   // float4 per lane: d <= 1024

template <int MODE>   // 0 = dis, 1 = div
__global__ void __launch_bounds__(256)
discrepancy_kernel(const float* __restrict__ emb, const float* __restrict__ qemb, const int32_t* __restrict__ nn,
                   const int32_t* __restrict__ cache, int64_t nq, int64_t m, int d, int kk, int kc, int k,
                   float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  const int d4 = d >> 2;
  for (int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < nq; row += warps) {
    float4 acc[kDiscMaxV4];
#pragma unroll
    for (int i = 0; i < kDiscMaxV4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    int count = 0;
    auto add_row = [&](int idx) {
      const float4* r = reinterpret_cast<const float4*>(emb + int64_t(idx) * d);
#pragma unroll
      for (int i = 0; i < kDiscMaxV4; ++i) {
        const int c = lane + 32 * i;
        if (c < d4) { const float4 y = __ldg(r + c); acc[i].x += y.x; acc[i].y += y.y; acc[i].z += y.z; acc[i].w += y.w; }
      }
      ++count;
    };
    for (int a = 0; a < kk; ++a) {
      const int j = nn[row * kk + a];
      if (j < 0 || int64_t(j) >= m) continue;
      if (MODE == 1) {
        add_row(j);
      } else {
        for (int b = 0; b < kc; ++b) {
          const int l = cache[int64_t(j) * kc + b];
          if (l < 0 || l == j || int64_t(l) >= m) continue;    // cache[i] = [j for j in cache[i] if j != i]  (:167-168)
          add_row(l);
        }
      }
    }
    float part = 0.f;
    if (MODE == 0) {
      const float4* qr = reinterpret_cast<const float4*>(qemb + row * d);
#pragma unroll
      for (int i = 0; i < kDiscMaxV4; ++i) {
        const int c = lane + 32 * i;
        if (c < d4) { const float4 x = qr[c]; part += x.x * acc[i].x + x.y * acc[i].y + x.z * acc[i].z + x.w * acc[i].w; }
      }
    } else {
#pragma unroll
      for (int i = 0; i < kDiscMaxV4; ++i) {
        const int c = lane + 32 * i;
        if (c < d4) part += acc[i].x * acc[i].x + acc[i].y * acc[i].y + acc[i].z * acc[i].z + acc[i].w * acc[i].w;
      }
    }
    part = warp_sum(part);
    if (lane == 0) {
      if (MODE == 0) out[row] = count > 0 ? (float(count) - part) / float(count) : __int_as_float(0x7fc00000);
      else out[row] = (float(count) * float(count) - part) / (float(k) * float(k));
    }
  }
}

}  // namespace lemon

extern "C" int lemon_discrepancy(lemon_ctx* ctx, const float* emb, const float* qemb, const int32_t* nn,
                                 const int32_t* cache, int64_t nq, int64_t m, int d, int kk, int kc, int k, int mode,
                                 float* out, void* stream) {
  using namespace lemon;
  if (!ctx) return LEMON_ERR_INVALID;
  if (!emb || !nn || !out || nq < 0 || m < 1 || d <= 0 || (d & 3) || d > 128 * kDiscMaxV4 || kk < 1 || k < 1 ||
      (mode == 0 && (!cache || !qemb || kc < 1)) || (mode != 0 && mode != 1))
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "discrepancy: bad args (d %% 4 == 0, d <= 1024)");
  if (nq == 0) return LEMON_OK;
  int64_t blocks = (nq + 7) / 8;
  const int64_t cap = int64_t(ctx->num_sms) * 16;
  if (blocks > cap) blocks = cap;
  if (mode == 0)
    discrepancy_kernel<0><<<unsigned(blocks), 256, 0, (cudaStream_t)stream>>>(emb, qemb, nn, cache, nq, m, d, kk, kc, k, out);
  else
    discrepancy_kernel<1><<<unsigned(blocks), 256, 0, (cudaStream_t)stream>>>(emb, qemb, nn, cache, nq, m, d, kk, kc, k, out);
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}
