// K0: row-wise L2 normalisation + fp16 operand copy + rounding-error statistics, and the
// row-wise cross-modal distance.  HBM-bound: d*4 B read, d*4 + d16*2 B written per row.
// Restates lib/utils/utils.py:39-40 (F.normalize p=2 dim=1 eps=1e-12) and
// run_lemon.py:169,173,250-253 (dists_tr, d_1).
#include "lemon_common.cuh"

namespace lemon {

__device__ __forceinline__ void atomic_max_pos(float* addr, float v) {
  // non-negative floats order like their bit patterns
  atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}

__global__ void __launch_bounds__(256)
normalize_cast_kernel(const float* __restrict__ in, float* __restrict__ out_f32, __half* __restrict__ out_f16,
                      float* __restrict__ row_stats, float* __restrict__ stats_max, int64_t n, int d, int d16,
                      int64_t in_stride, int do_normalize) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
  for (int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
    const float* x = in + row * in_stride;
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) { float v = x[c]; ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float denom = do_normalize ? fmaxf(sqrtf(ss), 1e-12f) : 1.0f;
    float sy = 0.f, sh = 0.f, se = 0.f;
    for (int c = lane; c < d16; c += 32) {
      float y = 0.f;
      if (c < d) {
        y = x[c] / denom;
        if (out_f32) out_f32[row * d + c] = y;
      }
      const __half h = __float2half_rn(y);
      const float hf = __half2float(h);
      const float e = y - hf;
      sy = fmaf(y, y, sy); sh = fmaf(hf, hf, sh); se = fmaf(e, e, se);
      if (out_f16) out_f16[row * d16 + c] = h;
    }
    sy = warp_sum(sy); sh = warp_sum(sh); se = warp_sum(se);
    const float ny = sqrtf(sy), nh = sqrtf(sh), ne = sqrtf(se);
    if (lane == 0 && row_stats) {
      reinterpret_cast<float4*>(row_stats)[row] = make_float4(ny, nh, ne, sy);
    }
    m0 = fmaxf(m0, ny); m1 = fmaxf(m1, nh); m2 = fmaxf(m2, ne); m3 = fmaxf(m3, fabsf(sy - 1.0f));
  }
  if (stats_max && lane == 0) {
    atomic_max_pos(stats_max + 0, m0); atomic_max_pos(stats_max + 1, m1);
    atomic_max_pos(stats_max + 2, m2); atomic_max_pos(stats_max + 3, m3);
  }
}

// Split-precision operands for the second tensor-core pass (rows the first pass could not certify).
//   x = x_hi + x_lo + x_e,  x_hi = fp16(x),  x_lo = fp16(x - x_hi)   (x_lo may be an fp16 subnormal: its absolute
//   rounding error is <= 2^-25 per component, so ||x_e|| ~ 2^-22 ||x|| instead of the 2^-11 ||x|| of one fp16 word)
// Queries are laid out [hi | lo | hi], database rows [lo | hi | hi] (each block d16 wide), so that ONE K-loop of the
// unchanged K1 kernel over 3*d16 columns accumulates  q_hi.b_lo + q_lo.b_hi + q_hi.b_hi  in fp32 -- the two small
// cross terms FIRST, while the partial sums are ~2^-10 of the result, so that only the last d16 additions round at
// full magnitude and the accumulation bound stays that of a single d16-long product.
// row_stats / stats_max are written in the layout lemon_rerank reads, with the residual ||x_e|| in the place of the
// one-word rounding error:  {||x||, ||x||, ||x_e||, ||x||^2}  and maxima  {||x||, ||x||, ||x_e||, | ||x||^2 - 1 |}.
__global__ void __launch_bounds__(256)
split_cast_kernel(const float* __restrict__ x, __half* __restrict__ out, float* __restrict__ row_stats,
                  float* __restrict__ stats_max, int64_t n, int d, int d16, int role) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float m0 = 0.f, m2 = 0.f, m3 = 0.f;
  for (int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
    const float* xr = x + row * d;
    __half* o = out + row * (3 * int64_t(d16));
    float sx = 0.f, se = 0.f;
    for (int c = lane; c < d16; c += 32) {
      const float v = c < d ? xr[c] : 0.f;
      const __half hi = __float2half_rn(v);
      const float r1 = v - __half2float(hi);            // exact in fp32
      const __half lo = __float2half_rn(r1);
      const float e = r1 - __half2float(lo);            // exact in fp32
      sx = fmaf(v, v, sx); se = fmaf(e, e, se);
      o[c] = role == 0 ? hi : lo;
      o[d16 + c] = role == 0 ? lo : hi;
      o[2 * d16 + c] = hi;
    }
    sx = warp_sum(sx); se = warp_sum(se);
    const float nx = sqrtf(sx), ne = sqrtf(se);
    if (lane == 0 && row_stats) reinterpret_cast<float4*>(row_stats)[row] = make_float4(nx, nx, ne, sx);
    m0 = fmaxf(m0, nx); m2 = fmaxf(m2, ne); m3 = fmaxf(m3, fabsf(sx - 1.0f));
  }
  if (stats_max && lane == 0) {
    atomic_max_pos(stats_max + 0, m0); atomic_max_pos(stats_max + 1, m0);
    atomic_max_pos(stats_max + 2, m2); atomic_max_pos(stats_max + 3, m3);
  }
}

template <int METRIC>
__global__ void __launch_bounds__(256)
rowwise_dist_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                    int64_t n, int d) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
    float v = warp_pair_value<METRIC>(a + row * d, b + row * d, d, lane);
    if (lane == 0) out[row] = (METRIC == LEMON_METRIC_IP) ? 1.0f - v : v;
  }
}

}  // namespace lemon

extern "C" int lemon_normalize_cast(lemon_ctx* ctx, const float* in, float* out_f32, void* out_f16,
                                    float* row_stats, float* stats_max, int64_t n, int d, int d16,
                                    int64_t in_stride, int do_normalize, void* stream) {
  if (!ctx) return LEMON_ERR_INVALID;
  if (!in || n < 0 || d <= 0) return lemon_set_error(ctx, LEMON_ERR_INVALID, "normalize_cast: bad args");
  if (out_f16 && (d16 < d || d16 % 64)) return lemon_set_error(ctx, LEMON_ERR_INVALID, "normalize_cast: d16 must be a multiple of 64 and >= d");
  if (in_stride == 0) in_stride = d;
  if (in_stride < d || (in_stride != d && out_f32 == in))
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "normalize_cast: in_stride must be >= d (and out_f32 must not alias a strided input)");
  if (!out_f16) d16 = d;
  if (n == 0) return LEMON_OK;
  if (stats_max) LEMON_CUDA_CHECK(ctx, cudaMemsetAsync(stats_max, 0, 4 * sizeof(float), (cudaStream_t)stream));
  const int threads = 256;
  int64_t blocks = (n + 7) / 8;
  const int64_t cap = int64_t(ctx->num_sms) * 16;
  if (blocks > cap) blocks = cap;
  lemon::normalize_cast_kernel<<<unsigned(blocks), threads, 0, (cudaStream_t)stream>>>(
      in, out_f32, (__half*)out_f16, row_stats, stats_max, n, d, d16, in_stride, do_normalize);
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}

extern "C" int lemon_rowwise_dist(lemon_ctx* ctx, const float* a, const float* b, float* out, int64_t n,
                                  int d, int metric, void* stream) {
  if (!ctx) return LEMON_ERR_INVALID;
  if (!a || !b || !out || n < 0 || d <= 0) return lemon_set_error(ctx, LEMON_ERR_INVALID, "rowwise_dist: bad args");
  if (n == 0) return LEMON_OK;
  int64_t blocks = (n + 7) / 8;
  const int64_t cap = int64_t(ctx->num_sms) * 16;
  if (blocks > cap) blocks = cap;
  if (metric == LEMON_METRIC_IP)
    lemon::rowwise_dist_kernel<LEMON_METRIC_IP><<<unsigned(blocks), 256, 0, (cudaStream_t)stream>>>(a, b, out, n, d);
  else
    lemon::rowwise_dist_kernel<LEMON_METRIC_L2><<<unsigned(blocks), 256, 0, (cudaStream_t)stream>>>(a, b, out, n, d);
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}

extern "C" int lemon_split_cast(lemon_ctx* ctx, const float* x, void* out_f16, float* row_stats, float* stats_max,
                                int64_t n, int d, int d16, int role, void* stream) {
  if (!ctx) return LEMON_ERR_INVALID;
  if (!x || !out_f16 || n < 0 || d <= 0 || d16 < d || d16 % 64 || (role != 0 && role != 1))
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "split_cast: bad args");
  if (stats_max) LEMON_CUDA_CHECK(ctx, cudaMemsetAsync(stats_max, 0, 4 * sizeof(float), (cudaStream_t)stream));
  if (n == 0) return LEMON_OK;
  int64_t blocks = (n + 7) / 8;
  const int64_t cap = int64_t(ctx->num_sms) * 16;
  if (blocks > cap) blocks = cap;
  lemon::split_cast_kernel<<<unsigned(blocks), 256, 0, (cudaStream_t)stream>>>(x, (__half*)out_f16, row_stats, stats_max, n, d,
                                                                             d16, role);
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}
