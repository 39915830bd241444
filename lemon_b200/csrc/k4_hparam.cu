// Hyper-parameter grid stage on the GPU (first "next" row).
// The reference evaluates every grid point with calc_scores_given_hparams_vectorized + optimize_f1_efficient on
// the CPU (lib/metrics/utils.py:117-121, 167-186, 286-296; 7056 points, ~31 min per run).  Here one thread block
// handles one grid point: score_i = d_1 + beta * s_n + gamma * s_m from the per-(tau) neighbour terms, then
// Brent's bounded minimiser (the algorithm behind scipy.optimize.fminbound, xtol 1e-8) on
// t -> -F1(y, score >= t), every function evaluation being a block-wide count.  Built with -fmad=false so the
// float64 arithmetic follows the same operation sequence as the CPU code.
#include "lemon_common.cuh"

namespace lemon {

constexpr int kF1Threads = 256;

struct Counts { int tp, pp; };

__device__ __forceinline__ Counts block_count(const double* __restrict__ s, const uint8_t* __restrict__ y, int n,
                                              double thr, int* red) {
  int tp = 0, pp = 0;
  for (int i = threadIdx.x; i < n; i += kF1Threads) {
    const bool p = s[i] >= thr;
    pp += p;
    tp += p && y[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { tp += __shfl_xor_sync(kFull, tp, o); pp += __shfl_xor_sync(kFull, pp, o); }
  const int w = threadIdx.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[2 * w] = tp; red[2 * w + 1] = pp; }
  __syncthreads();
  Counts c{0, 0};
#pragma unroll
  for (int i = 0; i < kF1Threads / 32; ++i) { c.tp += red[2 * i]; c.pp += red[2 * i + 1]; }
  return c;
}

__global__ void __launch_bounds__(kF1Threads)
f1_grid_kernel(const double* __restrict__ d1, const double* __restrict__ sn, const double* __restrict__ sm,
               const uint8_t* __restrict__ y, int n, const double* __restrict__ beta, const double* __restrict__ gamma,
               const int32_t* __restrict__ tidx, int G, double xatol, int maxfun, double* __restrict__ out_f1,
               double* __restrict__ out_thr, double* __restrict__ scratch) {
  __shared__ int red[2 * kF1Threads / 32];
  __shared__ double dred[2 * kF1Threads / 32];
  double* s = scratch + size_t(blockIdx.x) * n;
  for (int g = blockIdx.x; g < G; g += gridDim.x) {
    // ---- scores of this grid point, their range and the number of positives
    double lo = CUDART_INF, hi = -CUDART_INF;
    int P = 0;
    const double b = beta ? beta[g] : 0.0, c = gamma ? gamma[g] : 0.0;
    const size_t toff = sn ? size_t(tidx[g]) * n : 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kF1Threads) {
      double v = d1[i];
      if (sn) v = __dadd_rn(__dadd_rn(v, __dmul_rn(b, sn[toff + i])), __dmul_rn(c, sm[toff + i]));   // utils.py:77
      s[i] = v;
      lo = fmin(lo, v); hi = fmax(hi, v);
      P += y[i] != 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(kFull, lo, o)); hi = fmax(hi, __shfl_xor_sync(kFull, hi, o));
      P += __shfl_xor_sync(kFull, P, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { dred[2 * w] = lo; dred[2 * w + 1] = hi; red[2 * w] = P; }
    __syncthreads();
    lo = dred[0]; hi = dred[1]; P = 0;
#pragma unroll
    for (int i = 0; i < kF1Threads / 32; ++i) { lo = fmin(lo, dred[2 * i]); hi = fmax(hi, dred[2 * i + 1]); P += red[2 * i]; }
    __syncthreads();
    auto negf1 = [&](double t) -> double {
      const Counts k = block_count(s, y, n, t, red);
      const int den = k.pp + P;                                  // (TP + FP) + (TP + FN)
      return den > 0 ? -(2.0 * double(k.tp) / double(den)) : -0.0;
    };
    // ---- Brent's bounded minimiser (every thread carries the same scalar state)
    const double sqrt_eps = sqrt(2.2e-16);
    const double golden_mean = 0.5 * (3.0 - sqrt(5.0));
    double a = lo, bb = hi;
    double fulc = a + golden_mean * (bb - a);
    double nfc = fulc, xf = fulc;
    double rat = 0.0, e = 0.0;
    double x = xf;
    double fx = negf1(x);
    int num = 1;
    double ffulc = fx, fnfc = fx;
    double xm = 0.5 * (a + bb);
    double tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
    double tol2 = 2.0 * tol1;
    while (fabs(xf - xm) > (tol2 - 0.5 * (bb - a))) {
      bool golden = true;
      if (fabs(e) > tol1) {
        golden = false;
        double r = (xf - nfc) * (fx - ffulc);
        double q = (xf - fulc) * (fx - fnfc);
        double p = (xf - fulc) * q - (xf - nfc) * r;
        q = 2.0 * (q - r);
        if (q > 0.0) p = -p;
        q = fabs(q);
        r = e;
        e = rat;
        if (fabs(p) < fabs(0.5 * q * r) && p > q * (a - xf) && p < q * (bb - xf)) {
          rat = (p + 0.0) / q;
          x = xf + rat;
          if ((x - a) < tol2 || (bb - x) < tol2) {
            const double dd = xm - xf;
            const double si = (dd > 0.0 ? 1.0 : (dd < 0.0 ? -1.0 : 0.0)) + (dd == 0.0 ? 1.0 : 0.0);
            rat = tol1 * si;
          }
        } else {
          golden = true;
        }
      }
      if (golden) {
        e = (xf >= xm) ? (a - xf) : (bb - xf);
        rat = golden_mean * e;
      }
      const double si = (rat > 0.0 ? 1.0 : (rat < 0.0 ? -1.0 : 0.0)) + (rat == 0.0 ? 1.0 : 0.0);
      x = xf + si * fmax(fabs(rat), tol1);
      const double fu = negf1(x);
      ++num;
      if (fu <= fx) {
        if (x >= xf) a = xf; else bb = xf;
        fulc = nfc; ffulc = fnfc;
        nfc = xf; fnfc = fx;
        xf = x; fx = fu;
      } else {
        if (x < xf) a = x; else bb = x;
        if (fu <= fnfc || nfc == xf) {
          fulc = nfc; ffulc = fnfc;
          nfc = x; fnfc = fu;
        } else if (fu <= ffulc || fulc == xf || fulc == nfc) {
          fulc = x; ffulc = fu;
        }
      }
      xm = 0.5 * (a + bb);
      tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
      tol2 = 2.0 * tol1;
      if (num >= maxfun) break;
    }
    const double best = -negf1(xf);                              // utils.py:292: best_f1 = -neg_f1(best_thres)
    if (threadIdx.x == 0) { out_f1[g] = best; out_thr[g] = xf; }
  }
}

}  // namespace lemon

extern "C" int lemon_f1_grid(lemon_ctx* ctx, const double* d1, const double* sn, const double* sm, const uint8_t* y,
                             int64_t n, const double* beta, const double* gamma, const int32_t* tidx, int64_t n_points,
                             double xatol, int maxfun, double* out_f1, double* out_thr, double* scratch,
                             int64_t scratch_rows, void* stream) {
  using namespace lemon;
  if (!ctx) return LEMON_ERR_INVALID;
  if (!d1 || !y || !out_f1 || !out_thr || !scratch || n < 1 || n > (int64_t(1) << 30) || n_points < 1 || scratch_rows < 1 ||
      ((sn == nullptr) != (sm == nullptr)) || (sn && (!beta || !gamma || !tidx)))
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "f1_grid: bad args");
  int64_t blocks = n_points < scratch_rows ? n_points : scratch_rows;
  f1_grid_kernel<<<unsigned(blocks), kF1Threads, 0, (cudaStream_t)stream>>>(d1, sn, sm, y, int(n), beta, gamma, tidx,
                                                                           int(n_points), xatol, maxfun, out_f1, out_thr, scratch);
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}
