// Shared device helpers for liblemon_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math_constants.h>

#include "../../include/lemon_b200.h"

struct lemon_ctx {
  int device;
  int num_sms;
  int cc_major, cc_minor;
  int64_t launches;
  char err[512];
  // tensor-core kernel scratch: a ring of 16 progress arrays (256 int32 each) for K1's DB-walk pacing
  void* tc_scratch;
  size_t tc_scratch_bytes;
  unsigned tc_launch_seq;
  void* encode_tiled;   // cuTensorMapEncodeTiled, resolved lazily
  // K1 tuning knobs; -1 = library default.  Only builds with -DLEMON_TC_EXPERIMENT read them from the environment
  // (once, at ctx creation); the product build always runs the defaults.
  int tune_kres, tune_debug, tune_cert, tune_boot, tune_bn, tune_pace;
};

int lemon_set_error(lemon_ctx* ctx, int code, const char* fmt, ...);

#define LEMON_CUDA_CHECK(ctx, expr)                                                         \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return lemon_set_error((ctx), LEMON_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,         \
                             cudaGetErrorString(_e), __FILE__, __LINE__);                   \
  } while (0)

namespace lemon {

constexpr int kCap = 256;            // keys one warp sorts at once (8 per lane); K2a staging capacity
constexpr int kListCap = LEMON_LIST_CAP;   // slots per K1 candidate list
constexpr int kKeep = LEMON_KPRIME;  // survivors per compaction (k')
constexpr unsigned kFull = 0xffffffffu;

// ---- 64-bit sort key: (order-preserving float bits) << 32 | ~idx -------------------------
// larger key == better: larger value first, then LOWER index first.  key 0 == empty slot.
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  uint32_t u = o ^ ((o >> 31) ? 0x80000000u : 0xffffffffu);
  return __uint_as_float(u);
}
// bit pattern of the largest float strictly below f (f finite, not the most negative float)
__device__ __forceinline__ uint32_t f2ord_dec(float f) {
  const uint32_t o = f2ord(f) - 1u;
  return o ^ ((o >> 31) ? 0x80000000u : 0xffffffffu);
}
__device__ __forceinline__ uint64_t make_key(float v, uint32_t idx) {
  return (uint64_t(f2ord(v)) << 32) | uint64_t(~idx);
}
__device__ __forceinline__ float key_val(uint64_t k) { return ord2f(uint32_t(k >> 32)); }
__device__ __forceinline__ int32_t key_idx(uint64_t k) { return int32_t(~uint32_t(k)); }

__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int m) {
  uint32_t lo = __shfl_xor_sync(kFull, uint32_t(v), m);
  uint32_t hi = __shfl_xor_sync(kFull, uint32_t(v >> 32), m);
  return (uint64_t(hi) << 32) | lo;
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(kFull, uint32_t(v), src);
  uint32_t hi = __shfl_sync(kFull, uint32_t(v >> 32), src);
  return (uint64_t(hi) << 32) | lo;
}

// Warp-wide bitonic sort of 256 keys, DESCENDING.  Element e = lane*8 + r lives in key[r] of
// `lane`; afterwards lane L holds ranks 8L .. 8L+7 (rank 0 = largest key).
__device__ __forceinline__ void warp_sort256_desc(uint64_t (&key)[8], int lane) {
#pragma unroll
  for (int k = 2; k <= 256; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 8) {
        const int lm = j >> 3;
        const bool lower = (lane & lm) == 0;
        const bool desc = ((lane * 8) & k) == 0;   // k >= 16 here, so the bit is a lane bit
        const bool keep_max = (lower == desc);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          uint64_t o = shfl_xor_u64(key[r], lm);
          uint64_t mx = key[r] > o ? key[r] : o;
          uint64_t mn = key[r] > o ? o : key[r];
          key[r] = keep_max ? mx : mn;
        }
      } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          if ((r & j) == 0) {
            const int e = lane * 8 + r;
            const bool desc = (e & k) == 0;
            uint64_t a = key[r], b = key[r | j];
            uint64_t mx = a > b ? a : b;
            uint64_t mn = a > b ? b : a;
            key[r] = desc ? mx : mn;
            key[r | j] = desc ? mn : mx;
          }
        }
      }
    }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// fp32 "exact" pair value as THIS library defines it: lane-strided float4 partial sums
// (ascending column order per lane) followed by a xor-butterfly.  Used by every kernel that
// reports a similarity / distance so that values agree bit-for-bit across kernels.
//   metric IP: <a,b>      metric L2: sum (a-b)^2
template <int METRIC>
__device__ __forceinline__ float warp_pair_value(const float* __restrict__ a, const float* __restrict__ b,
                                                 int d, int lane) {
  float acc = 0.f;
  const int d4 = (d & 3) ? 0 : (d >> 2);   // rows are 16B-aligned only when d % 4 == 0
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  for (int c = lane; c < d4; c += 32) {
    float4 x = a4[c], y = __ldg(b4 + c);
    if (METRIC == LEMON_METRIC_IP) {
      acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc);
      acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
    } else {
      float t;
      t = x.x - y.x; acc = fmaf(t, t, acc); t = x.y - y.y; acc = fmaf(t, t, acc);
      t = x.z - y.z; acc = fmaf(t, t, acc); t = x.w - y.w; acc = fmaf(t, t, acc);
    }
  }
  for (int c = (d4 << 2) + lane; c < d; c += 32) {   // scalar path (d % 4 != 0)
    float x = a[c], y = __ldg(b + c);
    if (METRIC == LEMON_METRIC_IP) acc = fmaf(x, y, acc);
    else { float t = x - y; acc = fmaf(t, t, acc); }
  }
  return warp_sum(acc);
}

// Four pair values at once (same query row against four DB rows).  Per pair the operation order is EXACTLY
// that of warp_pair_value (bit-identical results); issuing the four rows' loads together gives the gather
// kernels 4x the memory-level parallelism.
template <int METRIC>
__device__ __forceinline__ void warp_pair_value4(const float* __restrict__ a, const float* __restrict__ b0,
                                                 const float* __restrict__ b1, const float* __restrict__ b2,
                                                 const float* __restrict__ b3, int d, int lane, float (&out)[4]) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float* bs[4] = {b0, b1, b2, b3};
  const int d4 = (d & 3) ? 0 : (d >> 2);
  const float4* a4 = reinterpret_cast<const float4*>(a);
  for (int c = lane; c < d4; c += 32) {
    const float4 x = a4[c];
    float4 y[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = __ldg(reinterpret_cast<const float4*>(bs[j]) + c);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (METRIC == LEMON_METRIC_IP) {
        acc[j] = fmaf(x.x, y[j].x, acc[j]); acc[j] = fmaf(x.y, y[j].y, acc[j]);
        acc[j] = fmaf(x.z, y[j].z, acc[j]); acc[j] = fmaf(x.w, y[j].w, acc[j]);
      } else {
        float t;
        t = x.x - y[j].x; acc[j] = fmaf(t, t, acc[j]); t = x.y - y[j].y; acc[j] = fmaf(t, t, acc[j]);
        t = x.z - y[j].z; acc[j] = fmaf(t, t, acc[j]); t = x.w - y[j].w; acc[j] = fmaf(t, t, acc[j]);
      }
    }
  }
  for (int c = (d4 << 2) + lane; c < d; c += 32) {
    const float x = a[c];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float yv = __ldg(bs[j] + c);
      if (METRIC == LEMON_METRIC_IP) acc[j] = fmaf(x, yv, acc[j]);
      else { const float t = x - yv; acc[j] = fmaf(t, t, acc[j]); }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) out[j] = warp_sum(acc[j]);
}

}  // namespace lemon
