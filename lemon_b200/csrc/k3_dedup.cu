// Exact-duplicate handling for the kNN database, entirely on the device.  Classification datasets have only C
// distinct text embeddings (run_lemon.py:117-119,140-143) and caption-noise injection duplicates captions
// (lib/datasets/noise_captioning.py:44-53), so thousands of DB rows can be bit-identical.  Identical rows are
// searched once.  Step 1 (every database): hash_rows -> count_dups (open-addressing table) gives the number of
// duplicate rows in two kernels; the host reads that one counter (the number of unique rows sizes the search
// operands, so one round trip per database is inherent).  Step 2 (only databases with >= 10 % duplicates):
//   hash_rows -> LSD radix sort of (hash, row) -> run flags -> scans -> bit-wise verification against the run's
//   first (= lowest) row -> groups renumbered by ascending representative row -> offsets / members
// all queued on one stream.  The search then runs on one representative per group and
// expand_groups turns every hit into its members in ascending DB index — the list the full search returns under
// the documented total order (value best-first, then index ascending).
//
// The radix sort (64-bit keys, 32-bit payload, 8 bits per pass, stable) and the exclusive scan are also used by
// lemon_keep_lowest, the CC3M filtering consumer of the scores (train_clip_from_scratch.py:110-113).
#include "lemon_common.cuh"

namespace lemon {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// ------------------------------------------------------------------------------------ hashing
__global__ void __launch_bounds__(256)
hash_rows_kernel(const uint32_t* __restrict__ x, int64_t n, int d, uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < n; row += warps) {
    const uint32_t* r = x + row * d;
    uint64_t h = 0x9e3779b97f4a7c15ull * uint64_t(lane + 1);
    for (int c = lane; c < d; c += 32) h = mix64(h ^ (uint64_t(r[c]) | (uint64_t(c) << 32)));
    // order-independent combine across lanes is fine: every lane's stream is position-tagged
    uint64_t acc = h;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += shfl_xor_u64(acc, o);
    if (lane == 0) {
      keys[row] = mix64(acc) >> 1;   // 63 bits: non-negative when read as int64
      if (vals) vals[row] = int32_t(row);
    }
  }
}

// Counts the rows whose 63-bit hash was already seen (open-addressing table, one atomicCAS per probe): n - count is
// the number of distinct hashes, i.e. the number of unique rows unless two different rows collide.  This is all the
// caller needs to decide whether the database is worth de-duplicating; the sort below only runs when it is.
__global__ void __launch_bounds__(256)
count_dups_kernel(const uint64_t* __restrict__ keys, int64_t n, unsigned long long* __restrict__ table, uint64_t mask,
                  int32_t* __restrict__ counters) {
  int local = 0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
    const unsigned long long tag = keys[i] | 0x8000000000000000ull;       // never 0 (= empty slot)
    uint64_t slot = (keys[i] * 0x9e3779b97f4a7c15ull) >> 20 & mask;
    for (uint64_t probes = 0; probes <= mask; ++probes) {
      const unsigned long long prev = atomicCAS(table + slot, 0ull, tag);
      if (prev == 0ull) break;
      if (prev == tag) { ++local; break; }
      slot = (slot + 1) & mask;
    }
  }
  local = __reduce_add_sync(kFull, local);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(counters, local);
}

// ------------------------------------------------------------------------------------ radix sort
constexpr int kRsThreads = 256;
constexpr int kRsItems = 8;                          // keys per thread
constexpr int kRsTile = kRsThreads * kRsItems;       // 2048 keys per block; warp w owns keys [w*256, (w+1)*256) of the tile
constexpr int kRsWarps = kRsThreads / 32;

__global__ void __launch_bounds__(kRsThreads)
rs_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift, int32_t* __restrict__ hist, int nblk) {
  __shared__ int32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = int64_t(blockIdx.x) * kRsTile;
#pragma unroll
  for (int i = 0; i < kRsItems; ++i) {
    const int64_t e = base + i * kRsThreads + threadIdx.x;
    if (e < n) atomicAdd(&h[int(keys[e] >> shift) & 255], 1);
  }
  __syncthreads();
  hist[int64_t(threadIdx.x) * nblk + blockIdx.x] = h[threadIdx.x];     // bin-major: a flat scan gives the global offsets
}

// exclusive scan of `len` int32 values in place by ONE block (len = 256 * nblk, a few hundred thousand at most)
__global__ void __launch_bounds__(1024)
scan_single_block_kernel(int32_t* __restrict__ a, int64_t len) {
  __shared__ int32_t part[1024];
  const int t = threadIdx.x;
  const int64_t per = (len + 1023) / 1024;
  const int64_t s = int64_t(t) * per, e = min(len, s + per);
  int32_t sum = 0;
  for (int64_t i = s; i < e; ++i) sum += a[i];
  part[t] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {            // Hillis-Steele inclusive scan of the partials
    const int32_t v = t >= o ? part[t - o] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int32_t run = part[t] - sum;
  for (int64_t i = s; i < e; ++i) { const int32_t v = a[i]; a[i] = run; run += v; }
}

__global__ void __launch_bounds__(kRsThreads)
rs_scatter_kernel(const uint64_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in, uint64_t* __restrict__ keys_out,
                  int32_t* __restrict__ vals_out, int64_t n, int shift, const int32_t* __restrict__ hist, int nblk) {
  __shared__ int32_t wh[kRsWarps][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < kRsWarps * 256; i += kRsThreads) (&wh[0][0])[i] = 0;
  __syncthreads();
  const int64_t base = int64_t(blockIdx.x) * kRsTile + warp * (kRsItems * 32);
  uint64_t key[kRsItems];
  int32_t val[kRsItems];
  unsigned peers[kRsItems];
#pragma unroll
  for (int i = 0; i < kRsItems; ++i) {
    const int64_t e = base + i * 32 + lane;
    const bool ok = e < n;
    key[i] = ok ? keys_in[e] : 0ull;
    val[i] = ok ? vals_in[e] : 0;
    const int digit = ok ? (int(key[i] >> shift) & 255) : 256 + lane;      // invalid lanes match nobody
    peers[i] = __match_any_sync(kFull, digit);
    if (ok && lane == __ffs(peers[i]) - 1) wh[warp][digit] += __popc(peers[i]);
    __syncwarp();
  }
  __syncthreads();
  {   // per digit: global offset of this block, then the warps of the block in order
    const int t = threadIdx.x;
    int32_t run = hist[int64_t(t) * nblk + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) { const int32_t c = wh[w][t]; wh[w][t] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kRsItems; ++i) {
    const int64_t e = base + i * 32 + lane;
    const bool ok = e < n;
    const int digit = int(key[i] >> shift) & 255;
    if (ok) {
      const int32_t pos = wh[warp][digit] + __popc(peers[i] & ((1u << lane) - 1u));
      keys_out[pos] = key[i];
      vals_out[pos] = val[i];
    }
    __syncwarp();
    if (ok && lane == __ffs(peers[i]) - 1) wh[warp][digit] += __popc(peers[i]);
    __syncwarp();
  }
}

static inline int rs_blocks(int64_t n) { return int((n + kRsTile - 1) / kRsTile); }

// Stable ascending sort of (keys, vals) on bits [0, 8*passes).  a = input and final output when `passes` is even
// (else the result is in b); hist: 256 * rs_blocks(n) int32.
static void radix_sort_pairs(uint64_t* ka, int32_t* va, uint64_t* kb, int32_t* vb, int32_t* hist, int64_t n, int passes,
                             cudaStream_t st, lemon_ctx* ctx) {
  const int nblk = rs_blocks(n);
  for (int p = 0; p < passes; ++p) {
    rs_hist_kernel<<<nblk, kRsThreads, 0, st>>>(ka, n, 8 * p, hist, nblk);
    scan_single_block_kernel<<<1, 1024, 0, st>>>(hist, int64_t(256) * nblk);
    rs_scatter_kernel<<<nblk, kRsThreads, 0, st>>>(ka, va, kb, vb, n, 8 * p, hist, nblk);
    ctx->launches += 3;
    uint64_t* tk = ka; ka = kb; kb = tk;
    int32_t* tv = va; va = vb; vb = tv;
  }
}

// ------------------------------------------------------------------------------------ exclusive scan (any length)
constexpr int kScanTile = 1024 * 4;
__global__ void __launch_bounds__(1024)
scan_tile_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n, int32_t* __restrict__ tile_sums) {
  __shared__ int32_t part[1024];
  const int t = threadIdx.x;
  const int64_t base = int64_t(blockIdx.x) * kScanTile + t * 4;
  int32_t v[4], sum = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[i] = base + i < n ? in[base + i] : 0; sum += v[i]; }
  part[t] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int32_t a = t >= o ? part[t - o] : 0;
    __syncthreads();
    part[t] += a;
    __syncthreads();
  }
  int32_t run = part[t] - sum;
#pragma unroll
  for (int i = 0; i < 4; ++i) { if (base + i < n) out[base + i] = run; run += v[i]; }
  if (t == 1023) tile_sums[blockIdx.x] = part[1023];
}
__global__ void __launch_bounds__(1024)
scan_add_kernel(int32_t* __restrict__ out, int64_t n, const int32_t* __restrict__ tile_offsets, int32_t* __restrict__ total,
                const int32_t* __restrict__ in_last) {
  const int64_t base = int64_t(blockIdx.x) * kScanTile + threadIdx.x * 4;
  const int32_t off = tile_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (base + i < n) {
      const int32_t v = out[base + i] + off;
      out[base + i] = v;
      if (base + i == n - 1 && total) *total = v + in_last[n - 1];
    }
  }
}
// out[i] = sum_{j<i} in[j] (in != out); *total = sum of all (device, may be NULL); tiles: scratch of >= n/4096 + 1 int32
static void exclusive_scan(const int32_t* in, int32_t* out, int64_t n, int32_t* tiles, int32_t* total, cudaStream_t st,
                           lemon_ctx* ctx) {
  const int nt = int((n + kScanTile - 1) / kScanTile);
  scan_tile_kernel<<<nt, 1024, 0, st>>>(in, out, n, tiles);
  scan_single_block_kernel<<<1, 1024, 0, st>>>(tiles, nt);
  scan_add_kernel<<<nt, 1024, 0, st>>>(out, n, tiles, total, in);
  ctx->launches += 3;
}

// ------------------------------------------------------------------------------------ grouping
// sorted position p: flag[p] = 1 when a new run of equal hashes starts
__global__ void run_flags_kernel(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ flag) {
  const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p < n) flag[p] = (p == 0 || keys[p] != keys[p - 1]) ? 1 : 0;
}
// run id of every sorted position, first position of every run
__global__ void run_starts_kernel(const int32_t* __restrict__ flag, int32_t* __restrict__ excl, int64_t n,
                                  int32_t* __restrict__ run_start) {
  const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int32_t g = excl[p] + flag[p] - 1;
  excl[p] = g;                                   // in place: exclusive scan -> run id
  if (flag[p]) run_start[g] = int32_t(p);
}
// bit-wise verification against the run's first (= lowest, the sort is stable) row; marks representatives
__global__ void __launch_bounds__(256)
verify_runs_kernel(const uint32_t* __restrict__ x, const int32_t* __restrict__ rows_sorted, const int32_t* __restrict__ run_id,
                   const int32_t* __restrict__ run_start, int64_t n, int d, int32_t* __restrict__ is_rep,
                   int32_t* __restrict__ counters) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t p = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; p < n; p += warps) {
    const int32_t row = rows_sorted[p];
    const int32_t rep = rows_sorted[run_start[run_id[p]]];
    if (lane == 0) is_rep[row] = rep == row;
    if (rep == row) continue;
    const uint32_t* a = x + int64_t(row) * d;
    const uint32_t* b = x + int64_t(rep) * d;
    bool diff = false;
    for (int c = lane; c < d; c += 32) diff |= a[c] != b[c];
    if (__any_sync(kFull, diff) && lane == 0) atomicOr(counters + 1, 1);      // 63-bit hash collision: no de-duplication
  }
}
// runs renumbered by ascending representative row: sizes and representatives under the new numbering
__global__ void run_sizes_kernel(const int32_t* __restrict__ rows_sorted, const int32_t* __restrict__ run_start,
                                 const int32_t* __restrict__ rep_rank, const int32_t* __restrict__ counters, int64_t n,
                                 int32_t* __restrict__ new_id, int32_t* __restrict__ sizes, int32_t* __restrict__ rep_rows) {
  const int64_t g = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int32_t n_u = counters[0];
  if (g >= n_u) return;
  const int32_t s = run_start[g], e = g + 1 < n_u ? run_start[g + 1] : int32_t(n);
  const int32_t rep = rows_sorted[s];
  const int32_t id = rep_rank[rep];
  new_id[g] = id;
  sizes[id] = e - s;
  rep_rows[id] = rep;
}
__global__ void members_kernel(const int32_t* __restrict__ rows_sorted, const int32_t* __restrict__ run_id,
                               const int32_t* __restrict__ run_start, const int32_t* __restrict__ new_id,
                               const int32_t* __restrict__ offs32, const int32_t* __restrict__ counters, int64_t n,
                               int32_t* __restrict__ members, int64_t* __restrict__ offsets) {
  const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int32_t g = run_id[p];
  const int32_t id = new_id[g];
  members[offs32[id] + (int32_t(p) - run_start[g])] = rows_sorted[p];       // ascending row index inside a group
  const int32_t n_u = counters[0];
  if (p < n_u) offsets[p] = offs32[p];
  if (p == 0) offsets[n_u] = n;
}

__global__ void gather_rows_kernel(const uint4* __restrict__ src, const int32_t* __restrict__ idx, const int32_t* __restrict__ n_idx,
                                   int64_t max_idx, int row_u4, uint4* __restrict__ dst) {
  const int64_t cnt = n_idx ? min(int64_t(*n_idx), max_idx) : max_idx;
  const int64_t total = cnt * row_u4;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int64_t r = i / row_u4, c = i - r * row_u4;
    dst[i] = __ldg(src + int64_t(idx[r]) * row_u4 + c);
  }
}

// top lists over the UNIQUE rows -> top lists over the original DB rows.  Consecutive unique rows with EQUAL values
// (distinct rows that tie exactly) are merged by ascending member index, as the total order demands.
template <int METRIC>
__global__ void __launch_bounds__(256)
expand_groups_kernel(const float* __restrict__ uval, const int32_t* __restrict__ uidx, const int64_t* __restrict__ offsets,
                     const int32_t* __restrict__ members, int64_t nq, int kp, float* __restrict__ top_val,
                     int32_t* __restrict__ top_idx) {
  const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= nq) return;
  const float* uv = uval + row * kp;
  const int32_t* ui = uidx + row * kp;
  float* ov = top_val + row * kp;
  int32_t* oi = top_idx + row * kp;
  int pos = 0;
  int r = 0;
  while (r < kp && pos < kp) {
    const int u = ui[r];
    if (u < 0) break;
    const float v = uv[r];
    int r2 = r + 1;
    while (r2 < kp && ui[r2] >= 0 && uv[r2] == v) ++r2;
    if (r2 == r + 1) {
      const int64_t e = offsets[u + 1];
      for (int64_t j = offsets[u]; j < e && pos < kp; ++j) { ov[pos] = v; oi[pos] = members[j]; ++pos; }
    } else {
      // tied unique rows r .. r2-1: k-way merge of their member lists (each ascending), lowest DB index first
      int64_t cur[LEMON_MAX_KP];
      for (int j = r; j < r2; ++j) cur[j] = offsets[ui[j]];
      while (pos < kp) {
        int best = -1;
        int32_t best_m = 0x7fffffff;
        for (int j = r; j < r2; ++j) {
          if (cur[j] < offsets[ui[j] + 1]) {
            const int32_t mj = members[cur[j]];
            if (mj < best_m) { best_m = mj; best = j; }
          }
        }
        if (best < 0) break;
        ov[pos] = v; oi[pos] = best_m; ++pos; ++cur[best];
      }
    }
    r = r2;
  }
  for (; pos < kp; ++pos) { ov[pos] = METRIC == LEMON_METRIC_IP ? -CUDART_INF_F : CUDART_INF_F; oi[pos] = -1; }
}

// ------------------------------------------------------------------------------------ keep_lowest (CC3M consumer)
__global__ void score_keys_kernel(const double* __restrict__ s, int64_t n, uint64_t* __restrict__ keys, int32_t* __restrict__ vals) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t u = uint64_t(__double_as_longlong(s[i]));
  keys[i] = (u >> 63) ? ~u : (u | 0x8000000000000000ull);     // ascending unsigned order == ascending double order (NaN last)
  vals[i] = int32_t(i);
}
__global__ void emit_lowest_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ vals, int64_t n_keep,
                                   int64_t* __restrict__ out_idx, double* __restrict__ out_score) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_keep) return;
  out_idx[i] = vals[i];
  if (out_score) {
    const uint64_t k = keys[i];
    const uint64_t u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    out_score[i] = __longlong_as_double((long long)u);
  }
}

struct SortWs {
  uint64_t *ka, *kb;
  int32_t *va, *vb, *hist;
  unsigned char* rest;
};
static inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
static inline size_t sort_ws_bytes(int64_t n) {
  return 2 * align256(size_t(n) * 8) + 2 * align256(size_t(n) * 4) + align256(size_t(256) * rs_blocks(n) * 4);
}
static inline SortWs carve_sort_ws(void* ws, int64_t n) {
  unsigned char* p = static_cast<unsigned char*>(ws);
  SortWs w;
  w.ka = reinterpret_cast<uint64_t*>(p); p += align256(size_t(n) * 8);
  w.kb = reinterpret_cast<uint64_t*>(p); p += align256(size_t(n) * 8);
  w.va = reinterpret_cast<int32_t*>(p); p += align256(size_t(n) * 4);
  w.vb = reinterpret_cast<int32_t*>(p); p += align256(size_t(n) * 4);
  w.hist = reinterpret_cast<int32_t*>(p); p += align256(size_t(256) * rs_blocks(n) * 4);
  w.rest = p;
  return w;
}

}  // namespace lemon

static inline uint64_t dup_table_slots(int64_t n) {
  uint64_t s = 1024;
  while (s < uint64_t(n) * 2) s <<= 1;
  return s;
}

extern "C" int64_t lemon_dedup_count_workspace_bytes(int64_t n) {
  return n < 1 ? 0 : int64_t(lemon::align256(size_t(n) * 8) + dup_table_slots(n) * 8);
}

extern "C" int lemon_dedup_count(lemon_ctx* ctx, const float* x, int64_t n, int d, void* workspace, int32_t* counters,
                                 void* stream) {
  using namespace lemon;
  if (!ctx) return LEMON_ERR_INVALID;
  if (!x || !workspace || !counters || n < 1 || d <= 0 || (uintptr_t(workspace) & 255))
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "dedup_count: bad args (workspace must be 256 B aligned)");
  cudaStream_t st = (cudaStream_t)stream;
  uint64_t* keys = static_cast<uint64_t*>(workspace);
  unsigned long long* table = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(workspace) + align256(size_t(n) * 8));
  const uint64_t slots = dup_table_slots(n);
  LEMON_CUDA_CHECK(ctx, cudaMemsetAsync(counters, 0, sizeof(int32_t), st));
  LEMON_CUDA_CHECK(ctx, cudaMemsetAsync(table, 0, slots * 8, st));
  int64_t wblocks = (n + 7) / 8;
  if (wblocks > int64_t(ctx->num_sms) * 16) wblocks = int64_t(ctx->num_sms) * 16;
  hash_rows_kernel<<<unsigned(wblocks), 256, 0, st>>>(reinterpret_cast<const uint32_t*>(x), n, d, keys, nullptr);
  int64_t blocks = (n + 255) / 256;
  if (blocks > int64_t(ctx->num_sms) * 8) blocks = int64_t(ctx->num_sms) * 8;
  count_dups_kernel<<<unsigned(blocks), 256, 0, st>>>(keys, n, table, slots - 1, counters);
  ctx->launches += 2;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}

extern "C" int64_t lemon_dedup_workspace_bytes(int64_t n) {
  using namespace lemon;
  if (n < 1) return 0;
  // sort buffers + 6 int32 arrays of n (+1) + scan tiles
  return int64_t(sort_ws_bytes(n) + 6 * align256(size_t(n + 1) * 4) + align256(size_t(n / kScanTile + 2) * 4));
}

extern "C" int lemon_dedup_build(lemon_ctx* ctx, const float* x, int64_t n, int d, void* workspace, int32_t* rep_rows,
                                 int32_t* members, int64_t* offsets, int32_t* counters, void* stream) {
  using namespace lemon;
  if (!ctx) return LEMON_ERR_INVALID;
  if (!x || !workspace || !rep_rows || !members || !offsets || !counters || n < 1 || n >= (int64_t(1) << 31) || d <= 0 ||
      (uintptr_t(workspace) & 255))
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "dedup_build: bad args (workspace must be 256 B aligned)");
  cudaStream_t st = (cudaStream_t)stream;
  SortWs w = carve_sort_ws(workspace, n);
  unsigned char* p = w.rest;
  auto take = [&](size_t elems) { int32_t* r = reinterpret_cast<int32_t*>(p); p += align256(elems * 4); return r; };
  int32_t* flag = take(n + 1);
  int32_t* run_id = take(n + 1);
  int32_t* run_start = take(n + 1);
  int32_t* is_rep = take(n + 1);
  int32_t* rep_rank = take(n + 1);     // later reused: exclusive scan of the group sizes
  int32_t* new_id = take(n + 1);
  int32_t* tiles = take(n / kScanTile + 2);
  int32_t* sizes = flag;               // flag is dead once the run ids exist
  const unsigned g256 = unsigned((n + 255) / 256);
  int64_t wblocks = (n + 7) / 8;
  if (wblocks > int64_t(ctx->num_sms) * 16) wblocks = int64_t(ctx->num_sms) * 16;

  LEMON_CUDA_CHECK(ctx, cudaMemsetAsync(counters, 0, 2 * sizeof(int32_t), st));
  hash_rows_kernel<<<unsigned(wblocks), 256, 0, st>>>(reinterpret_cast<const uint32_t*>(x), n, d, w.ka, w.va);
  radix_sort_pairs(w.ka, w.va, w.kb, w.vb, w.hist, n, 8, st, ctx);                   // 63-bit hash: 8 passes, result in ka / va
  run_flags_kernel<<<g256, 256, 0, st>>>(w.ka, n, flag);
  exclusive_scan(flag, run_id, n, tiles, counters, st, ctx);                         // counters[0] = number of runs = n_unique
  run_starts_kernel<<<g256, 256, 0, st>>>(flag, run_id, n, run_start);
  verify_runs_kernel<<<unsigned(wblocks), 256, 0, st>>>(reinterpret_cast<const uint32_t*>(x), w.va, run_id, run_start, n, d,
                                                        is_rep, counters);
  exclusive_scan(is_rep, rep_rank, n, tiles, nullptr, st, ctx);                      // rank of every representative among them
  LEMON_CUDA_CHECK(ctx, cudaMemsetAsync(sizes, 0, size_t(n + 1) * 4, st));
  run_sizes_kernel<<<g256, 256, 0, st>>>(w.va, run_start, rep_rank, counters, n, new_id, sizes, rep_rows);
  int32_t* offs32 = is_rep;            // is_rep is dead after its scan
  exclusive_scan(sizes, offs32, n, tiles, nullptr, st, ctx);
  members_kernel<<<g256, 256, 0, st>>>(w.va, run_id, run_start, new_id, offs32, counters, n, members, offsets);
  ctx->launches += 6;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}

extern "C" int lemon_gather_rows(lemon_ctx* ctx, const void* src, const int32_t* idx, const int32_t* n_idx, int64_t max_idx,
                                 int64_t row_bytes, void* dst, void* stream) {
  if (!ctx) return LEMON_ERR_INVALID;
  if (!src || !idx || !dst || max_idx < 0 || row_bytes <= 0 || row_bytes % 16 || (uintptr_t(src) & 15) || (uintptr_t(dst) & 15))
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "gather_rows: bad args (rows must be multiples of 16 B, 16 B aligned)");
  if (max_idx == 0) return LEMON_OK;
  const int64_t total = max_idx * (row_bytes / 16);
  int64_t blocks = (total + 255) / 256;
  if (blocks > int64_t(ctx->num_sms) * 32) blocks = int64_t(ctx->num_sms) * 32;
  lemon::gather_rows_kernel<<<unsigned(blocks), 256, 0, (cudaStream_t)stream>>>(
      static_cast<const uint4*>(src), idx, n_idx, max_idx, int(row_bytes / 16), static_cast<uint4*>(dst));
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}

extern "C" int64_t lemon_keep_lowest_workspace_bytes(int64_t n) { return n < 1 ? 0 : int64_t(lemon::sort_ws_bytes(n)); }

extern "C" int lemon_keep_lowest(lemon_ctx* ctx, const double* score, int64_t n, int64_t n_keep, void* workspace,
                                 int64_t* out_idx, double* out_score, void* stream) {
  using namespace lemon;
  if (!ctx) return LEMON_ERR_INVALID;
  if (!score || !workspace || !out_idx || n < 1 || n >= (int64_t(1) << 31) || n_keep < 0 || n_keep > n || (uintptr_t(workspace) & 255))
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "keep_lowest: bad args");
  if (n_keep == 0) return LEMON_OK;
  cudaStream_t st = (cudaStream_t)stream;
  SortWs w = carve_sort_ws(workspace, n);
  const unsigned g256 = unsigned((n + 255) / 256);
  score_keys_kernel<<<g256, 256, 0, st>>>(score, n, w.ka, w.va);
  radix_sort_pairs(w.ka, w.va, w.kb, w.vb, w.hist, n, 8, st, ctx);
  emit_lowest_kernel<<<unsigned((n_keep + 255) / 256), 256, 0, st>>>(w.ka, w.va, n_keep, out_idx, out_score);
  ctx->launches += 2;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}

extern "C" int lemon_hash_rows(lemon_ctx* ctx, const float* x, int64_t n, int d, int64_t* out, void* stream) {
  if (!ctx) return LEMON_ERR_INVALID;
  if (!x || !out || n < 0 || d <= 0) return lemon_set_error(ctx, LEMON_ERR_INVALID, "hash_rows: bad args");
  if (n == 0) return LEMON_OK;
  int64_t blocks = (n + 7) / 8;
  const int64_t cap = int64_t(ctx->num_sms) * 16;
  if (blocks > cap) blocks = cap;
  lemon::hash_rows_kernel<<<unsigned(blocks), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t*>(x), n, d,
                                                                            reinterpret_cast<uint64_t*>(out), nullptr);
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}

extern "C" int lemon_expand_groups(lemon_ctx* ctx, const float* uval, const int32_t* uidx, const int64_t* offsets,
                                   const int32_t* members, int64_t nq, int kp, int metric, float* top_val,
                                   int32_t* top_idx, void* stream) {
  if (!ctx) return LEMON_ERR_INVALID;
  if (!uval || !uidx || !offsets || !members || !top_val || !top_idx || nq < 0 || kp < 1 || kp > LEMON_MAX_KP)
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "expand_groups: bad args");
  if (nq == 0) return LEMON_OK;
  const unsigned blocks = unsigned((nq + 255) / 256);
  if (metric == LEMON_METRIC_IP)
    lemon::expand_groups_kernel<LEMON_METRIC_IP><<<blocks, 256, 0, (cudaStream_t)stream>>>(uval, uidx, offsets, members, nq, kp, top_val, top_idx);
  else
    lemon::expand_groups_kernel<LEMON_METRIC_L2><<<blocks, 256, 0, (cudaStream_t)stream>>>(uval, uidx, offsets, members, nq, kp, top_val, top_idx);
  ctx->launches++;
  LEMON_CUDA_CHECK(ctx, cudaGetLastError());
  return LEMON_OK;
}
