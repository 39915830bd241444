// K1: tensor-core candidate search.  Replaces faiss IndexFlatIP.search (run_lemon.py:235-236)
// for the bulk of the work: S = Q * DB^T on tcgen05 (fp16 operands, fp32 TMEM accumulators) with a
// streaming top-k fused into the epilogue, so the nq x m similarity matrix never reaches HBM.
//
// Structure (one persistent CTA pair with cta_group::2 -- or single CTA -- per SM, 384 threads):
//   warp 0    TMA producer : query tile (resident in SMEM for the whole DB scan) + DB tiles (ring)
//   warp 1    MMA issuer   : tcgen05.mma kind::f16, 128(x2) x BN x 16 per instruction, accumulators
//                             double-buffered in TMEM (512 columns)
//   warp 2    TMEM alloc / dealloc
//   warp 3    pacer (leader CTA): keeps the CTA pairs within an L2-sized window of the slowest pair's DB position,
//                             so that a DB tile is fetched from DRAM once and not once per pair ("DB-walk pacing" below)
//   warps 4-11 epilogue    : two groups of four warps; group g scans column half g of EVERY accumulator tile
//                             (so a buffer is held for half a scan time and no group idles while "its" buffer
//                             is being refilled).  tcgen05.ld 32x32b (thread == query row), FMNMX3 max tree
//                             against the row's threshold, survivors appended branch-free as 64-bit keys (raw float bits | ~row) to the
//                             row's 1024-slot list (which lives in the OUTPUT array); the row thresholds rise by
//                             a counting ladder (no sorting or compaction during the scan), are shared between
//                             the groups and bootstrapped from group maxima.
// Work item = (query row tile, DB segment); items are dealt round-robin so all CTAs walk the DB
// in the same order and DB tiles are served from L2.  The selection over the union of a row's lists
// is done by the re-rank kernel (k2_rerank.cu).
#include <cuda.h>
#ifdef LEMON_TC_PROFILE
#include <cstdio>   // in-kernel clock counters (build with LEMON_BUILD_DEFS=-DLEMON_TC_PROFILE; prints from CTA 0)
#define LEMON_PROF(x) x
#else
#define LEMON_PROF(x)
#endif

#include "lemon_common.cuh"

namespace lemon {

constexpr int kTcThreads = 128 + 2 * 128;      // 4 control warps + 2 epilogue groups of 4 warps
constexpr int kBM = 128;                       // query rows per CTA (TMEM lanes)
constexpr int kBK = 64;                        // K elements per smem chunk: 128 B rows, SWIZZLE_128B
constexpr int kAChunkBytes = kBM * kBK * 2;    // 16 KB
constexpr int kMaxSmem = 232448;               // 227 KB

struct TcParams {
  int64_t nq, m;
  int kchunks;          // d16 / 64
  int kres;             // query-tile K chunks resident in SMEM for the whole scan (the rest stream through the ring)
  int nseg;
  int64_t seg_len;      // columns per segment (multiple of BN)
  int nstage;
  int64_t n_items;      // row_tiles * nseg
  uint64_t* cand_keys;  // [nq_pad][nseg][2][kListCap]: the streaming key buffers ARE the output
  int32_t* cand_cnt;    // [nq_pad][nseg][2]
  float* cand_theta;    // [nq_pad][nseg][2]: every column not in the list has approximate value <= theta
  int cert;             // ladder: exceedance count that certifies a level (`keep` of the C ABI)
  int boot_tiles;       // 256-column tiles per item that only bootstrap the thresholds (8 or 16; 0 = off)
  int debug;            // experiments (-DLEMON_TC_EXPERIMENT builds): 1 = epilogue does no work, 2 = filter only
  int32_t* progress;    // [n_units] tile steps issued by every CTA pair (zeroed before the launch); nullptr = no pacing
  int pace_window;      // a pair may run at most this many tile steps ahead of the slowest pair
};

// ------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t local_addr) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(local_addr));
  return r;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 26)) __trap();   // seconds of spinning: fail loudly instead of hanging the GPU
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

template <int CG>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  if (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  } else {
    // executed by both CTAs of the pair; `bar` is the shared::cluster address of the LEADER's barrier
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  }
}

template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else         asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <int CG>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if (CG == 1) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  }
}
// tcgen05.commit: the mbarrier is arrived on when all MMAs issued so far by this thread are done.
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  } else {
    const uint16_t mask = 3;   // same barrier offset in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask)
                 : "memory");
  }
}

// tcgen05.ld 32x32b.x32: thread t of the warp receives columns [c, c+32) of TMEM lane (quadrant*32 + t).
// Issued asynchronously; tmem_wait_ld ties the registers to the wait so no use can be hoisted above it.
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.wait::ld.sync.aligned;"
      : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
        "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
        "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
        "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
      :
      : "memory");
}
// Branch-free survivor append: `if (v > theta) { *p++ = (bits(v) << 32) | ~idx; }` as straight-line instructions
// (compare, index word, predicated 64-bit store, predicated pointer bump).  The compiler's version of the same
// statement is a divergent branch around a 13-instruction body per element (BSSY / BRA / BSYNC), whose latency -- not
// its instruction count -- bounded the epilogue.  The high word is the RAW float bit pattern: the order-preserving
// transform every sort needs (two more instructions per element, executed for all 32 lanes whether they store or
// not) is applied by the readers instead (the re-rank kernel, and prune_exact below on the rare full list).
// Only the low word of the write pointer is bumped: a row's key list is 8 KB and 8 KB-aligned (checked on the
// host), so it never carries.
__device__ __forceinline__ void append_if_above(uint64_t& ptr, float v, float theta, uint32_t nidx) {
  asm volatile(
      "{\n .reg .pred p;\n .reg .b32 lo, ph;\n"
      " setp.gt.f32 p, %1, %2;\n"
      " @p st.global.v2.b32 [%0], {%4, %3};\n"
      " mov.b64 {lo, ph}, %0;\n"
      " @p add.u32 lo, lo, 8;\n"
      " mov.b64 %0, {lo, ph};\n}"
      : "+l"(ptr)
      : "f"(v), "f"(theta), "r"(__float_as_uint(v)), "r"(nidx)
      : "memory");
}
// list keys in memory carry raw float bits; registers that are sorted carry the order-preserving pattern
__device__ __forceinline__ uint64_t key_raw_to_ord(uint64_t k) {
  return k == 0ull ? 0ull : ((uint64_t(f2ord(__uint_as_float(uint32_t(k >> 32)))) << 32) | (k & 0xffffffffull));
}
__device__ __forceinline__ uint64_t key_ord_to_raw(uint64_t k) {
  return k == 0ull ? 0ull : ((uint64_t(__float_as_uint(ord2f(uint32_t(k >> 32)))) << 32) | (k & 0xffffffffull));
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (tile rows are 128 B, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return uint64_t((saddr & 0x3FFFF) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) |
         (uint64_t(2) << 61);
}

// ---------------------------------------------------------------- streaming top-k (epilogue)
constexpr int kBootTiles = 8;                  // 256-column tiles per item (4 per epilogue group) that bootstrap the row thresholds
constexpr int kBootMinTiles = 16;              // items shorter than this run without the bootstrap (their lists fill up once and are reduced exactly)
constexpr int kEpiGroups = 2;                  // epilogue warp groups; group g scans column half g of every accumulator tile

// Merge of two descending-sorted 64-key lists into the best 64, descending.  Lanes 0-7 hold list A
// (lane L: ranks 8L..8L+7); lanes 8-15 hold list B REVERSED (ascending over the 64 slots), so the 128 keys
// form a bitonic sequence; one half-cleaner step moves the best 64 into lanes 0-7 (still bitonic), six
// more steps sort them.
__device__ __forceinline__ void warp_merge_best64_desc(uint64_t (&key)[8], int lane) {
#pragma unroll
  for (int r = 0; r < 8; ++r) {                      // j = 64 elements = 8 lanes
    const uint64_t o = shfl_xor_u64(key[r], 8);
    const uint64_t mx = key[r] > o ? key[r] : o, mn = key[r] > o ? o : key[r];
    key[r] = (lane & 8) ? mn : mx;
  }
#pragma unroll
  for (int lm = 4; lm >= 1; lm >>= 1) {              // j = 32, 16, 8 elements
    const bool lower = (lane & lm) == 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const uint64_t o = shfl_xor_u64(key[r], lm);
      const uint64_t mx = key[r] > o ? key[r] : o, mn = key[r] > o ? o : key[r];
      key[r] = lower ? mx : mn;
    }
  }
#pragma unroll
  for (int j = 4; j >= 1; j >>= 1) {                 // in-lane
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if ((r & j) == 0) {
        const uint64_t a = key[r], b = key[r | j];
        key[r] = a > b ? a : b;
        key[r | j] = a > b ? b : a;
      }
    }
  }
}

// Threshold ladder.  A row's threshold must rise during the scan or every column would be appended, but finding
// the row's current 64th best (sort / select over its key list) costs thousands of cycles per row on one warp and
// made the slowest of the 16 warps that share an accumulator tile hold back the MMA pipe.  Instead every row
// keeps three trial levels t1 < t2 < t3 above its threshold and counts, per 32-column chunk, whether the chunk
// maximum exceeds each of them (one compare + one predicated add per level and chunk).  Every counted chunk holds
// a distinct column above the level, and every such column is in the row's list, so once a level has been counted
// kKeep times it is a certified threshold: theta moves up to it, the ladder shifts and a new top level is opened.
// The spacing adapts so that the count roughly halves from level to level.  Nothing is sorted, re-read or
// compacted during the scan; lists are simply long enough (kListCap) for the appends of a whole item, and the rare
// list that does fill up is reduced to its exact best kKeep below (which also re-seeds the ladder from exact ranks).
struct Ladder {
  float t1, t2, t3, delta;
  int c1, c2, c3;
};
__device__ __forceinline__ void ladder_restart(Ladder& ld, float theta) {
  ld.delta = fmaxf(ld.delta, 1e-6f * fmaxf(1.f, fabsf(theta)));
  ld.t1 = theta + ld.delta; ld.t2 = ld.t1 + ld.delta; ld.t3 = ld.t2 + ld.delta;
  ld.c1 = ld.c2 = ld.c3 = 0;
}
__device__ __forceinline__ void ladder_off(Ladder& ld) {
  ld.t1 = ld.t2 = ld.t3 = CUDART_INF_F; ld.delta = 0.f; ld.c1 = ld.c2 = ld.c3 = 0;
}
// per-lane slow path, entered when some lane has a certified or overtaken level
__device__ __forceinline__ void ladder_step(Ladder& ld, float& theta, int cert) {
  if (ld.c1 >= cert) {
    theta = fmaxf(theta, ld.t1);
    if (ld.c2 >= (3 * cert) / 4) ld.delta *= 1.5f;        // next level nearly certified too: levels too dense
    else if (ld.c2 < (5 * cert) / 16) ld.delta *= 0.75f;  // far from it: too sparse
    ld.delta = fmaxf(ld.delta, 1e-6f * fmaxf(1.f, fabsf(ld.t3)));
    ld.t1 = ld.t2; ld.c1 = ld.c2; ld.t2 = ld.t3; ld.c2 = ld.c3;
    ld.t3 = ld.t3 + ld.delta; ld.c3 = 0;
  }
  if (theta > -CUDART_INF_F && ld.delta > 0.f) {
    // levels overtaken by a threshold adopted from the other epilogue group: drop them but keep the counts of the
    // levels that are still above it; restart only when the whole ladder has been passed
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (ld.t1 <= theta) {
        ld.t1 = ld.t2; ld.c1 = ld.c2; ld.t2 = ld.t3; ld.c2 = ld.c3;
        ld.t3 = ld.t3 + ld.delta; ld.c3 = 0;
      }
    }
    if (ld.t1 <= theta) ladder_restart(ld, theta);
  }
}

// Exact reduction of one row's FULL key list `b` (cntL valid keys, up to kListCap) to its best kKeep = 64, written
// back sorted; thL becomes the cert-th value.  lv = values at ranks 3c/4, c/2, c/4 (ladder seeds with those exact
// counts).  The list is sorted 256 keys at a time and the running best 64 merged in.  All lanes pass the same arguments.
__device__ __forceinline__ float rank_value(const uint64_t (&sorted)[8], int rank, int lane) {   // rank < 64, lanes 0-7 hold 8 each
  uint64_t sel = 0ull;
#pragma unroll
  for (int i = 0; i < 8; ++i) if (i == (rank & 7)) sel = sorted[i];
  return key_val(shfl_u64(sel, rank >> 3));
}
__device__ __forceinline__ void prune_exact(uint64_t* b, int& cntL, float& thL, float (&lv)[3], int cert, int lane) {
  uint64_t best[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) best[r] = 0ull;
  for (int q0 = 0; q0 < cntL; q0 += kCap) {
    uint64_t key[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = q0 + lane * 8 + 2 * i;
      const ulonglong2 r = __ldcg(reinterpret_cast<const ulonglong2*>(b + e));
      key[2 * i] = e < cntL ? key_raw_to_ord(r.x) : 0ull;
      key[2 * i + 1] = e + 1 < cntL ? key_raw_to_ord(r.y) : 0ull;
    }
    warp_sort256_desc(key, lane);
    if (q0 > 0) {
      uint64_t rev[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) rev[r] = shfl_u64(key[7 - r], (15 - lane) & 31);   // lanes 8-15: this block's best 64, reversed
#pragma unroll
      for (int r = 0; r < 8; ++r) key[r] = lane < 8 ? best[r] : (lane < 16 ? rev[r] : 0ull);
      warp_merge_best64_desc(key, lane);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) best[r] = key[r];
  }
  __syncwarp();
  if (lane < kKeep / 8) {
#pragma unroll
    for (int i = 0; i < 8; i += 2)
      *reinterpret_cast<ulonglong2*>(b + lane * 8 + i) = make_ulonglong2(key_ord_to_raw(best[i]), key_ord_to_raw(best[i + 1]));
  }
  // cert columns >= the value at rank cert-1 stay in the list: that value is the new threshold
  const float vth = rank_value(best, cert - 1, lane);
  const float v1 = rank_value(best, (3 * cert) / 4 - 1, lane), v2 = rank_value(best, cert / 2 - 1, lane),
              v3 = rank_value(best, cert / 4 - 1, lane);
  if (cntL >= kKeep) {
    thL = fmaxf(thL, vth); cntL = kKeep;
    lv[0] = v1; lv[1] = v2; lv[2] = v3;
  } else {
    lv[0] = lv[1] = lv[2] = CUDART_INF_F;
  }
  __syncwarp();
}

// Reduces the lists of all lanes that are about to overflow (rare).
__device__ __forceinline__ void prune_full_rows(uint64_t* warp_keys, size_t row_stride, int& cnt, float& theta, Ladder& ld,
                                                int cert, int lane) {
  unsigned need = __ballot_sync(kFull, cnt > kListCap - 32);
  while (need) {
    const int L = __ffs(need) - 1;
    need &= need - 1;
    int cntL = __shfl_sync(kFull, cnt, L);
    float thL = __shfl_sync(kFull, theta, L);
    float lv[3];
    prune_exact(warp_keys + size_t(L) * row_stride, cntL, thL, lv, cert, lane);
    if (lane == L) {
      cnt = cntL; theta = thL;
      if (lv[0] < CUDART_INF_F) {
        ld.t1 = lv[0]; ld.t2 = lv[1]; ld.t3 = lv[2];
        ld.c1 = (3 * cert) / 4; ld.c2 = cert / 2; ld.c3 = cert / 4;
        ld.delta = fmaxf((lv[2] - lv[0]) * 0.625f, 1e-6f * fmaxf(1.f, fabsf(thL)));   // ranks 48 -> 16: 1.6 halvings
      }
    }
  }
  __syncwarp();
}

// Warp-wide bitonic sort of 32 floats (one per lane), descending: lane r ends up with rank r.
__device__ __forceinline__ float warp_sort32_desc_f(float x, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const float o = __shfl_xor_sync(kFull, x, j);
      const bool keep_max = ((lane & j) == 0) == ((lane & k) == 0);
      x = keep_max ? fmaxf(x, o) : fminf(x, o);
    }
  }
  return x;
}

// Threshold bootstrap.  During an item's first tiles a group only records the maxima of its 8-column groups
// (128 floats per row, kept in the row's still unused key buffer).  Every group maximum is a distinct DB
// column, so a value with at least 64 recorded maxima >= it is a valid threshold (at least 64 columns beat it),
// and it is nearly as tight as the true 64th best of those tiles.  The pivot comes from a sorted sample with
// an exact count.  The bootstrap tiles are re-scanned at the end of the item.
__device__ __forceinline__ void boot_select(const uint64_t* warp_keys, size_t row_stride, int nmax, int cert, float& theta, float& delta,
                                            int lane) {
  const bool two = nmax > 128;                       // 128 or 256 recorded maxima per row
  for (int L = 0; L < 32; ++L) {
    const float4* src = reinterpret_cast<const float4*>(warp_keys + size_t(L) * row_stride);
    const float4 v = __ldcg(src + lane);
    const float4 w = two ? __ldcg(src + 32 + lane) : make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    const float s = warp_sort32_desc_f(v.x, lane);   // systematic sample: every 4th (8th) maximum
    // sample ranks around the cert-th of the recorded maxima (cert = 64: 13/16/20 of 32 samples, 6/8/11 with 256 maxima)
    const int r0 = two ? cert / 8 - 2 : cert / 4 - 3, r1 = two ? cert / 8 : cert / 4, r2 = two ? cert / 8 + 3 : cert / 4 + 4;
    const float p0 = __shfl_sync(kFull, s, r0), p1 = __shfl_sync(kFull, s, r1), p2 = __shfl_sync(kFull, s, r2);
    int c0 = (v.x >= p0) + (v.y >= p0) + (v.z >= p0) + (v.w >= p0) + (w.x >= p0) + (w.y >= p0) + (w.z >= p0) + (w.w >= p0);
    int c1 = (v.x >= p1) + (v.y >= p1) + (v.z >= p1) + (v.w >= p1) + (w.x >= p1) + (w.y >= p1) + (w.z >= p1) + (w.w >= p1);
    int c2 = (v.x >= p2) + (v.y >= p2) + (v.z >= p2) + (v.w >= p2) + (w.x >= p2) + (w.y >= p2) + (w.z >= p2) + (w.w >= p2);
    c0 = __reduce_add_sync(kFull, c0); c1 = __reduce_add_sync(kFull, c1); c2 = __reduce_add_sync(kFull, c2);
    float mn = fminf(fminf(v.x, v.y), fminf(v.z, v.w));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(kFull, mn, o));
    const float th = c0 >= cert ? p0 : (c1 >= cert ? p1 : (c2 >= cert ? p2 : mn));   // mn: 128 maxima >= it
    // the filter keeps values STRICTLY above the threshold, and the >= 64 columns that certify `th` are only
    // collected later (re-scan): step one ulp down so that columns equal to `th` (mass ties!) are kept too
    const float th_open = __uint_as_float(f2ord_dec(th));
    // ladder spacing: between the threshold (>= 64 maxima above) and a high sample rank (~12-14 above) the tail
    // count drops ~4.6x = 2.2 halvings
    const float qv = __shfl_sync(kFull, s, two ? 1 : (cert >= 48 ? 3 : 2));
    if (lane == L) { theta = fmaxf(theta, th_open); delta = (qv - th) * (1.f / 2.2f); }
  }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------- DB-walk pacing
// All CTA pairs stream the same DB segment in the same order, and the DB (0.4 - 5 GB of fp16) does not fit in L2.
// Left alone the pairs drift apart by more tiles than L2 can hold, and then EVERY pair fetches EVERY tile from DRAM
// itself (ncu, 412 500 x 3.3 M x 768 launch: 6.75 TB of DRAM reads instead of ~0.11 TB, 37 % of the DRAM bandwidth --
// power that the 1000 W cap takes away from the tensor cores).  Pacing keeps the pairs within `pace_window` tile steps
// of the slowest one: every pair publishes the number of tile steps it has issued; the otherwise idle warp 3 of the
// leader CTA polls the minimum over all pairs and posts "issue up to here" in shared memory; the TMA producer
// checks that word before each tile.  It is a performance hint only: the producer's wait is bounded and on timeout
// the pair stops pacing for the rest of the launch (and reports "infinitely far", so nobody waits for it).
constexpr int kPaceWindow = 24;                // tile steps (d = 768: 24 x 393 KB = 9.4 MB of DB inside the window)
constexpr int kPaceMaxSpins = 4000;            // x ~100 ns: a pair waits at most ~0.4 ms per check before it gives up
constexpr int kPaceFar = 0x3fffffff;
__device__ __forceinline__ int ld_volatile_global(const int32_t* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_global(int32_t* p, int v) {
  asm volatile("st.volatile.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------ kernel
template <int CG, int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_db, const TcParams p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  if (smem_base & 1023u) __trap();   // SWIZZLE_128B tiles need 1024 B alignment
  // the shuffle tells ptxas that the warp index is warp-uniform (uniform registers / branches in the role code)
  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0;
  const int64_t unit = blockIdx.x / CG;            // CTA (pair) id
  const int64_t n_units = gridDim.x / CG;
  constexpr int kBRows = BN / CG;                  // DB rows this CTA stages per tile
  constexpr uint32_t kStageBytes = kBRows * kBK * 2;
  constexpr uint32_t kTmemCols = 512;              // all of TMEM: kNBuf accumulator buffers of BN columns
  constexpr uint32_t kNBuf = 512 / BN;             // 2 (BN=256) or 4 (BN=128)
  constexpr int kGrpCols = BN / kEpiGroups;        // accumulator columns each epilogue group scans per tile
  const int kBoot = p.boot_tiles * 256 / BN;       // bootstrap tiles per item: 128 (or 256) group maxima per epilogue group

  const uint32_t a_smem = smem_base;
  const uint32_t b_smem = a_smem + uint32_t(p.kres) * kAChunkBytes;
  const uint32_t bar_base = b_smem + uint32_t(p.nstage) * kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.nstage + s); };
  const uint32_t a_full = bar_base + 8u * (2 * p.nstage);
  const uint32_t a_empty = a_full + 8;
  const uint32_t tmem_full = a_full + 16;    // [4]
  const uint32_t tmem_empty = a_full + 48;   // [4]
  const uint32_t tmem_slot = a_full + 80;
  // per-row threshold exchange between the two epilogue groups: [2][128] x (item tag << 32 | float bits)
  volatile uint64_t* th_sh = reinterpret_cast<volatile uint64_t*>(smem_raw + (a_full + 96 - smem_u32(smem_raw)));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  volatile int* pace_go = reinterpret_cast<volatile int*>(smem_raw + (a_full + 88 - smem_u32(smem_raw)));     // issue limit
  volatile int* pace_done = pace_go + 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nstage; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int b = 0; b < int(kNBuf); ++b) { mbar_init(tmem_full + 8 * b, 1); mbar_init(tmem_empty + 8 * b, CG * 4 * kEpiGroups); }
    *pace_go = p.pace_window;
    *pace_done = 0;
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<CG>(tmem_slot, kTmemCols);
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      const uint32_t full0 = (CG == 2) ? mapa_rank0(full_bar(0)) : full_bar(0);   // leader's barriers
      const uint32_t afull_l = (CG == 2) ? mapa_rank0(a_full) : a_full;
      int stage = 0; uint32_t phase = 0; uint32_t it = 0;
      int gstep = 0;                                              // tile steps issued so far in this launch
      bool pace = p.progress != nullptr && cta_rank == 0;
      for (int64_t item = unit; item < p.n_items; item += n_units, ++it) {
        const int64_t rt = item / p.nseg, seg = item % p.nseg;
        const int64_t col0 = seg * p.seg_len;
        const int64_t col1 = min(p.m, col0 + p.seg_len);
        const int64_t ntiles = col1 > col0 ? (col1 - col0 + BN - 1) / BN : 0;
        if (ntiles == 0) { --it; continue; }
        const int row0 = int(rt * (kBM * CG) + cta_rank * kBM);
        // query tile: wait until the previous item's MMAs have drained it
        mbar_wait(a_empty, (it & 1) ^ 1);
        if (cta_rank == 0) mbar_arrive_expect_tx(a_full, uint32_t(p.kres) * kAChunkBytes * CG);
        for (int kc = 0; kc < p.kres; ++kc)
          tma_load_2d<CG>(a_smem + kc * kAChunkBytes, &map_q, afull_l, kc * kBK, row0);
        const int64_t nsteps = ntiles + (ntiles >= kBootMinTiles ? kBoot : 0);   // bootstrap tiles are scanned twice
        for (int64_t i = 0; i < nsteps; ++i, ++gstep) {
          if (pace) {                                             // stay within the window of the slowest pair
            st_volatile_global(p.progress + unit, gstep);
            int spins = 0;
            while (gstep >= *pace_go && spins < kPaceMaxSpins) { __nanosleep(100); ++spins; }
            if (spins >= kPaceMaxSpins) { pace = false; st_volatile_global(p.progress + unit, kPaceFar); }
          }
          const int64_t t = i < ntiles ? i : i - ntiles;
          const int dbrow = int(col0 + t * BN + cta_rank * kBRows);
          for (int kc = 0; kc < p.kchunks; ++kc) {
            if (kc >= p.kres) {   // non-resident query chunk: one ring stage (same size as a DB stage)
              mbar_wait(empty_bar(stage), phase ^ 1);
              if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), kAChunkBytes * CG);
              tma_load_2d<CG>(b_smem + stage * kStageBytes, &map_q, full0 + 8u * stage, kc * kBK, row0);
              if (++stage == p.nstage) { stage = 0; phase ^= 1; }
            }
            mbar_wait(empty_bar(stage), phase ^ 1);
            if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), kStageBytes * CG);
            tma_load_2d<CG>(b_smem + stage * kStageBytes, &map_db, full0 + 8u * stage, kc * kBK, dbrow);
            if (++stage == p.nstage) { stage = 0; phase ^= 1; }
          }
        }
      }
      if (p.progress != nullptr && cta_rank == 0) { st_volatile_global(p.progress + unit, kPaceFar); *pace_done = 1; }
    }
  } else if (warp == 3) {
    // =============================== pacer (leader CTA): slowest pair's progress -> issue limit ===============================
    if (p.progress != nullptr && cta_rank == 0) {
      while (__shfl_sync(kFull, *pace_done, 0) == 0) {          // lane 0's view: the exit must be warp-uniform
        int mn = kPaceFar;
        for (int u = lane; u < int(n_units); u += 32) mn = min(mn, ld_volatile_global(p.progress + u));
        mn = __reduce_min_sync(kFull, mn);
        if (lane == 0) *pace_go = mn >= kPaceFar ? 0x7fffffff : mn + p.pace_window;
        __nanosleep(400);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA) ===============================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = (1u << 4) | (uint32_t(BN >> 3) << 17) | (uint32_t((kBM * CG) >> 4) << 24);
      int stage = 0; uint32_t phase = 0; uint32_t it = 0; uint32_t tc = 0;
      LEMON_PROF(long long pf_empty = 0; long long pf_full = 0; const long long pf_t0 = clock64();)
      for (int64_t item = unit; item < p.n_items; item += n_units, ++it) {
        const int64_t seg = item % p.nseg;
        const int64_t col0 = seg * p.seg_len;
        const int64_t col1 = min(p.m, col0 + p.seg_len);
        const int64_t ntiles = col1 > col0 ? (col1 - col0 + BN - 1) / BN : 0;
        if (ntiles == 0) { --it; continue; }
        mbar_wait(a_full, it & 1);
        tc_fence_after();
        const int64_t nsteps = ntiles + (ntiles >= kBootMinTiles ? kBoot : 0);
        for (int64_t i = 0; i < nsteps; ++i, ++tc) {
          const uint32_t buf = tc % kNBuf, use = tc / kNBuf;
          LEMON_PROF(const long long pf_a = clock64();)
          mbar_wait(tmem_empty + 8 * buf, (use & 1) ^ 1);
          tc_fence_after();
          LEMON_PROF(pf_empty += clock64() - pf_a;)
          const uint32_t d_tmem = tmem_base + buf * BN;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            uint32_t a_addr = a_smem + kc * kAChunkBytes;
            int a_stage = -1;
            if (kc >= p.kres) {                      // streamed query chunk
              mbar_wait(full_bar(stage), phase);
              a_stage = stage;
              a_addr = b_smem + stage * kStageBytes;
              if (++stage == p.nstage) { stage = 0; phase ^= 1; }
            }
            LEMON_PROF(const long long pf_b = clock64();)
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            LEMON_PROF(pf_full += clock64() - pf_b;)
            const uint64_t adesc = make_smem_desc(a_addr);
            const uint64_t bdesc = make_smem_desc(b_smem + stage * kStageBytes);
#pragma unroll
            for (int k4 = 0; k4 < kBK / 16; ++k4)   // +32 B per UMMA_K inside the 128 B swizzle row
              umma_f16<CG>(d_tmem, adesc + uint64_t(2 * k4), bdesc + uint64_t(2 * k4), idesc, (kc | k4) != 0);
            if (a_stage >= 0) umma_commit<CG>(empty_bar(a_stage));
            umma_commit<CG>(empty_bar(stage));       // frees the DB stage (both CTAs)
            if (++stage == p.nstage) { stage = 0; phase ^= 1; }
          }
          umma_commit<CG>(tmem_full + 8 * buf);      // accumulator ready (both CTAs)
        }
        umma_commit<CG>(a_empty);                    // query tile drained (both CTAs)
      }
      LEMON_PROF(if (blockIdx.x == 0) printf("K1 MMA issuer: total %lld cycles, waiting for a free accumulator %lld, for a full stage %lld, tiles %u\n",
                                             clock64() - pf_t0, pf_empty, pf_full, tc);)
    }
  } else if (warp >= 4) {
    // =============================== epilogue: streaming top-k ===============================
    // Two groups of 4 warps; group g scans columns [g*BN/2, (g+1)*BN/2) of every accumulator tile.  (Groups that
    // alternate whole tiles cannot scan while their buffer is being refilled: the step time was T_mma + T_scan
    // per two tiles; with split tiles it is max(T_mma, T_scan/2 + pruning) per tile.)  Both groups track the
    // same query rows; each keeps its own key buffer and they exchange thresholds through shared memory (a
    // threshold certified by either group is a valid filter for both).
    const int grp = (warp - 4) >> 2;
    {
    const int quad = warp & 3;
    const int row_local = quad * 32 + lane;
    const uint32_t tmem_lane = uint32_t(quad * 32) << 16;
    const size_t row_stride = size_t(p.nseg) * kEpiGroups * kListCap;     // keys between consecutive query rows
    const uint32_t tempty0 = (CG == 2) ? mapa_rank0(tmem_empty) : tmem_empty;
    uint32_t tc = 0, it = 0;
    LEMON_PROF(long long pf_wait = 0; long long pf_scan = 0; long long pf_max = 0; const long long pf_t0 = clock64();)
    for (int64_t item = unit; item < p.n_items; item += n_units, ++it) {
      const int64_t rt = item / p.nseg, seg = item % p.nseg;
      const int64_t col0 = seg * p.seg_len;
      const int64_t col1 = min(p.m, col0 + p.seg_len);
      const int64_t ntiles = col1 > col0 ? (col1 - col0 + BN - 1) / BN : 0;
      const int64_t row = rt * (kBM * CG) + cta_rank * kBM + row_local;
      const uint64_t tag = uint64_t(uint32_t(item) + 1u) << 32;
      float theta = (p.debug & 2) ? CUDART_INF_F : -CUDART_INF_F;
      int cnt = 0;
      Ladder ld;
      ladder_off(ld);
      // this item's key buffers live in the output array: row-major [row][seg][group][kListCap]
      uint64_t* warp_keys = p.cand_keys + ((rt * (kBM * CG) + cta_rank * kBM + quad * 32) * p.nseg + seg) * (kEpiGroups * kListCap) +
                            size_t(grp) * kListCap;
      uint64_t* my_keys = warp_keys + size_t(lane) * row_stride;
      const uint32_t keys_lo = uint32_t(reinterpret_cast<uintptr_t>(my_keys));
      th_sh[grp * kBM + row_local] = tag | __float_as_uint(theta);
      const int nboot = ntiles >= kBootMinTiles ? kBoot : 0;
      const int64_t nsteps = ntiles + nboot;
      int bcount = 0;
      for (int64_t i = 0; i < nsteps; ++i, ++tc) {
        const int64_t t = i < ntiles ? i : i - ntiles;
        const uint32_t buf = tc % kNBuf, use = tc / kNBuf;
        LEMON_PROF(const long long pf_a = clock64();)
        mbar_wait(tmem_full + 8 * buf, use & 1);
        tc_fence_after();
        LEMON_PROF(const long long pf_b = clock64(); pf_wait += pf_b - pf_a;)
        if (i < nboot && bcount >= 0) {
          // ---- bootstrap tile: record the 8-column group maxima only (no appends, no prunes)
          const uint32_t taddr_b = tmem_base + tmem_lane + buf * BN + grp * kGrpCols;
          float4* bf = reinterpret_cast<float4*>(my_keys) + bcount * (kGrpCols / 32);
#pragma unroll 1
          for (int ch = 0; ch < kGrpCols / 32; ++ch) {
            uint32_t r[32];
            tmem_ld32_async(taddr_b + ch * 32, r);
            tmem_wait_ld(r);
            float gm[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float a = fmax3(__uint_as_float(r[8 * g]), __uint_as_float(r[8 * g + 1]), __uint_as_float(r[8 * g + 2]));
              const float bq = fmax3(__uint_as_float(r[8 * g + 3]), __uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
              gm[g] = fmax3(a, bq, fmaxf(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7])));
            }
            bf[ch] = make_float4(gm[0], gm[1], gm[2], gm[3]);
          }
          ++bcount;
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(tempty0 + 8 * buf);
            else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty + 8 * buf) : "memory");
          }
          if (bcount * (kGrpCols / 8) >= p.boot_tiles * 16) {   // all of this group's bootstrap tiles are recorded
            float delta = 0.f;
            boot_select(warp_keys, row_stride, p.boot_tiles * 16, p.cert, theta, delta, lane);
            if (!(p.debug & 2) && theta > -CUDART_INF_F) { ld.delta = fmaxf(delta, 0.f); ladder_restart(ld, theta); }
            __syncwarp();
            th_sh[grp * kBM + row_local] = tag | __float_as_uint(theta);
            bcount = -1000000;
          }
          continue;
        }
        {   // adopt the other group's threshold for this row if it is tighter (same item only)
          const uint64_t o = th_sh[(grp ^ 1) * kBM + row_local];
          if ((o >> 32) == (tag >> 32)) theta = fmaxf(theta, __uint_as_float(uint32_t(o)));
        }
        const int64_t colb = col0 + t * BN + grp * kGrpCols;          // this group's half of the tile
        const int valid = int(max(int64_t(0), min(int64_t(kGrpCols), col1 - colb)));
        const uint32_t taddr = tmem_base + tmem_lane + buf * BN + grp * kGrpCols;
        if (!(p.debug & 1)) {
          const int nchunks = (valid + 31) >> 5;
          const bool partial = valid < kGrpCols;     // only the last tile of a segment
#pragma unroll 1
          for (int ch = 0; ch < nchunks; ++ch) {
            const int c = ch * 32;
            uint32_t r[32];
            tmem_ld32_async(taddr + c, r);
            tmem_wait_ld(r);
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (partial) {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (c + j >= valid) v[j] = -CUDART_INF_F;
            }
            float gm[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const float a = fmax3(v[8 * g], v[8 * g + 1], v[8 * g + 2]);
              const float bq = fmax3(v[8 * g + 3], v[8 * g + 4], v[8 * g + 5]);
              gm[g] = fmax3(a, bq, fmaxf(v[8 * g + 6], v[8 * g + 7]));
            }
            const float mx = fmaxf(fmaxf(gm[0], gm[1]), fmaxf(gm[2], gm[3]));
            // warp-uniform gating (votes), so that groups without a survivor in ANY lane are really skipped
            if (__any_sync(kFull, mx > theta)) {
              ld.c1 += mx > ld.t1; ld.c2 += mx > ld.t2; ld.c3 += mx > ld.t3;
              if (__any_sync(kFull, ld.c1 >= p.cert || (ld.t1 <= theta && ld.delta > 0.f))) ladder_step(ld, theta, p.cert);
              const uint32_t nidx0 = ~uint32_t(colb + c);   // ~(idx0 + j) == ~idx0 - j
              uint64_t wp = reinterpret_cast<uint64_t>(my_keys + cnt);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                if (__any_sync(kFull, gm[g] > theta)) {
#pragma unroll
                  for (int j = 8 * g; j < 8 * g + 8; ++j) append_if_above(wp, v[j], theta, nidx0 - uint32_t(j));
                }
              }
              cnt = int((uint32_t(wp) - keys_lo) >> 3);
            }
            __syncwarp();
            if (__any_sync(kFull, cnt > kListCap - 32)) prune_full_rows(warp_keys, row_stride, cnt, theta, ld, p.cert, lane);   // must not overflow
          }
        }
        LEMON_PROF({ const long long pf_c = clock64() - pf_b; pf_scan += pf_c; if (pf_c > pf_max) pf_max = pf_c; })
        // publish this row's threshold and hand the accumulator buffer back to the MMA warp
        th_sh[grp * kBM + row_local] = tag | __float_as_uint(theta);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster(tempty0 + 8 * buf);
          else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty + 8 * buf) : "memory");
        }
      }
      // ---- item done: the row's key buffer already sits in the output; publish its length and threshold.
      // (The exact top-kp selection over the union of the lists happens in the re-rank kernel, where thousands
      // of warps hide its latency; here it would sit on the critical path of one persistent warp.)
      {
        const size_t li = (size_t(row) * p.nseg + seg) * kEpiGroups + grp;
        p.cand_cnt[li] = cnt;
        p.cand_theta[li] = theta;
      }
      __syncwarp();
    }
    LEMON_PROF(if (lane == 0 && blockIdx.x == 0) printf("K1 epilogue warp %d: total %lld cycles, waiting for an accumulator %lld, scanning %lld (longest tile %lld), tiles %u\n",
                                                       warp, clock64() - pf_t0, pf_wait, pf_scan, pf_max, tc);)
    }
  }

  // =============================== teardown ===============================
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(lemon_ctx* ctx, CUtensorMap* map, const void* base, int64_t rows, int d16, int box_rows) {
  if (!ctx->encode_tiled) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
      return lemon_set_error(ctx, LEMON_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    ctx->encode_tiled = fn;
  }
  const cuuint64_t gdim[2] = {cuuint64_t(d16), cuuint64_t(rows)};
  const cuuint64_t gstride[1] = {cuuint64_t(d16) * 2};
  const cuuint32_t box[2] = {cuuint32_t(kBK), cuuint32_t(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return lemon_set_error(ctx, LEMON_ERR_CUDA, "cuTensorMapEncodeTiled failed: %d", int(r));
  return LEMON_OK;
}

template <int CG, int BN>
static int launch_tc(lemon_ctx* ctx, const void* q16, const void* db16, int64_t nq, int64_t m, int d16, int nseg, int keep,
                     uint64_t* cand_keys, int32_t* cand_cnt, float* cand_theta, cudaStream_t stream) {
  const int kchunks = d16 / kBK;
  const uint32_t stage_bytes = (BN / CG) * kBK * 2;
  // Query-tile residency: all K chunks stay in SMEM when that leaves a ring of >= 4 DB stages; otherwise (d16 > 640)
  // the last chunks are re-streamed through the ring with every DB tile -- a 2-stage ring (32 KB in flight per CTA)
  // cannot cover the L2 latency, 4 stages can, at the price of 17 % more L2 -> SMEM traffic (d16 = 768: 10 of 12
  // chunks resident; measured 907 -> 1082 TFLOP/s on a 50000 x 400000 x 768 launch; 9 / 8 / 6 resident: 1064 / 1042 / 1016).
  int kres = kchunks;
  if (stage_bytes == uint32_t(kAChunkBytes)) {
    const int fit = int((int64_t(kMaxSmem) - 2368 - 4 * int64_t(stage_bytes)) / kAChunkBytes);
    if (fit < kres) kres = fit < 1 ? 1 : fit;
    if (ctx->tune_kres >= 1 && ctx->tune_kres <= kchunks) kres = ctx->tune_kres;
  }
  const uint32_t a_bytes = uint32_t(kres) * kAChunkBytes;
  // dynamic shared memory starts 1024-aligned (no static __shared__ in this kernel; checked on the device)
  const int64_t avail = int64_t(kMaxSmem) - 2368 /*barriers + threshold exchange*/ - a_bytes;
  int nstage = int(avail / stage_bytes);
  if (nstage > 8) nstage = 8;
  if (nstage < 2) return lemon_set_error(ctx, LEMON_ERR_INVALID, "knn_candidates: d16=%d leaves no room for a DB ring", d16);
  const size_t smem = a_bytes + size_t(nstage) * stage_bytes + 2368;

  TcParams p;
  p.nq = nq; p.m = m; p.kchunks = kchunks; p.kres = kres; p.nstage = nstage;
  int64_t seg_len = (m + nseg - 1) / nseg;
  seg_len = (seg_len + BN - 1) / BN * BN;
  // segments past the end of the DB are legal: they emit empty (-inf, -1) lists
  p.nseg = nseg; p.seg_len = seg_len;
  const int64_t row_tiles = (nq + kBM * CG - 1) / (kBM * CG);
  p.n_items = row_tiles * nseg;
  p.cand_keys = cand_keys; p.cand_cnt = cand_cnt; p.cand_theta = cand_theta;
  // pacing (see "DB-walk pacing" above): only when the DB is scanned by several pairs and is long enough to drift
  p.progress = nullptr;
  p.pace_window = ctx->tune_pace >= 0 ? ctx->tune_pace : kPaceWindow;
  // tuning knobs: library defaults unless the ctx was created by a -DLEMON_TC_EXPERIMENT build (capi.cu)
  p.debug = ctx->tune_debug > 0 ? ctx->tune_debug : 0;
  p.cert = (ctx->tune_cert >= 16 && ctx->tune_cert <= kKeep) ? ctx->tune_cert : keep;
  p.boot_tiles = (ctx->tune_boot == 0 || ctx->tune_boot == 8 || ctx->tune_boot == 16) ? ctx->tune_boot : kBootTiles;

  int64_t units = ctx->num_sms / CG;
  if (units > p.n_items) units = p.n_items;
  const unsigned grid = unsigned(units * CG);
  CUtensorMap map_q, map_db;
  int rc = make_map(ctx, &map_q, q16, nq, d16, kBM);
  if (rc) return rc;
  rc = make_map(ctx, &map_db, db16, m, d16, BN / CG);
  if (rc) return rc;

  if (p.pace_window > 0 && ctx->tc_scratch && units > 1 && nseg == 1 && seg_len / BN >= 4 * p.pace_window) {
    constexpr int kSlots = 16, kSlotInts = 256;          // a ring of progress arrays: launches in flight never share one
    p.progress = static_cast<int32_t*>(ctx->tc_scratch) + (ctx->tc_launch_seq++ % kSlots) * kSlotInts;
    LEMON_CUDA_CHECK(ctx, cudaMemsetAsync(p.progress, 0, kSlotInts * sizeof(int32_t), stream));
  }
  auto kern = knn_tc_kernel<CG, BN>;
  LEMON_CUDA_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  LEMON_CUDA_CHECK(ctx, cudaLaunchKernelEx(&cfg, kern, map_q, map_db, p));
  ctx->launches++;
  return LEMON_OK;
}

}  // namespace lemon

extern "C" int lemon_knn_candidates(lemon_ctx* ctx, const void* q16, const void* db16, int64_t nq, int64_t m,
                                    int d16, int nseg, int cta_group, int keep, uint64_t* cand_keys, int32_t* cand_cnt,
                                    float* cand_theta, void* stream) {
  using namespace lemon;
  if (!ctx) return LEMON_ERR_INVALID;
  if (!q16 || !db16 || !cand_keys || !cand_cnt || !cand_theta || nq < 0 || m < 1 || d16 < 64 || d16 % 64 || d16 > LEMON_MAX_D16_OPERAND ||
      nseg < 1 || nseg > 64 || keep < 0 || (keep > 0 && keep < 16) || keep > kKeep || m >= (int64_t(1) << 31) || (uintptr_t(q16) & 15) || (uintptr_t(db16) & 15) ||
      (uintptr_t(cand_keys) & (kListCap * 8 - 1)))
    return lemon_set_error(ctx, LEMON_ERR_INVALID, "knn_candidates: bad args (d16 %% 64 == 0, d16 <= %d, 16B-aligned operands, 8 KB-aligned cand_keys)", LEMON_MAX_D16_OPERAND);
  if (nq == 0) return LEMON_OK;
  if (keep == 0) keep = kKeep;
  if (cta_group == 0) cta_group = 2;   // CTA pairs: half the SMEM/L2 operand traffic per MMA
  cudaStream_t st = (cudaStream_t)stream;
  LEMON_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
  int bn = ctx->tune_bn > 0 ? ctx->tune_bn : 0;      // experiments: force the DB tile width (128 or 256)
  if (cta_group == 1) {
    if (bn == 0) bn = d16 <= 512 ? 256 : 128;
    if (bn == 256 && d16 <= 512) return launch_tc<1, 256>(ctx, q16, db16, nq, m, d16, nseg, keep, cand_keys, cand_cnt, cand_theta, st);
    return launch_tc<1, 128>(ctx, q16, db16, nq, m, d16, nseg, keep, cand_keys, cand_cnt, cand_theta, st);
  }
  if (cta_group == 2) {
    if (bn == 128) return launch_tc<2, 128>(ctx, q16, db16, nq, m, d16, nseg, keep, cand_keys, cand_cnt, cand_theta, st);
    return launch_tc<2, 256>(ctx, q16, db16, nq, m, d16, nseg, keep, cand_keys, cand_cnt, cand_theta, st);
  }
  return lemon_set_error(ctx, LEMON_ERR_INVALID, "knn_candidates: cta_group must be 0, 1 or 2");
}
