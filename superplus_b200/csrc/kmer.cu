// K1..K6: 2-bit packing, canonical k-mer extraction, open-addressed HBM table build,
// ONT search with anchor compaction, ONT-side multiplicity and statistics.
//
// Reference semantics restated (paths relative to /root/reference/gap_closer):
//   base code      bio.h:22-24            A=0 C=1 T=2 G=3 (N -> 3), complement = code ^ 2
//   rolling k-mer  kseq1.h:28-59          fwd = ((fwd<<2)&mask)|b ; rc = (rc>>2)|(comp(b)<<2(k-1))
//   canonical      kmer.c:86-94, ont.c:161-167   fwd < rc ? (fwd, flag 0) : (rc, flag REV)
//   table          kmer.c:124-152 -> hash.c:113-152   distinct keys with multiplicity
//   search         ont.c:141-204          anchor <=> key present with multi == 1
//   ONT counts     ont.c:230-254          multiplicity of each anchored k-mer over all reads
//   stats          kmer.c:265-312
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <functional>
#include <thread>

#include <mutex>
#include <immintrin.h>

#include "gcg_internal.cuh"
#include "host_par.h"
#include "kmer_dev.cuh"

// =============================================================================================
// K1  ASCII -> 2 bit
// =============================================================================================
__global__ void __launch_bounds__ (256)
k1_pack_kernel (const uint4 * __restrict__ ascii, uint64_t * __restrict__ packed, int64_t n_words)
{
  int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t w = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
    uint4 a = __ldg (ascii + 2 * w), b = __ldg (ascii + 2 * w + 1);
    uint32_t hi = (pack4 (a.x) << 24) | (pack4 (a.y) << 16) | (pack4 (a.z) << 8) | pack4 (a.w);
    uint32_t lo = (pack4 (b.x) << 24) | (pack4 (b.y) << 16) | (pack4 (b.z) << 8) | pack4 (b.w);
    packed[w] = ((uint64_t) hi << 32) | lo;
  }
}

// =============================================================================================
// K2+K3  contig chop + table insert
// =============================================================================================
// One thread per k-mer start position (lane j of a warp takes position j of the warp's word): the
// inserts of a word are independent chains of load -> CAS -> store, so they are spread over the
// lanes instead of being walked by one thread (a 4.6 Mb scaffold is only 144 k words).
// ILP > 1 (GCG_BUILD_ILP=2|4, an experiment kept for A/B timing, not the default): every thread carries
// ILP such chains at once (words w, w + stride, ...) — the bucket loads are issued together, then the
// claims, and only then is the first CAS result looked at; a chain that does not end with its first
// claim (duplicate, lost slot, full bucket) falls back to the general probe loop table_insert.  ncu
// (profiles/r01_ncu_summary.txt, k23) shows 87 % of the stall samples on the long scoreboard, most of
// them behind the CAS, which suggested more inserts in flight; measured, it is the opposite — cfg2
// 0.121 ms (ILP 1) -> 0.29 ms (2) -> 0.33 ms (4), cfg5s 1.45 -> 2.42 -> 2.94 ms: the kernel is not
// short of parallelism, it is at the random-sector rate of the memory system (7.2 M DRAM sectors per
// launch = 60 G sectors/s, above the 37-52 G/s the gather microbenchmark reaches on tables beyond the
// L2; the 106 MB table does not stay L2 resident while it is being written: 56 % sector hits), and
// more requests in flight only lengthen the queues and lose more claims.
template <int ILP>
__global__ void __launch_bounds__ (256, ILP == 1 ? 8 : ILP == 2 ? 6 : 4)
k23_build_kernel (const uint64_t * __restrict__ packed, const int64_t * __restrict__ woff,
                  const int32_t * __restrict__ len, const int32_t * __restrict__ tile_seq, int64_t n_seq, int64_t n_words, int k,
                  unsigned long long * __restrict__ keys, unsigned long long * __restrict__ vals, uint32_t n_bucket)
{
  const int lane = threadIdx.x & 31;
  const int64_t wstride = (int64_t) gridDim.x * (blockDim.x >> 5);
  for (int64_t w0 = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w0 < n_words; w0 += wstride * ILP) {
    unsigned long long key[ILP], val[ILP];
    uint32_t b[ILP];
    int act[ILP];                                   // 0 nothing, 1 claim slot idx, 2 general loop
    int idx[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      const int64_t w = w0 + u * wstride;
      act[u] = 0; idx[u] = 0; key[u] = 0; val[u] = 0; b[u] = 0;
      if (w < n_words) {
        int64_t s = find_seq_from (woff, n_seq, w, __ldg (tile_seq + (w >> 5)));
        int32_t L = __ldg (len + s);
        int32_t p0 = (int32_t) ((w - __ldg (woff + s)) << 5);
        int32_t nvalid = L - k + 1 - p0;             // number of valid starts in this word
        if (lane < nvalid) {
          bool fw;
          key[u] = key_at (__ldg (packed + w), __ldg (packed + w + 1), lane, k, &fw);
          val[u] = ((unsigned long long) s << 32) | ((unsigned long long) (uint32_t) (p0 + lane) << 1) | (fw ? 0ULL : 1ULL);
          b[u] = __umulhi (kmer_hash32 (key[u] - 1ULL), n_bucket);
          act[u] = 2;
        }
      }
    }
    if (ILP == 1) {
      if (act[0]) table_insert (keys, vals, n_bucket, key[0], val[0]);
      continue;
    }
    bucket4 q[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u)
      if (act[u]) q[u] = ld_bucket_cg (keys + 4ULL * b[u]);
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      if (!act[u]) continue;
      const unsigned long long cur[4] = {q[u].a, q[u].b, q[u].c, q[u].d};
#pragma unroll
      for (int i = 3; i >= 0; --i) {                // the FIRST empty slot or match decides (slots fill front to back)
        if (cur[i] == 0ULL) { act[u] = 1; idx[u] = i; }
        else if ((cur[i] & GCG_KEY_MASK) == key[u]) { act[u] = 2; idx[u] = i; }
      }
    }
    unsigned long long old[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      old[u] = 1ULL;
      if (act[u] == 1) old[u] = atomicCAS (keys + 4ULL * b[u] + idx[u], 0ULL, key[u]);
    }
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      if (act[u] == 1 && old[u] == 0ULL) vals[4ULL * b[u] + idx[u]] = val[u];
      else if (act[u]) table_insert (keys, vals, n_bucket, key[u], val[u]);      // duplicate, lost claim or full bucket
    }
  }
}

// =============================================================================================
// K4+K5  ONT chop + lookup + multi==1 filter
//   One thread per packed word (32 k-mer start positions), one warp per tile of 32 words.  Each
//   probe is ONE 256-bit load of a 32-byte bucket (LDG.E.256: one L1 wavefront and one L2/DRAM
//   sector per lookup), issued four at a time so every thread keeps four sectors in flight.  The
//   main loop is branch-free per probe: "present with multiplicity 1" is one masked 64-bit compare
//   per slot, and the few probes whose key may have overflowed into a later bucket (bit 62 of slot
//   `hash & 3` is set and the home bucket did not match) are only noted in a second mask; they are
//   finished after the loop, dealt round-robin over the warp's lanes, so no lane waits on another
//   lane's dependent load inside the hot loop.  The kernel writes a 32-bit anchor mask per word
//   (plain coalesced store); the anchors themselves, and the ONT-side multiplicity (ont.c:245), are
//   produced by hits_emit_kernel after a prefix sum over the masks, which re-probes just the
//   anchored positions (a few percent).
// =============================================================================================
#define K4_UNROLL 4
#define K4_WARPS 8

// FILTER (tables beyond the L2, where a probe is a random HBM access at ~40 G/s): one L2-resident
// 32-bit filter word is tested first and the bucket is only loaded when all of the key's bits are
// set.  The filter holds the keys that are present exactly once, i.e. the only ones that anchor, so
// with 96 % of ONT k-mers absent (sequencing errors) most probes never leave the L2.
// KC: k as a compile-time constant (25, the gc CLI's; 31, configs[3]) so that the rolling masks and
// shifts are immediates; 0 = any k at run time.
// (32 registers, 8 blocks per SM.  The kernel sits at 85 % of the L1 wavefront rate: a random 32-byte
// bucket load is one wavefront per lane, 138 M of them at one per clock and SM are 0.47 ms at cfg2.
// Asking for fewer, fatter blocks to keep four loads in flight per thread measured slower, 0.60 ms.)
// The probe of one 32-word tile by one warp: returns the lane's 32-bit anchor mask of word (tile << 5) + lane
// (bit j: the canonical k-mer at position p0 + j is in the table with multiplicity 1).  *seq_out / *p0_out
// receive the word's read index and its first position (undefined for words past the end).
struct k4_smem { uint32_t excl[K4_WARPS][33], pend[K4_WARPS][32], add[K4_WARPS][32]; };

// PART (remote-probe search of a hash-partitioned table): the key's owner partition picks the arrays — a peer GPU's,
// reached over NVLink — and the bucket count; only probes that pass the (local, replicated) union filter leave the GPU
// CG (streaming search): the packed words are read with ld.global.cg — they may be arriving by DMA while the launch runs
template <bool FILTER, int KC, bool PF, bool PART = false, bool CG = false>
__device__ __forceinline__ uint32_t
k4_probe_tile (k4_smem & sm, const unsigned long long * __restrict__ vals, const int64_t tile, const uint64_t * __restrict__ packed, const int64_t * __restrict__ woff,
               const int32_t * __restrict__ len, const int32_t * __restrict__ tile_seq, const int64_t n_seq, const int64_t n_words, const int k,
               const unsigned long long * __restrict__ keys, const uint32_t n_bucket,
               const uint32_t * __restrict__ filter, const uint32_t filter_words, const int filter_k3,
               uint64_t * pk_out, int32_t * seq_out, int32_t * p0_out,
               const gcg_part_desc * __restrict__ parts = nullptr, const uint32_t n_part = 1)
{
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t w = (tile << 5) + lane;
  uint32_t mymask = 0, pend = 0;
  int nvalid = 0;
  uint64_t pk = 0;
  int32_t sq = 0, p0 = 0;
  if (w < n_words) {
    int64_t s = find_seq_from (woff, n_seq, w, __ldg (tile_seq + tile));
    int32_t L = __ldg (len + s);
    p0 = (int32_t) ((w - __ldg (woff + s)) << 5);
    sq = (int32_t) s;
    GCG_DEV_ASSERT (s >= 0 && s < n_seq && p0 >= 0 && (p0 < L || L == 0) && __ldg (woff + s) <= w && w < __ldg (woff + s + 1));
    nvalid = L - k + 1 - p0;
    nvalid = nvalid < 0 ? 0 : (nvalid > 32 ? 32 : nvalid);
    pk = CG ? __ldcg (packed + w) : __ldg (packed + w);
  }
  *pk_out = pk; *seq_out = sq; *p0_out = p0;
  if (nvalid) {
    kroll r;
    r.init (pk, CG ? __ldcg (packed + w + 1) : __ldg (packed + w + 1), k);
    for (int j0 = 0; j0 < nvalid; j0 += K4_UNROLL) {
      unsigned long long key[K4_UNROLL];
      uint32_t fp[K4_UNROLL], bk[K4_UNROLL];
      bucket4 q[K4_UNROLL];
      if (FILTER) {
        uint32_t hh[K4_UNROLL], fw[K4_UNROLL];
#pragma unroll
        for (int u = 0; u < K4_UNROLL; ++u) {
          if (j0 + u) r.step ();
          key[u] = (r.fwd < r.rc ? r.fwd : r.rc) + 1ULL;
          hh[u] = kmer_hash32 (key[u] - 1ULL);
          fp[u] = hh[u] & 3u;
          fw[u] = __ldg (filter + __umulhi (kmer_hash32b (key[u] - 1ULL), filter_words));
        }
#pragma unroll
        for (int u = 0; u < K4_UNROLL; ++u) {
          const uint32_t m = filter_mask (hh[u], filter_k3);
          q[u].a = q[u].b = q[u].c = q[u].d = 0ULL;        // no match, no overflow mark
          if (PART) {
            if ((fw[u] & m) == m) {
              const gcg_part_desc pd = parts[kmer_owner (key[u] - 1ULL, n_part)];
              bk[u] = __umulhi (hh[u], pd.n_bucket);
              q[u] = ld_bucket (pd.keys + 4ULL * bk[u]);
            }
          } else {
            bk[u] = __umulhi (hh[u], n_bucket);
            if ((fw[u] & m) == m) q[u] = ld_bucket_keep (keys + 4ULL * bk[u]);
          }
        }
      } else {
#pragma unroll
        for (int u = 0; u < K4_UNROLL; ++u) {
          if (j0 + u) r.step ();                     // harmless past nvalid: state is discarded
          key[u] = (r.fwd < r.rc ? r.fwd : r.rc) + 1ULL;
          uint32_t h = kmer_hash32 (key[u] - 1ULL);
          fp[u] = h & 3u;
          bk[u] = __umulhi (h, n_bucket);
          q[u] = ld_bucket_keep (keys + 4ULL * bk[u]);
        }
      }
      // four result bits per group at constant positions, one variable shift per group; positions
      // past the end of the read are cleared once per word
      uint32_t hb = 0, ob = 0;
#pragma unroll
      for (int u = 0; u < K4_UNROLL; ++u) {
        const bool hit = bucket_has_unique (q[u], key[u]);           // multi == 1  (ont.c:171,195)
        hb |= hit ? (1u << u) : 0u;
        ob |= bucket_ovf_bit (q[u], fp[u], u);
        // the anchor's value word (the four of a bucket share one sector) is wanted a few microseconds from now,
        // when the tile's anchors are emitted: start it on its way into the L2
        if (PF && hit) prefetch_l2 (vals + 4ULL * bk[u]);
      }
      mymask |= hb << j0;
      pend |= (ob & ~hb) << j0;
    }
    const uint32_t vmask = nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u);
    mymask &= vmask; pend &= vmask;
  }
  // ---- the rare probes that have to look at later buckets, dealt round-robin over the lanes
  uint32_t c = __popc (pend), x = c;
  for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, x, o); if (lane >= o) x += y; }
  const uint32_t total = __shfl_sync (0xffffffffu, x, 31);
  if (total) {
    __syncwarp ();
    sm.excl[wid][lane] = x - c;
    sm.pend[wid][lane] = pend;
    sm.add[wid][lane] = 0;
    if (lane == 31) sm.excl[wid][32] = total;
    __syncwarp ();
    for (uint32_t h = lane; h < total; h += 32) {
      int lo = 0, hi = 32;                          // excl[lo] <= h < excl[hi]
#pragma unroll
      for (int it = 0; it < 5; ++it) { int mid = (lo + hi) >> 1; if (sm.excl[wid][mid] <= h) lo = mid; else hi = mid; }
      const int j = __fns (sm.pend[wid][lo], 0, (int) (h - sm.excl[wid][lo]) + 1);
      GCG_DEV_ASSERT (lo >= 0 && lo < 32 && j >= 0 && j < 32 && sm.excl[wid][lo] <= h && h < sm.excl[wid][lo + 1]);
      const int64_t ww = (tile << 5) + lo;
      bool fw;
      unsigned long long kw, key = CG ? key_at (__ldcg (packed + ww), __ldcg (packed + ww + 1), j, k, &fw) : key_at (__ldg (packed + ww), __ldg (packed + ww + 1), j, k, &fw);
      const unsigned long long * tk = keys;
      uint32_t nb = n_bucket;
      if (PART) { const gcg_part_desc pd = parts[kmer_owner (key - 1ULL, n_part)]; tk = pd.keys; nb = pd.n_bucket; }
      uint32_t hs = kmer_hash32 (key - 1ULL), b = __umulhi (hs, nb);
      b = (b + 1 == nb) ? 0 : b + 1;                // the home bucket has been looked at
      unsigned long long slot = table_lookup (tk, nb, b, ld_bucket (tk + 4ULL * b), key, hs & 3u, &kw);
      if (slot != ~0ULL && !(kw & GCG_KEY_MULTI)) atomicOr (&sm.add[wid][lo], 1u << j);
    }
    __syncwarp ();
    mymask |= sm.add[wid][lane];
  }
  return mymask;
}

template <bool FILTER, int KC>
__global__ void __launch_bounds__ (32 * K4_WARPS)
k45_search_kernel (const uint64_t * __restrict__ packed, const int64_t * __restrict__ woff,
                   const int32_t * __restrict__ len, const int32_t * __restrict__ tile_seq, int64_t n_seq, int64_t n_words, const int k_arg,
                   const unsigned long long * __restrict__ keys, uint32_t n_bucket, uint32_t * __restrict__ hitmask,
                   const uint32_t * __restrict__ filter, uint32_t filter_words, int filter_k3, const int64_t tile0)
{
  __shared__ k4_smem sm;
  const int k = KC > 0 ? KC : k_arg;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t n_tiles = (n_words + 31) >> 5;
  const int64_t wstride = (int64_t) gridDim.x * K4_WARPS;
  for (int64_t tile = tile0 + (int64_t) blockIdx.x * K4_WARPS + wid; tile < n_tiles; tile += wstride) {   // tiles [tile0, n_tiles)
    uint64_t pk; int32_t sq, p0;
    const uint32_t mymask = k4_probe_tile<FILTER, KC, false> (sm, nullptr, tile, packed, woff, len, tile_seq, n_seq, n_words, k, keys, n_bucket,
                                                              filter, filter_words, filter_k3, &pk, &sq, &p0);
    const int64_t w = (tile << 5) + lane;
    if (w < n_words) hitmask[w] = mymask;
  }
}

// =============================================================================================
// K4+K5 in ONE pass: probe, ordered anchor index, ONT-side multiplicity and anchor records
//   The two-pass form above (probe -> masks, prefix sum, hits_emit re-probing the anchored positions)
//   walks the masks three times and fetches every anchor's key bucket again after 126 MB of anchor
//   stream have pushed the table out of the L2 (ncu round 1: hits_emit 5.8 x its algorithmic DRAM
//   bytes).  Here the warp that probed a tile also emits its anchors, microseconds later, while the
//   buckets it hit are still in the L2:
//     * tiles are handed out in increasing order by an atomic counter, so every tile a warp may have
//       to wait for is already with a running warp (no co-residency assumption);
//     * the global index of a tile's first anchor comes from a chained scan over the tiles
//       (decoupled look-back: a tile publishes its anchor count, then sums its predecessors'
//       published counts back to the nearest one whose inclusive prefix is known) — one 8-byte state
//       word per tile instead of the mask / prefix arrays and three scan launches;
//     * the anchors of a tile are dealt round-robin to the lanes in (word, bit) order, so anchors leave
//       in (read,pos) order as coalesced stores, exactly the order of the two-pass form;
//     * the ONT-side multiplicity (ont.c:245) is recorded by the same atomicOr that fetches (tid,pos,flag).
//   Only anchors whose global index lies in [win_lo, win_hi) are materialised (record written,
//   multiplicity recorded): a result buffer sized from an estimate can overflow without side effects,
//   and the second launch that finishes the job (win_lo = old capacity) repeats nothing.
//   FMT 0: 16-byte gcg_hit {read, pos, tid, cpos << 2 | flags}.
//   FMT 1: 8-byte compact anchor  pos << 36 | (cbase[tid] + cpos) << 2 | flags  with read_off[read] =
//          index of the read's first anchor (SURVEY 8d's 8-byte hit record).
// =============================================================================================
#define SCANST_AGG 0x4000000000000000ULL
#define SCANST_INC 0x8000000000000000ULL
#define SCANST_VAL 0x3FFFFFFFFFFFFFFFULL

__device__ __forceinline__ unsigned long long ld_state (const unsigned long long * p)
{
  unsigned long long v;
  asm volatile ("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_state (unsigned long long * p, unsigned long long v)
{
  asm volatile ("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

struct k45f_args {
  const uint64_t * packed; const int64_t * woff; const int32_t * len; const int32_t * tile_seq;
  int64_t n_seq, n_words;
  int k;
  const unsigned long long * keys; unsigned long long * vals; uint32_t n_bucket;
  const uint32_t * filter; uint32_t filter_words; int filter_k3;
  unsigned long long * tile_ctr;            // zeroed before the launch
  unsigned long long * state;               // one word per tile, zeroed before the launch
  void * out;                               // gcg_hit[] or uint64_t[], indexed by the global anchor index
  unsigned long long win_lo, win_hi;
  int32_t read_base;                        // added to the read index of FMT 0 records
  const int64_t * cbase;                    // FMT 1: first base of every contig in the concatenated scaffold coordinate
  long long * read_off;                     // FMT 1 (optional for FMT 0): [n_seq], pre-set to -1; written for every read that owns a word
  unsigned long long * total_out;           // base + number of anchors of the launch (device or mapped host memory)
  const unsigned long long * base_in;       // optional: global index of the launch's first anchor (anchors of earlier chunks); must not alias total_out
  unsigned long long * done_out;            // optional (mapped host memory): receives base_in's value when the launch starts, i.e. the number of
                                            // anchors that are COMPLETE in `out` (every earlier launch has finished)
  const gcg_part_desc * parts;              // remote-probe search: the partitions of the table (device array), else NULL
  uint32_t n_part;
  // streaming search (one launch over reads that are still being uploaded, search_host_stream):
  const unsigned long long * ready;         // optional: number of words of `packed` that have arrived (device word, written in the upload
                                            // stream's order behind the words themselves); a warp waits for its tile's words + 1
  unsigned int * group_cnt;                 // optional: one counter per group of 2^group_shift block tiles, zeroed before the launch
  unsigned long long * group_done;          // (mapped host memory) [group]: 1 + number of anchors complete in `out` once every tile up to the
  int group_shift;                          // group's last one has been emitted AND all groups before it report too (the host reads in order)
};

// Build-time switches (A/B variants, scripts/build_variants.sh):
//   K45F_MINB      resident blocks per SM asked of ptxas (8 = 32 registers, full occupancy)
//   K45F_BLOCKSCAN 0: chained scan over warp tiles; 1: over BLOCK tiles (8 warp tiles share one state word; three
//                  barriers per round); 2 (default): block tiles, emit deferred by one round, no barriers (below).
//                  The first wave of a launch resolves its prefixes hop by hop, 32 states per hop — with
//                  9472 warp tiles in flight that is 296 dependent L2 round trips, with 1184 block tiles 37
//   K45F_PREFETCH  the probe loop starts the value sector of every hit on its way into the L2
#ifndef K45F_MINB
#define K45F_MINB 6
#endif
#ifndef K45F_BLOCKSCAN
#define K45F_BLOCKSCAN 2
#endif
#ifndef K45F_PREFETCH
#define K45F_PREFETCH 0
#endif
#ifndef K45F_DIAG
#define K45F_DIAG 0
#endif

// exclusive prefix of chain element `idx` (whose own count `total` is published here), by the calling warp;
// idx must have been handed out in increasing order (every predecessor is with a running warp or block)
__device__ __forceinline__ unsigned long long
chain_lookback (unsigned long long * __restrict__ state, const int64_t idx, const unsigned long long total, const int lane,
                const unsigned long long base0)
{
  unsigned long long base = base0;
  if (idx > 0) {
    base = 0;
    if (lane == 0) st_state (state + idx, SCANST_AGG | total);
    int64_t j = idx - 1;
    for (;;) {
      const int64_t at = j - lane;                    // lane 0 looks at the nearest predecessor
      const unsigned long long st = at >= 0 ? ld_state (state + at) : (SCANST_INC | base0);   // before element 0: what earlier launches emitted
      GCG_DEV_ASSERT (at < idx);
      const uint32_t inc = __ballot_sync (0xffffffffu, (st & SCANST_INC) != 0);
      const uint32_t none = __ballot_sync (0xffffffffu, (st & (SCANST_INC | SCANST_AGG)) == 0);
      const int f = inc ? __ffs (inc) - 1 : 32;       // nearest element with a known inclusive prefix
      const uint32_t need = f < 31 ? ((2u << f) - 1u) : 0xffffffffu;
      if (none & need) { __nanosleep (40); continue; }     // an element nearer than that has not published yet
      unsigned long long sum = (lane <= f) ? (st & SCANST_VAL) : 0ULL;
      for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync (0xffffffffu, sum, o);
      base += sum;
      if (f < 32) break;
      j -= 32;
    }
  }
  GCG_DEV_ASSERT (base >= base0 && base + total <= SCANST_VAL);
  if (lane == 0) st_state (state + idx, SCANST_INC | (base + total));
  return base;
}

// bounded spin: a protocol error must end in a trap with a message, not in a hung GPU
#define K45F_SPIN_LIMIT (1u << 24)
#define K45F_STREAM_TIMEOUT_NS 120000000000ULL        // streaming search: two minutes without the next words of the reads
#define K45F_SPIN_FAIL(what) do { printf ("k45_fused_kernel: %s never arrived (block %d warp %d)\n", what, (int) blockIdx.x, (int) (threadIdx.x >> 5)); __trap (); } while (0)

// the anchors of one staged tile, 32 per round, at their global indices base .. base + total
template <int FMT, bool PART>
__device__ __forceinline__ void
k45f_emit_tile (const k45f_args & A, const int k, const int lane, const unsigned long long base, const unsigned long long base0, const uint32_t total,
                const uint32_t * __restrict__ excl, const uint32_t * __restrict__ mask, const uint64_t * __restrict__ pk,
                const int32_t * __restrict__ seq, const int32_t * __restrict__ p0s)
{
  for (uint32_t h = lane; h < total; h += 32) {
    const unsigned long long g = base + h;
    if (g < A.win_lo || g >= A.win_hi) continue;
    int lo = 0, hi = 32;                          // excl[lo] <= h < excl[hi]
#pragma unroll
    for (int it = 0; it < 5; ++it) { int mid = (lo + hi) >> 1; if (excl[mid] <= h) lo = mid; else hi = mid; }
    const int j = __fns (mask[lo], 0, (int) (h - excl[lo]) + 1);
    GCG_DEV_ASSERT (lo >= 0 && lo < 32 && j >= 0 && j < 32 && ((mask[lo] >> j) & 1u) && excl[lo] <= h && h < excl[lo + 1]);
    bool fw;
    unsigned long long kw;
    const unsigned long long key = key_at (pk[lo], pk[lo + 1], j, k, &fw);
    const unsigned long long * tk = A.keys;
    unsigned long long * tv = A.vals;
    uint32_t nb = A.n_bucket;
    if (PART) { const gcg_part_desc pd = A.parts[kmer_owner (key - 1ULL, A.n_part)]; tk = pd.keys; tv = pd.vals; nb = pd.n_bucket; }
    const uint32_t hs = kmer_hash32 (key - 1ULL), b = __umulhi (hs, nb);
    const bucket4 q = PART ? ld_bucket (tk + 4ULL * b) : ld_bucket_keep (tk + 4ULL * b);
    const int f = bucket_find (q, key, &kw);
    const unsigned long long slot = f >= 0 ? 4ULL * b + (unsigned) f
                                           : table_lookup (tk, nb, b, q, key, hs & 3u, &kw);   // walks on to the overflow buckets
    GCG_DEV_ASSERT (slot != ~0ULL && slot < 4ULL * nb && (kw & GCG_KEY_MASK) == key && !(kw & GCG_KEY_MULTI));   // the mask bit said: present, once
    GCG_DEV_ASSERT (g >= base0 && g < A.win_hi);
    // one atomic both records the anchor (ONT-side multiplicity, ont.c:245) and returns (tid, pos, flag); with a
    // partitioned table it travels over NVLink to the owner, who thereby sees the hits of every rank
#if K45F_DIAG == 1
    const unsigned long long v = slot;               // (diagnostic build: no value access — WRONG results, timing only)
#else
    const unsigned long long v = atomicOr (tv + slot, GCG_VAL_ONT1);
    if ((v & (GCG_VAL_ONT1 | GCG_VAL_ONT2)) == GCG_VAL_ONT1) atomicOr (tv + slot, GCG_VAL_ONT2);
#endif
    const uint32_t tid = (uint32_t) (v >> 32) & 0x7FFFFFFFu, cpos = (uint32_t) (v >> 1) & 0x3FFFFFFFu;
    const uint32_t flags = (uint32_t) (v & 1ULL) | (fw ? 0u : 2u);
    if (FMT == 0) {
      int4 hh;                                    // gcg_hit {read, pos, tid, cpos_flags} as one 16-byte store
      hh.x = seq[lo] + A.read_base;
      hh.y = p0s[lo] + j;
      hh.z = (int32_t) tid;
      hh.w = (int32_t) ((cpos << 2) | flags);
      __stcs (reinterpret_cast<int4 *> (A.out) + g, hh);
    } else {
      const unsigned long long gpos = (unsigned long long) __ldg (A.cbase + tid) + cpos;
      __stcs (reinterpret_cast<unsigned long long *> (A.out) + g,
              ((unsigned long long) (uint32_t) (p0s[lo] + j) << 36) | (gpos << 2) | flags);
    }
  }
}

#if K45F_BLOCKSCAN == 2
// ---- the pipelined form ------------------------------------------------------------------------
// A block works in ROUNDS: in round r its eight warps probe the eight warp tiles of block tile bt_r (handed out in
// increasing order by the global counter).  The chain element is the block tile.  What made the plain block form slow
// (ncu round 2: 28 % of the stall samples at its three barriers) is that a tile's anchors were emitted right after its
// probe: every block then waits, at a barrier, for ALL block tiles before it to publish their counts — a convoy of
// 1184 blocks moving at the pace of the slowest.  Here the emit of round r is DEFERRED to round r + 1: by then the
// counts of every earlier block tile were published a whole probe (tens of microseconds) ago and the look-back is a
// couple of L2 reads; and nothing in a round is a barrier:
//   * the first warp to finish its probe of round r claims the block tile of round r + 1 (so nobody waits for it);
//   * every warp adds its anchor count into a shared counter; the warp that completes the eight publishes the block
//     tile's count (SCANST_AGG) to the chain;
//   * the first warp that wants to emit round r - 1 resolves that block tile's prefix (look-back) for the block, the
//     others find it in shared memory;
//   * per-round scalars live in four rotating slots (a warp can be at most two rounds ahead of another: it cannot
//     resolve round r + 1 before every warp of the block has finished probing it), a warp's staged tile in two.
// Every spin is bounded and traps with a message instead of hanging the GPU.
#define K45F_SLOTS 8        // rotating per-round scalars (a warp is at most two rounds ahead of another; writers touch round + 1)
// STREAM: the launch of the streaming search (A.ready / A.group_* are set); the device-resident search carries none of it
template <bool FILTER, int KC, int FMT, bool PART = false, bool STREAM = false>
__global__ void __launch_bounds__ (32 * K4_WARPS, K45F_MINB)
k45_fused_kernel (const k45f_args A)
{
  __shared__ k4_smem sm;
  __shared__ uint32_t s_excl[2][K4_WARPS][33], s_mask[2][K4_WARPS][32];
  __shared__ uint64_t s_pk[2][K4_WARPS][33];
  __shared__ int32_t s_seq[2][K4_WARPS][32], s_p0[2][K4_WARPS][32];
  __shared__ unsigned long long s_bt[K45F_SLOTS], s_bbase[K45F_SLOTS], s_acc[K45F_SLOTS];   // block tile of a round; its exclusive prefix; (arrivals << 48) | anchors
  __shared__ uint32_t s_wtot[K45F_SLOTS][K4_WARPS];
  __shared__ int s_btgen[K45F_SLOTS], s_lock[K45F_SLOTS], s_basegen[K45F_SLOTS];            // round + 1 when valid (0 = never)
  const int k = KC > 0 ? KC : A.k;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t n_tiles = (A.n_words + 31) >> 5;
  const int64_t n_bt = (n_tiles + K4_WARPS - 1) / K4_WARPS;
  const unsigned long long base0 = A.base_in ? *A.base_in : 0ULL;
  if (A.done_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *A.done_out = base0;
  if (threadIdx.x < K45F_SLOTS) { s_acc[threadIdx.x] = 0; s_btgen[threadIdx.x] = 0; s_lock[threadIdx.x] = 0; s_basegen[threadIdx.x] = 0; }
  __syncthreads ();
  if (threadIdx.x == 0) { s_bt[0] = atomicAdd (A.tile_ctr, 1ULL); s_btgen[0] = 1; }
  __syncthreads ();

  uint32_t prev_total = 0;                             // anchors of the tile this warp staged in the previous round
  int64_t prev_bt = -1;
  for (int r = 0;; ++r) {
    const int qs = r & (K45F_SLOTS - 1), q2 = r & 1;
    // ---- the block tile of this round (claimed by the first warp that finished the previous round's probe)
    if (lane == 0) { uint32_t spin = 0; while (*(volatile int *) &s_btgen[qs] != r + 1) if (++spin > K45F_SPIN_LIMIT) K45F_SPIN_FAIL ("the block tile of a round"); }
    __syncwarp ();
    const int64_t bt = (int64_t) *(volatile unsigned long long *) &s_bt[qs];
    const bool live = bt < n_bt;
    uint32_t total = 0;
    if (live) {
      const int64_t tile = bt * K4_WARPS + wid;          // may be past the end in the last block tile: probes nothing
      if (STREAM && A.ready != nullptr && (tile << 5) < A.n_words) {
        // streaming: this tile's words (and the one behind them, which the last lane's k-mers run into) must have arrived
        if (lane == 0) {
          const unsigned long long need = (unsigned long long) (((tile + 1) << 5) < A.n_words ? ((tile + 1) << 5) + 1 : A.n_words + 1);
          if (*(volatile const unsigned long long *) A.ready < need) {
            unsigned long long t0;
            asm volatile ("mov.u64 %0, %%globaltimer;" : "=l" (t0));
            while (*(volatile const unsigned long long *) A.ready < need) {
              __nanosleep (200);
              unsigned long long t1;
              asm volatile ("mov.u64 %0, %%globaltimer;" : "=l" (t1));
              if (t1 - t0 > K45F_STREAM_TIMEOUT_NS) K45F_SPIN_FAIL ("the upload of a tile's words");
            }
          }
          __threadfence ();
        }
        __syncwarp ();
      }
      uint64_t pk; int32_t sq, p0;
      const uint32_t mymask = k4_probe_tile<FILTER, KC, K45F_PREFETCH != 0 && !PART, PART, STREAM> (sm, A.vals, tile, A.packed, A.woff, A.len, A.tile_seq, A.n_seq, A.n_words, k, A.keys, A.n_bucket,
                                                                                            A.filter, A.filter_words, A.filter_k3, &pk, &sq, &p0, A.parts, A.n_part);
      const int64_t w = (tile << 5) + lane;
      const uint32_t c = __popc (mymask);
      uint32_t x = c;
      for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, x, o); if (lane >= o) x += y; }
      total = __shfl_sync (0xffffffffu, x, 31);
      GCG_DEV_ASSERT (tile < n_tiles || mymask == 0);
      // stage the tile for the emit one round later
      s_excl[q2][wid][lane] = x - c;
      s_mask[q2][wid][lane] = mymask;
      s_pk[q2][wid][lane] = pk;
      s_seq[q2][wid][lane] = w < A.n_words ? sq : -1;
      s_p0[q2][wid][lane] = p0;
      if (lane == 31) { s_excl[q2][wid][32] = total; s_pk[q2][wid][32] = w + 1 <= A.n_words ? (STREAM ? __ldcg (A.packed + w + 1) : __ldg (A.packed + w + 1)) : 0ULL; }   // (slack word at n_words)
      // ---- count in; the first arrival claims the next round's block tile, the last one publishes this tile's count
      if (lane == 0) {
        s_wtot[qs][wid] = total;
        __threadfence_block ();
        const unsigned long long old = atomicAdd (&s_acc[qs], (1ULL << 48) | (unsigned long long) total);
        const int arrived = (int) (old >> 48);
        GCG_DEV_ASSERT (arrived < K4_WARPS);
        if (arrived == 0) {
          const int ns = (r + 1) & (K45F_SLOTS - 1);
          s_acc[ns] = 0;                                 // (that slot served round r - 7: everybody left it long ago)
          s_bt[ns] = atomicAdd (A.tile_ctr, 1ULL);
          __threadfence_block ();
          *(volatile int *) &s_btgen[ns] = r + 2;
        }
        if (arrived == K4_WARPS - 1) {
          const unsigned long long bsum = (old + total) & 0xFFFFFFFFFFFFULL;
          __threadfence ();                              // (every warp's s_wtot of this round is written and ordered before the count becomes visible)
          if (bt > 0) st_state (A.state + bt, SCANST_AGG | bsum);
          else {
            st_state (A.state, SCANST_INC | (base0 + bsum));                 // the first element needs no look-back
            if (n_bt == 1) *A.total_out = base0 + bsum;
            s_bbase[qs] = base0;
            __threadfence_block ();
            *(volatile int *) &s_basegen[qs] = r + 1;
          }
        }
      }
      __syncwarp ();
    }
    // ---- emit the tile staged one round ago
    if (prev_bt >= 0) {
      const int ps = (r - 1) & (K45F_SLOTS - 1), p2 = (r - 1) & 1;
      int ready = 0;
      if (lane == 0) ready = *(volatile int *) &s_basegen[ps] == r;
      ready = __shfl_sync (0xffffffffu, ready, 0);
      if (!ready) {
        // the first warp to get here resolves the block tile's prefix for the whole block
        int mine = 0;
        if (lane == 0) mine = atomicMax (&s_lock[ps], r) < r;
        mine = __shfl_sync (0xffffffffu, mine, 0);
        if (mine) {
          // the tile's own count is published when the last warp of the block has finished that probe; the counts of the
          // block tiles before it were published about a whole probe ago
          uint32_t spin = 0;
          unsigned long long own;
          do { own = ld_state (A.state + prev_bt); if (++spin > K45F_SPIN_LIMIT) K45F_SPIN_FAIL ("this block tile's own count"); } while ((own & (SCANST_AGG | SCANST_INC)) == 0);
          __threadfence ();
          unsigned long long bb = base0;
          if (!(own & SCANST_INC)) {                     // (block tile 0 is published with its inclusive prefix: nothing before it)
            const unsigned long long bsum = own & SCANST_VAL;
            bb = 0;
            int64_t j = prev_bt - 1;
            for (;;) {
              const int64_t at = j - lane;
              const unsigned long long st = at >= 0 ? ld_state (A.state + at) : (SCANST_INC | base0);
              const uint32_t inc = __ballot_sync (0xffffffffu, (st & SCANST_INC) != 0);
              const uint32_t none = __ballot_sync (0xffffffffu, (st & (SCANST_INC | SCANST_AGG)) == 0);
              const int f = inc ? __ffs (inc) - 1 : 32;
              const uint32_t need = f < 31 ? ((2u << f) - 1u) : 0xffffffffu;
              if (none & need) { if (++spin > K45F_SPIN_LIMIT) K45F_SPIN_FAIL ("an earlier block tile's count"); __nanosleep (40); continue; }
              unsigned long long sum = (lane <= f) ? (st & SCANST_VAL) : 0ULL;
              for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync (0xffffffffu, sum, o);
              bb += sum;
              if (f < 32) break;
              j -= 32;
            }
            GCG_DEV_ASSERT (prev_bt > 0 && bb >= base0 && bb + bsum <= SCANST_VAL);
            if (lane == 0) {
              st_state (A.state + prev_bt, SCANST_INC | (bb + bsum));
              if (prev_bt == n_bt - 1) *A.total_out = bb + bsum;
            }
          } else GCG_DEV_ASSERT (prev_bt == 0);
          if (lane == 0) {
            s_bbase[ps] = bb;
            __threadfence_block ();
            *(volatile int *) &s_basegen[ps] = r;
          }
        } else if (lane == 0) {
          uint32_t spin = 0;
          while (*(volatile int *) &s_basegen[ps] != r) if (++spin > K45F_SPIN_LIMIT) K45F_SPIN_FAIL ("the block tile's prefix");
        }
        __syncwarp ();
      }
      unsigned long long base = *(volatile unsigned long long *) &s_bbase[ps];
      for (int i = 0; i < wid; ++i) base += *(volatile uint32_t *) &s_wtot[ps][i];
      // first anchor of every read that starts in this tile
      if (A.read_off != nullptr && s_seq[p2][wid][lane] >= 0 && s_p0[p2][wid][lane] == 0) {
        GCG_DEV_ASSERT (s_seq[p2][wid][lane] < A.n_seq);
        A.read_off[s_seq[p2][wid][lane]] = (long long) (base + s_excl[p2][wid][lane]);
      }
#if K45F_DIAG != 2
      if (prev_total) k45f_emit_tile<FMT, PART> (A, k, lane, base, base0, prev_total, s_excl[p2][wid], s_mask[p2][wid], s_pk[p2][wid], s_seq[p2][wid], s_p0[p2][wid]);
#endif
      __syncwarp ();
      if (STREAM && A.group_cnt != nullptr) {
        // streaming: the last warp to finish a group of block tiles tells the host how many anchors are complete up to
        // the group's end (the inclusive prefix of its last block tile: resolved before any of its warps emitted)
        __threadfence ();
        if (lane == 0) {
          const int64_t g = prev_bt >> A.group_shift;
          const int64_t g_end = ((g + 1) << A.group_shift) < n_bt ? ((g + 1) << A.group_shift) : n_bt;
          const unsigned int want = (unsigned int) (g_end - (g << A.group_shift)) * K4_WARPS;
          if (atomicAdd (A.group_cnt + g, 1u) + 1u == want) {
            __threadfence ();
            const unsigned long long inc = ld_state (A.state + g_end - 1);
            GCG_DEV_ASSERT ((inc & SCANST_INC) != 0);
            *(volatile unsigned long long *) (A.group_done + g) = (inc & SCANST_VAL) + 1ULL;
            __threadfence_system ();
          }
        }
      }
    }
    if (!live) break;
    prev_total = total;
    prev_bt = bt;
  }
}
#else
template <bool FILTER, int KC, int FMT>
__global__ void __launch_bounds__ (32 * K4_WARPS, K45F_MINB)
k45_fused_kernel (const k45f_args A)
{
  __shared__ k4_smem sm;
  __shared__ uint32_t s_excl[K4_WARPS][33], s_mask[K4_WARPS][32];
  __shared__ uint64_t s_pk[K4_WARPS][33];
  __shared__ int32_t s_seq[K4_WARPS][32], s_p0[K4_WARPS][32];
#if K45F_BLOCKSCAN
  __shared__ unsigned long long s_bt, s_bbase;
  __shared__ uint32_t s_wtot[K4_WARPS];
#endif
  const int k = KC > 0 ? KC : A.k;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t n_tiles = (A.n_words + 31) >> 5;
  const unsigned long long base0 = A.base_in ? *A.base_in : 0ULL;
  if (A.done_out != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *A.done_out = base0;
  for (;;) {
#if K45F_BLOCKSCAN
    // ---- a block takes K4_WARPS consecutive warp tiles; chain element = block tile
    const int64_t n_bt = (n_tiles + K4_WARPS - 1) / K4_WARPS;
    __syncthreads ();                                  // (the previous round's s_bt / s_wtot / s_bbase have been read)
    if (threadIdx.x == 0) s_bt = atomicAdd (A.tile_ctr, 1ULL);
    __syncthreads ();
    const int64_t bt = (int64_t) s_bt;
    if (bt >= n_bt) break;
    const int64_t tile = bt * K4_WARPS + wid;          // may be past the end in the last block tile: probes nothing
#else
    unsigned long long t_ = 0;
    if (lane == 0) t_ = atomicAdd (A.tile_ctr, 1ULL);
    const int64_t tile = (int64_t) __shfl_sync (0xffffffffu, t_, 0);
    if (tile >= n_tiles) break;
#endif
    uint64_t pk; int32_t sq, p0;
    const uint32_t mymask = k4_probe_tile<FILTER, KC, K45F_PREFETCH != 0> (sm, A.vals, tile, A.packed, A.woff, A.len, A.tile_seq, A.n_seq, A.n_words, k, A.keys, A.n_bucket,
                                                                           A.filter, A.filter_words, A.filter_k3, &pk, &sq, &p0);
    const int64_t w = (tile << 5) + lane;
    const uint32_t c = __popc (mymask);
    uint32_t x = c;
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, x, o); if (lane >= o) x += y; }
    const uint32_t total = __shfl_sync (0xffffffffu, x, 31);
    // ---- chained scan: global index of this tile's first anchor
    unsigned long long base;
#if K45F_BLOCKSCAN
    if (lane == 0) s_wtot[wid] = total;
    __syncthreads ();
    if (wid == 0) {
      uint32_t bsum = 0;
#pragma unroll
      for (int i = 0; i < K4_WARPS; ++i) bsum += s_wtot[i];
      const unsigned long long bb = chain_lookback (A.state, bt, bsum, lane, base0);
      if (lane == 0) {
        s_bbase = bb;
        if (bt == n_bt - 1) *A.total_out = bb + bsum;
      }
    }
    __syncthreads ();
    base = s_bbase;
    for (int i = 0; i < wid; ++i) base += s_wtot[i];
#else
    base = chain_lookback (A.state, tile, total, lane, base0);
    if (lane == 0 && tile == n_tiles - 1) *A.total_out = base + total;
#endif
    GCG_DEV_ASSERT (tile < n_tiles || mymask == 0);
    if (A.read_off != nullptr && w < A.n_words && p0 == 0) { GCG_DEV_ASSERT (sq >= 0 && sq < A.n_seq); A.read_off[sq] = (long long) (base + (x - c)); }
    if (total == 0) continue;
#if K45F_DIAG == 2
    continue;                                          // (diagnostic build: probe + scan only — NO anchors, timing only)
#endif
    // ---- emit: the tile's anchors in (word, bit) order, 32 per round
    __syncwarp ();
    s_excl[wid][lane] = x - c;
    s_mask[wid][lane] = mymask;
    s_pk[wid][lane] = pk;
    s_seq[wid][lane] = sq;
    s_p0[wid][lane] = p0;
    if (lane == 31) { s_excl[wid][32] = total; s_pk[wid][32] = w + 1 <= A.n_words ? __ldg (A.packed + w + 1) : 0ULL; }   // (slack word at n_words)
    __syncwarp ();
    k45f_emit_tile<FMT, false> (A, k, lane, base, base0, total, s_excl[wid], s_mask[wid], s_pk[wid], s_seq[wid], s_p0[wid]);
  }
}
#endif

// ---- ordered compaction: prefix popcount over the per-word hit masks, then scatter ----------
#define SCAN_ITEMS 8
#define SCAN_BLOCK 256
#define SCAN_TILE (SCAN_ITEMS * SCAN_BLOCK)

__global__ void __launch_bounds__ (SCAN_BLOCK)
scan_reduce_kernel (const uint32_t * __restrict__ mask, int64_t n, uint32_t * __restrict__ block_sum)
{
  __shared__ uint32_t s[SCAN_BLOCK / 32];
  int64_t base = (int64_t) blockIdx.x * SCAN_TILE;
  uint32_t v = 0;
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    int64_t idx = base + (int64_t) i * SCAN_BLOCK + threadIdx.x;
    if (idx < n) v += __popc (mask[idx]);
  }
  v = __reduce_add_sync (0xffffffffu, v);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads ();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int i = 0; i < SCAN_BLOCK / 32; ++i) t += s[i];
    block_sum[blockIdx.x] = t;
  }
}

// single block: exclusive scan of block sums in place
__global__ void __launch_bounds__ (1024)
scan_blocksums_kernel (uint32_t * __restrict__ block_sum, int64_t nb, unsigned long long * __restrict__ total)
{
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads ();
  for (int64_t base = 0; base < nb; base += 1024) {
    int64_t idx = base + threadIdx.x;
    uint32_t v = idx < nb ? block_sum[idx] : 0, x = v;
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads ();
    if (threadIdx.x < 32) {
      uint32_t wv = s_warp[threadIdx.x], wx = wv;
      for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, wx, o); if (threadIdx.x >= o) wx += y; }
      s_warp[threadIdx.x] = wx - wv;
    }
    __syncthreads ();
    uint32_t excl = x - v + s_warp[threadIdx.x >> 5] + s_carry;
    if (idx < nb) block_sum[idx] = excl;
    __syncthreads ();
    if (threadIdx.x == 1023) s_carry = excl + v;
    __syncthreads ();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

__global__ void __launch_bounds__ (SCAN_BLOCK)
scan_apply_kernel (const uint32_t * __restrict__ mask, int64_t n, const uint32_t * __restrict__ block_off,
                   uint32_t * __restrict__ prefix)
{
  // items are laid out blocked-by-thread here so that each thread scans SCAN_ITEMS consecutive words
  __shared__ uint32_t s_warp[SCAN_BLOCK / 32];
  int64_t base = (int64_t) blockIdx.x * SCAN_TILE + (int64_t) threadIdx.x * SCAN_ITEMS;
  uint32_t c[SCAN_ITEMS], t = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { int64_t idx = base + i; c[i] = idx < n ? __popc (mask[idx]) : 0; t += c[i]; }
  uint32_t x = t;
  for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
  __syncthreads ();
  if (threadIdx.x < 32) {
    uint32_t wv = threadIdx.x < SCAN_BLOCK / 32 ? s_warp[threadIdx.x] : 0, wx = wv;
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, wx, o); if (threadIdx.x >= o) wx += y; }
    if (threadIdx.x < SCAN_BLOCK / 32) s_warp[threadIdx.x] = wx - wv;
  }
  __syncthreads ();
  uint32_t run = x - t + s_warp[threadIdx.x >> 5] + block_off[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { int64_t idx = base + i; if (idx < n) prefix[idx] = run; run += c[i]; }
}

// anchors in (read,pos) order.  A warp owns 32 consecutive words; the set bits of their masks
// are dealt round-robin to the lanes (balanced, independent loads, coalesced output).  Only the
// anchored positions (a few percent) are re-probed here to fetch (tid,pos,flag).
__global__ void __launch_bounds__ (256)
hits_emit_kernel (const uint64_t * __restrict__ packed, const int64_t * __restrict__ woff, const int32_t * __restrict__ tile_seq,
                  int64_t n_seq, int64_t n_words, int k, const unsigned long long * __restrict__ keys,
                  unsigned long long * __restrict__ vals, uint32_t n_bucket,
                  const uint32_t * __restrict__ mask, const uint32_t * __restrict__ prefix, int32_t read_base,
                  gcg_hit * __restrict__ out, const int64_t tile0)
{
  __shared__ uint32_t s_excl[8][33];
  __shared__ uint32_t s_mask[8][32];
  __shared__ uint64_t s_pk[8][33];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int64_t n_tiles = (n_words + 31) >> 5;
  int64_t wstride = (int64_t) gridDim.x * (blockDim.x >> 5);
  for (int64_t tile = tile0 + (int64_t) blockIdx.x * (blockDim.x >> 5) + wid; tile < n_tiles; tile += wstride) {   // tiles [tile0, n_tiles)
    int64_t w = (tile << 5) + lane;
    uint32_t m = w < n_words ? mask[w] : 0;
    uint32_t c = __popc (m), x = c;
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, x, o); if (lane >= o) x += y; }
    uint32_t total = __shfl_sync (0xffffffffu, x, 31);
    if (total == 0) continue;
    uint32_t base = prefix[tile << 5];
    const int64_t s_hint = __ldg (tile_seq + tile);
    __syncwarp ();
    s_excl[wid][lane] = x - c;
    s_mask[wid][lane] = m;
    // the tile's 33 packed words, one coalesced load (the anchors below pick theirs from shared memory:
    // one global round trip less in every anchor's dependent chain)
    s_pk[wid][lane] = w <= n_words ? __ldg (packed + w) : 0ULL;            // (the array has a slack word at n_words)
    if (lane == 31) { s_excl[wid][32] = total; s_pk[wid][32] = w + 1 <= n_words ? __ldg (packed + w + 1) : 0ULL; }
    __syncwarp ();
    for (uint32_t h = lane; h < total; h += 32) {
      int lo = 0, hi = 32;                          // s_excl[lo] <= h < s_excl[hi]
#pragma unroll
      for (int it = 0; it < 5; ++it) { int mid = (lo + hi) >> 1; if (s_excl[wid][mid] <= h) lo = mid; else hi = mid; }
      uint32_t mm = s_mask[wid][lo];
      int j = __fns (mm, 0, (int) (h - s_excl[wid][lo]) + 1);
      int64_t ww = (tile << 5) + lo;
      bool fw;
      unsigned long long kw, key = key_at (s_pk[wid][lo], s_pk[wid][lo + 1], j, k, &fw);
      uint32_t hs = kmer_hash32 (key - 1ULL), b = __umulhi (hs, n_bucket);
      const bucket4 q = ld_bucket (keys + 4ULL * b);
      const int f = bucket_find (q, key, &kw);
      const unsigned long long slot = f >= 0 ? 4ULL * b + (unsigned) f
                                             : table_lookup (keys, n_bucket, b, q, key, hs & 3u, &kw);   // walks on to the overflow buckets
      int64_t s = find_seq_from (woff, n_seq, ww, s_hint);
      int32_t p0 = (int32_t) ((ww - __ldg (woff + s)) << 5);
      // The ONT-side multiplicity state (ont.c:245: anchored once / more than once) lives in two spare
      // bits of the value word, so ONE atomic both records the anchor and returns (tid, pos, flag):
      // two random sectors per anchor (key bucket, value word) instead of three.
      unsigned long long v = atomicOr (vals + slot, GCG_VAL_ONT1);        // slot is valid: the mask bit says the key is present
      if ((v & (GCG_VAL_ONT1 | GCG_VAL_ONT2)) == GCG_VAL_ONT1) atomicOr (vals + slot, GCG_VAL_ONT2);
      int4 hh;                                      // gcg_hit {read, pos, tid, cpos_flags} as one 16-byte store
      hh.x = (int32_t) s + read_base;
      hh.y = p0 + j;
      hh.z = (int32_t) ((v >> 32) & 0x7FFFFFFFu);
      hh.w = (int32_t) ((((uint32_t) (v >> 1) & 0x3FFFFFFFu) << 2) | (uint32_t) (v & 1ULL) | (fw ? 0u : 2u));
      reinterpret_cast<int4 *> (out)[base + h] = hh;
    }
  }
}

// =============================================================================================
// K6  statistics
// =============================================================================================
// A thread takes a whole bucket per step: the four key words as one 256-bit load and, when the bucket
// holds a key that is present once, its four value words as another (n_slot is a multiple of 4, both
// arrays are 32-byte aligned).  One 8-byte load per thread and step ran at 1.8 TB/s (0.059 ms at cfg2).
__global__ void __launch_bounds__ (256)
k6_stats_kernel (const unsigned long long * __restrict__ keys, const unsigned long long * __restrict__ vals, uint64_t n_slot,
                 unsigned long long * __restrict__ out4)
{
  unsigned int total = 0, uniq = 0, ot = 0, ou = 0;          // per thread: at most 4 per step, far below 2^32
  const int64_t n_bucket = (int64_t) (n_slot >> 2), stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; b < n_bucket; b += stride) {
    const bucket4 q = ld_bucket (keys + 4 * b);
    const unsigned long long kw[4] = {q.a, q.b, q.c, q.d};
    unsigned int once = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (kw[i] & GCG_KEY_MASK) { ++total; if (!(kw[i] & GCG_KEY_MULTI)) once |= 1u << i; }
    if (once) {
      uniq += __popc (once);
      const bucket4 v = ld_bucket (vals + 4 * b);          // (only keys present once can anchor)
      const unsigned long long vw[4] = {v.a, v.b, v.c, v.d};
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if ((once >> i & 1u) && (vw[i] & GCG_VAL_ONT1)) { ++ot; if (!(vw[i] & GCG_VAL_ONT2)) ++ou; }
    }
  }
  unsigned long long t4[4] = {total, uniq, ot, ou};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    unsigned long long x = t4[c];
    for (int o = 16; o; o >>= 1) x += __shfl_down_sync (0xffffffffu, x, o);
    if ((threadIdx.x & 31) == 0 && x) atomicAdd (out4 + c, x);
  }
}

// ONT-side multiplicity of a replica folded into the primary table (gcg_table_merge_ont): the state
// of a slot is 0 / once (ONT1) / more than once (ONT1|ONT2); the sum of two states saturates at the
// last.  Slots without a key hold unspecified value words on both sides and are left alone.
__global__ void __launch_bounds__ (256)
ont_merge_kernel (const unsigned long long * __restrict__ keys, unsigned long long * __restrict__ vals,
                  const unsigned long long * __restrict__ other, int64_t n)
{
  const int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long b = other[i] & (GCG_VAL_ONT1 | GCG_VAL_ONT2);
    if (b == 0 || (keys[i] & GCG_KEY_MASK) == 0) continue;
    const unsigned long long a = vals[i];
    unsigned long long m = a | b;
    if (a & b & GCG_VAL_ONT1) m |= GCG_VAL_ONT2;
    if (m != a) vals[i] = m;
  }
}

// ... and the replica's own state cleared once it has been folded: what it collects next is added on top
__global__ void __launch_bounds__ (256)
ont_clear_kernel (const unsigned long long * __restrict__ keys, unsigned long long * __restrict__ vals, int64_t n)
{
  const int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned long long a = vals[i];
    if ((a & (GCG_VAL_ONT1 | GCG_VAL_ONT2)) != 0 && (keys[i] & GCG_KEY_MASK) != 0) vals[i] = a & ~(GCG_VAL_ONT1 | GCG_VAL_ONT2);
  }
}

__global__ void __launch_bounds__ (256)
table_dump_kernel (const unsigned long long * __restrict__ keys, const unsigned long long * __restrict__ vals,
                   uint64_t n_slot, unsigned long long * __restrict__ counter, int64_t cap,
                   uint64_t * __restrict__ key_out, int32_t * __restrict__ multi, int32_t * __restrict__ tid,
                   int32_t * __restrict__ pos, uint8_t * __restrict__ rev)
{
  int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t) n_slot; i += stride) {
    unsigned long long kw = keys[i];
    if (!(kw & GCG_KEY_MASK)) continue;
    unsigned long long o = atomicAdd (counter, 1ULL);
    if ((int64_t) o >= cap) continue;
    unsigned long long v = vals[i];
    key_out[o] = (kw & GCG_KEY_MASK) - 1ULL;
    multi[o] = (kw & GCG_KEY_MULTI) ? 2 : 1;
    tid[o] = (int32_t) ((v >> 32) & 0x7FFFFFFFu);
    pos[o] = (int32_t) ((v >> 1) & 0x3FFFFFFFu);            // (bits 31 and 63 hold the ONT-side state)
    rev[o] = (uint8_t) (v & 1ULL);
  }
}

// =============================================================================================
// contig chop to kmer_t records (def.h:58-66) for the unchanged host consumers
// =============================================================================================
__constant__ uint32_t c_crc_table[256];

struct __align__ (8) kmer_rec { uint64_t kseq; int32_t hs_id, tid, pos; uint16_t flag; int16_t kmer_len; };

__global__ void __launch_bounds__ (256)
chop_records_kernel (const uint64_t * __restrict__ packed, const int64_t * __restrict__ woff,
                     const int32_t * __restrict__ len, const int64_t * __restrict__ koff, int64_t n_seq,
                     int64_t n_words, int k, uint32_t n_thread, kmer_rec * __restrict__ out)
{
  int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t w = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
    int64_t s = find_seq (woff, n_seq, w);
    int32_t L = __ldg (len + s);
    int32_t p0 = (int32_t) ((w - __ldg (woff + s)) << 5);
    int32_t nvalid = L - k + 1 - p0;
    if (nvalid <= 0) continue;
    if (nvalid > 32) nvalid = 32;
    kroll r;
    r.init (packed[w], packed[w + 1], k);
    kmer_rec * dst = out + __ldg (koff + s) + p0;
    for (int j = 0; j < nvalid; ++j) {
      if (j) r.step ();
      bool fw = r.fwd < r.rc;
      uint64_t key = fw ? r.fwd : r.rc;
      // crc32.h:70-81 over the 8 little-endian bytes of the key (kmer.h:46-49)
      uint32_t crc = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < 8; ++b) crc = c_crc_table[(crc ^ (uint32_t) (key >> (8 * b))) & 0xff] ^ (crc >> 8);
      crc ^= 0xffffffffu;
      kmer_rec rec;
      rec.kseq = key;
      rec.hs_id = (int32_t) (crc % n_thread);
      rec.tid = (int32_t) s;
      rec.pos = p0 + j;
      rec.flag = fw ? 0 : 1;
      rec.kmer_len = (int16_t) k;
      dst[j] = rec;
    }
  }
}

// =============================================================================================
// host side
// =============================================================================================
static int grid_for (gcg_ctx * ctx, int64_t n_threads_wanted, int block, int max_blocks_per_sm)
{
  int64_t nb = (n_threads_wanted + block - 1) / block;
  int64_t cap = (int64_t) ctx->sm_count * max_blocks_per_sm;
  if (nb > cap) nb = cap;
  if (nb < 1) nb = 1;
  return (int) nb;
}

struct seq_src {
  const char * const * ptrs = nullptr;   // pointer-array form
  const char * buf = nullptr;            // concatenated form
  const int64_t * off = nullptr;
  const char * at (int64_t i) const { return ptrs ? ptrs[i] : buf + off[i]; }
};

static int make_layout (const int32_t * len, const int64_t * off, int64_t n, std::vector<int64_t> & woff,
                        std::vector<int32_t> & hlen, int64_t & n_words, int64_t & n_bases)
{
  woff.resize ((size_t) n + 1);
  hlen.resize ((size_t) n);
  int64_t w = 0, b = 0;
  for (int64_t i = 0; i < n; ++i) {
    int64_t l = len ? (int64_t) len[i] : off[i + 1] - off[i];
    GCG_CHECK (l >= 0 && l <= 0x7FFFFFFF, GCG_ERANGE, "sequence %lld has length %lld outside int32", (long long) i, (long long) l);
    woff[(size_t) i] = w;
    hlen[(size_t) i] = (int32_t) l;
    w += (l + 31) >> 5;
    b += l;
  }
  woff[(size_t) n] = w;
  n_words = w;
  n_bases = b;
  return GCG_OK;
}

// copy the ASCII bytes that fall into words [w0,w1) into dst (dst[0] is byte 32*w0)
static void gather_range (const seq_src & src, const std::vector<int64_t> & woff, const std::vector<int32_t> & len,
                          int64_t s_begin, int64_t s_end, int64_t w0, int64_t w1, char * dst)
{
  for (int64_t s = s_begin; s < s_end; ++s) {
    int64_t sw0 = woff[(size_t) s], l = len[(size_t) s];
    if (l == 0) continue;
    int64_t b0 = sw0 * 32, b1 = b0 + l;                 // byte range of the sequence in the flat layout
    int64_t c0 = std::max (b0, w0 * 32), c1 = std::min (b1, w1 * 32);
    if (c1 <= c0) continue;
    gcg_copy_stream (dst + (c0 - w0 * 32), src.at (s) + (c0 - b0), (size_t) (c1 - c0));
  }
}

// Streams the flat ASCII layout to the device in ring-sized chunks (16 MiB) through the pinned ring and calls
// `consume(d_chunk, w0, nw)` (enqueued on ctx->stream) for each chunk.
template <class F>
static int stream_ascii (gcg_ctx * ctx, const seq_src & src, const std::vector<int64_t> & woff,
                         const std::vector<int32_t> & len, int64_t n, int64_t n_words, F consume)
{
  int rc = gcg_stage_reserve (ctx);
  if (rc) return rc;
  const int64_t chunk_words = (int64_t) (ctx->stage.cap / 32);
  int64_t s_lo = 0;
  double t_wait = 0, t_gather = 0;
  auto now = [] () { return std::chrono::steady_clock::now (); };
  auto ms = [] (std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli> (b - a).count (); };
  for (int64_t w0 = 0; w0 < n_words; w0 += chunk_words) {
    int64_t w1 = std::min (n_words, w0 + chunk_words);
    int slot = ctx->stage.next;
    ctx->stage.next ^= 1;
    // the pinned slot may still feed a copy enqueued by an earlier call on this context
    auto t0 = now ();
    if (ctx->stage.busy[slot]) { GCG_CUDA (cudaEventSynchronize (ctx->stage.ev[slot])); ctx->stage.busy[slot] = false; }
    auto t1 = now ();
    t_wait += ms (t0, t1);
    // sequences overlapping [w0,w1)
    while (s_lo + 1 < n && woff[(size_t) s_lo + 1] <= w0) ++s_lo;
    int64_t s_hi = std::upper_bound (woff.begin () + s_lo, woff.begin () + n, w1 - 1) - woff.begin ();
    if (s_hi > n) s_hi = n;
    char * dst = ctx->stage.h[slot];
    int64_t ns = s_hi - s_lo;
    int nt = ctx->host_threads;
    if ((w1 - w0) * 32 < (1 << 20) || ns < 2 * nt) nt = 1;
    if (nt <= 1) gather_range (src, woff, len, s_lo, s_hi, w0, w1, dst);
    else {
      const int64_t n_task = std::min<int64_t> (ns, 4 * nt);
      gcg_workers_run (gcg_ctx_workers (ctx), n_task, [&] (int64_t t) {
        gather_range (src, woff, len, s_lo + ns * t / n_task, s_lo + ns * (t + 1) / n_task, w0, w1, dst);
        gcg_copy_fence ();
      });
    }
    gcg_copy_fence ();
    t_gather += ms (t1, now ());
    GCG_CUDA (cudaMemcpyAsync (ctx->stage.d[slot], dst, (size_t) (w1 - w0) * 32, cudaMemcpyHostToDevice, ctx->stream));
    rc = consume (ctx->stage.d[slot], w0, w1 - w0);
    if (rc) return rc;
    GCG_CUDA (cudaEventRecord (ctx->stage.ev[slot], ctx->stream));
    ctx->stage.busy[slot] = true;
  }
  if (ctx->trace && n_words * 32 > (1 << 20))
    fprintf (stderr, "[gcg]   staging %.1f MB: host gather %.3f ms (%d threads), waits on the ring %.3f ms\n", n_words * 32 / 1e6, t_gather, ctx->host_threads, t_wait);
  return GCG_OK;
}

extern "C" int gcg_host_pack_2bit (const char * seq, int64_t len, uint64_t * words_out)
{
  GCG_CHECK (len >= 0 && (len == 0 || (seq && words_out)), GCG_EINVAL, "gcg_host_pack_2bit: bad argument");
  if (len > 0) { gcg_pack_stream (words_out, seq, (size_t) len); gcg_copy_fence (); }
  return GCG_OK;
}

static int launch_pack (gcg_ctx * ctx, const char * d_ascii, uint64_t * d_packed, int64_t nw)
{
  if (nw <= 0) return GCG_OK;
  gcg_kscope ks (ctx, "k1_pack");
  k1_pack_kernel<<<grid_for (ctx, nw, 256, 8), 256, 0, ctx->stream>>> ((const uint4 *) d_ascii, d_packed, nw);
  GCG_CUDA (cudaGetLastError ());
  return GCG_OK;
}

// tile -> sequence hints (largest s with woff[s] <= 32*tile), uploaded to *d_tseq
static int tseq_upload (gcg_ctx * ctx, const std::vector<int64_t> & h_woff, int64_t n, int64_t n_words, int32_t ** d_tseq)
{
  const int64_t n_tiles = (n_words + 31) >> 5;
  std::vector<int32_t> tseq ((size_t) std::max<int64_t> (n_tiles, 1), 0);
  int64_t cur = 0;
  for (int64_t t = 0; t < n_tiles; ++t) {
    while (cur + 1 < n && h_woff[(size_t) cur + 1] <= (t << 5)) ++cur;
    tseq[(size_t) t] = (int32_t) cur;
  }
  GCG_CUDA (gcg_dmalloc (ctx, d_tseq, tseq.size () * 4));
  // (pageable source: the copy is staged by the runtime before the call returns)
  GCG_CUDA (cudaMemcpyAsync (*d_tseq, tseq.data (), tseq.size () * 4, cudaMemcpyHostToDevice, ctx->stream));
  return GCG_OK;
}

static int seqs_alloc (gcg_ctx * ctx, gcg_seqs * s)
{
  GCG_CUDA (gcg_dmalloc (ctx, &s->d_packed, (size_t) (s->n_words + 2) * 8));
  GCG_CUDA (cudaMemsetAsync (s->d_packed + s->n_words, 0, 16, ctx->stream));
  GCG_CUDA (gcg_dmalloc (ctx, &s->d_woff, (size_t) (s->n + 1) * 8));
  GCG_CUDA (gcg_dmalloc (ctx, &s->d_len, (size_t) std::max<int64_t> (s->n, 1) * 4));
  GCG_CUDA (cudaMemcpyAsync (s->d_woff, s->h_woff.data (), (size_t) (s->n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (s->n) GCG_CUDA (cudaMemcpyAsync (s->d_len, s->h_len.data (), (size_t) s->n * 4, cudaMemcpyHostToDevice, ctx->stream));
  return tseq_upload (ctx, s->h_woff, s->n, s->n_words, &s->d_tseq);
}

// wait == false: the call returns when the caller's strings have been read (gathered into the pinned ring); the copies
// and the pack kernel are queued on ctx->stream, where everything that uses the sequences is ordered behind them
static int seqs_upload_impl (gcg_ctx * ctx, const seq_src & src, const int32_t * len, int64_t n, gcg_seqs ** out, bool wait = true)
{
  GCG_CHECK (ctx && out && n >= 0, GCG_EINVAL, "gcg_seqs_upload: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  gcg_seqs * s = new gcg_seqs ();
  s->ctx = ctx; s->n = n;
  int rc = make_layout (len, src.off, n, s->h_woff, s->h_len, s->n_words, s->n_bases);
  if (!rc) rc = seqs_alloc (ctx, s);
  if (!rc) rc = stream_ascii (ctx, src, s->h_woff, s->h_len, n, s->n_words,
                              [&] (char * d, int64_t w0, int64_t nw) { return launch_pack (ctx, d, s->d_packed + w0, nw); });
  if (!rc && wait && cudaStreamSynchronize (ctx->stream) != cudaSuccess) { gcg_set_error ("gcg_seqs_upload: stream sync failed"); rc = GCG_ECUDA; }
  if (rc) { gcg_seqs_free (s); return rc; }
  *out = s;
  return GCG_OK;
}

extern "C" int gcg_seqs_upload (gcg_ctx * ctx, const char * const * seq, const int32_t * len, int64_t n, gcg_seqs ** out)
{
  GCG_CHECK (n == 0 || (seq && len), GCG_EINVAL, "gcg_seqs_upload: NULL input");
  seq_src src; src.ptrs = seq;
  return seqs_upload_impl (ctx, src, len, n, out);
}

extern "C" int gcg_seqs_upload_concat (gcg_ctx * ctx, const char * buf, const int64_t * off, int64_t n, gcg_seqs ** out)
{
  GCG_CHECK (off && (buf || n == 0), GCG_EINVAL, "gcg_seqs_upload_concat: NULL input");
  seq_src src; src.buf = buf; src.off = off;
  return seqs_upload_impl (ctx, src, nullptr, n, out);
}

extern "C" int gcg_ascii_upload_concat (gcg_ctx * ctx, const char * buf, const int64_t * off, int64_t n, gcg_ascii ** out)
{
  GCG_CHECK (ctx && out && off && n >= 0, GCG_EINVAL, "gcg_ascii_upload_concat: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  gcg_ascii * a = new gcg_ascii ();
  a->ctx = ctx; a->n = n;
  seq_src src; src.buf = buf; src.off = off;
  int rc = make_layout (nullptr, off, n, a->h_woff, a->h_len, a->n_words, a->n_bases);
  if (rc) { delete a; return rc; }
  GCG_CUDA (gcg_dmalloc (ctx, &a->d_ascii, (size_t) std::max<int64_t> (a->n_words, 1) * 32));
  GCG_CUDA (gcg_dmalloc (ctx, &a->d_woff, (size_t) (n + 1) * 8));
  GCG_CUDA (gcg_dmalloc (ctx, &a->d_len, (size_t) std::max<int64_t> (n, 1) * 4));
  GCG_CUDA (cudaMemcpyAsync (a->d_woff, a->h_woff.data (), (size_t) (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (n) GCG_CUDA (cudaMemcpyAsync (a->d_len, a->h_len.data (), (size_t) n * 4, cudaMemcpyHostToDevice, ctx->stream));
  rc = tseq_upload (ctx, a->h_woff, n, a->n_words, &a->d_tseq);
  if (rc) { gcg_ascii_free (a); return rc; }
  rc = stream_ascii (ctx, src, a->h_woff, a->h_len, n, a->n_words, [&] (char * d, int64_t w0, int64_t nw) {
    GCG_CUDA (cudaMemcpyAsync (a->d_ascii + w0 * 32, d, (size_t) nw * 32, cudaMemcpyDeviceToDevice, ctx->stream));
    return GCG_OK;
  });
  if (!rc) GCG_CUDA (cudaStreamSynchronize (ctx->stream));
  if (rc) { gcg_ascii_free (a); return rc; }
  *out = a;
  return GCG_OK;
}

extern "C" int gcg_seqs_pack (gcg_ctx * ctx, const gcg_ascii * a, gcg_seqs ** out)
{
  GCG_CHECK (ctx && a && out, GCG_EINVAL, "gcg_seqs_pack: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  gcg_seqs * s = new gcg_seqs ();
  s->ctx = ctx; s->n = a->n; s->n_words = a->n_words; s->n_bases = a->n_bases;
  s->h_woff = a->h_woff; s->h_len = a->h_len;
  // the layout arrays are already on the device: copy them there (no host staging, no stream stall)
  const size_t n_tiles = (size_t) std::max<int64_t> ((s->n_words + 31) >> 5, 1);
  cudaError_t e;
  if ((e = gcg_dmalloc (ctx, &s->d_packed, (size_t) (s->n_words + 2) * 8)) != cudaSuccess ||
      (e = gcg_dmalloc (ctx, &s->d_woff, (size_t) (s->n + 1) * 8)) != cudaSuccess ||
      (e = gcg_dmalloc (ctx, &s->d_len, (size_t) std::max<int64_t> (s->n, 1) * 4)) != cudaSuccess ||
      (e = gcg_dmalloc (ctx, &s->d_tseq, n_tiles * 4)) != cudaSuccess ||
      (e = cudaMemsetAsync (s->d_packed + s->n_words, 0, 16, ctx->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync (s->d_woff, a->d_woff, (size_t) (s->n + 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream)) != cudaSuccess ||
      (s->n && (e = cudaMemcpyAsync (s->d_len, a->d_len, (size_t) s->n * 4, cudaMemcpyDeviceToDevice, ctx->stream)) != cudaSuccess) ||
      (e = cudaMemcpyAsync (s->d_tseq, a->d_tseq, n_tiles * 4, cudaMemcpyDeviceToDevice, ctx->stream)) != cudaSuccess) {
    gcg_set_error ("gcg_seqs_pack: %s", cudaGetErrorString (e));
    gcg_seqs_free (s);
    return GCG_ECUDA;
  }
  int rc = launch_pack (ctx, a->d_ascii, s->d_packed, s->n_words);
  if (rc) { gcg_seqs_free (s); return rc; }
  *out = s;
  return GCG_OK;
}

extern "C" void gcg_ascii_free (gcg_ascii * a)
{
  if (!a) return;
  gcg_dfree (a->ctx, a->d_ascii); gcg_dfree (a->ctx, a->d_woff); gcg_dfree (a->ctx, a->d_len); gcg_dfree (a->ctx, a->d_tseq);
  delete a;
}

extern "C" void gcg_seqs_free (gcg_seqs * s)
{
  if (!s) return;
  gcg_dfree (s->ctx, s->d_packed); gcg_dfree (s->ctx, s->d_woff); gcg_dfree (s->ctx, s->d_len); gcg_dfree (s->ctx, s->d_tseq);
  delete s;
}

extern "C" int64_t gcg_seqs_count (const gcg_seqs * s) { return s ? s->n : 0; }
extern "C" int64_t gcg_seqs_bases (const gcg_seqs * s) { return s ? s->n_bases : 0; }
extern "C" int64_t gcg_seqs_kmers (const gcg_seqs * s, int k)
{
  if (!s) return 0;
  int64_t t = 0;
  for (int32_t l : s->h_len) if (l >= k) t += l - k + 1;
  return t;
}

// ---- table ------------------------------------------------------------------------------------
extern "C" void gcg_table_free (gcg_table * t)
{
  if (!t) return;
  if (t->shared_block) { cudaSetDevice (t->ctx->device); cudaStreamSynchronize (t->ctx->stream); cudaFree (t->shared_block); }
  else { gcg_dfree (t->ctx, t->d_keys); gcg_dfree (t->ctx, t->d_vals); }
  gcg_dfree (t->ctx, t->d_filter); gcg_dfree (t->ctx, t->d_cbase);
  delete t;
}

// zeroed table sized for n_kmers inserted occurrences (load factor <= 0.5 over 4-slot buckets)
int gcg_table_alloc (gcg_ctx * ctx, int64_t n_kmers, int k, gcg_table ** out)
{
  GCG_CHECK (k >= 1 && k <= 31, GCG_ERANGE, "gcg_table: k=%d outside [1,31] (kseq1_t is one uint64, kseq1.h:20,73)", k);
  GCG_CUDA (cudaSetDevice (ctx->device));
  gcg_table * t = new gcg_table ();
  t->ctx = ctx; t->k = k;
  // buckets of four slots at load factor <= 0.5 (GCG_TABLE_LOAD=<percent> for A/B runs: a denser table is smaller —
  // closer to the L2 — but more of its buckets overflow)
  static const int load_pct = [] () { const char * e = getenv ("GCG_TABLE_LOAD"); int v = e ? atoi (e) : 50; return v >= 10 && v <= 90 ? v : 50; } ();
  int64_t nb = (n_kmers * 25 + load_pct - 1) / load_pct + 64;
  if (nb >= 0xFFFFFFFFLL) { delete t; gcg_set_error ("gcg_table: %lld k-mers exceed the bucket index range", (long long) n_kmers); return GCG_ERANGE; }
  t->n_bucket = (uint32_t) nb;
  t->n_slot = (uint64_t) nb * 4;
  t->n_inserted = n_kmers;
  cudaError_t e;
  if ((e = gcg_dmalloc (ctx, &t->d_keys, t->n_slot * 8)) != cudaSuccess || (e = gcg_dmalloc (ctx, &t->d_vals, t->n_slot * 8)) != cudaSuccess) {
    gcg_set_error ("gcg_table: cudaMalloc of %llu slots failed: %s", (unsigned long long) t->n_slot, cudaGetErrorString (e));
    gcg_table_free (t);
    return GCG_ENOMEM;
  }
  GCG_CUDA (cudaMemsetAsync (t->d_keys, 0, t->n_slot * 8, ctx->stream));
  *out = t;
  return GCG_OK;
}

extern "C" int gcg_table_build_seqs (gcg_ctx * ctx, const gcg_seqs * contigs, int k, gcg_table ** out)
{
  GCG_CHECK (ctx && contigs && out, GCG_EINVAL, "gcg_table_build: bad argument");
  GCG_CHECK (k >= 1 && k <= 31, GCG_ERANGE, "gcg_table_build: k=%d outside [1,31] (kseq1_t is one uint64, kseq1.h:20,73)", k);
  GCG_CHECK (contigs->n < 0x7FFFFFFF, GCG_ERANGE, "gcg_table_build: too many contigs");
  for (int32_t l : contigs->h_len)
    GCG_CHECK (l <= (1 << 30), GCG_ERANGE, "gcg_table_build: contig longer than 2^30 bases");
  int64_t n_kmers = gcg_seqs_kmers (contigs, k);
  gcg_table * t = nullptr;
  int rc = gcg_table_alloc (ctx, n_kmers, k, &t);
  if (rc) return rc;
  {
    // scaffold coordinate of the compact anchors: contig i starts at base cbase[i] of the concatenated contigs
    std::vector<int64_t> cb ((size_t) contigs->n + 1);
    int64_t run = 0;
    for (int64_t i = 0; i < contigs->n; ++i) { cb[(size_t) i] = run; run += contigs->h_len[(size_t) i]; }
    cb[(size_t) contigs->n] = run;
    t->n_contig = contigs->n; t->n_cbases = run;
    cudaError_t e;
    if ((e = gcg_dmalloc (ctx, &t->d_cbase, cb.size () * 8)) != cudaSuccess ||
        (e = cudaMemcpyAsync (t->d_cbase, cb.data (), cb.size () * 8, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) {
      gcg_set_error ("gcg_table_build: contig offsets: %s", cudaGetErrorString (e));
      gcg_table_free (t);
      return GCG_ECUDA;
    }
  }
  if (contigs->n_words > 0 && n_kmers > 0) {
    gcg_kscope ks (ctx, "k23_build");
    // GCG_BUILD_ILP=1|2|4: inserts in flight per thread (A/B of the kernel variants; 1 is the default and the fastest)
    static const int ilp = [] () { const char * e = getenv ("GCG_BUILD_ILP"); int v = e ? atoi (e) : 1; return v == 2 || v == 4 ? v : 1; } ();
    const int grid = grid_for (ctx, contigs->n_words * 32, 256, 8);
#define K23_ARGS contigs->d_packed, contigs->d_woff, contigs->d_len, contigs->d_tseq, contigs->n, contigs->n_words, k, t->d_keys, t->d_vals, t->n_bucket
    if (ilp == 2) k23_build_kernel<2><<<grid, 256, 0, ctx->stream>>> (K23_ARGS);
    else if (ilp == 4) k23_build_kernel<4><<<grid, 256, 0, ctx->stream>>> (K23_ARGS);
    else k23_build_kernel<1><<<grid, 256, 0, ctx->stream>>> (K23_ARGS);
#undef K23_ARGS
    GCG_CUDA (cudaGetLastError ());
  }
  *out = t;
  return GCG_OK;
}

extern "C" int gcg_table_build (gcg_ctx * ctx, const char * const * contig_seq, const int32_t * contig_len,
                                int32_t n_contig, int k, gcg_table ** out)
{
  gcg_seqs * s = nullptr;
  GCG_CHECK (n_contig == 0 || (contig_seq && contig_len), GCG_EINVAL, "gcg_table_build: NULL input");
  seq_src src; src.ptrs = contig_seq;
  // no wait for the upload, the pack or the insert kernel: the caller's strings have been read when the upload returns
  // (the gather into the pinned ring is a host pass), everything that uses the table is ordered behind the insert on
  // the context's stream, and the packed contigs go back to the context's block cache, which only hands them to work
  // queued later on that stream — so the host can start on the reads (gcg_search*) while the table is being built
  int rc = seqs_upload_impl (ctx, src, contig_len, n_contig, &s, false);
  if (rc) return rc;
  rc = gcg_table_build_seqs (ctx, s, k, out);
  gcg_seqs_free (s);
  return rc;
}

extern "C" int gcg_table_stats (gcg_ctx * ctx, gcg_table * t, int64_t out[4])
{
  GCG_CHECK (ctx && t && out, GCG_EINVAL, "gcg_table_stats: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  GCG_CUDA (cudaMemsetAsync (ctx->d_counters, 0, 4 * 8, ctx->stream));
  {
    gcg_kscope ks (ctx, "k6_stats");
    k6_stats_kernel<<<grid_for (ctx, (int64_t) t->n_slot, 256, 8), 256, 0, ctx->stream>>> (t->d_keys, t->d_vals, t->n_slot, ctx->d_counters);
    GCG_CUDA (cudaGetLastError ());
  }
  GCG_CUDA (cudaMemcpyAsync (ctx->h_counters, ctx->d_counters, 4 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  GCG_CUDA (cudaStreamSynchronize (ctx->stream));
  for (int i = 0; i < 4; ++i) out[i] = (int64_t) ctx->h_counters[i];
  return GCG_OK;
}

extern "C" int64_t gcg_table_size (gcg_ctx * ctx, gcg_table * t)
{
  int64_t st[4];
  if (gcg_table_stats (ctx, t, st)) return -1;
  return st[0];
}

extern "C" int gcg_table_clone (gcg_ctx * dst_ctx, gcg_table * src, gcg_table ** out)
{
  GCG_CHECK (dst_ctx && src && out, GCG_EINVAL, "gcg_table_clone: bad argument");
  // everything queued on the source's stream (build, searches) must have landed before the copy
  GCG_CUDA (cudaSetDevice (src->ctx->device));
  GCG_CUDA (cudaStreamSynchronize (src->ctx->stream));
  GCG_CUDA (cudaSetDevice (dst_ctx->device));
  gcg_table * t = new gcg_table ();
  t->ctx = dst_ctx; t->k = src->k; t->n_bucket = src->n_bucket; t->n_slot = src->n_slot; t->n_inserted = src->n_inserted;
  t->n_contig = src->n_contig; t->n_cbases = src->n_cbases;
  cudaError_t e;
  if ((e = gcg_dmalloc (dst_ctx, &t->d_keys, t->n_slot * 8)) != cudaSuccess || (e = gcg_dmalloc (dst_ctx, &t->d_vals, t->n_slot * 8)) != cudaSuccess ||
      (src->d_cbase && (e = gcg_dmalloc (dst_ctx, &t->d_cbase, (size_t) (t->n_contig + 1) * 8)) != cudaSuccess)) {
    gcg_set_error ("gcg_table_clone: cudaMalloc of %llu slots failed: %s", (unsigned long long) t->n_slot, cudaGetErrorString (e));
    gcg_table_free (t);
    return GCG_ENOMEM;
  }
  // (the pre-filter is not copied: gcg_table_filter_ensure rebuilds it on the first search that wants it)
  if ((e = cudaMemcpyPeerAsync (t->d_keys, dst_ctx->device, src->d_keys, src->ctx->device, t->n_slot * 8, dst_ctx->stream)) != cudaSuccess ||
      (e = cudaMemcpyPeerAsync (t->d_vals, dst_ctx->device, src->d_vals, src->ctx->device, t->n_slot * 8, dst_ctx->stream)) != cudaSuccess ||
      (src->d_cbase && (e = cudaMemcpyPeerAsync (t->d_cbase, dst_ctx->device, src->d_cbase, src->ctx->device, (size_t) (t->n_contig + 1) * 8, dst_ctx->stream)) != cudaSuccess) ||
      (e = cudaStreamSynchronize (dst_ctx->stream)) != cudaSuccess) {
    gcg_set_error ("gcg_table_clone: device %d -> device %d copy failed: %s", src->ctx->device, dst_ctx->device, cudaGetErrorString (e));
    gcg_table_free (t);
    return GCG_ECUDA;
  }
  {
    // a replica counts only what IT finds: whatever the source had already collected stays with the source
    gcg_kscope ks (dst_ctx, "ont_clear");
    ont_clear_kernel<<<grid_for (dst_ctx, (int64_t) t->n_slot, 256, 8), 256, 0, dst_ctx->stream>>> (t->d_keys, t->d_vals, (int64_t) t->n_slot);
    if (cudaGetLastError () != cudaSuccess || cudaStreamSynchronize (dst_ctx->stream) != cudaSuccess) {
      gcg_set_error ("gcg_table_clone: clearing the replica's ONT state failed");
      gcg_table_free (t);
      return GCG_ECUDA;
    }
  }
  *out = t;
  return GCG_OK;
}

extern "C" int gcg_table_merge_ont (gcg_ctx * ctx, gcg_table * dst, gcg_table * src)
{
  GCG_CHECK (ctx && dst && src && dst->ctx == ctx, GCG_EINVAL, "gcg_table_merge_ont: bad argument (dst must belong to ctx)");
  GCG_CHECK (dst != src, GCG_EINVAL, "gcg_table_merge_ont: a table cannot be merged into itself");
  GCG_CHECK (dst->n_slot == src->n_slot && dst->k == src->k && dst->n_inserted == src->n_inserted, GCG_EINVAL,
             "gcg_table_merge_ont: the tables are not clones of each other (%llu / %llu slots)", (unsigned long long) dst->n_slot, (unsigned long long) src->n_slot);
  GCG_CUDA (cudaSetDevice (src->ctx->device));
  GCG_CUDA (cudaStreamSynchronize (src->ctx->stream));          // the replica's searches have finished
  GCG_CUDA (cudaSetDevice (ctx->device));
  // value words of the replica come over in pieces of at most 512 MB (a whole cfg5 value array is 4 GB)
  const int64_t n = (int64_t) dst->n_slot, piece = std::min<int64_t> (n, (int64_t) 64 << 20);
  unsigned long long * tmp = nullptr;
  const bool same_dev = src->ctx->device == ctx->device;
  if (!same_dev) GCG_CUDA (gcg_dmalloc (ctx, &tmp, (size_t) piece * 8));
  for (int64_t at = 0; at < n; at += piece) {
    const int64_t m = std::min (piece, n - at);
    const unsigned long long * other = src->d_vals + at;
    if (!same_dev) {
      cudaError_t e = cudaMemcpyPeerAsync (tmp, ctx->device, src->d_vals + at, src->ctx->device, (size_t) m * 8, ctx->stream);
      if (e != cudaSuccess) { gcg_dfree (ctx, tmp); gcg_set_error ("gcg_table_merge_ont: peer copy failed: %s", cudaGetErrorString (e)); return GCG_ECUDA; }
      other = tmp;
    }
    gcg_kscope ks (ctx, "ont_merge");
    ont_merge_kernel<<<grid_for (ctx, m, 256, 8), 256, 0, ctx->stream>>> (dst->d_keys + at, dst->d_vals + at, other, m);
    GCG_CUDA (cudaGetLastError ());
  }
  GCG_CUDA (cudaStreamSynchronize (ctx->stream));
  if (tmp) gcg_dfree (ctx, tmp);
  // the counts have moved: the replica starts its next batch from zero
  GCG_CUDA (cudaSetDevice (src->ctx->device));
  {
    gcg_kscope ks (src->ctx, "ont_clear");
    ont_clear_kernel<<<grid_for (src->ctx, n, 256, 8), 256, 0, src->ctx->stream>>> (src->d_keys, src->d_vals, n);
    GCG_CUDA (cudaGetLastError ());
  }
  GCG_CUDA (cudaStreamSynchronize (src->ctx->stream));
  GCG_CUDA (cudaSetDevice (ctx->device));
  return GCG_OK;
}

extern "C" int gcg_table_dump (gcg_ctx * ctx, gcg_table * t, int64_t cap, uint64_t * key, int32_t * multi_out,
                               int32_t * tid, int32_t * pos, uint8_t * rev)
{
  GCG_CHECK (ctx && t && cap >= 0, GCG_EINVAL, "gcg_table_dump: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  uint64_t * dk; int32_t * dm, * dt, * dp; uint8_t * dr;
  size_t c = (size_t) std::max<int64_t> (cap, 1);
  GCG_CUDA (gcg_dmalloc (ctx, &dk, c * 8)); GCG_CUDA (gcg_dmalloc (ctx, &dm, c * 4)); GCG_CUDA (gcg_dmalloc (ctx, &dt, c * 4));
  GCG_CUDA (gcg_dmalloc (ctx, &dp, c * 4)); GCG_CUDA (gcg_dmalloc (ctx, &dr, c));
  GCG_CUDA (cudaMemsetAsync (ctx->d_counters + 8, 0, 8, ctx->stream));
  {
    gcg_kscope ks (ctx, "table_dump");
    table_dump_kernel<<<grid_for (ctx, (int64_t) t->n_slot, 256, 8), 256, 0, ctx->stream>>> (
        t->d_keys, t->d_vals, t->n_slot, ctx->d_counters + 8, cap, dk, dm, dt, dp, dr);
    GCG_CUDA (cudaGetLastError ());
  }
  GCG_CUDA (cudaStreamSynchronize (ctx->stream));
  GCG_CUDA (cudaMemcpy (key, dk, (size_t) cap * 8, cudaMemcpyDeviceToHost));
  GCG_CUDA (cudaMemcpy (multi_out, dm, (size_t) cap * 4, cudaMemcpyDeviceToHost));
  GCG_CUDA (cudaMemcpy (tid, dt, (size_t) cap * 4, cudaMemcpyDeviceToHost));
  GCG_CUDA (cudaMemcpy (pos, dp, (size_t) cap * 4, cudaMemcpyDeviceToHost));
  GCG_CUDA (cudaMemcpy (rev, dr, (size_t) cap, cudaMemcpyDeviceToHost));
  gcg_dfree (ctx, dk); gcg_dfree (ctx, dm); gcg_dfree (ctx, dt); gcg_dfree (ctx, dp); gcg_dfree (ctx, dr);
  return GCG_OK;
}

// ---- pre-filter for tables beyond the L2 ------------------------------------------------------
__global__ void __launch_bounds__ (256)
filter_build_kernel (const unsigned long long * __restrict__ keys, uint64_t n_slot, uint32_t * __restrict__ filter,
                     uint32_t filter_words, int k3)
{
  int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t) n_slot; i += stride) {
    const unsigned long long kw = keys[i];
    if (!(kw & GCG_KEY_MASK) || (kw & GCG_KEY_MULTI)) continue;      // only keys present exactly once can anchor
    const unsigned long long key = (kw & GCG_KEY_MASK) - 1ULL;
    atomicOr (filter + __umulhi (kmer_hash32b (key), filter_words), filter_mask (kmer_hash32 (key), k3));
  }
}

static void filter_shape (int64_t n_keys, int64_t * n_words, int * k3)
{
  int64_t max_mb = 64;
  if (const char * e = getenv ("GCG_FILTER_MAX_MB")) max_mb = std::max<int64_t> (1, atoll (e));
  const int64_t bytes = std::min<int64_t> (std::max<int64_t> (n_keys, 1 << 20), max_mb << 20);
  *n_words = bytes / 4;
  *k3 = bytes * 8 >= 4 * std::max<int64_t> (n_keys, 1);       // three bits per key from 4 bits of filter per key up
}

// A table whose key array does not fit the L2 gets a filter of about one byte per inserted k-mer,
// capped so that the filter itself stays L2 resident (GCG_FILTER_MAX_MB, default 64).
// GCG_FILTER=0 / 1 forces it off / on for any size.
int gcg_table_filter_ensure (gcg_ctx * ctx, gcg_table * t)
{
  int want = t->n_slot * 8 > ((uint64_t) 80 << 20);
  if (const char * e = getenv ("GCG_FILTER")) want = atoi (e) != 0;
  if (!want) { t->filter_valid = false; t->filter_words = 0; return GCG_OK; }
  if (t->filter_valid) return GCG_OK;
  if (!t->d_filter) {
    int64_t words = 0;
    filter_shape (t->n_inserted, &words, &t->filter_k3);
    t->filter_words = (uint32_t) words;
    GCG_CUDA (gcg_dmalloc (ctx, &t->d_filter, (size_t) t->filter_words * 4));
  }
  GCG_CUDA (cudaMemsetAsync (t->d_filter, 0, (size_t) t->filter_words * 4, ctx->stream));
  {
    gcg_kscope ks (ctx, "filter_build");
    filter_build_kernel<<<grid_for (ctx, (int64_t) t->n_slot, 256, 8), 256, 0, ctx->stream>>> (t->d_keys, t->n_slot, t->d_filter, t->filter_words, t->filter_k3);
    GCG_CUDA (cudaGetLastError ());
  }
  t->filter_valid = true;
  return GCG_OK;
}

// ---- the same filter in a caller-owned buffer: the union over the partitions of a partitioned table
extern "C" int gcg_filter_shape (int64_t n_keys, int64_t * n_words, int * k3)
{
  GCG_CHECK (n_keys >= 0 && n_words && k3, GCG_EINVAL, "gcg_filter_shape: bad argument");
  filter_shape (n_keys, n_words, k3);
  return GCG_OK;
}

extern "C" int gcg_filter_add_table (gcg_ctx * ctx, gcg_table * t, void * d_words, int64_t n_words, int k3)
{
  GCG_CHECK (ctx && t && d_words && n_words > 0 && n_words < 0xFFFFFFFFLL, GCG_EINVAL, "gcg_filter_add_table: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  gcg_kscope ks (ctx, "filter_build");
  filter_build_kernel<<<grid_for (ctx, (int64_t) t->n_slot, 256, 8), 256, 0, ctx->stream>>> (t->d_keys, t->n_slot, (uint32_t *) d_words, (uint32_t) n_words, k3 != 0);
  GCG_CUDA (cudaGetLastError ());
  return GCG_OK;
}

__global__ void __launch_bounds__ (256)
filter_or_kernel (uint32_t * __restrict__ dst, const uint32_t * __restrict__ src, int64_t n)
{
  int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] |= src[i];
}

extern "C" int gcg_filter_or (gcg_ctx * ctx, void * d_words, const void * d_other, int64_t n_words)
{
  GCG_CHECK (ctx && d_words && d_other && n_words > 0, GCG_EINVAL, "gcg_filter_or: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  gcg_kscope ks (ctx, "filter_or");
  filter_or_kernel<<<grid_for (ctx, n_words, 256, 8), 256, 0, ctx->stream>>> ((uint32_t *) d_words, (const uint32_t *) d_other, n_words);
  GCG_CUDA (cudaGetLastError ());
  return GCG_OK;
}

// probes the words of the tiles [tile0, ceil(n_words / 32)); masks, woff and tile hints are indexed by the global word / tile
static void launch_k45 (gcg_ctx * ctx, const gcg_table * t, const uint64_t * d_packed, const int64_t * d_woff, const int32_t * d_len,
                        const int32_t * d_tseq, int64_t n_seq, int64_t n_words, int k, uint32_t * d_mask, int64_t tile0 = 0)
{
  gcg_kscope ks (ctx, "k45_search");
  const int grid = grid_for (ctx, (((n_words + 31) >> 5) - tile0) * 32, 32 * K4_WARPS, 8);
  const bool flt = t->filter_valid && t->filter_words;
  auto fn = flt ? (k == 25 ? k45_search_kernel<true, 25> : k == 31 ? k45_search_kernel<true, 31> : k45_search_kernel<true, 0>)
                : (k == 25 ? k45_search_kernel<false, 25> : k == 31 ? k45_search_kernel<false, 31> : k45_search_kernel<false, 0>);
  fn<<<grid, 32 * K4_WARPS, 0, ctx->stream>>> (d_packed, d_woff, d_len, d_tseq, n_seq, n_words, k, t->d_keys, t->n_bucket, d_mask,
                                              flt ? t->d_filter : nullptr, flt ? t->filter_words : 0u, flt ? t->filter_k3 : 0, tile0);
}

// ---- search -----------------------------------------------------------------------------------
// exclusive prefix sum of popcount(mask[w]) into prefix[w]; *total = number of set bits.  bsum holds
// (n_words + SCAN_TILE - 1) / SCAN_TILE words of scratch.  Synchronises the stream.
int64_t gcg_mask_scan_blocks (int64_t n_words) { return (n_words + SCAN_TILE - 1) / SCAN_TILE; }

// launches only: the total lands in *d_total (device, 8 bytes)
static int mask_scan_launch (gcg_ctx * ctx, const uint32_t * d_mask, int64_t n_words, uint32_t * d_prefix, uint32_t * d_bsum,
                             unsigned long long * d_total)
{
  int64_t nb = gcg_mask_scan_blocks (n_words);
  { gcg_kscope ks (ctx, "scan_reduce");
    scan_reduce_kernel<<<(unsigned) nb, SCAN_BLOCK, 0, ctx->stream>>> (d_mask, n_words, d_bsum); }
  { gcg_kscope ks (ctx, "scan_blocksums");
    scan_blocksums_kernel<<<1, 1024, 0, ctx->stream>>> (d_bsum, nb, d_total); }
  { gcg_kscope ks (ctx, "scan_apply");
    scan_apply_kernel<<<(unsigned) nb, SCAN_BLOCK, 0, ctx->stream>>> (d_mask, n_words, d_bsum, d_prefix); }
  GCG_CUDA (cudaGetLastError ());
  return GCG_OK;
}

int gcg_mask_scan (gcg_ctx * ctx, const uint32_t * d_mask, int64_t n_words, uint32_t * d_prefix, uint32_t * d_bsum, int64_t * total)
{
  int rc = mask_scan_launch (ctx, d_mask, n_words, d_prefix, d_bsum, ctx->d_counters + 4);
  if (rc) return rc;
  GCG_CUDA (cudaMemcpyAsync (ctx->h_counters + 4, ctx->d_counters + 4, 8, cudaMemcpyDeviceToHost, ctx->stream));
  GCG_CUDA (cudaStreamSynchronize (ctx->stream));
  *total = (int64_t) ctx->h_counters[4];
  return GCG_OK;
}

extern "C" void gcg_hits_free (gcg_hits * h)
{
  if (!h) return;
  gcg_dfree (h->ctx, h->d_hits); gcg_dfree (h->ctx, h->d_read_off);
  delete h;
}

extern "C" int64_t gcg_hits_count (const gcg_hits * h) { return h ? h->n : 0; }

// ---- the one-pass search: launch helper -------------------------------------------------------
// (streaming search: where the launch finds out which words have arrived and where it reports finished groups of block tiles)
struct fused_stream { const unsigned long long * ready; unsigned int * group_cnt; unsigned long long * group_done; int group_shift; };
// d_state holds n_tiles + 1 words (the last one is the tile counter); both are zeroed here.
static int launch_fused (gcg_ctx * ctx, const gcg_table * t, const uint64_t * d_packed, const int64_t * d_woff, const int32_t * d_len,
                         const int32_t * d_tseq, int64_t n_seq, int64_t n_words, int k, int fmt, void * d_out,
                         unsigned long long win_lo, unsigned long long win_hi, int32_t read_base, long long * d_read_off,
                         unsigned long long * d_state, unsigned long long * total_out,
                         const unsigned long long * base_in = nullptr, bool preset_read_off = false, unsigned long long * done_out = nullptr,
                         const fused_stream * st = nullptr)
{
  const int64_t n_tiles = (n_words + 31) >> 5;
  GCG_CUDA (cudaMemsetAsync (d_state, 0, (size_t) (n_tiles + 1) * 8, ctx->stream));
  if (d_read_off && !preset_read_off) GCG_CUDA (cudaMemsetAsync (d_read_off, 0xFF, (size_t) std::max<int64_t> (n_seq, 1) * 8, ctx->stream));
  k45f_args A;
  A.packed = d_packed; A.woff = d_woff; A.len = d_len; A.tile_seq = d_tseq; A.n_seq = n_seq; A.n_words = n_words; A.k = k;
  A.keys = t->d_keys; A.vals = t->d_vals; A.n_bucket = t->n_bucket;
  const bool flt = t->filter_valid && t->filter_words;
  A.filter = flt ? t->d_filter : nullptr; A.filter_words = flt ? t->filter_words : 0u; A.filter_k3 = flt ? t->filter_k3 : 0;
  A.tile_ctr = d_state + n_tiles; A.state = d_state; A.out = d_out; A.win_lo = win_lo; A.win_hi = win_hi;
  A.read_base = read_base; A.cbase = t->d_cbase; A.read_off = d_read_off; A.total_out = total_out; A.base_in = base_in; A.done_out = done_out;
  A.parts = t->d_parts; A.n_part = (uint32_t) t->n_part;
  A.ready = st ? st->ready : nullptr; A.group_cnt = st ? st->group_cnt : nullptr; A.group_done = st ? st->group_done : nullptr; A.group_shift = st ? st->group_shift : 0;
#if K45F_BLOCKSCAN != 2
  GCG_CHECK (st == nullptr, GCG_EINVAL, "gcg_search: this build's search kernel has no streaming form");
#endif
  const int grid = grid_for (ctx, n_tiles * 32, 32 * K4_WARPS, 8);
#if K45F_BLOCKSCAN == 2
  if (t->d_parts) {
    // remote-probe search of a hash-partitioned table: filter locally, probe the owner's partition over NVLink
    GCG_CHECK (flt && fmt == 0 && t->n_part >= 1 && t->n_part <= GCG_MAX_PART, GCG_EINVAL, "gcg_search (partitioned view): needs the union filter and 16-byte records");
    gcg_kscope ks (ctx, "k45_fused_remote");
    auto fnp = k == 25 ? k45_fused_kernel<true, 25, 0, true> : k == 31 ? k45_fused_kernel<true, 31, 0, true> : k45_fused_kernel<true, 0, 0, true>;
    fnp<<<grid, 32 * K4_WARPS, 0, ctx->stream>>> (A);
    GCG_CUDA (cudaGetLastError ());
    return GCG_OK;
  }
#else
  GCG_CHECK (t->d_parts == nullptr, GCG_EINVAL, "gcg_search (partitioned view): this build's search kernel has no remote-probe form");
#endif
  gcg_kscope ks (ctx, "k45_fused");
#if K45F_BLOCKSCAN == 2
#define K45F(F, KC) (st ? (fmt ? k45_fused_kernel<F, KC, 1, false, true> : k45_fused_kernel<F, KC, 0, false, true>) : (fmt ? k45_fused_kernel<F, KC, 1> : k45_fused_kernel<F, KC, 0>))
#else
#define K45F(F, KC) (fmt ? k45_fused_kernel<F, KC, 1> : k45_fused_kernel<F, KC, 0>)
#endif
  auto fn = flt ? (k == 25 ? K45F (true, 25) : k == 31 ? K45F (true, 31) : K45F (true, 0))
                : (k == 25 ? K45F (false, 25) : k == 31 ? K45F (false, 31) : K45F (false, 0));
#undef K45F
  fn<<<grid, 32 * K4_WARPS, 0, ctx->stream>>> (A);
  GCG_CUDA (cudaGetLastError ());
  return GCG_OK;
}

static size_t anchor_bytes (int fmt) { return fmt ? 8 : sizeof (gcg_hit); }

static int compact_limits_ok (const gcg_table * t, int64_t max_read_len)
{
  GCG_CHECK (t->d_cbase != nullptr, GCG_EINVAL, "compact anchors need a table built from contigs (gcg_table_build*), not an owner-side partition");
  GCG_CHECK (t->n_cbases < (1LL << 34) && max_read_len < (1LL << 28), GCG_ERANGE,
             "compact anchors hold 34 bits of scaffold coordinate and 28 bits of read position (%lld contig bases, longest read %lld): use the 16-byte form",
             (long long) t->n_cbases, (long long) max_read_len);
  return GCG_OK;
}

// two-pass form (probe -> masks, prefix sum, emit), kept behind GCG_SEARCH_FUSED=0 for A/B timing (16-byte records only)
static int search_seqs_two_pass (gcg_ctx * ctx, gcg_table * t, const gcg_seqs * reads, int k, gcg_hits * h)
{
  const int64_t n_words = reads->n_words;
  uint32_t * d_mask = nullptr, * d_prefix = nullptr, * d_bsum = nullptr;
  const int64_t nb = gcg_mask_scan_blocks (n_words);
  int rc = GCG_OK;
  cudaError_t e;
  if ((e = gcg_dmalloc (ctx, &d_mask, (size_t) n_words * 4)) != cudaSuccess || (e = gcg_dmalloc (ctx, &d_prefix, (size_t) n_words * 4)) != cudaSuccess ||
      (e = gcg_dmalloc (ctx, &d_bsum, (size_t) nb * 4)) != cudaSuccess) {
    gcg_set_error ("gcg_search: cudaMalloc failed: %s", cudaGetErrorString (e));
    rc = GCG_ENOMEM;
  }
  while (!rc) {
    launch_k45 (ctx, t, reads->d_packed, reads->d_woff, reads->d_len, reads->d_tseq, reads->n, n_words, k, d_mask, 0);
    if (cudaGetLastError () != cudaSuccess) { gcg_set_error ("gcg_search: kernel launch failed"); rc = GCG_ECUDA; break; }
    int64_t n_hit = 0;
    if ((rc = gcg_mask_scan (ctx, d_mask, n_words, d_prefix, d_bsum, &n_hit)) != 0) break;
    if (n_hit > 0 && (e = gcg_dmalloc (ctx, &h->d_hits, (size_t) n_hit * sizeof (gcg_hit))) != cudaSuccess) {
      gcg_set_error ("gcg_search: cudaMalloc of %lld anchors failed: %s", (long long) n_hit, cudaGetErrorString (e));
      rc = GCG_ENOMEM;
      break;
    }
    if (n_hit > 0) {
      gcg_kscope ks (ctx, "hits_emit");
      hits_emit_kernel<<<grid_for (ctx, ((n_words + 31) >> 5) * 32, 256, 8), 256, 0, ctx->stream>>> (
          reads->d_packed, reads->d_woff, reads->d_tseq, reads->n, n_words, k, t->d_keys, t->d_vals, t->n_bucket, d_mask, d_prefix, 0, h->d_hits, 0);
      if (cudaGetLastError () != cudaSuccess) { gcg_set_error ("gcg_search: emit launch failed"); rc = GCG_ECUDA; break; }
    }
    h->n = n_hit;
    break;
  }
  gcg_dfree (ctx, d_mask); gcg_dfree (ctx, d_prefix); gcg_dfree (ctx, d_bsum);
  return rc;
}

// The anchors stay on the device; everything that reads them (download, statistics, the next search)
// is ordered behind the kernels on the context's stream.  The result buffer is sized from the anchor
// density of the previous search on this context (first call: one anchor per 8 k-mers); if the reads
// anchor more densely than that, the launch materialises what fits and a second launch into a buffer
// of the now known size finishes the rest (k45_fused_kernel's window).
static int search_seqs_impl (gcg_ctx * ctx, gcg_table * t, const gcg_seqs * reads, int k, int fmt, gcg_hits ** out)
{
  GCG_CHECK (ctx && t && reads && out, GCG_EINVAL, "gcg_search: bad argument");
  GCG_CHECK (k == t->k, GCG_EINVAL, "gcg_search: k=%d but the table was built with k=%d", k, t->k);
  GCG_CHECK (reads->n < 0x7FFFFFFF, GCG_ERANGE, "gcg_search: too many reads");
  GCG_CUDA (cudaSetDevice (ctx->device));
  if (fmt) {
    int32_t mx = 0;
    for (int32_t l : reads->h_len) mx = std::max (mx, l);
    int rcl = compact_limits_ok (t, mx);
    if (rcl) return rcl;
  }
  gcg_hits * h = new gcg_hits ();
  h->ctx = ctx; h->fmt = fmt; h->n_seq = reads->n;
  const int64_t n_words = reads->n_words, n_kmers = gcg_seqs_kmers (reads, k);
  if (n_words == 0 || n_kmers == 0) { *out = h; return GCG_OK; }
  if (n_kmers >= 0xFFFFFFFFLL) {
    gcg_set_error ("gcg_search: %lld positions in one call exceed the 32-bit anchor index; split the read set (gcg_search does)", (long long) n_kmers);
    gcg_hits_free (h);
    return GCG_ERANGE;
  }
  int rc = t->d_parts ? GCG_OK : gcg_table_filter_ensure (ctx, t);      // (a partitioned view brings its union filter)
  static const bool two_pass = [] () { const char * e = getenv ("GCG_SEARCH_FUSED"); return e && atoi (e) == 0; } ();
  if (!rc && two_pass && fmt == 0 && !t->d_parts) {
    rc = search_seqs_two_pass (ctx, t, reads, k, h);
    if (rc) { gcg_hits_free (h); return rc; }
    *out = h;
    return GCG_OK;
  }
  const int64_t n_tiles = (n_words + 31) >> 5;
  unsigned long long * d_state = nullptr;
  cudaError_t e = cudaSuccess;
  const size_t rec = anchor_bytes (fmt);
  double frac = ctx->last_anchor_frac > 0 ? ctx->last_anchor_frac * 1.125 : 0.125;
  if (const char * ef = getenv ("GCG_SEARCH_CAP_FRAC")) frac = atof (ef);        // test hook: force the second launch
  int64_t cap = std::min<int64_t> (n_kmers, (int64_t) ((double) n_kmers * frac) + 4096);
  if (!rc && ((e = gcg_dmalloc (ctx, &d_state, (size_t) (n_tiles + 1) * 8)) != cudaSuccess ||
              (e = gcg_dmalloc (ctx, &h->d_hits, (size_t) cap * rec)) != cudaSuccess ||
              (fmt && (e = gcg_dmalloc (ctx, &h->d_read_off, (size_t) (reads->n + 1) * 8)) != cudaSuccess))) {
    gcg_set_error ("gcg_search: cudaMalloc failed (%lld anchors): %s", (long long) cap, cudaGetErrorString (e));
    rc = GCG_ENOMEM;
  }
  int64_t total = 0;
  if (!rc) rc = launch_fused (ctx, t, reads->d_packed, reads->d_woff, reads->d_len, reads->d_tseq, reads->n, n_words, k, fmt, h->d_hits,
                              0ULL, (unsigned long long) cap, 0, h->d_read_off, d_state, ctx->d_counters + 4);
  if (!rc && (cudaMemcpyAsync (ctx->h_counters + 4, ctx->d_counters + 4, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
              cudaStreamSynchronize (ctx->stream) != cudaSuccess)) {
    gcg_set_error ("gcg_search: probe failed: %s", cudaGetErrorString (cudaGetLastError ()));
    rc = GCG_ECUDA;
  }
  if (!rc) total = (int64_t) ctx->h_counters[4];
  if (!rc && total > cap) {
    // denser than estimated: the first `cap` anchors are in place (and counted); finish the rest in a buffer of the right size
    void * bigger = nullptr;
    if ((e = gcg_dmalloc (ctx, &bigger, (size_t) total * rec)) != cudaSuccess) {
      gcg_set_error ("gcg_search: cudaMalloc of %lld anchors failed: %s", (long long) total, cudaGetErrorString (e));
      rc = GCG_ENOMEM;
    } else {
      if (cudaMemcpyAsync (bigger, h->d_hits, (size_t) cap * rec, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) { gcg_set_error ("gcg_search: anchor copy failed"); rc = GCG_ECUDA; }
      gcg_dfree (ctx, h->d_hits);                     // (parked, and reused only by work queued behind the copy)
      h->d_hits = (gcg_hit *) bigger;
      if (!rc) rc = launch_fused (ctx, t, reads->d_packed, reads->d_woff, reads->d_len, reads->d_tseq, reads->n, n_words, k, fmt, h->d_hits,
                                  (unsigned long long) cap, (unsigned long long) total, 0, h->d_read_off, d_state, ctx->d_counters + 4);
    }
  }
  if (!rc && fmt) {
    // read_off[n] = total; reads without a word (length 0) are filled in on download
    const long long tot = total;
    if (cudaMemcpyAsync (h->d_read_off + reads->n, &tot, 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { gcg_set_error ("gcg_search: read offset copy failed"); rc = GCG_ECUDA; }
  }
  gcg_dfree (ctx, d_state);
  if (rc) { gcg_hits_free (h); return rc; }
  h->n = total;
  ctx->last_anchor_frac = (double) total / (double) n_kmers;
  *out = h;
  return GCG_OK;
}

extern "C" int gcg_search_seqs (gcg_ctx * ctx, gcg_table * t, const gcg_seqs * reads, int k, gcg_hits ** out)
{
  return search_seqs_impl (ctx, t, reads, k, 0, out);
}

// ---- hash-partitioned table, probed where it lies ----------------------------------------------------------------
// The routed form of the partitioned search (part.cu) walks every read position four times — count by owner, write
// the keys, then two collect passes — and moves 8 bytes out and 8 bytes back per surviving k-mer through windows and
// barriers.  Here the partitions stay where their owners built them and the ONE search kernel of the single-GPU path
// reads them directly: a k-mer that passes the union filter (replicated, L2 resident) loads its 32-byte bucket from the
// owner's key array over NVLink / NVSwitch peer memory, an anchor's atomicOr travels to the owner's value array and
// comes back with (tid, pos, flag) — so the owner still sees the hits of every rank (ont.c:245) and the statistics
// are an all-reduce of the partitions', as before.  No routing, no exchange buffers, no collective on the data path;
// the ranks meet only to know that every partition is built and, afterwards, that every rank has finished probing.
extern "C" int gcg_table_create_shared (gcg_ctx * ctx, int64_t n_records, int k, gcg_table ** out)
{
  GCG_CHECK (ctx && out && n_records >= 0, GCG_EINVAL, "gcg_table_create_shared: bad argument");
  GCG_CHECK (k >= 1 && k <= 31, GCG_ERANGE, "gcg_table_create_shared: k=%d outside [1,31]", k);
  GCG_CUDA (cudaSetDevice (ctx->device));
  int64_t nb = (n_records + 1) / 2 + 64;
  GCG_CHECK (nb < 0xFFFFFFFFLL, GCG_ERANGE, "gcg_table_create_shared: %lld k-mers exceed the bucket index range", (long long) n_records);
  gcg_table * t = new gcg_table ();
  t->ctx = ctx; t->k = k;
  t->n_bucket = (uint32_t) nb; t->n_slot = (uint64_t) nb * 4; t->n_inserted = n_records;
  t->shared_cap_slots = (t->n_slot + t->n_slot / 8 + 31) & ~(uint64_t) 31;     // head room: a rebuild of about the same size reuses the block (values stay 256-byte aligned)
  if (cudaMalloc (&t->shared_block, (size_t) t->shared_cap_slots * 16) != cudaSuccess) {
    gcg_set_error ("gcg_table_create_shared: cudaMalloc of %llu slots failed: %s", (unsigned long long) t->shared_cap_slots, cudaGetErrorString (cudaGetLastError ()));
    delete t;
    return GCG_ENOMEM;
  }
  t->d_keys = (unsigned long long *) t->shared_block;
  t->d_vals = t->d_keys + t->shared_cap_slots;                       // (values start at the CAPACITY: the layout does not move when the table is rebuilt)
  GCG_CUDA (cudaMemsetAsync (t->d_keys, 0, t->n_slot * 8, ctx->stream));
  *out = t;
  return GCG_OK;
}

// empties a shared table for a rebuild of n_records occurrences; GCG_ERANGE (nothing changed) if the block is too small
extern "C" int gcg_table_reset_shared (gcg_ctx * ctx, gcg_table * t, int64_t n_records)
{
  GCG_CHECK (ctx && t && t->shared_block && n_records >= 0, GCG_EINVAL, "gcg_table_reset_shared: not a shared table");
  const int64_t nb = (n_records + 1) / 2 + 64;
  GCG_CHECK ((uint64_t) nb * 4 <= t->shared_cap_slots, GCG_ERANGE, "gcg_table_reset_shared: %lld records do not fit the block", (long long) n_records);
  GCG_CUDA (cudaSetDevice (ctx->device));
  t->n_bucket = (uint32_t) nb; t->n_slot = (uint64_t) nb * 4; t->n_inserted = n_records;
  t->filter_valid = false;
  GCG_CUDA (cudaMemsetAsync (t->d_keys, 0, t->n_slot * 8, ctx->stream));
  return GCG_OK;
}

// what a peer needs to probe the partition: the block (IPC handle for another process, pointer inside this one),
// the offset of the value array in it, and the bucket count
extern "C" int gcg_table_shared_info (gcg_table * t, void ** d_block, int64_t * vals_offset_bytes, uint32_t * n_bucket, void * ipc_handle64)
{
  GCG_CHECK (t && t->shared_block, GCG_EINVAL, "gcg_table_shared_info: not a shared table");
  if (d_block) *d_block = t->shared_block;
  if (vals_offset_bytes) *vals_offset_bytes = (int64_t) t->shared_cap_slots * 8;
  if (n_bucket) *n_bucket = t->n_bucket;
  if (ipc_handle64) {
    static_assert (sizeof (cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    GCG_CUDA (cudaSetDevice (t->ctx->device));
    GCG_CUDA (cudaIpcGetMemHandle (&h, t->shared_block));
    memcpy (ipc_handle64, &h, 64);
  }
  return GCG_OK;
}

// d_keys[p] / d_vals[p]: partition p's arrays as THIS context can address them; n_bucket[p] its bucket count;
// the filter is the union of the partitions' anchoring keys (gcg_filter_add_table + gcg_filter_or)
extern "C" int gcg_search_seqs_remote (gcg_ctx * ctx, const gcg_seqs * reads, int k, int n_part, const void * const * d_keys, void * const * d_vals,
                                       const uint32_t * n_bucket, const void * d_filter, int64_t filter_words, int filter_k3, gcg_hits ** out)
{
  GCG_CHECK (ctx && reads && out && d_keys && d_vals && n_bucket && d_filter && n_part >= 1 && n_part <= GCG_MAX_PART && filter_words > 0 && filter_words < 0xFFFFFFFFLL,
             GCG_EINVAL, "gcg_search_seqs_remote: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  gcg_part_desc h_parts[GCG_MAX_PART];
  memset (h_parts, 0, sizeof h_parts);
  for (int p = 0; p < n_part; ++p) {
    GCG_CHECK (d_keys[p] && d_vals[p] && n_bucket[p] > 0, GCG_EINVAL, "gcg_search_seqs_remote: partition %d is not mapped", p);
    h_parts[p].keys = (const unsigned long long *) d_keys[p]; h_parts[p].vals = (unsigned long long *) d_vals[p]; h_parts[p].n_bucket = n_bucket[p];
  }
  gcg_part_desc * d_parts = nullptr;
  GCG_CUDA (gcg_dmalloc (ctx, &d_parts, sizeof h_parts));
  GCG_CUDA (cudaMemcpyAsync (d_parts, h_parts, sizeof h_parts, cudaMemcpyHostToDevice, ctx->stream));     // (pageable source: staged before the call returns)
  gcg_table view;
  view.ctx = ctx; view.k = k;
  view.d_filter = (uint32_t *) d_filter; view.filter_words = (uint32_t) filter_words; view.filter_k3 = filter_k3; view.filter_valid = true;
  view.d_parts = d_parts; view.n_part = n_part;
  int rc = search_seqs_impl (ctx, &view, reads, k, 0, out);
  gcg_dfree (ctx, d_parts);
  return rc;
}

extern "C" int gcg_search_seqs_compact (gcg_ctx * ctx, gcg_table * t, const gcg_seqs * reads, int k, gcg_hits ** out)
{
  return search_seqs_impl (ctx, t, reads, k, 1, out);
}

// reads that own no anchor index of their own (length 0: no word) start where the next read starts
static void read_off_fill (int64_t * read_off, int64_t n_read, int64_t total)
{
  read_off[n_read] = total;
  for (int64_t r = n_read - 1; r >= 0; --r) if (read_off[r] < 0) read_off[r] = read_off[r + 1];
}

extern "C" int gcg_hits_download_compact (gcg_ctx * ctx, const gcg_hits * h, uint64_t * anchors, int64_t cap, int64_t * read_off)
{
  GCG_CHECK (ctx && h && h->fmt == 1 && (anchors || cap == 0) && read_off, GCG_EINVAL, "gcg_hits_download_compact: bad argument (or not a compact anchor list)");
  const int64_t n = std::min (cap, h->n);
  if (h->d_read_off) GCG_CUDA (cudaMemcpyAsync (read_off, h->d_read_off, (size_t) (h->n_seq + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  else for (int64_t r = 0; r <= h->n_seq; ++r) read_off[r] = -1;
  if (n > 0) GCG_CUDA (cudaMemcpyAsync (anchors, h->d_hits, (size_t) n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  GCG_CUDA (cudaStreamSynchronize (ctx->stream));
  read_off_fill (read_off, h->n_seq, h->n);
  return GCG_OK;
}

extern "C" int gcg_hits_download (gcg_ctx * ctx, const gcg_hits * h, gcg_hit * dst, int64_t cap)
{
  GCG_CHECK (ctx && h && h->fmt == 0 && (dst || cap == 0), GCG_EINVAL, "gcg_hits_download: bad argument (or a compact anchor list: gcg_hits_download_compact)");
  int64_t n = std::min (cap, h->n);
  if (n > 0) {
    GCG_CUDA (cudaMemcpyAsync (dst, h->d_hits, (size_t) n * sizeof (gcg_hit), cudaMemcpyDeviceToHost, ctx->stream));
    GCG_CUDA (cudaStreamSynchronize (ctx->stream));
  }
  return GCG_OK;
}

// ---- host-buffer search: a stream pipeline ------------------------------------------------------
// The reads are cut into chunks of whole reads (8 MiB of bases by default).  For chunk c
//   host        gathers the reads into a pinned slot and packs them to 2 bits on the way (worker pool)
//   `up`        copies slot -> HBM                                     (PCIe, host -> device)
//   ctx->stream probes the chunk and emits its anchors in one launch (k45_fused_kernel)
// and the anchors reach the caller's pinned result in one of two ways:
//   direct (default): ONE device result for the whole call; the kernel stores every anchor record — and
//       every read's offset — at its FINAL index: the chained scan of a chunk starts from the running
//       anchor count of the chunks before it, which lives in device memory.  Every launch begins by
//       posting that count into mapped host memory: "anchors [0, n) are complete" (all earlier
//       launches have finished).  The host loop just reads the word as it passes by and queues D2H
//       copies of whatever has become complete on the `down` stream — it never waits for the GPU
//       between a chunk's kernel and its output, only to reuse a slot.  A result sized from an estimate
//       can overflow: anchors past the capacity are neither stored nor counted (the kernel's window),
//       the call learns the true total at the end and a second pass materialises exactly the missing tail.
//       (GCG_SEARCH_ZEROCOPY=1 lets the kernel store straight into the mapped pinned result instead of a
//       device result + copies: measured slower, 3.0 against 2.6 ms at cfg2 — the posted PCIe writes of the
//       SMs reach about half the copy engine's rate and back up into the probing warps.)
//   staged (GCG_SEARCH_DIRECT=0, the round-1 form kept for A/B timing): the kernel fills the slot's device
//       buffer (sized for the worst case), the host WAITS for the chunk's anchor count one chunk late and
//       queues a D2H copy to the final place.
// Chunks hold far fewer than 2^32 positions, so any read set size works (cfg5: 5 G positions);
// anchors come out in (read,pos) order because chunks are consecutive.  The ONT multiplicity state
// accumulates in the table across chunks.
#define PIPE_SLOTS 5
#define PIPE_LAG 2          // staged mode: most chunks between a submit and the host looking at its anchor count (GCG_SEARCH_LAG, default 1)

struct pipe_slot {
  uint64_t * h_packed = nullptr;                     // pinned, cap_words words: the chunk 2-bit packed by the host gather
  char * h_meta = nullptr, * d_meta = nullptr;       // woff | len | tile_seq of the chunk
  uint64_t * d_packed = nullptr;
  unsigned long long * d_state = nullptr;            // chained-scan state of the chunk's tiles + the tile counter
  long long * d_read_off = nullptr;                  // staged mode, compact form: first anchor of every read of the chunk
  gcg_hit * d_hits = nullptr;                        // staged mode: cap_words * 32 anchors of 16 bytes (every position could anchor)
  cudaEvent_t ev_up = nullptr, ev_emit = nullptr, ev_free = nullptr;
  bool busy = false, pending = false;                // ev_free recorded / chunk waiting for its download
  int64_t first_read = 0, n_read = 0;
};

// direct mode: the anchors of everything launched so far are complete when this runs (stream order behind the chunk's
// search kernel) — tell the host, which queues their download (a kernel start also posts, but the next kernel may wait
// a long time for its upload)
__global__ void post_count_kernel (const unsigned long long * __restrict__ total, unsigned long long * done_out)
{
  *done_out = *total;
}

struct gcg_pipe {
  int64_t cap_words = 0;
  size_t meta_cap = 0;
  bool staged = false;                               // the staged-mode buffers exist
  cudaStream_t up = nullptr, down = nullptr;
  pipe_slot s[PIPE_SLOTS];
  unsigned long long * h_count = nullptr;            // pinned + mapped, one per slot: the search kernel stores the chunk's anchor
  unsigned long long * hd_count = nullptr;           // count straight into host memory (device alias of h_count) — a copy of
                                                     // 8 bytes would queue behind the anchor downloads on the D2H copy engine
  cudaEvent_t ev_down[2] = {nullptr, nullptr};       // direct mode: behind the last two queued pieces of the download
  // one-launch search: what its UPLOAD stream writes (ready word, read offsets / lengths, packed words) lives in a block of
  // its own — blocks of the context's cache may have been released a moment ago with readers still queued on ctx->stream
  char * st_buf = nullptr;
  size_t st_cap = 0;
  unsigned long long * d_run = nullptr;              // zero copy: running anchor count, two words used alternately (a launch
                                                     // reads one and writes the other: late blocks must not see their own total)
  int64_t last_total = 0;                            // anchors of the previous call: sizes the next result buffer
};

void gcg_pipe_free (gcg_ctx * ctx)
{
  gcg_pipe * p = ctx->pipe;
  if (!p) return;
  if (p->up) cudaStreamSynchronize (p->up);
  if (p->down) cudaStreamSynchronize (p->down);
  for (pipe_slot & q : p->s) {
    if (q.h_packed) cudaFreeHost (q.h_packed);
    if (q.h_meta) cudaFreeHost (q.h_meta);
    cudaFree (q.d_meta); cudaFree (q.d_packed); cudaFree (q.d_state); cudaFree (q.d_read_off); cudaFree (q.d_hits);
    for (cudaEvent_t e : {q.ev_up, q.ev_emit, q.ev_free}) if (e) cudaEventDestroy (e);
  }
  if (p->st_buf) cudaFree (p->st_buf);
  if (p->h_count) cudaFreeHost (p->h_count);
  for (cudaEvent_t e : p->ev_down) if (e) cudaEventDestroy (e);
  cudaFree (p->d_run);
  if (p->up) cudaStreamDestroy (p->up);
  if (p->down) cudaStreamDestroy (p->down);
  delete p;
  ctx->pipe = nullptr;
}

// 0 staged, 1 direct (device result + lazy copies), 2 zero copy (kernel stores into the mapped pinned result)
static int pipe_mode (void)
{
  const char * z = getenv ("GCG_SEARCH_ZEROCOPY"), * d = getenv ("GCG_SEARCH_DIRECT");
  if (z && atoi (z) == 1) return 2;
  if (d && atoi (d) == 0) return 0;
  return 1;
}

static int pipe_reserve (gcg_ctx * ctx, int64_t cap_words, bool staged)
{
  if (ctx->pipe && ctx->pipe->cap_words >= cap_words && (ctx->pipe->staged || !staged)) return GCG_OK;
  GCG_CUDA (cudaStreamSynchronize (ctx->stream));
  if (ctx->pipe) { cap_words = std::max (cap_words, ctx->pipe->cap_words); staged = staged || ctx->pipe->staged; }
  gcg_pipe_free (ctx);
  gcg_pipe * p = new gcg_pipe ();
  ctx->pipe = p;
  p->cap_words = cap_words;
  p->staged = staged;
  p->meta_cap = (size_t) 2 << 20;
  const size_t tiles = (size_t) ((cap_words + 31) >> 5);
  if (p->meta_cap < tiles * 4 + (1 << 20)) p->meta_cap = tiles * 4 + (1 << 20);
  {
    // uploads are on the critical path of a search, downloads are not: ask for the upload stream to be served first
    int prio_lo = 0, prio_hi = 0;
    GCG_CUDA (cudaDeviceGetStreamPriorityRange (&prio_lo, &prio_hi));
    const bool prio = getenv ("GCG_SEARCH_NO_PRIO") == nullptr;
    GCG_CUDA (cudaStreamCreateWithPriority (&p->up, cudaStreamNonBlocking, prio ? prio_hi : prio_lo));
    GCG_CUDA (cudaStreamCreateWithPriority (&p->down, cudaStreamNonBlocking, prio_lo));
  }
  for (cudaEvent_t & e : p->ev_down) GCG_CUDA (cudaEventCreateWithFlags (&e, cudaEventDisableTiming));
  GCG_CUDA (cudaHostAlloc (&p->h_count, PIPE_SLOTS * sizeof (unsigned long long), cudaHostAllocMapped));
  GCG_CUDA (cudaHostGetDevicePointer (&p->hd_count, p->h_count, 0));
  GCG_CUDA (cudaMalloc (&p->d_run, 2 * sizeof (unsigned long long)));
  for (pipe_slot & q : p->s) {
    GCG_CUDA (cudaHostAlloc (&q.h_packed, (size_t) cap_words * 8, cudaHostAllocDefault));
    GCG_CUDA (cudaHostAlloc (&q.h_meta, p->meta_cap, cudaHostAllocDefault));
    GCG_CUDA (cudaMalloc (&q.d_meta, p->meta_cap));
    GCG_CUDA (cudaMalloc (&q.d_packed, (size_t) (cap_words + 2) * 8));
    GCG_CUDA (cudaMalloc (&q.d_state, (tiles + 1) * 8));
    if (staged) {
      GCG_CUDA (cudaMalloc (&q.d_read_off, (p->meta_cap / 12 + 2) * 8));       // a chunk's reads fit its meta block at 12 bytes each
      GCG_CUDA (cudaMalloc (&q.d_hits, (size_t) cap_words * 32 * sizeof (gcg_hit)));
    }
    GCG_CUDA (cudaMemset (q.d_packed, 0, (size_t) (cap_words + 2) * 8));
    for (cudaEvent_t * e : {&q.ev_up, &q.ev_emit, &q.ev_free}) GCG_CUDA (cudaEventCreateWithFlags (e, cudaEventDisableTiming));
  }
  return GCG_OK;
}

extern "C" int gcg_warmup (gcg_ctx * ctx)
{
  GCG_CHECK (ctx, GCG_EINVAL, "gcg_warmup: ctx == NULL");
  GCG_CUDA (cudaSetDevice (ctx->device));
  int rc = gcg_stage_reserve (ctx);
  if (!rc) rc = pipe_reserve (ctx, ((int64_t) 8 << 20) / 32, pipe_mode () == 0);
  return rc;
}

struct search_result {
  gcg_ctx * ctx;
  int fmt = 0;
  size_t rec = sizeof (gcg_hit);
  char * buf = nullptr;                              // pinned: gcg_hit[cap] or uint64_t[cap]
  int64_t cap = 0, n = 0;
  int64_t * read_off = nullptr;                      // compact form: pinned, [n_read + 1]
  std::vector<std::pair<int64_t, int64_t>> chunk_base;   // staged mode: (first read, anchors before the chunk)
};

// staged mode: place the anchors of the chunk in slot `q` behind the ones already placed
static int pipe_download (gcg_ctx * ctx, pipe_slot & q, int slot, search_result & res)
{
  gcg_pipe * p = ctx->pipe;
  GCG_CUDA (cudaEventSynchronize (q.ev_emit));
  const int64_t n = (int64_t) p->h_count[slot];
  if (res.n + n > res.cap) {
    // estimate too small: move what is there into a larger block (the copies in flight target the old one)
    GCG_CUDA (cudaStreamSynchronize (p->down));
    int64_t ncap = std::max<int64_t> (res.cap * 2, res.n + n + (res.n + n) / 4);
    char * nb = (char *) gcg_pinned_alloc ((size_t) ncap * res.rec);
    GCG_CHECK (nb != nullptr, GCG_ENOMEM, "gcg_search: pinned alloc of %lld anchors failed", (long long) ncap);
    // (the pool may be busy gathering the next chunk: gcg_workers_run then copies on this thread, host_par.cpp)
    if (res.n) gcg_par_memcpy (ctx, nb, res.buf, (size_t) res.n * res.rec);
    gcg_free (res.buf);
    res.buf = nb; res.cap = ncap;
  }
  if (n > 0) GCG_CUDA (cudaMemcpyAsync (res.buf + (size_t) res.n * res.rec, q.d_hits, (size_t) n * res.rec, cudaMemcpyDeviceToHost, p->down));
  if (res.fmt && q.n_read > 0) {
    GCG_CUDA (cudaMemcpyAsync (res.read_off + q.first_read, q.d_read_off, (size_t) q.n_read * 8, cudaMemcpyDeviceToHost, p->down));
    res.chunk_base.emplace_back (q.first_read, res.n);
  }
  GCG_CUDA (cudaEventRecord (q.ev_free, p->down));
  q.busy = true; q.pending = false;
  res.n += n;
  return GCG_OK;
}

// the anchors of a host-buffer search left on the device (gcg_search_runs reduces them there): compact anchors
// and n_read + 1 filled read offsets, both blocks of the context's cache (gcg_dfree)
struct search_dev_keep { void * d_anchors = nullptr; long long * d_read_off = nullptr; };

// fmt 0: *hits_out = gcg_hit[*n_hit]; fmt 1: *hits_out = uint64_t[*n_hit] and *read_off_out = int64_t[n_read + 1]
// keep != NULL (fmt 1 only): the anchors are not downloaded, *hits_out stays NULL
// where the reads of a host-buffer search come from: ASCII strings (rseq_t.b of the reference), or words the caller
// has already 2-bit packed in the library's layout (gcg_host_pack_2bit; the FASTQ loader of superplus_b200/gap_closer
// packs while it copies the bases): read i = words [pwoff[i], pwoff[i] + (len[i] + 31) / 32)
struct read_source { const char * const * seq = nullptr; const uint64_t * packed = nullptr; const int64_t * pwoff = nullptr; };

// ---- streaming search: ONE launch over reads that are still arriving ----------------------------------------------
// The chunked pipeline below pays for every chunk's launch: the first wave of a launch resolves its chained scan hop by
// hop, the last wave drains, and a chunk cannot start before its whole upload has landed — 60-70 us per launch, nine
// launches per cfg2 search, and at the end of the gather the GPU is 0.5 ms behind (GCG_TIMELINE).  Here the search
// kernel is launched ONCE, right after the table build, over device arrays sized for the whole read set; the host
// gathers + packs the reads piece by piece (or takes the caller's packed words as they lie), every piece is followed on
// the upload stream by an 8-byte copy that moves the kernel's `ready` word forward, and a warp whose tile has not
// arrived yet polls that word.  Tiles are handed out in order, so whatever a warp waits for is the next thing the
// host sends.  Finished groups of block tiles report their anchor count into mapped host memory; the host queues the
// download of everything complete while it waits for the gather.  A result that overflows its estimate is finished by a
// second launch over the words that are already on the device.
__global__ void tile_seq_kernel (const int64_t * __restrict__ woff, int64_t n_seq, int64_t n_tiles, int32_t * __restrict__ tseq)
{
  const int64_t tile = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (tile >= n_tiles) return;
  const int64_t w = tile << 5;
  int64_t lo = 0, hi = n_seq;                        // the last read with woff <= w (empty reads share their successor's offset)
  while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (__ldg (woff + mid) <= w) lo = mid; else hi = mid; }
  tseq[tile] = (int32_t) lo;
}

// The streaming search needs the host to keep working AFTER the launch call: it uploads the words the launch is
// waiting for.  Under a tool that makes launches synchronous (ncu and compute-sanitizer serialise kernels and return
// from the launch call when the kernel has finished; CUDA_LAUNCH_BLOCKING=1) that would be a dead wait, so the first
// search of a context finds out: a one-thread kernel waits up to 20 ms for a host flag that is set right after the launch
// call returns.  If the kernel gave up, launches are synchronous and this context keeps the chunked pipeline.
__global__ void launch_probe_kernel (volatile unsigned long long * flag, unsigned long long timeout_ns, unsigned long long * out)
{
  unsigned long long t0, t1;
  asm volatile ("mov.u64 %0, %%globaltimer;" : "=l" (t0));
  while (*flag == 0ULL) {
    asm volatile ("mov.u64 %0, %%globaltimer;" : "=l" (t1));
    if (t1 - t0 > timeout_ns) { *out = 1ULL; return; }
    __nanosleep (500);
  }
  *out = 2ULL;
}

static bool launches_are_async (gcg_ctx * ctx)
{
  if (ctx->launch_async >= 0) return ctx->launch_async != 0;
  gcg_pipe * p = ctx->pipe;                          // (pipe_reserve has run)
  volatile unsigned long long * h = p->h_count;
  h[PIPE_SLOTS - 1] = 0ULL; h[PIPE_SLOTS - 2] = 0ULL;
  launch_probe_kernel<<<1, 1, 0, ctx->stream>>> (p->hd_count + (PIPE_SLOTS - 1), 20000000ULL, p->hd_count + (PIPE_SLOTS - 2));
  h[PIPE_SLOTS - 1] = 1ULL;
  const bool ok = cudaGetLastError () == cudaSuccess && cudaStreamSynchronize (ctx->stream) == cudaSuccess;
  ctx->launch_async = ok && h[PIPE_SLOTS - 2] == 2ULL ? 1 : 0;
  if (ctx->trace || !ctx->launch_async)
    fprintf (stderr, "[gcg] kernel launches are %s: the host-buffer search runs %s\n", ctx->launch_async ? "asynchronous" : "SYNCHRONOUS (a profiler or CUDA_LAUNCH_BLOCKING)",
             ctx->launch_async ? "as one launch over arriving reads" : "one launch per chunk");
  return ctx->launch_async != 0;
}

// read_off[r] of a read without a word of its own (empty) or without anchors is the offset of the next read that has
// one (read_off_fill's rule), read_off[n_read] = total: a suffix minimum over offsets that never decrease.  One block,
// 1024 reads per step from the last read down — n_read is thousands to a few hundred thousand.
__global__ void __launch_bounds__ (1024)
read_off_fill_kernel (long long * __restrict__ roff, int64_t n_read, long long total)
{
  __shared__ long long s_w[32];
  __shared__ long long s_carry;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) { roff[n_read] = total; s_carry = total; }
  __syncthreads ();
  for (int64_t hi = n_read; hi > 0; hi -= 1024) {
    const int64_t i = hi - 1 - (int64_t) threadIdx.x;          // thread 0 takes the highest index of the step
    long long v = 0x7FFFFFFFFFFFFFFFLL;
    if (i >= 0) { const long long x = roff[i]; if (x >= 0) v = x; }
    for (int o = 1; o < 32; o <<= 1) { const long long y = __shfl_up_sync (0xffffffffu, v, o); if (lane >= o && y < v) v = y; }
    if (lane == 31) s_w[wid] = v;
    __syncthreads ();
    long long m = s_carry;
    for (int w = 0; w < wid; ++w) if (s_w[w] < m) m = s_w[w];
    if (v < m) m = v;
    if (i >= 0) roff[i] = m;
    __syncthreads ();
    if (threadIdx.x == 1023) s_carry = m;
    __syncthreads ();
  }
}

static bool stream_mode_on (void)
{
#if K45F_BLOCKSCAN != 2
  return false;
#else
  const char * e = getenv ("GCG_SEARCH_STREAM");
  return !(e && atoi (e) == 0);
#endif
}

static int search_host_stream (gcg_ctx * ctx, gcg_table * t, const read_source & src, const int32_t * read_len,
                               int64_t n_read, int k, int fmt, void ** hits_out, int64_t * n_hit, int64_t ** read_off_out,
                               search_dev_keep * keep, bool * handled)
{
  *handled = false;
  if (!stream_mode_on () || t->d_parts != nullptr || n_read <= 0 || (!keep && pipe_mode () != 1)) return GCG_OK;
  int64_t total_words = 0, total_kmers = 0, max_words = 0;
  int32_t max_len = 0;
  for (int64_t r = 0; r < n_read; ++r) {
    if (read_len[r] < 0) return GCG_OK;              // (the chunked path reports it)
    const int64_t w = ((int64_t) read_len[r] + 31) >> 5;
    total_words += w; max_words = std::max (max_words, w); max_len = std::max (max_len, read_len[r]);
    if (read_len[r] >= k) total_kmers += (int64_t) read_len[r] - k + 1;
  }
  int64_t limit_mb = 8192;
  if (const char * e = getenv ("GCG_SEARCH_STREAM_MAX_MB")) limit_mb = std::max<int64_t> (0, atoll (e));
  if (total_kmers == 0 || total_words * 8 > (limit_mb << 20) || (fmt && compact_limits_ok (t, max_len) != 0)) return GCG_OK;
  *handled = true;

  GCG_CUDA (cudaSetDevice (ctx->device));
  gcg_trace_mark (ctx, nullptr);
  bool packed_pinned = false;
  if (src.packed) {
    cudaPointerAttributes at;
    packed_pinned = cudaPointerGetAttributes (&at, src.packed) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError ();
  }
  // pieces: 2, 4, 8, 8, ... MiB of bases (whole reads); a piece is gathered into one of the pipeline's pinned slots
  int64_t piece_words = ((int64_t) 8 << 20) / 32, first_words = ((int64_t) 2 << 20) / 32, last_words = ((int64_t) 2 << 20) / 32;
  if (const char * e = getenv ("GCG_SEARCH_PIECE_MB")) piece_words = std::max<int64_t> (1, atoll (e)) * (1 << 20) / 32;
  if (const char * e = getenv ("GCG_SEARCH_PIECE_FIRST_MB")) first_words = std::max<int64_t> (1, atoll (e)) * (1 << 20) / 32;
  if (const char * e = getenv ("GCG_SEARCH_PIECE_LAST_MB")) last_words = std::max<int64_t> (1, atoll (e)) * (1 << 20) / 32;
  if (const char * e = getenv ("GCG_SEARCH_CHUNK_BYTES")) piece_words = first_words = last_words = std::max<int64_t> (1, atoll (e) / 32);      // (tests: many tiny pieces)
  int rc = pipe_reserve (ctx, std::max (((int64_t) 32 << 20) / 32, std::max (piece_words, max_words)), false);
  if (!rc) rc = gcg_table_filter_ensure (ctx, t);
  if (rc) return rc;
  gcg_pipe * p = ctx->pipe;
  if (!launches_are_async (ctx)) { *handled = false; return GCG_OK; }
  const int64_t cap_words = p->cap_words;
  const int64_t n_tiles = (total_words + 31) >> 5, n_bt = (n_tiles + K4_WARPS - 1) / K4_WARPS;
  const int group_shift = 8;                          // 256 block tiles = 2 M k-mers, about 1 MB of anchors at ONT error rates
  const int64_t n_group = (n_bt + (1 << group_shift) - 1) >> group_shift;

  search_result res;
  res.ctx = ctx; res.fmt = fmt; res.rec = anchor_bytes (fmt);
  res.cap = std::max<int64_t> (p->last_total + p->last_total / 8, total_kmers / 10) + 4096;
  if (const char * e = getenv ("GCG_SEARCH_RES_CAP")) res.cap = std::max<int64_t> (1, atoll (e));     // test hook: force the second pass

  // ---- host blocks (parked pinned blocks: no driver call in the steady state)
  const size_t meta_bytes = (size_t) (n_read + 1) * 8 + (size_t) n_read * 4;
  char * h_meta = (char *) gcg_pinned_alloc (meta_bytes);
  unsigned long long * h_group = (unsigned long long *) gcg_pinned_alloc ((size_t) n_group * 8);
  const int64_t max_pieces = total_words / std::max<int64_t> (1, std::min (first_words, piece_words)) + n_read + 8;
  unsigned long long * h_ready = (unsigned long long *) gcg_pinned_alloc ((size_t) std::min<int64_t> (max_pieces, total_words + 8) * 8 + 64);
  if (fmt) res.read_off = (int64_t *) gcg_pinned_alloc ((size_t) (n_read + 1) * 8);
  if (!keep) res.buf = (char *) gcg_pinned_alloc ((size_t) res.cap * res.rec);
  // ---- device blocks
  uint64_t * d_packed = nullptr; char * d_meta = nullptr; int32_t * d_tseq = nullptr; unsigned long long * d_state = nullptr, * d_ready = nullptr;
  unsigned int * d_group = nullptr; void * d_res = nullptr; long long * d_roff = nullptr; unsigned long long * hd_group = nullptr;
  auto release = [&] () {
    gcg_dfree (ctx, d_tseq); gcg_dfree (ctx, d_state); gcg_dfree (ctx, d_group);
    if (p->st_cap > ((size_t) 2 << 30)) { cudaFree (p->st_buf); p->st_buf = nullptr; p->st_cap = 0; }      // (a very large read set: not kept)
    if (d_res) gcg_dfree (ctx, d_res);
    if (d_roff) gcg_dfree (ctx, d_roff);
    gcg_free (h_meta); gcg_free (h_group); gcg_free (h_ready);
  };
  auto fail = [&] (int code) { release (); gcg_free (res.buf); gcg_free (res.read_off); return code; };
  if (!h_meta || !h_group || !h_ready || (fmt && !res.read_off) || (!keep && !res.buf)) { gcg_set_error ("gcg_search: pinned host memory for %lld reads / %lld anchors", (long long) n_read, (long long) res.cap); return fail (GCG_ENOMEM); }
  // (the upload stream's targets: one block of the pipeline's own, kept between calls up to 2 GiB)
  const size_t meta_pad = (meta_bytes + 255) & ~(size_t) 255, st_need = 256 + meta_pad + (size_t) (total_words + 2) * 8;
  cudaError_t e = cudaSuccess;
  if (p->st_cap < st_need) {
    if (p->st_buf) { cudaFree (p->st_buf); p->st_buf = nullptr; p->st_cap = 0; }
    const size_t want = st_need + st_need / 8;
    if ((e = cudaMalloc (&p->st_buf, want)) == cudaSuccess) p->st_cap = want;
    else { cudaGetLastError (); if ((e = cudaMalloc (&p->st_buf, st_need)) == cudaSuccess) p->st_cap = st_need; }
  }
  if (e == cudaSuccess) {
    d_ready = (unsigned long long *) p->st_buf;
    d_meta = p->st_buf + 256;
    d_packed = (uint64_t *) (p->st_buf + 256 + meta_pad);
  }
  if (e == cudaSuccess) e = gcg_dmalloc (ctx, &d_tseq, (size_t) n_tiles * 4);
  if (e == cudaSuccess) e = gcg_dmalloc (ctx, &d_state, (size_t) (n_tiles + 1) * 8);
  if (e == cudaSuccess) e = gcg_dmalloc (ctx, &d_group, (size_t) n_group * 4);
  if (e == cudaSuccess) e = gcg_dmalloc (ctx, &d_res, (size_t) res.cap * res.rec);
  if (e == cudaSuccess && fmt) e = gcg_dmalloc (ctx, &d_roff, (size_t) (n_read + 1) * 8);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer ((void **) &hd_group, h_group, 0);
  if (e != cudaSuccess) { gcg_set_error ("gcg_search: device memory for %lld words / %lld anchors: %s", (long long) total_words, (long long) res.cap, cudaGetErrorString (e)); cudaGetLastError (); return fail (GCG_ENOMEM); }

  // ---- meta: word offset and length of every read; the first read of every tile is found on the device
  int64_t * woff = (int64_t *) h_meta;
  int32_t * len = (int32_t *) (h_meta + (size_t) (n_read + 1) * 8);
  {
    int64_t w = 0;
    for (int64_t r = 0; r < n_read; ++r) { woff[r] = w; len[r] = read_len[r]; w += ((int64_t) read_len[r] + 31) >> 5; }
    woff[n_read] = w;
  }
  const int64_t * d_woff = (const int64_t *) d_meta;
  const int32_t * d_len = (const int32_t *) (d_meta + (size_t) (n_read + 1) * 8);
  pipe_slot & q0 = p->s[0];
  h_ready[0] = 0;
  cudaStream_t up = p->up, down = p->down, cs = ctx->stream;
  // (The upload stream writes only into the pipeline's own block.  Blocks of the context's cache may have been released a
  //  moment ago with readers still queued on ctx->stream — gcg_table_build releases the packed contigs, their offsets and
  //  lengths while its insert kernel is pending —, which is safe for work queued on ctx->stream and for nothing else.)
  if ((e = cudaMemcpyAsync (d_meta, h_meta, meta_bytes, cudaMemcpyHostToDevice, up)) != cudaSuccess ||
      (e = cudaMemcpyAsync (d_ready, h_ready, 8, cudaMemcpyHostToDevice, up)) != cudaSuccess ||
      (e = cudaMemsetAsync (d_packed + total_words, 0, 16, up)) != cudaSuccess ||
      (e = cudaEventRecord (q0.ev_up, up)) != cudaSuccess || (e = cudaStreamWaitEvent (cs, q0.ev_up, 0)) != cudaSuccess) {
    gcg_set_error ("gcg_search: meta upload: %s", cudaGetErrorString (e)); return fail (GCG_ECUDA);
  }
  {
    gcg_kscope ks (ctx, "tile_seq");
    tile_seq_kernel<<<(unsigned) ((n_tiles + 255) / 256), 256, 0, cs>>> (d_woff, n_read, n_tiles, d_tseq);
    if (cudaGetLastError () != cudaSuccess) { gcg_set_error ("gcg_search: tile_seq launch failed"); return fail (GCG_ECUDA); }
  }

  double t_gather = 0, t_wait = 0, t_submit = 0, t_drain = 0;
  auto now = [] () { return std::chrono::steady_clock::now (); };
  auto ms = [] (std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli> (b - a).count (); };
  const auto t_call = now ();
  // GCG_TIMELINE=1: device clock of the last piece's arrival and of the end of the launch, next to the host's times
  const bool timeline = getenv ("GCG_TIMELINE") != nullptr;
  cudaEvent_t tl_ref = nullptr, tl_up = nullptr, tl_kern = nullptr;
  if (timeline) { cudaEventCreate (&tl_ref); cudaEventCreate (&tl_up); cudaEventCreate (&tl_kern); cudaEventRecord (tl_ref, up); }
  gcg_workers * pool = gcg_ctx_workers (ctx);
  std::function<void (int64_t)> gather_fn;
  unsigned long long win_lo = 0, win_hi = 0;
  int64_t copied = 0, done_cnt = 0, next_g = 0, total = 0, n_piece = 0, n_flag = 0;
  int down_n = 0;
  const size_t DOWN_PIECE = (size_t) 2 << 20;
  bool launched = false;
  double t_kernel_seen = 0, t_loop_end = 0;
  int64_t left_at_end = 0;

  // everything complete so far (groups report in any order, the host reads them in order), and its download
  auto poll = [&] (size_t min_bytes) -> int {
    while (next_g < n_group) {
      const unsigned long long v = *(volatile unsigned long long *) (h_group + next_g);
      if (v == 0) break;
      done_cnt = (int64_t) (v - 1); ++next_g;
    }
    if (keep) return GCG_OK;
    for (;;) {
      const int64_t done = std::min<int64_t> (done_cnt, (int64_t) win_hi), from = std::max<int64_t> (copied, (int64_t) win_lo);
      if (done <= from || (size_t) (done - from) * res.rec < min_bytes) return GCG_OK;
      if (down_n >= 2) {
        if (cudaEventQuery (p->ev_down[down_n & 1]) == cudaErrorNotReady) return GCG_OK;
        cudaGetLastError ();
      }
      const int64_t to = std::min<int64_t> (done, from + (int64_t) (DOWN_PIECE / res.rec));
      GCG_CUDA (cudaMemcpyAsync (res.buf + (size_t) from * res.rec, (char *) d_res + (size_t) from * res.rec, (size_t) (to - from) * res.rec, cudaMemcpyDeviceToHost, down));
      GCG_CUDA (cudaEventRecord (p->ev_down[down_n & 1], down));
      ++down_n;
      copied = to;
    }
  };
  // if the host gives up while the launch is waiting for words: let it run through whatever is in the buffer
  auto unblock = [&] () {
    h_ready[0] = ~0ULL;
    cudaMemcpyAsync (d_ready, h_ready, 8, cudaMemcpyHostToDevice, up);
    cudaStreamSynchronize (up); cudaStreamSynchronize (cs); cudaStreamSynchronize (down);
    cudaGetLastError ();
  };

  struct piece_desc { int64_t r0 = 0, r1 = 0, w0 = 0, nw = 0; int slot = 0; bool direct = false; };
  auto plan = [&] (int64_t r0, piece_desc & d) {
    // growing from the first piece, and shrinking again towards the end: what is left to do when the last piece has
    // been gathered — its upload and its tiles — is proportional to its size
    int64_t target = std::min (piece_words, first_words << std::min<int64_t> (n_piece, 8));
    target = std::min (target, std::max (last_words, (total_words - woff[r0]) / 2));
    target = std::min (cap_words, std::max (max_words, target));
    int64_t r1 = r0, nw = 0;
    while (r1 < n_read) {
      const int64_t w = woff[r1 + 1] - woff[r1];
      if (r1 > r0 && nw + w > target) break;
      nw += w; ++r1;
    }
    d.r0 = r0; d.r1 = r1; d.w0 = woff[r0]; d.nw = nw; d.slot = (int) (n_piece % PIPE_SLOTS);
    d.direct = packed_pinned;
    if (d.direct) for (int64_t i = r0; i < r1; ++i) if (src.pwoff[i] - src.pwoff[r0] != woff[i] - woff[r0]) { d.direct = false; break; }
    ++n_piece;
  };
  auto start_gather = [&] (const piece_desc & d) -> int {
    pipe_slot & q = p->s[d.slot];
    auto t0 = now ();
    if (q.busy) { GCG_CUDA (cudaEventSynchronize (q.ev_free)); q.busy = false; }      // (its previous upload has left the slot)
    t_wait += ms (t0, now ());
    const int64_t nr = d.r1 - d.r0, r0 = d.r0, w0 = d.w0;
    const int64_t n_task = d.nw * 32 < (1 << 18) ? 1 : std::min<int64_t> (nr, 4 * (int64_t) ctx->host_threads);
    uint64_t * dst = q.h_packed;
    const char * const * read_seq = src.seq;
    const uint64_t * pk = src.packed; const int64_t * pwoff = src.pwoff;
    gather_fn = [=] (int64_t tk) {
      for (int64_t i = r0 + nr * tk / n_task; i < r0 + nr * (tk + 1) / n_task; ++i) {
        if (len[i] <= 0) continue;
        if (pk) gcg_copy_stream (dst + (woff[i] - w0), pk + pwoff[i], (size_t) (woff[i + 1] - woff[i]) * 8);
        else gcg_pack_stream (dst + (woff[i] - w0), read_seq[i], (size_t) len[i]);
      }
      gcg_copy_fence ();
    };
    gcg_workers_start (pool, n_task, gather_fn);
    return GCG_OK;
  };
  auto submit = [&] (const piece_desc & d) -> int {
    pipe_slot & q = p->s[d.slot];
    if (d.nw > 0)
      GCG_CUDA (cudaMemcpyAsync (d_packed + d.w0, d.direct ? (const void *) (src.packed + src.pwoff[d.r0]) : (const void *) q.h_packed, (size_t) d.nw * 8, cudaMemcpyHostToDevice, up));
    const int64_t slot_i = ++n_flag;                   // (one pinned word per flag copy: it must keep its value until the copy has run)
    h_ready[slot_i] = d.r1 == n_read ? (unsigned long long) total_words + 1ULL : (unsigned long long) (d.w0 + d.nw);
    GCG_CUDA (cudaMemcpyAsync (d_ready, h_ready + slot_i, 8, cudaMemcpyHostToDevice, up));
    if (!d.direct) { GCG_CUDA (cudaEventRecord (q.ev_free, up)); q.busy = true; }
    return GCG_OK;
  };

  // the host thread takes gather tasks between its polls when the pool is small (one GPU of eight gets 4 host threads:
  // 3 workers + this one); with 16 threads it was measured neutral to slightly slower (the polls come late)
  const bool help = getenv ("GCG_SEARCH_HELP") != nullptr ? atoi (getenv ("GCG_SEARCH_HELP")) != 0 : ctx->host_threads <= 8;
  fused_stream st;
  st.ready = d_ready; st.group_cnt = d_group; st.group_done = hd_group; st.group_shift = group_shift;
  for (int pass = 0; pass < 2 && !rc; ++pass) {
    win_hi = (unsigned long long) res.cap;
    memset (h_group, 0, (size_t) n_group * 8);
    copied = 0; done_cnt = 0; next_g = 0; down_n = 0;
    cudaError_t e2 = cudaMemsetAsync (d_group, 0, (size_t) n_group * 4, cs);
    if (e2 == cudaSuccess && fmt && pass == 0) e2 = cudaMemsetAsync (d_roff, 0xFF, (size_t) (n_read + 1) * 8, cs);
    if (e2 != cudaSuccess) { gcg_set_error ("gcg_search: %s", cudaGetErrorString (e2)); rc = GCG_ECUDA; break; }
    rc = launch_fused (ctx, t, d_packed, d_woff, d_len, d_tseq, n_read, total_words, k, fmt, d_res, win_lo, win_hi, 0, fmt ? d_roff : nullptr,
                       d_state, p->d_run, nullptr, true, nullptr, &st);
    if (rc) break;
    launched = true;
    if (pass == 0) {
      // ---- the reads, piece by piece: the next piece is gathered while this one is sent
      piece_desc cur, nxt;
      plan (0, cur);
      bool gathering = !cur.direct;
      if (gathering) rc = start_gather (cur);
      while (!rc) {
        const bool more = cur.r1 < n_read;
        if (more) plan (cur.r1, nxt);
        auto t0 = now ();
        if (gathering) {
          // (the host thread feeds the download stream and, in between, takes gather tasks like any worker)
          while (!rc && !gcg_workers_idle (pool)) { rc = poll ((size_t) 1 << 19); if (!help || !gcg_workers_help (pool)) for (int i = 0; i < 32; ++i) _mm_pause (); }
          gcg_workers_wait (pool);
        }
        t_gather += ms (t0, now ());
        gathering = false;
        if (rc) break;
        if (more && !nxt.direct) { rc = start_gather (nxt); gathering = !rc; }
        auto t1 = now ();
        if (!rc) rc = submit (cur);
        if (!rc) rc = poll ((size_t) 1 << 19);
        t_submit += ms (t1, now ());
        if (!more || rc) break;
        cur = nxt;
      }
      if (gathering) gcg_workers_wait (pool);
      if (rc) { unblock (); break; }
    }
    // ---- drain: the launch finishes on its own now; keep the download stream fed
    auto t_d0 = now ();
    t_loop_end = ms (t_call, t_d0);
    if (timeline && pass == 0) { cudaEventRecord (tl_up, up); cudaEventRecord (tl_kern, cs); }
    cudaError_t qe;
    while (!rc && (qe = cudaStreamQuery (cs)) == cudaErrorNotReady) { rc = poll ((size_t) 1 << 19); for (int i = 0; i < 64; ++i) _mm_pause (); }
    cudaGetLastError ();
    if (!rc) { t_kernel_seen = ms (t_call, now ()); left_at_end = done_cnt - copied; }
    if (!rc && cudaStreamSynchronize (cs) != cudaSuccess) { gcg_set_error ("gcg_search: %s", cudaGetErrorString (cudaGetLastError ())); rc = GCG_ECUDA; }
    if (rc) { unblock (); break; }
    rc = poll (1);
    if (!rc && next_g != n_group) { gcg_set_error ("gcg_search: the launch ended with %lld of %lld groups reported", (long long) next_g, (long long) n_group); rc = GCG_ECUDA; }
    if (rc) break;
    total = done_cnt;
    // the tail the loop has not queued yet (pieces may still be in flight on the download stream), and the read offsets
    if (!keep) {
      const int64_t done = std::min<int64_t> (total, res.cap), from = std::max<int64_t> (copied, (int64_t) win_lo);
      cudaError_t e3 = cudaSuccess;
      if (done > from) e3 = cudaMemcpyAsync (res.buf + (size_t) from * res.rec, (char *) d_res + (size_t) from * res.rec, (size_t) (done - from) * res.rec, cudaMemcpyDeviceToHost, down);
      if (e3 == cudaSuccess && fmt && total <= res.cap) e3 = cudaMemcpyAsync (res.read_off, d_roff, (size_t) n_read * 8, cudaMemcpyDeviceToHost, down);
      if (e3 == cudaSuccess) e3 = cudaStreamSynchronize (down);
      if (e3 != cudaSuccess) { gcg_set_error ("gcg_search: result download: %s", cudaGetErrorString (e3)); rc = GCG_ECUDA; break; }
    }    // (keep: the offsets stay on the device too and are completed there, below)
    t_drain += ms (t_d0, now ());
    if (total <= res.cap) break;
    // denser than estimated: anchors [0, cap) are in place; a result of the right size takes them over and a second
    // launch over the words already on the device materialises [cap, total) only
    if (pass != 0) { gcg_set_error ("gcg_search: anchor count changed between passes (%lld > %lld)", (long long) total, (long long) res.cap); rc = GCG_ECUDA; break; }
    if (ctx->trace) fprintf (stderr, "[gcg]   streaming search: result sized for %lld anchors, %lld found: second launch for the tail\n", (long long) res.cap, (long long) total);
    void * d_new = nullptr;
    cudaError_t e4 = gcg_dmalloc (ctx, &d_new, (size_t) total * res.rec);
    if (e4 == cudaSuccess) e4 = cudaMemcpyAsync (d_new, d_res, (size_t) res.cap * res.rec, cudaMemcpyDeviceToDevice, cs);
    if (e4 != cudaSuccess) { if (d_new) gcg_dfree (ctx, d_new); gcg_set_error ("gcg_search: device result of %lld anchors: %s", (long long) total, cudaGetErrorString (e4)); rc = GCG_ENOMEM; break; }
    gcg_dfree (ctx, d_res);                            // (stream ordered behind the copy: parked blocks are reused by later work on the same stream)
    d_res = d_new;
    if (!keep) {
      char * nb = (char *) gcg_pinned_alloc ((size_t) total * res.rec);
      if (nb == nullptr) { gcg_set_error ("gcg_search: pinned alloc of %lld anchors failed", (long long) total); rc = GCG_ENOMEM; break; }
      gcg_par_memcpy (ctx, nb, res.buf, (size_t) res.cap * res.rec);
      gcg_free (res.buf);
      res.buf = nb;
    }
    win_lo = (unsigned long long) res.cap;
    res.cap = total;
  }
  (void) launched;
  if (ctx->trace)
    fprintf (stderr, "[gcg]   search pipeline (streaming, one launch): %lld pieces, waits for the gather %.3f ms (%d threads), for a free slot %.3f ms; "
             "host: submit %.3f drain %.3f ms; last piece sent at %.3f, launch seen finished at %.3f (%lld reported anchors not yet queued for download), all done at %.3f ms\n",
             (long long) n_piece, t_gather, ctx->host_threads, t_wait, t_submit, t_drain, t_loop_end, t_kernel_seen, (long long) left_at_end, ms (t_call, now ()));
  if (timeline) {
    float t_up = -1, t_k = -1;
    cudaEventSynchronize (tl_kern);
    cudaEventElapsedTime (&t_up, tl_ref, tl_up); cudaEventElapsedTime (&t_k, tl_ref, tl_kern);
    fprintf (stderr, "[gcg]   timeline (device clock, ms after the start): last piece arrived at %.3f, launch ended at %.3f\n", t_up, t_k);
    cudaEventDestroy (tl_ref); cudaEventDestroy (tl_up); cudaEventDestroy (tl_kern);
  }
  gcg_trace_mark (ctx, "search: reads -> anchors (streaming)");
  for (pipe_slot & q : p->s) { q.busy = false; q.pending = false; }
  if (rc) return fail (rc);
  res.n = total;
  p->last_total = total;
  if (fmt && !keep) read_off_fill (res.read_off, n_read, total);
  if (keep) {
    // the reduction kernels read read_off[r + 1] of every read: complete the offsets where they are (the host copy
    // a kept search hands back is only a placeholder its callers release)
    {
      gcg_kscope ks (ctx, "read_off_fill");
      read_off_fill_kernel<<<1, 1024, 0, cs>>> (d_roff, n_read, (long long) total);
    }
    if (cudaGetLastError () != cudaSuccess) { gcg_set_error ("gcg_search: read offset fill launch failed"); return fail (GCG_ECUDA); }
    if (fmt) memset (res.read_off, 0, (size_t) (n_read + 1) * 8);
    keep->d_anchors = d_res; keep->d_read_off = d_roff; d_res = nullptr; d_roff = nullptr;
  }
  release ();
  if (fmt) *read_off_out = res.read_off;
  *n_hit = total;
  if (keep) return GCG_OK;
  if (total == 0) { gcg_free (res.buf); return GCG_OK; }
  *hits_out = res.buf;
  return GCG_OK;
}

static int search_host_impl (gcg_ctx * ctx, gcg_table * t, const read_source & src, const int32_t * read_len,
                             int64_t n_read, int k, int fmt, void ** hits_out, int64_t * n_hit, int64_t ** read_off_out,
                             search_dev_keep * keep = nullptr)
{
  const char * const * read_seq = src.seq;
  GCG_CHECK (ctx && t && hits_out && n_hit && n_read >= 0 && (n_read == 0 || ((read_seq || (src.packed && src.pwoff)) && read_len)), GCG_EINVAL, "gcg_search: bad argument");
  // packed words in page-locked memory go over PCIe from where they lie, chunk by chunk (no host pass at all)
  bool packed_pinned = false;
  if (src.packed) {
    cudaPointerAttributes at;
    packed_pinned = cudaPointerGetAttributes (&at, src.packed) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError ();
  }
  GCG_CHECK (k == t->k, GCG_EINVAL, "gcg_search: k=%d but the table was built with k=%d", k, t->k);
  GCG_CHECK (n_read < 0x7FFFFFFF, GCG_ERANGE, "gcg_search: too many reads");
  *hits_out = nullptr;
  *n_hit = 0;
  if (read_off_out) *read_off_out = nullptr;
  {
    // one launch over the whole read set while it is being uploaded (default), else the chunked pipeline below
    bool handled = false;
    const int src_rc = search_host_stream (ctx, t, src, read_len, n_read, k, fmt, hits_out, n_hit, read_off_out, keep, &handled);
    if (handled) return src_rc;
  }
  GCG_CUDA (cudaSetDevice (ctx->device));
  gcg_trace_mark (ctx, nullptr);
  // Chunks GROW: 4, 8, 16, 32, 32, ... MiB of bases (and shrink again towards the end).  A chunk's kernel runs the deferred-emit pipeline of
  // k45_fused_kernel, which needs several rounds per block to hide its look-back and emit phases: a launch over 8 MiB is
  // a single round (measured 100-116 us per 8 MiB chunk against 50 us per 8 MiB inside one launch over the whole read
  // set).  Large chunks fix that, a small first chunk keeps the time until the GPU has something to do short.
  // GCG_SEARCH_CHUNK_BYTES fixes one size for all chunks (tests; A/B timing).
  int64_t chunk_words = ((int64_t) 32 << 20) / 32, first_words = ((int64_t) 4 << 20) / 32, last_words = ((int64_t) 8 << 20) / 32, max_words = 0, total_kmers = 0;
  if (const char * e = getenv ("GCG_SEARCH_CHUNK_MAX_MB")) chunk_words = std::max<int64_t> (1, atoll (e)) * (1 << 20) / 32;      // (A/B timing of the schedule)
  if (const char * e = getenv ("GCG_SEARCH_CHUNK_FIRST_MB")) first_words = std::max<int64_t> (1, atoll (e)) * (1 << 20) / 32;
  if (const char * e = getenv ("GCG_SEARCH_CHUNK_LAST_MB")) last_words = std::max<int64_t> (1, atoll (e)) * (1 << 20) / 32;
  if (const char * e = getenv ("GCG_SEARCH_CHUNK_BYTES")) chunk_words = first_words = last_words = std::max<int64_t> (1, atoll (e) / 32);
  int64_t total_words = 0, words_done = 0;
  int32_t max_len = 0;
  for (int64_t r = 0; r < n_read; ++r) {
    GCG_CHECK (read_len[r] >= 0, GCG_ERANGE, "gcg_search: read %lld has a negative length", (long long) r);
    max_words = std::max<int64_t> (max_words, ((int64_t) read_len[r] + 31) >> 5);
    max_len = std::max (max_len, read_len[r]);
    total_words += ((int64_t) read_len[r] + 31) >> 5;
    if (read_len[r] >= k) total_kmers += (int64_t) read_len[r] - k + 1;
  }
  int rc = GCG_OK;
  if (fmt && (rc = compact_limits_ok (t, max_len)) != 0) return rc;
  const int mode = keep ? 1 : pipe_mode ();
  const bool zc = mode != 0;                         // anchors land at their final index (direct or zero copy)
  search_result res;
  res.ctx = ctx; res.fmt = fmt; res.rec = anchor_bytes (fmt);
  if (fmt) {
    res.read_off = (int64_t *) gcg_pinned_alloc ((size_t) (n_read + 1) * 8);
    GCG_CHECK (res.read_off != nullptr, GCG_ENOMEM, "gcg_search: pinned alloc of %lld read offsets failed", (long long) n_read + 1);
    memset (res.read_off, 0xFF, (size_t) (n_read + 1) * 8);
  }
  if (total_kmers == 0) {
    if (fmt) { read_off_fill (res.read_off, n_read, 0); *read_off_out = res.read_off; }
    return GCG_OK;
  }
  rc = pipe_reserve (ctx, std::max (chunk_words, max_words), !zc);
  if (!rc) rc = gcg_table_filter_ensure (ctx, t);
  if (rc) { gcg_free (res.read_off); return rc; }
  gcg_pipe * p = ctx->pipe;
  const int64_t cap_words = std::max (chunk_words, max_words);       // <= p->cap_words

  // (10 %-error ONT reads anchor 6.3 % of their 25-mers; an estimate that is too small costs a second pass)
  res.cap = std::max<int64_t> (p->last_total + p->last_total / 8, total_kmers / 10) + 4096;
  if (const char * e = getenv ("GCG_SEARCH_RES_CAP")) res.cap = std::max<int64_t> (1, atoll (e));     // test hook: force the grow path
  if (!keep) {
    res.buf = (char *) gcg_pinned_alloc ((size_t) res.cap * res.rec);
    if (res.buf == nullptr) { gcg_free (res.read_off); gcg_set_error ("gcg_search: pinned alloc of %lld anchors failed", (long long) res.cap); return GCG_ENOMEM; }
  }

  // work on ctx->stream enqueued by earlier calls (the table build) precedes the first probe by stream order
  struct chunk_desc { int64_t r0 = 0, r1 = 0, nr = 0, nw = 0, kmers = 0, n_tiles = 0; size_t tseq_off = 0; int slot = 0; bool direct = false; };
  int inflight[PIPE_SLOTS], n_inflight = 0;       // staged mode: submitted, not yet downloaded, oldest first
  int64_t c = 0, n_sub = 0;
  unsigned long long win_lo = 0, win_hi = 0;      // zero copy: window of global anchor indices this pass materialises
  void * d_out = nullptr;                         // direct: the device result; zero copy: device alias of the pinned result
  long long * d_roff = nullptr;
  void * d_res = nullptr; long long * d_res_roff = nullptr;     // direct: device blocks behind d_out / d_roff
  int64_t copied = 0;                             // direct: anchors already queued for download
  double t_gather = 0, t_wait = 0, t_prepare = 0, t_submit = 0, t_download = 0, t_drain = 0;
  auto now = [] () { return std::chrono::steady_clock::now (); };
  auto ms = [] (std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli> (b - a).count (); };
  gcg_workers * pool = gcg_ctx_workers (ctx);
  std::function<void (int64_t)> gather_fn;
  // GCG_TIMELINE=1: per chunk, when its gather was done and it was submitted (host clock) and when its upload and its
  // kernel ended (device clock), all in ms after the start of the pass
  struct tl_rec { double gathered = 0, submitted = 0; cudaEvent_t up = nullptr, kern = nullptr; int64_t words = 0; };
  const bool timeline = getenv ("GCG_TIMELINE") != nullptr;
  std::vector<tl_rec> tlv;
  cudaEvent_t tl_ref = nullptr;
  std::chrono::steady_clock::time_point tl_t0;

  // chunk starting at read r0: whole reads, at most cap_words words, meta within the slot's meta block;
  // waits for its slot and writes the meta block (host work that overlaps the gather of the chunk before)
  auto plan = [&] (int64_t r0, chunk_desc & d) -> int {
    int64_t r1 = r0, nw = 0;
    // this chunk's size: growing from the first, and shrinking again towards the end (what comes after the last
    // kernel — the download of its anchors — is proportional to the last chunk)
    int64_t target = std::min (cap_words, first_words << std::min<int64_t> (c, 8));
    target = std::min (target, std::max (last_words, (total_words - words_done) / 2));
    target = std::max (max_words, target);
    while (r1 < n_read) {
      const int64_t w = ((int64_t) read_len[r1] + 31) >> 5;
      if (r1 > r0 && nw + w > target) break;
      const int64_t nr = r1 - r0 + 1;
      if (r1 > r0 && (size_t) ((nr + 2) * 12 + ((nw + w + 31) >> 5) * 4 + 64) > p->meta_cap) break;
      nw += w; ++r1;
    }
    d.r0 = r0; d.r1 = r1; d.nr = r1 - r0; d.nw = nw; d.n_tiles = (nw + 31) >> 5; d.kmers = 0;
    words_done += nw;
    d.slot = (int) (c % PIPE_SLOTS);
    ++c;
    pipe_slot & q = p->s[d.slot];
    auto t0 = now ();
    if (q.pending) { gcg_set_error ("gcg_search: pipeline slot reused before its download"); return GCG_ECUDA; }   // PIPE_SLOTS > PIPE_LAG + 1
    if (q.busy) { GCG_CUDA (cudaEventSynchronize (q.ev_free)); q.busy = false; }
    t_wait += ms (t0, now ());
    if (nw == 0) return GCG_OK;
    const int64_t nr = d.nr;
    int64_t * woff = (int64_t *) q.h_meta;
    int32_t * len = (int32_t *) (q.h_meta + (size_t) (nr + 1) * 8);
    d.tseq_off = ((size_t) (nr + 1) * 8 + (size_t) nr * 4 + 7) & ~(size_t) 7;
    int32_t * tseq = (int32_t *) (q.h_meta + d.tseq_off);
    int64_t w = 0;
    d.direct = packed_pinned;
    for (int64_t i = 0; i < nr; ++i) {
      woff[i] = w; len[i] = read_len[r0 + i];
      if (d.direct && src.pwoff[r0 + i] - src.pwoff[r0] != w) d.direct = false;      // (the caller's words of this chunk are not one run)
      w += ((int64_t) read_len[r0 + i] + 31) >> 5;
      if (read_len[r0 + i] >= k) d.kmers += (int64_t) read_len[r0 + i] - k + 1;
    }
    woff[nr] = w;
    int64_t cur = 0;
    for (int64_t tl = 0; tl < d.n_tiles; ++tl) {
      while (cur + 1 < nr && woff[cur + 1] <= (tl << 5)) ++cur;
      tseq[tl] = (int32_t) cur;
    }
    return GCG_OK;
  };
  // starts the gather of a planned chunk on the pool (one job at a time: after the wait for the previous one)
  auto start_gather = [&] (const chunk_desc & d) {
    pipe_slot & q = p->s[d.slot];
    const int64_t nr = d.nr, nw = d.nw, r0 = d.r0;
    const int64_t * woff = (const int64_t *) q.h_meta;
    const int32_t * len = (const int32_t *) (q.h_meta + (size_t) (nr + 1) * 8);
    const int64_t n_task = nw * 32 < (1 << 18) ? 1 : std::min<int64_t> (nr, 4 * (int64_t) ctx->host_threads);
    uint64_t * dst = q.h_packed;
    if (src.packed) {
      const uint64_t * pk = src.packed; const int64_t * pwoff = src.pwoff;
      gather_fn = [=] (int64_t tk) {                 // the caller's packed words, read by read (8 bytes per 32 bases)
        for (int64_t i = nr * tk / n_task; i < nr * (tk + 1) / n_task; ++i)
          if (len[i] > 0) gcg_copy_stream (dst + woff[i], pk + pwoff[r0 + i], (size_t) (((int64_t) len[i] + 31) >> 5) * 8);
        gcg_copy_fence ();
      };
    } else
    gather_fn = [=] (int64_t tk) {                   // gather + 2-bit pack in one pass over the caller's strings
      for (int64_t i = nr * tk / n_task; i < nr * (tk + 1) / n_task; ++i)
        if (len[i] > 0) gcg_pack_stream (dst + woff[i], read_seq[r0 + i], (size_t) len[i]);
      gcg_copy_fence ();
    };
    gcg_workers_start (pool, n_task, gather_fn);
  };

  // direct mode: queue the download of the anchors that have become complete since the last look (at least min_bytes),
  // in pieces of at most DOWN_PIECE bytes, two in flight: the uploads of the next chunks are on the critical path, and
  // behind a burst of queued downloads they slow to a third (measured: 0.5 MB up in 155 us behind 8 MB of queued
  // pieces), while one piece at a time leaves the download stream idle between polls
  const size_t DOWN_PIECE = (size_t) 2 << 20;
  int down_n = 0;                                   // pieces queued so far (piece i is followed by ev_down[i & 1])
  auto download_ready = [&] (size_t min_bytes) -> int {
    for (;;) {
      const int64_t done = std::min<int64_t> ((int64_t) *(volatile unsigned long long *) p->h_count, (int64_t) win_hi);
      const int64_t from = std::max<int64_t> (copied, (int64_t) win_lo);
      if (done <= from || (size_t) (done - from) * res.rec < min_bytes) return GCG_OK;
      if (down_n >= 2) {                            // the piece before the last one must have landed
        if (cudaEventQuery (p->ev_down[down_n & 1]) == cudaErrorNotReady) return GCG_OK;
        cudaGetLastError ();
      }
      const int64_t to = std::min<int64_t> (done, from + (int64_t) (DOWN_PIECE / res.rec));
      GCG_CUDA (cudaMemcpyAsync (res.buf + (size_t) from * res.rec, (char *) d_res + (size_t) from * res.rec, (size_t) (to - from) * res.rec, cudaMemcpyDeviceToHost, p->down));
      GCG_CUDA (cudaEventRecord (p->ev_down[down_n & 1], p->down));
      ++down_n;
      copied = to;
    }
  };

  // copy the gathered chunk to the device and enqueue its kernel: probe, ordered anchor index, ONT-side
  // multiplicity and anchor records in one launch (k45_fused_kernel)
  auto submit = [&] (const chunk_desc & d) -> int {
    pipe_slot & q = p->s[d.slot];
    const int64_t nr = d.nr, nw = d.nw;
    const size_t meta_bytes = d.tseq_off + (size_t) d.n_tiles * 4;
    GCG_CUDA (cudaMemcpyAsync (q.d_packed, d.direct ? (const void *) (src.packed + src.pwoff[d.r0]) : (const void *) q.h_packed, (size_t) nw * 8, cudaMemcpyHostToDevice, p->up));
    GCG_CUDA (cudaMemcpyAsync (q.d_meta, q.h_meta, meta_bytes, cudaMemcpyHostToDevice, p->up));
    GCG_CUDA (cudaEventRecord (q.ev_up, p->up));
    GCG_CUDA (cudaStreamWaitEvent (ctx->stream, q.ev_up, 0));
    if (timeline) {
      tl_rec & tr = tlv.back ();
      tr.words = nw; tr.submitted = ms (tl_t0, now ());
      cudaEventCreate (&tr.up); cudaEventCreate (&tr.kern);
      cudaEventRecord (tr.up, p->up);
    }
    const int64_t * d_woff = (const int64_t *) q.d_meta;
    const int32_t * d_len = (const int32_t *) (q.d_meta + (size_t) (nr + 1) * 8);
    const int32_t * d_tseq = (const int32_t *) (q.d_meta + d.tseq_off);
    int e;
    if (zc) {
      // anchors and read offsets go straight to their final place in the pinned result; the running count
      // alternates between two device words
      e = launch_fused (ctx, t, q.d_packed, d_woff, d_len, d_tseq, nr, nw, k, fmt, d_out, win_lo, win_hi, (int32_t) d.r0,
                        fmt ? d_roff + d.r0 : nullptr, q.d_state, p->d_run + ((n_sub + 1) & 1), p->d_run + (n_sub & 1), true,
                        mode == 1 ? p->hd_count : nullptr);
      ++n_sub;
      if (e) return e;
      GCG_CUDA (cudaEventRecord (q.ev_free, ctx->stream));         // the slot is free again when its kernel has finished
      if (timeline) cudaEventRecord (tlv.back ().kern, ctx->stream);
      if (mode == 1 && !keep) {
        post_count_kernel<<<1, 1, 0, ctx->stream>>> (p->d_run + (n_sub & 1), p->hd_count);      // (n_sub was advanced: the word this launch wrote)
        GCG_CUDA (cudaGetLastError ());
      }
      q.busy = true;
      if (mode == 1 && !keep) return download_ready ((size_t) 1 << 20);     // whatever earlier launches have completed by now goes down
      return GCG_OK;
    }
    e = launch_fused (ctx, t, q.d_packed, d_woff, d_len, d_tseq, nr, nw, k, fmt, q.d_hits, 0ULL, ~0ULL, (int32_t) d.r0,
                      fmt ? q.d_read_off : nullptr, q.d_state, p->hd_count + d.slot);
    if (e) return e;
    GCG_CUDA (cudaEventRecord (q.ev_emit, ctx->stream));
    GCG_CUDA (cudaStreamWaitEvent (p->down, q.ev_emit, 0));
    q.pending = true;
    q.first_read = d.r0; q.n_read = nr;
    return GCG_OK;
  };

  int pipe_lag = 1;                                 // (measured on cfg2: 8 MiB chunks, lag 1: 3.0 ms; 16 MiB, lag 2: 3.35 ms)
  if (const char * e = getenv ("GCG_SEARCH_LAG")) pipe_lag = std::min (PIPE_LAG, std::max (0, atoi (e)));

  // one pass over all chunks
  auto run_pass = [&] () -> int {
    int prc = GCG_OK;
    chunk_desc cur_c, next_c;
    c = 0; n_sub = 0; n_inflight = 0; words_done = 0;
    if (zc) GCG_CUDA (cudaMemsetAsync (p->d_run, 0, 2 * sizeof (unsigned long long), ctx->stream));
    p->h_count[0] = 0;
    copied = 0;
    down_n = 0;
    if (timeline) { tlv.clear (); cudaEventCreate (&tl_ref); cudaEventRecord (tl_ref, p->up); tl_t0 = now (); }
    prc = plan (0, cur_c);
    bool gathering = !prc && cur_c.kmers > 0 && !cur_c.direct;
    if (gathering) start_gather (cur_c);
    while (!prc) {
      // the next chunk is planned while this one is gathered, and gathered while this one is copied and probed
      const bool more = cur_c.r1 < n_read;
      if (more) {
        auto t1 = now ();
        prc = plan (cur_c.r1, next_c);
        t_prepare += ms (t1, now ());
      }
      auto t0 = now ();
      if (gathering) {
        // (the download stream is fed while the pool gathers: the host thread has nothing else to do)
        if (mode == 1 && !keep)
          while (!prc && !gcg_workers_idle (pool)) { prc = download_ready ((size_t) 1 << 19); for (int i = 0; i < 32; ++i) _mm_pause (); }
        gcg_workers_wait (pool);
      }
      t_gather += ms (t0, now ());
      gathering = false;
      if (timeline) { tlv.emplace_back (); tlv.back ().gathered = ms (tl_t0, now ()); }
      if (prc) break;
      if (more && next_c.kmers > 0 && !next_c.direct) { start_gather (next_c); gathering = true; }
      if (cur_c.kmers > 0) {
        auto t1 = now ();
        prc = submit (cur_c);
        t_submit += ms (t1, now ());
        if (prc) break;
        if (!zc) {
          // place the anchors of the chunk submitted PIPE_LAG chunks ago: its count is there by now
          inflight[n_inflight++] = cur_c.slot;
          if (n_inflight > pipe_lag) {
            auto t2 = now ();
            prc = pipe_download (ctx, p->s[inflight[0]], inflight[0], res);
            t_download += ms (t2, now ());
            if (prc) break;
            for (int i = 1; i < n_inflight; ++i) inflight[i - 1] = inflight[i];
            --n_inflight;
          }
        }
      }
      if (!more) break;
      cur_c = next_c;
    }
    if (gathering) gcg_workers_wait (pool);
    // ---- drain
    auto t_d0 = now ();
    for (int i = 0; i < n_inflight && !prc; ++i) prc = pipe_download (ctx, p->s[inflight[i]], inflight[i], res);
    if (mode == 1 && !keep && !prc) {
      // the kernels still queued finish one after the other: keep the download stream fed while they do
      while (!prc && cudaStreamQuery (ctx->stream) == cudaErrorNotReady) {
        prc = download_ready ((size_t) 1 << 19);
        for (int i = 0; i < 64; ++i) _mm_pause ();
      }
      cudaGetLastError ();
    }
    bool have_total = false;
    if (mode == 1 && !keep && !prc && n_sub > 0 && cudaStreamQuery (ctx->stream) == cudaSuccess) {
      ctx->h_counters[4] = *(volatile unsigned long long *) p->h_count;      // the last launch's post_count_kernel has run
      have_total = true;
    }
    cudaGetLastError ();
    if (zc && !prc && n_sub > 0 && !have_total &&
        cudaMemcpyAsync (ctx->h_counters + 4, p->d_run + (n_sub & 1), 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) { gcg_set_error ("gcg_search: count copy failed"); prc = GCG_ECUDA; }
    if (cudaStreamSynchronize (p->down) != cudaSuccess || cudaStreamSynchronize (ctx->stream) != cudaSuccess) {
      if (!prc) { gcg_set_error ("gcg_search: %s", cudaGetErrorString (cudaGetLastError ())); prc = GCG_ECUDA; }
    }
    t_drain += ms (t_d0, now ());
    if (timeline) {
      fprintf (stderr, "[gcg]   timeline (ms after the start of the pass; drain began at %.3f, pass ended at %.3f)\n", ms (tl_t0, t_d0), ms (tl_t0, now ()));
      for (size_t i = 0; i < tlv.size (); ++i) {
        tl_rec & tr = tlv[i];
        float up = -1, kn = -1;
        if (tr.up) { cudaEventElapsedTime (&up, tl_ref, tr.up); cudaEventElapsedTime (&kn, tl_ref, tr.kern); cudaEventDestroy (tr.up); cudaEventDestroy (tr.kern); }
        fprintf (stderr, "[gcg]     chunk %2zu %6.2f MiB: gathered %.3f submitted %.3f uploaded %.3f kernel done %.3f\n", i, (double) tr.words * 32 / (1 << 20), tr.gathered, tr.submitted, up, kn);
      }
      cudaEventDestroy (tl_ref);
    }
    for (pipe_slot & q : p->s) { q.busy = false; q.pending = false; }
    return prc;
  };

  if (zc) {
    int64_t total = 0;
    win_lo = 0;
    for (int pass = 0; pass < 2 && !rc; ++pass) {
      win_hi = (unsigned long long) res.cap;
      if (mode == 2) {
        if (cudaHostGetDevicePointer (&d_out, res.buf, 0) != cudaSuccess || (fmt && cudaHostGetDevicePointer ((void **) &d_roff, res.read_off, 0) != cudaSuccess)) {
          gcg_set_error ("gcg_search: the pinned result is not mapped into the device's address space: %s", cudaGetErrorString (cudaGetLastError ()));
          rc = GCG_ECUDA;
          break;
        }
      } else {
        // the device result of this pass (parked blocks of the context: no driver call in the steady state)
        void * d_prev = d_res;                             // (kept anchors of a first pass that overflowed)
        d_res = nullptr;
        cudaError_t e = gcg_dmalloc (ctx, &d_res, (size_t) res.cap * res.rec);
        if (e == cudaSuccess && d_prev) e = cudaMemcpyAsync (d_res, d_prev, (size_t) win_lo * res.rec, cudaMemcpyDeviceToDevice, ctx->stream);
        if (d_prev) gcg_dfree (ctx, d_prev);
        if (e == cudaSuccess && fmt) e = gcg_dmalloc (ctx, &d_res_roff, (size_t) (n_read + 1) * 8);
        if (e == cudaSuccess && fmt) e = cudaMemsetAsync (d_res_roff, 0xFF, (size_t) (n_read + 1) * 8, ctx->stream);
        if (e != cudaSuccess) { gcg_set_error ("gcg_search: device result of %lld anchors: %s", (long long) res.cap, cudaGetErrorString (e)); rc = GCG_ENOMEM; break; }
        d_out = d_res; d_roff = d_res_roff;
      }
      rc = run_pass ();
      if (rc) break;
      total = (int64_t) ctx->h_counters[4];
      if (mode == 1) {
        // the tail the loop has not queued yet, and the read offsets
        const int64_t done = std::min<int64_t> (total, res.cap), from = std::max<int64_t> (copied, (int64_t) win_lo);
        cudaError_t e = cudaSuccess;
        if (!keep && done > from) e = cudaMemcpyAsync (res.buf + (size_t) from * res.rec, (char *) d_res + (size_t) from * res.rec, (size_t) (done - from) * res.rec, cudaMemcpyDeviceToHost, p->down);
        if (e == cudaSuccess && fmt) e = cudaMemcpyAsync (res.read_off, d_res_roff, (size_t) n_read * 8, cudaMemcpyDeviceToHost, p->down);
        if (e == cudaSuccess) e = cudaStreamSynchronize (p->down);
        if (!keep) { gcg_dfree (ctx, d_res); d_res = nullptr; }
        if (!keep || total > res.cap) { gcg_dfree (ctx, d_res_roff); d_res_roff = nullptr; }
        if (e != cudaSuccess) { gcg_set_error ("gcg_search: result download: %s", cudaGetErrorString (e)); rc = GCG_ECUDA; break; }
      }
      if (total <= res.cap) break;
      if (keep) {      // the second pass writes [cap, total) behind the kept anchors [0, cap) in a device result of the right size
        GCG_CHECK (pass == 0, GCG_ECUDA, "gcg_search: anchor count changed between passes (%lld > %lld)", (long long) total, (long long) res.cap);
        win_lo = (unsigned long long) res.cap;
        res.cap = total;
        continue;
      }
      // denser than estimated: anchors [0, cap) are in place and counted; a result of the right size takes them over
      // and a second pass over the reads materialises [cap, total) only
      GCG_CHECK (pass == 0, GCG_ECUDA, "gcg_search: anchor count changed between passes (%lld > %lld)", (long long) total, (long long) res.cap);
      char * nb = (char *) gcg_pinned_alloc ((size_t) total * res.rec);
      if (nb == nullptr) { gcg_set_error ("gcg_search: pinned alloc of %lld anchors failed", (long long) total); rc = GCG_ENOMEM; break; }
      gcg_par_memcpy (ctx, nb, res.buf, (size_t) res.cap * res.rec);
      gcg_free (res.buf);
      win_lo = (unsigned long long) res.cap;
      res.buf = nb; res.cap = total;
      if (ctx->trace) fprintf (stderr, "[gcg]   search pipeline: result sized for %lld anchors, %lld found: second pass for the tail\n", (long long) win_lo, (long long) total);
    }
    if (keep && !rc) {
      // filled offsets go back up: the reduction kernels read read_off[r + 1] of every read
      read_off_fill (res.read_off, n_read, total);
      if (cudaMemcpyAsync (d_res_roff, res.read_off, (size_t) (n_read + 1) * 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
          cudaStreamSynchronize (ctx->stream) != cudaSuccess) { gcg_set_error ("gcg_search: read offsets upload failed"); rc = GCG_ECUDA; }
      else { keep->d_anchors = d_res; keep->d_read_off = d_res_roff; d_res = nullptr; d_res_roff = nullptr; }
    }
    if (d_res) gcg_dfree (ctx, d_res);
    if (d_res_roff) gcg_dfree (ctx, d_res_roff);
    res.n = total;
  } else {
    rc = run_pass ();
  }
  if (ctx->trace)
    fprintf (stderr, "[gcg]   search pipeline (%s): %lld chunks, waits for the gather %.3f ms (%d threads), for a free slot %.3f ms; "
             "host: prepare %.3f submit %.3f download %.3f drain %.3f ms\n", mode == 2 ? "zero copy" : mode == 1 ? "direct" : "staged",
             (long long) c, t_gather, ctx->host_threads, t_wait, t_prepare, t_submit, t_download, t_drain);
  gcg_trace_mark (ctx, "search: reads -> anchors (pipelined)");
  if (rc) { gcg_free (res.buf); gcg_free (res.read_off); return rc; }
  p->last_total = res.n;
  if (fmt) {
    // staged mode: chunk-relative offsets -> offsets into the whole anchor array
    for (size_t ci = 0; ci < res.chunk_base.size (); ++ci) {
      const int64_t r0 = res.chunk_base[ci].first, base = res.chunk_base[ci].second;
      const int64_t r1 = ci + 1 < res.chunk_base.size () ? res.chunk_base[ci + 1].first : n_read;
      if (base) for (int64_t r = r0; r < r1; ++r) if (res.read_off[r] >= 0) res.read_off[r] += base;
    }
    read_off_fill (res.read_off, n_read, res.n);
    *read_off_out = res.read_off;
  }
  if (keep) { *n_hit = res.n; return GCG_OK; }
  if (res.n == 0) { gcg_free (res.buf); return GCG_OK; }
  *hits_out = res.buf;
  *n_hit = res.n;
  return GCG_OK;
}

int gcg_runs_from_anchors (gcg_ctx * ctx, const gcg_table * t, const void * d_anchors, const long long * d_read_off, int64_t n_read,
                           gcg_run ** runs_out, int64_t ** run_off_out, int64_t * n_run);      // runs.cu

// N3: search + the anchor grouping of map_ont2contigs (ctg_graph.c:600-656) on the device; neither the anchors
// nor anything per ONT base reaches the host
static int search_runs_impl (gcg_ctx * ctx, gcg_table * t, const char * const * read_seq, const uint64_t * packed, const int64_t * pwoff, const int32_t * read_len,
                             int64_t n_read, int k, gcg_run ** runs_out, int64_t ** run_off_out, int64_t * n_run, int64_t * n_anchor)
{
  GCG_CHECK (runs_out && run_off_out && n_run && n_anchor, GCG_EINVAL, "gcg_search_runs: bad argument");
  *runs_out = nullptr; *run_off_out = nullptr; *n_run = 0; *n_anchor = 0;
  search_dev_keep keep;
  void * none = nullptr;
  int64_t * read_off = nullptr;
  read_source src; src.seq = read_seq; src.packed = packed; src.pwoff = pwoff;
  int rc = search_host_impl (ctx, t, src, read_len, n_read, k, 1, &none, n_anchor, &read_off, &keep);
  if (rc) return rc;
  if (keep.d_read_off == nullptr) {
    // nothing was searched (no read holds a k-mer): every read has zero runs
    int64_t * ro = (int64_t *) gcg_pinned_alloc ((size_t) (n_read + 1) * 8);
    gcg_free (read_off);
    GCG_CHECK (ro != nullptr, GCG_ENOMEM, "gcg_search_runs: pinned alloc failed");
    memset (ro, 0, (size_t) (n_read + 1) * 8);
    *run_off_out = ro;
    return GCG_OK;
  }
  gcg_free (read_off);
  gcg_trace_mark (ctx, nullptr);
  rc = gcg_runs_from_anchors (ctx, t, keep.d_anchors, keep.d_read_off, n_read, runs_out, run_off_out, n_run);
  gcg_trace_mark (ctx, "search: anchors -> runs");
  gcg_dfree (ctx, keep.d_anchors); gcg_dfree (ctx, keep.d_read_off);
  return rc;
}

extern "C" int gcg_search_runs (gcg_ctx * ctx, gcg_table * t, const char * const * read_seq, const int32_t * read_len,
                                int64_t n_read, int k, gcg_run ** runs_out, int64_t ** run_off_out, int64_t * n_run, int64_t * n_anchor)
{
  return search_runs_impl (ctx, t, read_seq, nullptr, nullptr, read_len, n_read, k, runs_out, run_off_out, n_run, n_anchor);
}

extern "C" int gcg_search_runs_packed (gcg_ctx * ctx, gcg_table * t, const uint64_t * packed, const int64_t * woff, const int32_t * read_len,
                                       int64_t n_read, int k, gcg_run ** runs_out, int64_t ** run_off_out, int64_t * n_run, int64_t * n_anchor)
{
  return search_runs_impl (ctx, t, nullptr, packed, woff, read_len, n_read, k, runs_out, run_off_out, n_run, n_anchor);
}

extern "C" int gcg_search (gcg_ctx * ctx, gcg_table * t, const char * const * read_seq, const int32_t * read_len,
                           int64_t n_read, int k, gcg_hit ** hits_out, int64_t * n_hit)
{
  void * buf = nullptr;
  read_source src; src.seq = read_seq;
  int rc = search_host_impl (ctx, t, src, read_len, n_read, k, 0, &buf, n_hit, nullptr);
  if (hits_out) *hits_out = (gcg_hit *) buf;
  return rc;
}

extern "C" int gcg_search_compact (gcg_ctx * ctx, gcg_table * t, const char * const * read_seq, const int32_t * read_len,
                                   int64_t n_read, int k, uint64_t ** anchors_out, int64_t ** read_off_out, int64_t * n_anchor)
{
  GCG_CHECK (anchors_out && read_off_out, GCG_EINVAL, "gcg_search_compact: bad argument");
  void * buf = nullptr;
  read_source src; src.seq = read_seq;
  int rc = search_host_impl (ctx, t, src, read_len, n_read, k, 1, &buf, n_anchor, read_off_out);
  *anchors_out = (uint64_t *) buf;
  return rc;
}

extern "C" int gcg_search_compact_packed (gcg_ctx * ctx, gcg_table * t, const uint64_t * packed, const int64_t * woff, const int32_t * read_len,
                                          int64_t n_read, int k, uint64_t ** anchors_out, int64_t ** read_off_out, int64_t * n_anchor)
{
  GCG_CHECK (anchors_out && read_off_out && (n_read == 0 || (packed && woff)), GCG_EINVAL, "gcg_search_compact_packed: bad argument");
  void * buf = nullptr;
  read_source src; src.packed = packed; src.pwoff = woff;
  int rc = search_host_impl (ctx, t, src, read_len, n_read, k, 1, &buf, n_anchor, read_off_out);
  *anchors_out = (uint64_t *) buf;
  return rc;
}

// ---- contig chop to host kmer_t arrays --------------------------------------------------------
static std::mutex g_crc_mu;
static uint64_t g_crc_ready = 0;                  // bit d: device d holds the table (__constant__ memory is per device)
static int crc_table_upload (int device)
{
  std::lock_guard<std::mutex> lk (g_crc_mu);
  if (device < 64 && (g_crc_ready >> device & 1)) return GCG_OK;
  uint32_t tbl[256];
  for (uint32_t i = 0; i < 256; ++i) {          // reflected CRC-32, polynomial 0xEDB88320 (== crc32.h:15-68)
    uint32_t c = i;
    for (int j = 0; j < 8; ++j) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
    tbl[i] = c;
  }
  GCG_CUDA (cudaMemcpyToSymbol (c_crc_table, tbl, sizeof tbl));
  if (device < 64) g_crc_ready |= 1ULL << device;
  return GCG_OK;
}

extern "C" int gcg_chop_contigs (gcg_ctx * ctx, const gcg_seqs * contigs, int k, int n_thread,
                                 void * const * kmers_out, int32_t * n_kmer_out)
{
  GCG_CHECK (ctx && contigs && kmers_out && n_kmer_out, GCG_EINVAL, "gcg_chop_contigs: bad argument");
  GCG_CHECK (k >= 1 && k <= 31 && n_thread >= 1, GCG_ERANGE, "gcg_chop_contigs: k=%d n_thread=%d out of range", k, n_thread);
  GCG_CUDA (cudaSetDevice (ctx->device));
  int rc = crc_table_upload (ctx->device);
  if (rc) return rc;
  int64_t n = contigs->n;
  std::vector<int64_t> koff ((size_t) n + 1);
  int64_t tot = 0;
  for (int64_t i = 0; i < n; ++i) {
    koff[(size_t) i] = tot;
    int32_t l = contigs->h_len[(size_t) i];
    n_kmer_out[i] = l >= k ? l - k + 1 : 0;
    tot += n_kmer_out[i];
  }
  koff[(size_t) n] = tot;
  if (tot == 0) return GCG_OK;
  int64_t * d_koff; kmer_rec * d_rec;
  GCG_CUDA (gcg_dmalloc (ctx, &d_koff, (size_t) (n + 1) * 8));
  GCG_CUDA (gcg_dmalloc (ctx, &d_rec, (size_t) tot * sizeof (kmer_rec)));
  GCG_CUDA (cudaMemcpyAsync (d_koff, koff.data (), (size_t) (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  {
    gcg_kscope ks (ctx, "chop_records");
    chop_records_kernel<<<grid_for (ctx, contigs->n_words, 256, 8), 256, 0, ctx->stream>>> (
        contigs->d_packed, contigs->d_woff, contigs->d_len, d_koff, n, contigs->n_words, k, (uint32_t) n_thread, d_rec);
    GCG_CUDA (cudaGetLastError ());
  }
  // One download of all records into a pinned block (a blocking cudaMemcpy per contig into pageable memory costs tens
  // of microseconds each: seconds for an assembly of 10^5 - 10^6 contigs), then the host threads scatter them into
  // the callers' per-contig arrays.
  kmer_rec * h_rec = (kmer_rec *) gcg_pinned_alloc ((size_t) tot * sizeof (kmer_rec));
  if (h_rec == nullptr) { gcg_dfree (ctx, d_koff); gcg_dfree (ctx, d_rec); gcg_set_error ("gcg_chop_contigs: pinned alloc of %lld records failed", (long long) tot); return GCG_ENOMEM; }
  cudaError_t ce = cudaMemcpyAsync (h_rec, d_rec, (size_t) tot * sizeof (kmer_rec), cudaMemcpyDeviceToHost, ctx->stream);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize (ctx->stream);
  gcg_dfree (ctx, d_koff); gcg_dfree (ctx, d_rec);
  if (ce != cudaSuccess) { gcg_free (h_rec); gcg_set_error ("gcg_chop_contigs: %s", cudaGetErrorString (ce)); return GCG_ECUDA; }
  const int64_t n_task = std::min<int64_t> (n, std::max<int64_t> (1, 8 * (int64_t) ctx->host_threads));
  gcg_workers_run (gcg_ctx_workers (ctx), n_task, [&] (int64_t tk) {
    for (int64_t i = n * tk / n_task; i < n * (tk + 1) / n_task; ++i)
      if (n_kmer_out[i] > 0) memcpy (kmers_out[i], h_rec + koff[(size_t) i], (size_t) n_kmer_out[i] * sizeof (kmer_rec));
  });
  gcg_free (h_rec);
  return GCG_OK;
}
