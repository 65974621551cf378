// Hash-partitioned contig k-mer table: the device half of the multi-GPU exchange
// (BASELINE configs[3]: "k-mer table hash-partitioned over 8x B200, NVLink all-to-all").
//
// The reference partitions its tables the same way on the host: every k-mer goes to table
// crc32(kseq) % n_thread (kmer.c:88,124-152; ont.c:169,193) and every thread scans all k-mers for
// its own share.  Here a partition is one GPU.  Which hash picks the owner is unobservable
// (SURVEY F5), so the owner is a 64-bit mix that is independent of the in-table bucket hash.
//
//   route  : stable partition of the k-mers of a tile range by owner.  A warp owns a tile of 32
//            packed words (1024 k-mer start positions); per-lane per-owner counts live in shared
//            memory, a warp scan turns them into cursors, so the position of every k-mer inside
//            its owner's segment is a pure function of (tile, word, j): segment order is
//            (read, pos) order, the collect pass can recompute it, nothing is tagged or sorted.
//   insert : owner side of the build, 16-byte {key+1, tid<<32 | pos<<1 | rev} records.
//   lookup : owner side of the search, 8-byte keys in, 8-byte answers out (the value word of a
//            key present exactly once, ont.c:171,195; all ones otherwise); also counts the ONT-side
//            multiplicity (ont.c:245) — at the owner it sees the hits of every rank.
//   collect: answers (in the order the keys were sent) -> anchors in (read,pos) order.
#include <algorithm>
#include <stdio.h>
#include <string.h>

#include "gcg_internal.cuh"
#include "kmer_dev.cuh"

#define RT_WARPS 8
#define GCG_ANS_MISS 0xFFFFFFFFFFFFFFFFULL

struct gcg_route {
  gcg_ctx * ctx = nullptr;
  const gcg_seqs * seqs = nullptr;          // borrowed
  int k = 0, n_part = 0;
  int64_t tile0 = 0, n_tiles = 0;           // tiles [tile0, tile0 + n_tiles) of 32 words
  int64_t n_kmers = 0;                      // k-mer start positions in the range
  int64_t n_routed = 0;                     // of those, the ones that are routed (all, or the pre-filter's survivors)
  uint32_t * d_pass = nullptr;              // filtered plans: per word of the range, bit j = position j is routed
  uint32_t * d_off = nullptr;               // [n_part][n_tiles]: count, then exclusive offset inside the owner's segment
  uint32_t * d_bsum = nullptr;
  int64_t * d_seg = nullptr;                // [n_part + 1] segment starts (records)
  int64_t counts[GCG_MAX_PART] = {0};
};

// ---- owner of a canonical k-mer: kmer_owner () in kmer_dev.cuh -----------------------------------
extern "C" int gcg_kmer_owner (uint64_t canonical_kmer, int n_part)
{
  if (n_part < 1 || n_part > GCG_MAX_PART) return -1;
  return (int) kmer_owner (canonical_kmer, (uint32_t) n_part);
}

extern "C" int64_t gcg_seqs_tiles (const gcg_seqs * s) { return s ? (s->n_words + 31) >> 5 : 0; }

// ---- the shared walk over one tile ---------------------------------------------------------------
// number of valid k-mer starts in word w and, through *s_out / *p0_out, its sequence and first position
__device__ __forceinline__ int word_valid (const int64_t * __restrict__ woff, const int32_t * __restrict__ len,
                                           const int32_t * __restrict__ tile_seq, int64_t n_seq, int64_t n_words,
                                           int64_t tile, int64_t w, int k, int64_t * s_out, int32_t * p0_out)
{
  if (w >= n_words) return 0;
  int64_t s = find_seq_from (woff, n_seq, w, __ldg (tile_seq + tile));
  int32_t p0 = (int32_t) ((w - __ldg (woff + s)) << 5);
  int nvalid = __ldg (len + s) - k + 1 - p0;
  *s_out = s; *p0_out = p0;
  return nvalid < 0 ? 0 : (nvalid > 32 ? 32 : nvalid);
}

// per-lane counts of this word's k-mers by owner, into cnt[d][lane]
// Calls f (j, canonical k-mer, forward?) for the routed positions of word w, in position order.
// Without a pre-filter that is every valid position and the k-mers are rolled (kseq1.h:48-59); with
// one only a few percent of the positions survive, and extracting just those from the two packed
// words is cheaper than rolling through all 32.
template <class F>
__device__ __forceinline__ void for_each_routed (const uint64_t * __restrict__ packed, int64_t w, int nvalid, int k, uint32_t pass,
                                                 bool sparse, F f)
{
  if (!nvalid) return;
  const uint64_t hi = __ldg (packed + w), lo = __ldg (packed + w + 1);
  if (sparse) {
    uint32_t m = nvalid >= 32 ? pass : (pass & ((1u << nvalid) - 1u));
    while (m) {
      const int j = __ffs ((int) m) - 1;
      m &= m - 1u;
      bool fw;
      const unsigned long long key = key_at (hi, lo, j, k, &fw) - 1ULL;
      f (j, key, fw);
    }
  } else {
    kroll r;
    r.init (hi, lo, k);
    for (int j = 0; j < nvalid; ++j) {
      if (j) r.step ();
      const bool fw = r.fwd < r.rc;
      f (j, fw ? r.fwd : r.rc, fw);
    }
  }
}

// (`pass`: bit j set = position j takes part; all ones when the plan has no pre-filter)
__device__ __forceinline__ void lane_counts (const uint64_t * __restrict__ packed, int64_t w, int nvalid, int k, uint32_t n_part,
                                             uint32_t (* cnt)[32], int lane, uint32_t pass, bool sparse)
{
  for (uint32_t d = 0; d < n_part; ++d) cnt[d][lane] = 0;
  for_each_routed (packed, w, nvalid, k, pass, sparse, [&] (int, unsigned long long key, bool) { ++cnt[kmer_owner (key, n_part)][lane]; });
}

// Pre-filter of a routed search: the union, over all partitions, of the keys that can anchor (same
// blocked Bloom layout as the table's own pre-filter, kmer.cu).  A position whose k-mer fails it is
// a miss without asking anybody, so only the survivors (true anchors plus a few percent of false
// positives) are routed, cross NVLink and are looked up.  Four loads in flight per lane.
__device__ __forceinline__ uint32_t lane_filter_mask (const uint64_t * __restrict__ packed, int64_t w, int nvalid, int k,
                                                      const uint32_t * __restrict__ filter, uint32_t filter_words, int k3)
{
  uint32_t pass = 0;
  if (!nvalid) return 0;
  kroll r;
  r.init (__ldg (packed + w), __ldg (packed + w + 1), k);
  for (int j0 = 0; j0 < nvalid; j0 += 4) {
    uint32_t hh[4], fw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (j0 + u) r.step ();
      const unsigned long long key = r.fwd < r.rc ? r.fwd : r.rc;
      hh[u] = kmer_hash32 (key);
      fw[u] = __ldg (filter + __umulhi (kmer_hash32b (key), filter_words));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t m = filter_mask (hh[u], k3);
      pass |= (uint32_t) ((j0 + u < nvalid) && (fw[u] & m) == m) << (j0 + u);
    }
  }
  return pass;
}

// cnt[d][lane] <- off[d][tile] + exclusive warp scan of cnt[d][.]  (the lane's cursor inside segment d)
__device__ __forceinline__ void lane_cursors (uint32_t (* cnt)[32], const uint32_t * __restrict__ off, int64_t n_tiles, int64_t t,
                                              uint32_t n_part, int lane)
{
  for (uint32_t d = 0; d < n_part; ++d) {
    uint32_t c = cnt[d][lane], x = c;
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, x, o); if (lane >= o) x += y; }
    cnt[d][lane] = __ldg (off + (int64_t) d * n_tiles + t) + x - c;
  }
}

// ---- pre-filtered tiles: the survivors dealt over the lanes ------------------------------------------
// With the pre-filter a few percent of a tile's 1024 positions are routed, unevenly spread over the
// words: a loop "every lane walks the set bits of its own word" runs as long as the fullest word while
// most lanes idle (ncu: 8 of 32 lanes active).  Instead the warp lists the survivors of the tile in
// position order in shared memory and deals them out 32 per round.  The place of a k-mer inside its
// owner's segment is still "its rank among the tile's k-mers of that owner, in position order": lanes
// of a round that hold the same owner find each other with __match_any_sync, and a running count per
// owner carries over the rounds — the same layout the per-lane cursors produce, so the owner side and
// the unfiltered path are untouched.
struct sparse_tile {                     // one per warp, shared memory
  uint64_t pk[33];                       // the tile's packed words (+ the one after)
  unsigned short list[1024];             // (word in tile) << 5 | position in word, in position order
  uint32_t run[GCG_MAX_PART];            // k-mers of owner d handed out so far
  uint32_t m[32];                        // collect: anchor mask per word
};

// returns the number of survivors of the tile (the same in every lane)
__device__ __forceinline__ uint32_t sparse_build (sparse_tile & st, const uint64_t * __restrict__ packed, int64_t w, int64_t n_words,
                                                  int nvalid, uint32_t pass, uint32_t n_part, int lane)
{
  const uint32_t m = nvalid >= 32 ? pass : (nvalid > 0 ? (pass & ((1u << nvalid) - 1u)) : 0u);
  const uint32_t c = __popc (m);
  uint32_t x = c;
  for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, x, o); if (lane >= o) x += y; }
  const uint32_t total = __shfl_sync (0xffffffffu, x, 31);
  __syncwarp ();                                    // the previous tile's rounds are done with the arrays
  if (total == 0) return 0;
  st.pk[lane] = w <= n_words ? __ldg (packed + w) : 0ULL;           // (index n_words is the array's slack word)
  if (lane == 31) st.pk[32] = w + 1 <= n_words ? __ldg (packed + w + 1) : 0ULL;
  if ((uint32_t) lane < n_part) st.run[lane] = 0;
  st.m[lane] = 0;
  uint32_t at = x - c, mm = m;
  while (mm) { const int j = __ffs ((int) mm) - 1; mm &= mm - 1u; st.list[at++] = (unsigned short) ((lane << 5) | j); }
  __syncwarp ();
  return total;
}

// survivor h of the tile: its word, position, canonical key and strand
__device__ __forceinline__ void sparse_item (const sparse_tile & st, uint32_t h, int k, int * wl, int * j, unsigned long long * key, bool * fw)
{
  const uint32_t code = st.list[h];
  *wl = (int) (code >> 5); *j = (int) (code & 31u);
  *key = key_at (st.pk[*wl], st.pk[*wl + 1], *j, k, fw) - 1ULL;
}

// rank of this lane's k-mer inside owner d's segment part of the tile; every lane of `active` calls it
__device__ __forceinline__ uint32_t sparse_rank (sparse_tile & st, uint32_t active, uint32_t d, int lane)
{
  const uint32_t peers = __match_any_sync (active, d);
  const uint32_t r = st.run[d] + __popc (peers & ((1u << lane) - 1u));
  __syncwarp (active);
  if (lane == __ffs ((int) peers) - 1) st.run[d] += __popc (peers);
  __syncwarp (active);
  return r;
}

__global__ void __launch_bounds__ (32 * RT_WARPS)
route_count_kernel (const uint64_t * __restrict__ packed, const int64_t * __restrict__ woff, const int32_t * __restrict__ len,
                    const int32_t * __restrict__ tile_seq, int64_t n_seq, int64_t n_words, int k, uint32_t n_part,
                    int64_t tile0, int64_t n_tiles, uint32_t * __restrict__ cnt_out,
                    const uint32_t * __restrict__ filter, uint32_t filter_words, int filter_k3, uint32_t * __restrict__ pass_out)
{
  __shared__ uint32_t s_cnt[RT_WARPS][GCG_MAX_PART][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t wstride = (int64_t) gridDim.x * RT_WARPS;
  for (int64_t t = (int64_t) blockIdx.x * RT_WARPS + wid; t < n_tiles; t += wstride) {
    const int64_t tile = tile0 + t, w = (tile << 5) + lane;
    int64_t s; int32_t p0;
    const int nvalid = word_valid (woff, len, tile_seq, n_seq, n_words, tile, w, k, &s, &p0);
    uint32_t pass = 0xffffffffu;
    if (filter != nullptr) {
      pass = lane_filter_mask (packed, w, nvalid, k, filter, filter_words, filter_k3);
      if (w < n_words) pass_out[(t << 5) + lane] = pass;
    }
    // (the survivors are counted by the lane that owns their word: next to the filter pass of this
    // kernel the uneven loop costs nothing, dealing them out through shared memory cost 0.5 ms)
    lane_counts (packed, w, nvalid, k, n_part, s_cnt[wid], lane, pass, filter != nullptr);
    for (uint32_t d = 0; d < n_part; ++d) {
      uint32_t tot = __reduce_add_sync (0xffffffffu, s_cnt[wid][d][lane]);
      if (lane == 0) cnt_out[(int64_t) d * n_tiles + t] = tot;
    }
  }
}

// ---- staging through shared memory ------------------------------------------------------------
// Inside an owner's segment the k-mers of one tile form ONE contiguous run (lane 0's, then lane
// 1's ...), so a tile's keys / answers are moved between HBM and shared memory as n_part coalesced
// runs and the per-lane scatter / gather (8 bytes at 32 different places per warp instruction if
// done on global memory) happens in shared memory.
//   s_cnt[d][lane]  per-lane count, then the lane's cursor INSIDE THE TILE'S STAGE
//   s_run[d]        start of owner d's run in the stage (exclusive scan of the tile totals), s_run[n_part] = total
//   s_glob[d]       address of the first element of owner d's run (the owner's segment may be a
//                   peer GPU's exchange window: the copy-out is then a coalesced store over NVLink)
__device__ __forceinline__ void tile_layout (uint32_t (* cnt)[32], uint32_t * s_run, unsigned long long ** s_glob,
                                             unsigned long long * const * s_base,
                                             const uint32_t * __restrict__ off, int64_t n_tiles, int64_t t, uint32_t n_part, int lane)
{
  // tile totals per owner: lane d keeps tot_d
  uint32_t mine = 0;
  for (uint32_t d = 0; d < n_part; ++d) {
    const uint32_t tot = __reduce_add_sync (0xffffffffu, cnt[d][lane]);
    if ((uint32_t) lane == d) mine = tot;
  }
  uint32_t x = mine;
  for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, x, o); if (lane >= o) x += y; }
  __syncwarp ();
  if ((uint32_t) lane < n_part) {
    s_run[lane] = x - mine;
    s_glob[lane] = s_base[lane] + __ldg (off + (int64_t) lane * n_tiles + t);
  }
  if ((uint32_t) lane == n_part - 1) s_run[n_part] = x;
  __syncwarp ();
  // per-lane cursors inside the stage
  for (uint32_t d = 0; d < n_part; ++d) {
    uint32_t c = cnt[d][lane], e = c;
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, e, o); if (lane >= o) e += y; }
    cnt[d][lane] = s_run[d] + e - c;
  }
}

// owner whose run holds stage position i (n_part <= 16: a short scan of s_run)
__device__ __forceinline__ uint32_t run_of (const uint32_t * s_run, uint32_t n_part, uint32_t i)
{
  uint32_t d = 0;
  while (d + 1 < n_part && s_run[d + 1] <= i) ++d;
  return d;
}

#define RS_WARPS 4      // warps per block of the staged kernels (8 KB of stage per warp)

// where owner d's segment starts: inside one send buffer (exchange by all-to-all) or inside the
// owner's own exchange window (direct stores over NVLink / NVSwitch peer memory)
struct route_dst { unsigned long long * base[GCG_MAX_PART]; };

// 8-byte keys (key + 1) of the range, grouped by owner
__global__ void __launch_bounds__ (32 * RS_WARPS)
route_keys_kernel (const uint64_t * __restrict__ packed, const int64_t * __restrict__ woff, const int32_t * __restrict__ len,
                   const int32_t * __restrict__ tile_seq, int64_t n_seq, int64_t n_words, int k, uint32_t n_part,
                   int64_t tile0, int64_t n_tiles, const uint32_t * __restrict__ off, const __grid_constant__ route_dst dst,
                   const uint32_t * __restrict__ pass_in)
{
  __shared__ unsigned long long s_stage[RS_WARPS][1024];
  __shared__ uint32_t s_cnt[RS_WARPS][GCG_MAX_PART][32];
  __shared__ uint32_t s_run[RS_WARPS][GCG_MAX_PART + 1];
  __shared__ unsigned long long * s_glob[RS_WARPS][GCG_MAX_PART];
  __shared__ unsigned long long * s_base[GCG_MAX_PART];
  static_assert (sizeof (sparse_tile) <= sizeof (unsigned long long) * 1024, "the sparse tile lives in the warp's stage");
  if (threadIdx.x < n_part) s_base[threadIdx.x] = dst.base[threadIdx.x];
  __syncthreads ();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t wstride = (int64_t) gridDim.x * RS_WARPS;
  for (int64_t t = (int64_t) blockIdx.x * RS_WARPS + wid; t < n_tiles; t += wstride) {
    const int64_t tile = tile0 + t, w = (tile << 5) + lane;
    int64_t s; int32_t p0;
    const int nvalid = word_valid (woff, len, tile_seq, n_seq, n_words, tile, w, k, &s, &p0);
    __syncwarp ();                                    // the previous tile's copy-out has finished reading the stage
    const uint32_t pass = (pass_in != nullptr && w < n_words) ? __ldg (pass_in + (t << 5) + lane) : 0xffffffffu;
    if (pass_in != nullptr) {
      // pre-filtered: the few survivors go straight to their owners' segments, 32 per round; the lanes of
      // a round that share an owner write one contiguous run (the stage is not needed: it holds the list)
      sparse_tile & st = * reinterpret_cast<sparse_tile *> (s_stage[wid]);
      const uint32_t total = sparse_build (st, packed, w, n_words, nvalid, pass, n_part, lane);
      for (uint32_t h0 = 0; h0 < total; h0 += 32) {
        const uint32_t h = h0 + lane, active = __ballot_sync (0xffffffffu, h < total);
        if (h < total) {
          int wl, j; unsigned long long key; bool fw;
          sparse_item (st, h, k, &wl, &j, &key, &fw);
          const uint32_t d = kmer_owner (key, n_part), r = sparse_rank (st, active, d, lane);
          s_base[d][__ldg (off + (int64_t) d * n_tiles + t) + r] = key + 1ULL;
        }
      }
      continue;
    }
    lane_counts (packed, w, nvalid, k, n_part, s_cnt[wid], lane, pass, false);
    tile_layout (s_cnt[wid], s_run[wid], s_glob[wid], s_base, off, n_tiles, t, n_part, lane);
    for_each_routed (packed, w, nvalid, k, pass, false, [&] (int, unsigned long long key, bool) {
      s_stage[wid][s_cnt[wid][kmer_owner (key, n_part)][lane]++] = key + 1ULL;
    });
    __syncwarp ();
    const uint32_t total = s_run[wid][n_part];
    for (uint32_t i = lane; i < total; i += 32) {
      const uint32_t d = run_of (s_run[wid], n_part, i);
      s_glob[wid][d][i - s_run[wid][d]] = s_stage[wid][i];
    }
  }
}

// 16-byte records {key + 1, tid << 32 | pos << 1 | rev} of the range, grouped by owner (build side:
// the contigs are small next to the reads, written straight from the lanes' cursors)
__global__ void __launch_bounds__ (32 * RT_WARPS)
route_records_kernel (const uint64_t * __restrict__ packed, const int64_t * __restrict__ woff, const int32_t * __restrict__ len,
                      const int32_t * __restrict__ tile_seq, int64_t n_seq, int64_t n_words, int k, uint32_t n_part,
                      int64_t tile0, int64_t n_tiles, const uint32_t * __restrict__ off, const int64_t * __restrict__ seg,
                      ulonglong2 * __restrict__ out)
{
  __shared__ uint32_t s_cnt[RT_WARPS][GCG_MAX_PART][32];
  __shared__ int64_t s_seg[GCG_MAX_PART];
  if (threadIdx.x < n_part) s_seg[threadIdx.x] = seg[threadIdx.x];
  __syncthreads ();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t wstride = (int64_t) gridDim.x * RT_WARPS;
  for (int64_t t = (int64_t) blockIdx.x * RT_WARPS + wid; t < n_tiles; t += wstride) {
    const int64_t tile = tile0 + t, w = (tile << 5) + lane;
    int64_t s; int32_t p0;
    const int nvalid = word_valid (woff, len, tile_seq, n_seq, n_words, tile, w, k, &s, &p0);
    lane_counts (packed, w, nvalid, k, n_part, s_cnt[wid], lane, 0xffffffffu, false);
    lane_cursors (s_cnt[wid], off, n_tiles, t, n_part, lane);
    if (nvalid) {
      kroll r;
      r.init (__ldg (packed + w), __ldg (packed + w + 1), k);
      for (int j = 0; j < nvalid; ++j) {
        if (j) r.step ();
        const bool fw = r.fwd < r.rc;
        const unsigned long long key = fw ? r.fwd : r.rc;
        const uint32_t d = kmer_owner (key, n_part);
        ulonglong2 rec;
        rec.x = key + 1ULL;
        rec.y = ((unsigned long long) s << 32) | ((unsigned long long) (uint32_t) (p0 + j) << 1) | (fw ? 0ULL : 1ULL);
        out[s_seg[d] + s_cnt[wid][d][lane]++] = rec;
      }
    }
  }
}

// EMIT 0: mask[w] = bit j set <=> the answer of k-mer j of word w is a hit
// EMIT 1: anchors written at hits[prefix[w] ...] in position order
// Answers are read straight from global memory by the lane that owns them: a lane's answers for one
// owner are consecutive (four per 32-byte sector), and the pass is latency bound, so the full
// occupancy of the small-footprint kernel beats staging the tile through shared memory
// (measured 0.65 ms against 1.2 ms per pass at cfg2).
template <int EMIT>
__global__ void __launch_bounds__ (32 * RT_WARPS)
route_collect_kernel (const uint64_t * __restrict__ packed, const int64_t * __restrict__ woff, const int32_t * __restrict__ len,
                      const int32_t * __restrict__ tile_seq, int64_t n_seq, int64_t n_words, int k, uint32_t n_part,
                      int64_t tile0, int64_t n_tiles, const uint32_t * __restrict__ off, const int64_t * __restrict__ seg,
                      const unsigned long long * __restrict__ ans, uint32_t * __restrict__ mask,
                      const uint32_t * __restrict__ prefix, gcg_hit * __restrict__ hits, const uint32_t * __restrict__ pass_in)
{
  __shared__ uint32_t s_cnt[RT_WARPS][GCG_MAX_PART][32];
  __shared__ int64_t s_seg[GCG_MAX_PART];
  __shared__ sparse_tile s_sp[RT_WARPS];
  if (threadIdx.x < n_part) s_seg[threadIdx.x] = seg[threadIdx.x];
  __syncthreads ();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t wstride = (int64_t) gridDim.x * RT_WARPS;
  for (int64_t t = (int64_t) blockIdx.x * RT_WARPS + wid; t < n_tiles; t += wstride) {
    const int64_t tile = tile0 + t, w = (tile << 5) + lane, wl = (t << 5) + lane;    // wl: word index inside the range
    int64_t s; int32_t p0;
    const int nvalid = word_valid (woff, len, tile_seq, n_seq, n_words, tile, w, k, &s, &p0);
    if (EMIT) {
      // nothing to do for a tile without anchors
      const uint32_t m = w < n_words ? __ldg (mask + wl) : 0u;
      if (!__any_sync (0xffffffffu, m != 0u)) continue;
    }
    const uint32_t pass = (pass_in != nullptr && w < n_words) ? __ldg (pass_in + wl) : 0xffffffffu;
    if (pass_in != nullptr) {
      // pre-filtered: answers of the survivors, 32 per round (same ranks as route_keys_kernel)
      sparse_tile & st = s_sp[wid];
      const uint32_t total = sparse_build (st, packed, w, n_words, nvalid, pass, n_part, lane);
      uint32_t n_before = 0;                          // anchors of the tile found in earlier rounds
      const uint32_t tile_at = EMIT && total ? __ldg (prefix + (t << 5)) : 0u;
      const int64_t s0 = __shfl_sync (0xffffffffu, s, 0);      // (lanes past the last word hold no sequence)
      for (uint32_t h0 = 0; h0 < total; h0 += 32) {
        const uint32_t h = h0 + lane, active = __ballot_sync (0xffffffffu, h < total);
        bool hit = false;
        int wl2 = 0, j = 0; unsigned long long key = 0, v = 0; bool fw = false;
        if (h < total) {
          sparse_item (st, h, k, &wl2, &j, &key, &fw);
          const uint32_t d = kmer_owner (key, n_part), r = sparse_rank (st, active, d, lane);
          v = __ldg (ans + s_seg[d] + __ldg (off + (int64_t) d * n_tiles + t) + r);
          hit = v != GCG_ANS_MISS;
          if (!EMIT && hit) atomicOr (&st.m[wl2], 1u << j);
        }
        if (EMIT) {
          const uint32_t hb = __ballot_sync (0xffffffffu, hit);
          if (hit) {
            // sequence and first position of word wl2 of the tile: recomputed from the word index
            const int64_t ww = (tile << 5) + wl2;
            const int64_t sq = find_seq_from (woff, n_seq, ww, s0 >= 0 ? s0 : 0);
            int4 hh;                                  // gcg_hit {read, pos, tid, cpos_flags}
            hh.x = (int32_t) sq;
            hh.y = (int32_t) ((ww - __ldg (woff + sq)) << 5) + j;
            hh.z = (int32_t) ((v >> 32) & 0x7FFFFFFFu);
            hh.w = (int32_t) ((((uint32_t) (v >> 1) & 0x3FFFFFFFu) << 2) | (uint32_t) (v & 1ULL) | (fw ? 0u : 2u));
            reinterpret_cast<int4 *> (hits)[tile_at + n_before + __popc (hb & ((1u << lane) - 1u))] = hh;
          }
          n_before += __popc (hb);
        }
      }
      if (!EMIT) { __syncwarp (); if (w < n_words) mask[wl] = total ? st.m[lane] : 0u; }
      continue;
    }
    lane_counts (packed, w, nvalid, k, n_part, s_cnt[wid], lane, pass, false);
    lane_cursors (s_cnt[wid], off, n_tiles, t, n_part, lane);
    uint32_t m = 0;
    uint32_t at_hit = (EMIT && nvalid) ? __ldg (prefix + wl) : 0u;
    for_each_routed (packed, w, nvalid, k, pass, false, [&] (int j, unsigned long long key, bool fw) {
      const uint32_t d = kmer_owner (key, n_part);
      const unsigned long long v = __ldg (ans + s_seg[d] + s_cnt[wid][d][lane]++);
      if (v == GCG_ANS_MISS) return;
      m |= 1u << j;
      if (EMIT) {
        int4 hh;                                    // gcg_hit {read, pos, tid, cpos_flags}
        hh.x = (int32_t) s;
        hh.y = p0 + j;
        hh.z = (int32_t) ((v >> 32) & 0x7FFFFFFFu);
        hh.w = (int32_t) ((((uint32_t) (v >> 1) & 0x3FFFFFFFu) << 2) | (uint32_t) (v & 1ULL) | (fw ? 0u : 2u));
        reinterpret_cast<int4 *> (hits)[at_hit++] = hh;
      }
    });
    if (!EMIT && w < n_words) mask[wl] = m;
  }
}

// ---- owner side -------------------------------------------------------------------------------------
__global__ void __launch_bounds__ (256)
insert_records_kernel (const ulonglong2 * __restrict__ recs, int64_t n, unsigned long long * __restrict__ keys,
                       unsigned long long * __restrict__ vals, uint32_t n_bucket)
{
  const int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    ulonglong2 r = __ldg (recs + i);
    table_insert (keys, vals, n_bucket, r.x, r.y);
  }
}

// keys [first[r], first[r+1]) came from requester r; their answers go to base[r][i - first[r]] — the
// requester's own answer buffer (or, exchanged directly, its window on the peer GPU)
struct lookup_dst { unsigned long long * base[GCG_MAX_PART]; long long first[GCG_MAX_PART + 1]; int n_src; };

#define LK_UNROLL 4
__global__ void __launch_bounds__ (256)
lookup_keys_kernel (const unsigned long long * __restrict__ qkeys, int64_t n, const unsigned long long * __restrict__ keys,
                    unsigned long long * __restrict__ vals, uint32_t n_bucket,
                    const __grid_constant__ lookup_dst dst)
{
  __shared__ unsigned long long * s_base[GCG_MAX_PART];
  __shared__ long long s_first[GCG_MAX_PART + 1];
  if ((int) threadIdx.x < dst.n_src) s_base[threadIdx.x] = dst.base[threadIdx.x];
  if ((int) threadIdx.x <= dst.n_src) s_first[threadIdx.x] = dst.first[threadIdx.x];
  __syncthreads ();
  const int n_src = dst.n_src;
  const int64_t stride = (int64_t) gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * LK_UNROLL) {
    unsigned long long key[LK_UNROLL];
    uint32_t hs[LK_UNROLL];
    bucket4 q[LK_UNROLL];
#pragma unroll
    for (int u = 0; u < LK_UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      key[u] = i < n ? __ldg (qkeys + i) : 1ULL;
      hs[u] = kmer_hash32 (key[u] - 1ULL);
      q[u] = ld_bucket (keys + 4ULL * __umulhi (hs[u], n_bucket));
    }
#pragma unroll
    for (int u = 0; u < LK_UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= n) continue;
      unsigned long long kw, a = GCG_ANS_MISS;
      const unsigned long long slot = table_lookup (keys, n_bucket, __umulhi (hs[u], n_bucket), q[u], key[u], hs[u] & 3u, &kw);
      if (slot != ~0ULL && !(kw & GCG_KEY_MULTI)) {            // multi == 1  (ont.c:171,195)
        // one atomic returns the value word and records the anchor (ONT-side multiplicity, ont.c:245,
        // saturating at 2, in the word's two spare bits)
        const unsigned long long old = atomicOr (vals + slot, GCG_VAL_ONT1);
        if ((old & (GCG_VAL_ONT1 | GCG_VAL_ONT2)) == GCG_VAL_ONT1) atomicOr (vals + slot, GCG_VAL_ONT2);
        a = old & ~(GCG_VAL_ONT1 | GCG_VAL_ONT2);
      }
      int r = 0;
      while (r + 1 < n_src && s_first[r + 1] <= i) ++r;
      s_base[r][i - s_first[r]] = a;
    }
  }
}

// ---- batched exclusive scan of u32 rows (n_part rows of n values) ---------------------------------
#define PS_ITEMS 8
#define PS_BLOCK 256
#define PS_TILE (PS_ITEMS * PS_BLOCK)

__global__ void __launch_bounds__ (PS_BLOCK)
rows_reduce_kernel (const uint32_t * __restrict__ v, int64_t n, int64_t nb, uint32_t * __restrict__ bsum)
{
  __shared__ uint32_t s[PS_BLOCK / 32];
  const uint32_t * row = v + (int64_t) blockIdx.y * n;
  int64_t base = (int64_t) blockIdx.x * PS_TILE;
  uint32_t a = 0;
  for (int i = 0; i < PS_ITEMS; ++i) {
    int64_t idx = base + (int64_t) i * PS_BLOCK + threadIdx.x;
    if (idx < n) a += row[idx];
  }
  a = __reduce_add_sync (0xffffffffu, a);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
  __syncthreads ();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int i = 0; i < PS_BLOCK / 32; ++i) t += s[i];
    bsum[(int64_t) blockIdx.y * nb + blockIdx.x] = t;
  }
}

// one block per row: exclusive scan of the row's block sums in place, row total to totals[row]
__global__ void __launch_bounds__ (1024)
rows_blocksums_kernel (uint32_t * __restrict__ bsum, int64_t nb, unsigned long long * __restrict__ totals)
{
  __shared__ uint32_t s_warp[32];
  __shared__ unsigned long long s_carry;
  uint32_t * row = bsum + (int64_t) blockIdx.x * nb;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads ();
  for (int64_t base = 0; base < nb; base += 1024) {
    int64_t idx = base + threadIdx.x;
    uint32_t v = idx < nb ? row[idx] : 0, x = v;
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads ();
    if (threadIdx.x < 32) {
      uint32_t wv = s_warp[threadIdx.x], wx = wv;
      for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, wx, o); if (threadIdx.x >= o) wx += y; }
      s_warp[threadIdx.x] = wx - wv;
    }
    __syncthreads ();
    unsigned long long excl = (unsigned long long) (x - v + s_warp[threadIdx.x >> 5]) + s_carry;
    if (idx < nb) row[idx] = (uint32_t) excl;
    __syncthreads ();
    if (threadIdx.x == 1023) s_carry = excl + v;
    __syncthreads ();
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = s_carry;
}

__global__ void __launch_bounds__ (PS_BLOCK)
rows_apply_kernel (uint32_t * __restrict__ v, int64_t n, int64_t nb, const uint32_t * __restrict__ bsum)
{
  __shared__ uint32_t s_warp[PS_BLOCK / 32];
  uint32_t * row = v + (int64_t) blockIdx.y * n;
  int64_t base = (int64_t) blockIdx.x * PS_TILE + (int64_t) threadIdx.x * PS_ITEMS;
  uint32_t c[PS_ITEMS], t = 0;
#pragma unroll
  for (int i = 0; i < PS_ITEMS; ++i) { int64_t idx = base + i; c[i] = idx < n ? row[idx] : 0; t += c[i]; }
  uint32_t x = t;
  for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
  __syncthreads ();
  if (threadIdx.x < 32) {
    uint32_t wv = threadIdx.x < PS_BLOCK / 32 ? s_warp[threadIdx.x] : 0, wx = wv;
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, wx, o); if (threadIdx.x >= o) wx += y; }
    if (threadIdx.x < PS_BLOCK / 32) s_warp[threadIdx.x] = wx - wv;
  }
  __syncthreads ();
  uint32_t run = x - t + s_warp[threadIdx.x >> 5] + bsum[(int64_t) blockIdx.y * nb + blockIdx.x];
#pragma unroll
  for (int i = 0; i < PS_ITEMS; ++i) { int64_t idx = base + i; if (idx < n) row[idx] = run; run += c[i]; }
}

// =============================================================================================
// host side
// =============================================================================================
static int warp_grid (gcg_ctx * ctx, int64_t n_tiles)
{
  int64_t nb = (n_tiles + RT_WARPS - 1) / RT_WARPS, cap = (int64_t) ctx->sm_count * 8;
  return (int) std::max<int64_t> (1, std::min (nb, cap));
}

static int flat_grid (gcg_ctx * ctx, int64_t n, int per_thread)
{
  int64_t nb = (n + 256LL * per_thread - 1) / (256LL * per_thread), cap = (int64_t) ctx->sm_count * 8;
  return (int) std::max<int64_t> (1, std::min (nb, cap));
}

extern "C" void gcg_route_free (gcg_route * r)
{
  if (!r) return;
  gcg_dfree (r->ctx, r->d_off); gcg_dfree (r->ctx, r->d_bsum); gcg_dfree (r->ctx, r->d_seg); gcg_dfree (r->ctx, r->d_pass);
  delete r;
}

static int route_plan_impl (gcg_ctx * ctx, const gcg_seqs * s, int k, int n_part, int64_t tile_begin, int64_t tile_end,
                            const uint32_t * d_filter, uint32_t filter_words, int filter_k3, gcg_route ** out, int64_t * counts)
{
  GCG_CHECK (ctx && s && out && counts, GCG_EINVAL, "gcg_route_plan: bad argument");
  GCG_CHECK (k >= 1 && k <= 31, GCG_ERANGE, "gcg_route_plan: k=%d outside [1,31]", k);
  GCG_CHECK (n_part >= 1 && n_part <= GCG_MAX_PART, GCG_ERANGE, "gcg_route_plan: %d partitions outside [1,%d]", n_part, GCG_MAX_PART);
  const int64_t all_tiles = (s->n_words + 31) >> 5;
  GCG_CHECK (tile_begin >= 0 && tile_begin <= tile_end && tile_end <= all_tiles, GCG_EINVAL,
             "gcg_route_plan: tile range [%lld,%lld) outside [0,%lld]", (long long) tile_begin, (long long) tile_end, (long long) all_tiles);
  GCG_CHECK (s->n < 0x7FFFFFFF, GCG_ERANGE, "gcg_route_plan: too many sequences");
  GCG_CUDA (cudaSetDevice (ctx->device));
  gcg_route * r = new gcg_route ();
  r->ctx = ctx; r->seqs = s; r->k = k; r->n_part = n_part; r->tile0 = tile_begin; r->n_tiles = tile_end - tile_begin;
  // positions in the range (host side: sequences are word aligned, so whole words belong to one sequence)
  {
    const int64_t w_lo = tile_begin << 5, w_hi = std::min (s->n_words, tile_end << 5);
    int64_t tot = 0;
    size_t i = (size_t) (std::upper_bound (s->h_woff.begin (), s->h_woff.end (), w_lo) - s->h_woff.begin ());
    i = i ? i - 1 : 0;
    for (; i < (size_t) s->n && s->h_woff[i] < w_hi; ++i) {
      const int64_t nk = (int64_t) s->h_len[i] - k + 1;                       // valid starts are positions [0, nk)
      const int64_t a = std::max<int64_t> ((w_lo - s->h_woff[i]) << 5, 0), b = std::min<int64_t> ((w_hi - s->h_woff[i]) << 5, nk);
      if (b > a) tot += b - a;
    }
    r->n_kmers = tot;
  }
  if (r->n_kmers >= 0xFFFFFFFFLL) {
    gcg_set_error ("gcg_route_plan: %lld positions in one range exceed the 32-bit segment offsets; split the range", (long long) r->n_kmers);
    delete r;
    return GCG_ERANGE;
  }
  for (int d = 0; d < n_part; ++d) counts[d] = 0;
  const int64_t nt = r->n_tiles;
  int rc = GCG_OK;
  cudaError_t e;
  const int64_t nb = (nt + PS_TILE - 1) / PS_TILE;
  if ((d_filter != nullptr && (e = gcg_dmalloc (ctx, &r->d_pass, (size_t) std::max<int64_t> (nt, 1) * 32 * 4)) != cudaSuccess) ||
      (e = gcg_dmalloc (ctx, &r->d_off, (size_t) std::max<int64_t> (nt, 1) * n_part * 4)) != cudaSuccess ||
      (e = gcg_dmalloc (ctx, &r->d_bsum, (size_t) std::max<int64_t> (nb, 1) * n_part * 4)) != cudaSuccess ||
      (e = gcg_dmalloc (ctx, &r->d_seg, (GCG_MAX_PART + 1) * 8)) != cudaSuccess) {
    gcg_set_error ("gcg_route_plan: cudaMalloc failed: %s", cudaGetErrorString (e));
    gcg_route_free (r);
    return GCG_ENOMEM;
  }
  int64_t seg[GCG_MAX_PART + 1] = {0};
  if (nt > 0 && r->n_kmers > 0) {
    { gcg_kscope ks (ctx, "route_count");
      route_count_kernel<<<warp_grid (ctx, nt), 32 * RT_WARPS, 0, ctx->stream>>> (
          s->d_packed, s->d_woff, s->d_len, s->d_tseq, s->n, s->n_words, k, (uint32_t) n_part, r->tile0, nt, r->d_off,
          d_filter, filter_words, filter_k3, r->d_pass); }
    { gcg_kscope ks (ctx, "route_scan");
      rows_reduce_kernel<<<dim3 ((unsigned) nb, (unsigned) n_part), PS_BLOCK, 0, ctx->stream>>> (r->d_off, nt, nb, r->d_bsum); }
    { gcg_kscope ks (ctx, "route_scan");
      rows_blocksums_kernel<<<n_part, 1024, 0, ctx->stream>>> (r->d_bsum, nb, ctx->d_counters + 16); }
    { gcg_kscope ks (ctx, "route_scan");
      rows_apply_kernel<<<dim3 ((unsigned) nb, (unsigned) n_part), PS_BLOCK, 0, ctx->stream>>> (r->d_off, nt, nb, r->d_bsum); }
    if (cudaGetLastError () != cudaSuccess) { gcg_set_error ("gcg_route_plan: kernel launch failed"); rc = GCG_ECUDA; }
    if (!rc && (cudaMemcpyAsync (ctx->h_counters + 16, ctx->d_counters + 16, (size_t) n_part * 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                cudaStreamSynchronize (ctx->stream) != cudaSuccess)) {
      gcg_set_error ("gcg_route_plan: %s", cudaGetErrorString (cudaGetLastError ())); rc = GCG_ECUDA; }
    if (!rc) {
      int64_t tot = 0;
      for (int d = 0; d < n_part; ++d) { counts[d] = r->counts[d] = (int64_t) ctx->h_counters[16 + d]; tot += counts[d]; }
      r->n_routed = tot;
      if (d_filter ? tot > r->n_kmers : tot != r->n_kmers) {
        gcg_set_error ("gcg_route_plan: counted %lld k-mers, expected %lld", (long long) tot, (long long) r->n_kmers); rc = GCG_ECUDA; }
    }
  }
  if (!rc) {
    for (int d = 0; d < n_part; ++d) seg[d + 1] = seg[d] + r->counts[d];
    // (pageable source: staged by the runtime before the call returns)
    if (cudaMemcpyAsync (r->d_seg, seg, sizeof seg, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) { gcg_set_error ("gcg_route_plan: copy failed"); rc = GCG_ECUDA; }
  }
  if (rc) { gcg_route_free (r); return rc; }
  *out = r;
  return GCG_OK;
}

extern "C" int gcg_route_plan (gcg_ctx * ctx, const gcg_seqs * s, int k, int n_part, int64_t tile_begin, int64_t tile_end,
                               gcg_route ** out, int64_t * counts)
{
  return route_plan_impl (ctx, s, k, n_part, tile_begin, tile_end, nullptr, 0, 0, out, counts);
}

extern "C" int gcg_route_plan_filtered (gcg_ctx * ctx, const gcg_seqs * s, int k, int n_part, int64_t tile_begin, int64_t tile_end,
                                        const void * d_filter, int64_t filter_words, int filter_k3, gcg_route ** out, int64_t * counts)
{
  GCG_CHECK (d_filter != nullptr && filter_words > 0 && filter_words < 0xFFFFFFFFLL, GCG_EINVAL, "gcg_route_plan_filtered: bad filter");
  return route_plan_impl (ctx, s, k, n_part, tile_begin, tile_end, (const uint32_t *) d_filter, (uint32_t) filter_words, filter_k3 != 0, out, counts);
}

// number of k-mers the plan routes (= buffer elements): all positions of the range, or the pre-filter's survivors
extern "C" int64_t gcg_route_kmers (const gcg_route * r) { return r ? r->n_routed : 0; }
extern "C" int64_t gcg_route_positions (const gcg_route * r) { return r ? r->n_kmers : 0; }

static int staged_grid (gcg_ctx * ctx, int64_t n_tiles)
{
  int64_t nb = (n_tiles + RS_WARPS - 1) / RS_WARPS, cap = (int64_t) ctx->sm_count * 5;    // ~41 KB of shared memory per block
  return (int) std::max<int64_t> (1, std::min (nb, cap));
}

static int route_keys_launch (gcg_ctx * ctx, gcg_route * r, const route_dst & dst)
{
  const gcg_seqs * s = r->seqs;
  gcg_kscope ks (ctx, "route_keys");
  route_keys_kernel<<<staged_grid (ctx, r->n_tiles), 32 * RS_WARPS, 0, ctx->stream>>> (
      s->d_packed, s->d_woff, s->d_len, s->d_tseq, s->n, s->n_words, r->k, (uint32_t) r->n_part, r->tile0, r->n_tiles, r->d_off, dst, r->d_pass);
  GCG_CUDA (cudaGetLastError ());
  return GCG_OK;
}

extern "C" int gcg_route_keys (gcg_ctx * ctx, gcg_route * r, void * d_send)
{
  GCG_CHECK (ctx && r && (d_send || r->n_routed == 0), GCG_EINVAL, "gcg_route_keys: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  if (r->n_routed == 0) return GCG_OK;
  route_dst dst;
  int64_t at = 0;
  for (int d = 0; d < GCG_MAX_PART; ++d) {
    dst.base[d] = (unsigned long long *) d_send + at;
    if (d < r->n_part) at += r->counts[d];
  }
  return route_keys_launch (ctx, r, dst);
}

extern "C" int gcg_route_keys_direct (gcg_ctx * ctx, gcg_route * r, void * const * d_owner_base, const int64_t * owner_off)
{
  GCG_CHECK (ctx && r && d_owner_base && owner_off, GCG_EINVAL, "gcg_route_keys_direct: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  if (r->n_routed == 0) return GCG_OK;
  route_dst dst;
  for (int d = 0; d < GCG_MAX_PART; ++d) {
    dst.base[d] = nullptr;
    if (d >= r->n_part) continue;
    GCG_CHECK (d_owner_base[d] != nullptr || r->counts[d] == 0, GCG_EINVAL, "gcg_route_keys_direct: no window for owner %d", d);
    GCG_CHECK (owner_off[d] >= 0, GCG_EINVAL, "gcg_route_keys_direct: negative offset for owner %d", d);
    dst.base[d] = (unsigned long long *) d_owner_base[d] + owner_off[d];
  }
  return route_keys_launch (ctx, r, dst);
}

extern "C" int gcg_route_records (gcg_ctx * ctx, gcg_route * r, void * d_send)
{
  GCG_CHECK (ctx && r && (d_send || r->n_kmers == 0), GCG_EINVAL, "gcg_route_records: bad argument");
  GCG_CHECK (r->d_pass == nullptr, GCG_EINVAL, "gcg_route_records: the plan was made with a pre-filter (search side only)");
  GCG_CUDA (cudaSetDevice (ctx->device));
  if (r->n_kmers == 0) return GCG_OK;
  const gcg_seqs * s = r->seqs;
  gcg_kscope ks (ctx, "route_records");
  route_records_kernel<<<warp_grid (ctx, r->n_tiles), 32 * RT_WARPS, 0, ctx->stream>>> (
      s->d_packed, s->d_woff, s->d_len, s->d_tseq, s->n, s->n_words, r->k, (uint32_t) r->n_part, r->tile0, r->n_tiles,
      r->d_off, r->d_seg, (ulonglong2 *) d_send);
  GCG_CUDA (cudaGetLastError ());
  return GCG_OK;
}

extern "C" int gcg_route_collect (gcg_ctx * ctx, gcg_route * r, const void * d_answers, gcg_hits ** out)
{
  GCG_CHECK (ctx && r && out && (d_answers || r->n_routed == 0), GCG_EINVAL, "gcg_route_collect: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  gcg_hits * h = new gcg_hits ();
  h->ctx = ctx;
  if (r->n_routed == 0) { *out = h; return GCG_OK; }
  const gcg_seqs * s = r->seqs;
  const int64_t n_words = std::min (s->n_words, (r->tile0 + r->n_tiles) << 5) - (r->tile0 << 5);
  uint32_t * d_mask = nullptr, * d_prefix = nullptr, * d_bsum = nullptr;
  int rc = GCG_OK;
  cudaError_t e;
  if ((e = gcg_dmalloc (ctx, &d_mask, (size_t) n_words * 4)) != cudaSuccess || (e = gcg_dmalloc (ctx, &d_prefix, (size_t) n_words * 4)) != cudaSuccess ||
      (e = gcg_dmalloc (ctx, &d_bsum, (size_t) gcg_mask_scan_blocks (n_words) * 4)) != cudaSuccess) {
    gcg_set_error ("gcg_route_collect: cudaMalloc failed: %s", cudaGetErrorString (e));
    rc = GCG_ENOMEM;
  }
  while (!rc) {
    { gcg_kscope ks (ctx, "route_collect_mask");
      route_collect_kernel<0><<<warp_grid (ctx, r->n_tiles), 32 * RT_WARPS, 0, ctx->stream>>> (
          s->d_packed, s->d_woff, s->d_len, s->d_tseq, s->n, s->n_words, r->k, (uint32_t) r->n_part, r->tile0, r->n_tiles,
          r->d_off, r->d_seg, (const unsigned long long *) d_answers, d_mask, nullptr, nullptr, r->d_pass); }
    if (cudaGetLastError () != cudaSuccess) { gcg_set_error ("gcg_route_collect: kernel launch failed"); rc = GCG_ECUDA; break; }
    int64_t n_hit = 0;
    if ((rc = gcg_mask_scan (ctx, d_mask, n_words, d_prefix, d_bsum, &n_hit)) != 0) break;
    h->n = n_hit;
    if (n_hit > 0) {
      if ((e = gcg_dmalloc (ctx, &h->d_hits, (size_t) n_hit * sizeof (gcg_hit))) != cudaSuccess) {
        gcg_set_error ("gcg_route_collect: cudaMalloc of %lld anchors failed: %s", (long long) n_hit, cudaGetErrorString (e));
        rc = GCG_ENOMEM;
        break;
      }
      { gcg_kscope ks (ctx, "route_collect_emit");
        route_collect_kernel<1><<<warp_grid (ctx, r->n_tiles), 32 * RT_WARPS, 0, ctx->stream>>> (
            s->d_packed, s->d_woff, s->d_len, s->d_tseq, s->n, s->n_words, r->k, (uint32_t) r->n_part, r->tile0, r->n_tiles,
            r->d_off, r->d_seg, (const unsigned long long *) d_answers, d_mask, d_prefix, h->d_hits, r->d_pass); }
      if (cudaGetLastError () != cudaSuccess || cudaStreamSynchronize (ctx->stream) != cudaSuccess) {
        gcg_set_error ("gcg_route_collect: emit failed: %s", cudaGetErrorString (cudaGetLastError ())); rc = GCG_ECUDA; }
    }
    break;
  }
  gcg_dfree (ctx, d_mask); gcg_dfree (ctx, d_prefix); gcg_dfree (ctx, d_bsum);
  if (rc) { gcg_hits_free (h); return rc; }
  *out = h;
  return GCG_OK;
}

// ---- owner side ------------------------------------------------------------------------------------
extern "C" int gcg_table_create (gcg_ctx * ctx, int64_t n_records, int k, gcg_table ** out)
{
  GCG_CHECK (ctx && out && n_records >= 0, GCG_EINVAL, "gcg_table_create: bad argument");
  return gcg_table_alloc (ctx, n_records, k, out);
}

extern "C" int gcg_table_insert_records (gcg_ctx * ctx, gcg_table * t, const void * d_records, int64_t n)
{
  GCG_CHECK (ctx && t && n >= 0 && (d_records || n == 0), GCG_EINVAL, "gcg_table_insert_records: bad argument");
  GCG_CHECK (n <= t->n_inserted, GCG_ERANGE, "gcg_table_insert_records: %lld records into a table created for %lld", (long long) n, (long long) t->n_inserted);
  GCG_CUDA (cudaSetDevice (ctx->device));
  if (n == 0) return GCG_OK;
  t->filter_valid = false;
  gcg_kscope ks (ctx, "part_insert");
  insert_records_kernel<<<flat_grid (ctx, n, 1), 256, 0, ctx->stream>>> ((const ulonglong2 *) d_records, n, t->d_keys, t->d_vals, t->n_bucket);
  GCG_CUDA (cudaGetLastError ());
  return GCG_OK;
}

static int lookup_launch (gcg_ctx * ctx, gcg_table * t, const void * d_keys, int64_t n, const lookup_dst & dst)
{
  gcg_kscope ks (ctx, "part_lookup");
  lookup_keys_kernel<<<flat_grid (ctx, n, LK_UNROLL), 256, 0, ctx->stream>>> (
      (const unsigned long long *) d_keys, n, t->d_keys, t->d_vals, t->n_bucket, dst);
  GCG_CUDA (cudaGetLastError ());
  return GCG_OK;
}

extern "C" int gcg_table_lookup_keys (gcg_ctx * ctx, gcg_table * t, const void * d_keys, int64_t n, void * d_answers)
{
  GCG_CHECK (ctx && t && n >= 0 && ((d_keys && d_answers) || n == 0), GCG_EINVAL, "gcg_table_lookup_keys: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  if (n == 0) return GCG_OK;
  lookup_dst dst;
  memset (&dst, 0, sizeof dst);
  dst.n_src = 1; dst.base[0] = (unsigned long long *) d_answers; dst.first[0] = 0; dst.first[1] = n;
  return lookup_launch (ctx, t, d_keys, n, dst);
}

extern "C" int gcg_table_lookup_keys_direct (gcg_ctx * ctx, gcg_table * t, const void * d_keys, int n_src, const int64_t * src_count,
                                             void * const * d_answer_base, const int64_t * answer_off)
{
  GCG_CHECK (ctx && t && src_count && d_answer_base && answer_off && n_src >= 1 && n_src <= GCG_MAX_PART, GCG_EINVAL,
             "gcg_table_lookup_keys_direct: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  lookup_dst dst;
  memset (&dst, 0, sizeof dst);
  dst.n_src = n_src;
  int64_t n = 0;
  for (int r = 0; r < n_src; ++r) {
    GCG_CHECK (src_count[r] >= 0 && answer_off[r] >= 0 && (d_answer_base[r] || src_count[r] == 0), GCG_EINVAL,
               "gcg_table_lookup_keys_direct: bad source %d", r);
    dst.first[r] = n;
    dst.base[r] = (unsigned long long *) d_answer_base[r] + answer_off[r];
    n += src_count[r];
  }
  dst.first[n_src] = n;
  if (n == 0) return GCG_OK;
  GCG_CHECK (d_keys != nullptr, GCG_EINVAL, "gcg_table_lookup_keys_direct: no keys");
  return lookup_launch (ctx, t, d_keys, n, dst);
}

// ---- exchange windows: plain cudaMalloc blocks that other processes map through CUDA IPC and other
//      GPUs reach over NVLink (peer access)
struct gcg_window { gcg_ctx * ctx = nullptr; void * p = nullptr; int64_t bytes = 0; };

extern "C" int gcg_window_create (gcg_ctx * ctx, int64_t bytes, gcg_window ** out)
{
  GCG_CHECK (ctx && out && bytes >= 0, GCG_EINVAL, "gcg_window_create: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  gcg_window * w = new gcg_window ();
  w->ctx = ctx; w->bytes = std::max<int64_t> (bytes, 256);
  cudaError_t e = cudaMalloc (&w->p, (size_t) w->bytes);
  if (e != cudaSuccess) {
    // the parked blocks of the context may be what is in the way
    cudaGetLastError ();
    gcg_dcache_release (ctx);
    cudaStreamSynchronize (ctx->stream);
    e = cudaMalloc (&w->p, (size_t) w->bytes);
  }
  if (e != cudaSuccess) { gcg_set_error ("gcg_window_create: cudaMalloc of %lld bytes failed: %s", (long long) w->bytes, cudaGetErrorString (e)); delete w; return GCG_ENOMEM; }
  *out = w;
  return GCG_OK;
}

extern "C" void * gcg_window_ptr (gcg_window * w) { return w ? w->p : nullptr; }
extern "C" int64_t gcg_window_bytes (gcg_window * w) { return w ? w->bytes : 0; }

extern "C" int gcg_window_export (gcg_window * w, void * handle64)
{
  GCG_CHECK (w && handle64, GCG_EINVAL, "gcg_window_export: bad argument");
  static_assert (sizeof (cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  GCG_CUDA (cudaSetDevice (w->ctx->device));
  cudaIpcMemHandle_t h;
  GCG_CUDA (cudaIpcGetMemHandle (&h, w->p));
  memcpy (handle64, &h, 64);
  return GCG_OK;
}

extern "C" int gcg_window_open (gcg_ctx * ctx, const void * handle64, void ** d_peer)
{
  GCG_CHECK (ctx && handle64 && d_peer, GCG_EINVAL, "gcg_window_open: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  cudaIpcMemHandle_t h;
  memcpy (&h, handle64, 64);
  GCG_CUDA (cudaIpcOpenMemHandle (d_peer, h, cudaIpcMemLazyEnablePeerAccess));
  return GCG_OK;
}

extern "C" int gcg_window_close (gcg_ctx * ctx, void * d_peer)
{
  GCG_CHECK (ctx, GCG_EINVAL, "gcg_window_close: ctx == NULL");
  if (!d_peer) return GCG_OK;
  GCG_CUDA (cudaSetDevice (ctx->device));
  GCG_CUDA (cudaIpcCloseMemHandle (d_peer));
  return GCG_OK;
}

extern "C" void gcg_window_free (gcg_window * w)
{
  if (!w) return;
  cudaSetDevice (w->ctx->device);
  cudaStreamSynchronize (w->ctx->stream);
  cudaFree (w->p);
  delete w;
}

// same-process peers (one process driving several GPUs): let this context's device reach `peer_device`
extern "C" int gcg_peer_enable (gcg_ctx * ctx, int peer_device)
{
  GCG_CHECK (ctx, GCG_EINVAL, "gcg_peer_enable: ctx == NULL");
  if (peer_device == ctx->device) return GCG_OK;
  GCG_CUDA (cudaSetDevice (ctx->device));
  int can = 0;
  GCG_CUDA (cudaDeviceCanAccessPeer (&can, ctx->device, peer_device));
  GCG_CHECK (can, GCG_ECUDA, "gcg_peer_enable: device %d cannot access device %d", ctx->device, peer_device);
  cudaError_t e = cudaDeviceEnablePeerAccess (peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError (); e = cudaSuccess; }
  GCG_CUDA (e);
  return GCG_OK;
}

