// Device-side helpers shared by the k-mer translation units (kmer.cu, part.cu): 2-bit k-mer
// arithmetic, the bucket probe, sequence lookup.  Reference semantics cited in kmer.cu.
#pragma once
#include "gcg_internal.cuh"

// build-time switches of the search kernels (A/B variants: scripts/build_variants.sh)
#ifndef K45_EVICT_LAST
#define K45_EVICT_LAST 0
#endif
// GCG_CHECKED=1: bounds and invariant checks inside the kernels (index ranges of every scattered store and of the
// chained scan, monotonic prefixes, slot ranges); a violation prints its location and traps.  compute-sanitizer is not
// available on the GPU pool this was developed on (profiles/r02_sanitizer_refused.log), so the GPU test-suite is also run
// against this build (scripts/build_variants.sh checked "-DGCG_CHECKED=1"; profiles/r02_checked_build.log).
#ifndef GCG_CHECKED
#define GCG_CHECKED 0
#endif
#if GCG_CHECKED
#include <stdio.h>
#define GCG_DEV_ASSERT(c) do { if (!(c)) { printf ("GCG_CHECKED violation: %s  (%s:%d, block %d thread %d)\n", #c, __FILE__, __LINE__, (int) blockIdx.x, (int) threadIdx.x); __trap (); } } while (0)
#else
#define GCG_DEV_ASSERT(c) do { } while (0)
#endif

// =============================================================================================
// device helpers
// =============================================================================================
__device__ __forceinline__ uint32_t pack4 (uint32_t x)
{
  // four ASCII bytes (first base in the low byte) -> 8 bits, first base in the top two bits
  return (((x >> 1) & 0x03030303u) * 0x40100401u) >> 24;
}

__device__ __forceinline__ uint64_t revcomp64 (uint64_t x, int k)
{
  // kseq1.h:37-46: complement (^2 per base), reverse the 2-bit groups, drop the unused tail
  x ^= 0xAAAAAAAAAAAAAAAAULL;
  x = ((x & 0x3333333333333333ULL) << 2) | ((x >> 2) & 0x3333333333333333ULL);
  x = ((x & 0x0F0F0F0F0F0F0F0FULL) << 4) | ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL);
  uint32_t lo = (uint32_t) x, hi = (uint32_t) (x >> 32);
  lo = __byte_perm (lo, 0, 0x0123);
  hi = __byte_perm (hi, 0, 0x0123);
  x = ((uint64_t) lo << 32) | hi;
  return x >> (64 - 2 * k);
}

__device__ __forceinline__ uint32_t kmer_hash32 (uint64_t key)
{
  uint32_t h = (uint32_t) key ^ ((uint32_t) (key >> 32) * 0x9E3779B1u);
  h ^= h >> 16; h *= 0x85EBCA6Bu;
  h ^= h >> 13; h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}

// second, independent 32-bit hash (word index of the pre-filter) and the filter bits of a key
__device__ __forceinline__ uint32_t kmer_hash32b (uint64_t key)
{
  uint32_t h = (uint32_t) (key >> 32) * 0x85EBCA6Bu ^ (uint32_t) key * 0xC2B2AE35u;
  h ^= h >> 15; h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  return h;
}

__device__ __forceinline__ uint32_t filter_mask (uint32_t h, int k3)
{
  uint32_t m = (1u << (h & 31u)) | (1u << ((h >> 5) & 31u));
  if (k3) m |= 1u << ((h >> 10) & 31u);
  return m;
}

// owner partition of a canonical k-mer (hash-partitioned table, part.cu / the remote-probe search of kmer.cu): a 64-bit
// mix that is independent of the in-table bucket hash (which hash picks the owner is unobservable)
__host__ __device__ __forceinline__ uint32_t kmer_owner (uint64_t key, uint32_t n_part)
{
  uint64_t x = key * 0x9E3779B97F4A7C15ULL;
  x ^= x >> 29;
  x *= 0xBF58476D1CE4E5B9ULL;
  const uint32_t hi = (uint32_t) (x >> 32);          // (the usual closing x ^= x >> 32 only touches the low half)
#ifdef __CUDA_ARCH__
  return __umulhi (hi, n_part);
#else
  return (uint32_t) (((uint64_t) hi * (uint64_t) n_part) >> 32);
#endif
}

// one partition of a table as a searching GPU sees it: its key and value arrays — local memory, or a peer GPU's,
// mapped through CUDA IPC / peer access and reached over NVLink — and its bucket count
struct gcg_part_desc { const unsigned long long * keys; unsigned long long * vals; uint32_t n_bucket; uint32_t pad; };

// largest s in [0,n) with woff[s] <= w  (woff has n+1 entries, woff[n] > w)
__device__ __forceinline__ int64_t find_seq (const int64_t * __restrict__ woff, int64_t n, int64_t w)
{
  int64_t lo = 0, hi = n;          // invariant: woff[lo] <= w < woff[hi]
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg (woff + mid) <= w) lo = mid; else hi = mid;
  }
  return lo;
}

// same, starting from a hint h <= answer (the sequence of the first word of the 32-word tile, built
// on the host): reads are hundreds of words long, so this is almost always zero or one step
__device__ __forceinline__ int64_t find_seq_from (const int64_t * __restrict__ woff, int64_t n, int64_t w, int64_t h)
{
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    if (h + 1 >= n || __ldg (woff + h + 1) > w) return h;
    ++h;
  }
  int64_t lo = h, hi = n;            // many short or empty sequences inside one tile: finish by bisection
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg (woff + mid) <= w) lo = mid; else hi = mid;
  }
  return lo;
}

// Per-lane rolling state over the 32 k-mer start positions of one packed word.
struct kroll {
  uint64_t fwd, rc, nxt, mask;
  int shift_rc;
  __device__ __forceinline__ void init (uint64_t hi, uint64_t lo, int k)
  {
    mask = (1ULL << (2 * k)) - 1;       // k <= 31
    fwd = hi >> (64 - 2 * k);
    rc = revcomp64 (fwd, k);
    nxt = (hi << (2 * k)) | (lo >> (64 - 2 * k));
    shift_rc = 2 * (k - 1);
  }
  __device__ __forceinline__ void step ()
  {
    uint64_t b = nxt >> 62;
    nxt <<= 2;
    fwd = ((fwd << 2) | b) & mask;
    rc = (rc >> 2) | ((b ^ 2ULL) << shift_rc);
  }
};


struct __align__ (32) bucket4 { unsigned long long a, b, c, d; };

__device__ __forceinline__ bucket4 ld_bucket (const unsigned long long * p)
{
  bucket4 r;
  asm volatile ("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d) : "l"(p));
  return r;
}

// the same load for the probe loop of the search: the table's key sectors are asked to stay in the L2
// (evict-last) while the packed reads, the anchor records and the value words stream past them
__device__ __forceinline__ bucket4 ld_bucket_keep (const unsigned long long * p)
{
  bucket4 r;
#if K45_EVICT_LAST
  asm volatile ("ld.global.nc.L1::no_allocate.L2::evict_last.v4.u64 {%0,%1,%2,%3}, [%4];"
                : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d) : "l"(p));
#else
  asm volatile ("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
                : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d) : "l"(p));
#endif
  return r;
}

__device__ __forceinline__ void prefetch_l2 (const void * p)
{
  asm volatile ("prefetch.global.L2 [%0];" :: "l"(p));
}

// slot index (0..3) of `key` in the bucket or -1; *kw receives the matching key word
__device__ __forceinline__ int bucket_find (const bucket4 & q, unsigned long long key, unsigned long long * kw)
{
  if ((q.a & GCG_KEY_MASK) == key) { *kw = q.a; return 0; }
  if ((q.b & GCG_KEY_MASK) == key) { *kw = q.b; return 1; }
  if ((q.c & GCG_KEY_MASK) == key) { *kw = q.c; return 2; }
  if ((q.d & GCG_KEY_MASK) == key) { *kw = q.d; return 3; }
  return -1;
}

// has a key with fingerprint fp (hash & 3) ever been pushed out of this bucket?
__device__ __forceinline__ bool bucket_ovf (const bucket4 & q, uint32_t fp)
{
  uint32_t h0 = (uint32_t) (q.a >> 32), h1 = (uint32_t) (q.b >> 32), h2 = (uint32_t) (q.c >> 32), h3 = (uint32_t) (q.d >> 32);
  uint32_t h = (fp & 2u) ? ((fp & 1u) ? h3 : h2) : ((fp & 1u) ? h1 : h0);
  return (h & 0x40000000u) != 0;
}

// the same test as a bit at position u (u < 7, a compile-time constant in the search loop), without
// branches: the top bytes of the four slots side by side, bit 6 of byte fp is the mark
__device__ __forceinline__ uint32_t bucket_ovf_bit (const bucket4 & q, uint32_t fp, int u)
{
  const uint32_t t01 = __byte_perm ((uint32_t) (q.a >> 32), (uint32_t) (q.b >> 32), 0x0073);
  const uint32_t t23 = __byte_perm ((uint32_t) (q.c >> 32), (uint32_t) (q.d >> 32), 0x0073);
  const uint32_t w = __byte_perm (t01, t23, 0x5410);
  return (w >> (fp * 8u + (uint32_t) (6 - u))) & (1u << u);
}

// key present with multiplicity 1 in this bucket (key has bits 62/63 clear, so one compare does both)
__device__ __forceinline__ bool bucket_has_unique (const bucket4 & q, unsigned long long key)
{
  return ((q.a & ~GCG_KEY_OVF) == key) | ((q.b & ~GCG_KEY_OVF) == key) | ((q.c & ~GCG_KEY_OVF) == key) | ((q.d & ~GCG_KEY_OVF) == key);
}

// full probe sequence starting from an already loaded bucket; returns slot index or ~0
__device__ __forceinline__ unsigned long long table_lookup (const unsigned long long * __restrict__ keys, uint32_t n_bucket,
                                                           uint32_t b, bucket4 q, unsigned long long key, uint32_t fp, unsigned long long * kw)
{
  for (;;) {
    GCG_DEV_ASSERT (b < n_bucket);
    int f = bucket_find (q, key, kw);
    if (f >= 0) return 4ULL * b + f;
    if (!bucket_ovf (q, fp)) return ~0ULL;           // no key of this fingerprint ever left the bucket: absent
    b = (b + 1 == n_bucket) ? 0 : b + 1;
    q = ld_bucket (keys + 4ULL * b);
  }
}

// canonical key (+1) of the k-mer starting at bit offset 2j of the 128-bit window (wh, wl)
__device__ __forceinline__ unsigned long long key_at (uint64_t wh, uint64_t wl, int j, int k, bool * fw)
{
  uint64_t xw = j ? ((wh << (2 * j)) | (wl >> (64 - 2 * j))) : wh;
  uint64_t fwd = xw >> (64 - 2 * k), rc = revcomp64 (fwd, k);
  *fw = fwd < rc;
  return (*fw ? fwd : rc) + 1ULL;
}


// coherent 256-bit bucket load for the build (the table is being written by other threads: L2 is
// the point of coherence, so .cg; one request instead of up to four dependent 8-byte loads)
__device__ __forceinline__ bucket4 ld_bucket_cg (const unsigned long long * p)
{
  bucket4 r;
  asm volatile ("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];"
                : "=l"(r.a), "=l"(r.b), "=l"(r.c), "=l"(r.d) : "l"(p) : "memory");
  return r;
}

// insert one (key+1, val) pair: claim an empty slot with atomicCAS, or flag the key as seen more
// than once (kmer.c:124-152 -> hash.c:113-152 counts occurrences; only {1, >=2} is observable)
__device__ __forceinline__ void table_insert (unsigned long long * __restrict__ keys, unsigned long long * __restrict__ vals,
                                              uint32_t n_bucket, unsigned long long key, unsigned long long val)
{
  const uint32_t hsh = kmer_hash32 (key - 1ULL), fp = hsh & 3u;
  uint32_t b = __umulhi (hsh, n_bucket);
  for (;;) {
    GCG_DEV_ASSERT (b < n_bucket);
    unsigned long long * slot = keys + 4ULL * b;
    const bucket4 q = ld_bucket_cg (slot);
    unsigned long long cur[4] = {q.a, q.b, q.c, q.d};
    bool done = false;
#pragma unroll
    for (int i = 0; i < 4 && !done; ++i) {
      if (cur[i] == 0ULL) {
        // slots fill front to back, so an empty slot ends the bucket: claim it (or see who did)
        unsigned long long old = atomicCAS (slot + i, 0ULL, key);
        if (old == 0ULL) { vals[4ULL * b + i] = val; done = true; break; }
        cur[i] = old;
      }
      if ((cur[i] & GCG_KEY_MASK) == key) {
        if (!(cur[i] & GCG_KEY_MULTI)) atomicOr (slot + i, GCG_KEY_MULTI);
        done = true;
      }
    }
    if (done) return;
    // bucket is full of other keys: leave the key's overflow mark (bit 62 of slot `fp`) and move on
    const unsigned long long cf = (fp & 2u) ? ((fp & 1u) ? cur[3] : cur[2]) : ((fp & 1u) ? cur[1] : cur[0]);
    if (!(cf & GCG_KEY_OVF)) atomicOr (slot + fp, GCG_KEY_OVF);
    b = (b + 1 == n_bucket) ? 0 : b + 1;
  }
}
