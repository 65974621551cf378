// N3: the anchor grouping of map_ont2contigs (ctg_graph.c:600-656) on the device.
//
// The reference walks the dense okmers[] array of every read (16 bytes per ONT base) and cuts the
// read's anchors, in position order, into RUNS of consecutive anchors on the same contig; per run it
// counts the anchors that agree with the scaffold's strand (ONT_KMER_REV == KMER_REV, ctg_graph.c:617-623,
// 642-648: "FORW") and those that do not ("BACK"), and ont_node_init (ctg_graph.c:93-181) keeps, of
// the whole run, only: the contig, the two counts, and the first and the last anchor of the
// majority direction.  That is what a run record below holds — a few dozen bytes per run instead of
// 8 or 16 per anchor, so neither the anchors nor a dense per-base array ever reach the host.
//
//   gcg_run { tid, n_fwd, n_bwd, first_fwd, last_fwd, first_bwd, last_bwd }   anchors as compact words
//
// One warp per read, two passes (count runs per read, prefix sum, write): the records of a read are
// consecutive and in read order, reads in input order — deterministic, no atomics.
#include <algorithm>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gcg_internal.cuh"
#include "kmer_dev.cuh"

struct run_acc {
  int32_t tid, n_fwd, n_bwd;
  unsigned long long ff, lf, fb, lb;
};

// largest c in [0, n_ctg) with cbase[c] <= gpos (cbase has n_ctg + 1 entries; empty contigs share a base with
// their successor, and the successor is the one that holds the base)
__device__ __forceinline__ int32_t contig_of (const int64_t * __restrict__ cbase, int32_t n_ctg, int64_t gpos)
{
  int32_t lo = 0, hi = n_ctg;
  while (hi - lo > 1) {
    const int32_t mid = (lo + hi) >> 1;
    if (__ldg (cbase + mid) <= gpos) lo = mid; else hi = mid;
  }
  return lo;
}

template <bool EMIT>
__global__ void __launch_bounds__ (256)
runs_kernel (const unsigned long long * __restrict__ anchors, const long long * __restrict__ read_off, int64_t n_read,
             const int64_t * __restrict__ cbase, int32_t n_ctg, uint32_t * __restrict__ run_count,
             const uint32_t * __restrict__ run_base, gcg_run * __restrict__ out)
{
  const int lane = threadIdx.x & 31;
  const int64_t wstride = (int64_t) gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t) blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_read; r += wstride) {
    const long long beg = read_off[r], end = read_off[r + 1];
    GCG_DEV_ASSERT (beg >= 0 && end >= beg && end <= read_off[n_read]);
    run_acc acc;                                    // the open run (replicated in every lane)
    acc.tid = -1; acc.n_fwd = acc.n_bwd = 0; acc.ff = acc.lf = acc.fb = acc.lb = 0;
    uint32_t n_runs = 0;
    gcg_run * dst = EMIT ? out + run_base[r] : nullptr;
    auto close_run = [&] () {
      if (acc.tid < 0) return;
      if (EMIT && lane == 0) {
        gcg_run g;
        g.tid = acc.tid; g.n_fwd = acc.n_fwd; g.n_bwd = acc.n_bwd; g.pad = 0;
        g.first_fwd = acc.ff; g.last_fwd = acc.lf; g.first_bwd = acc.fb; g.last_bwd = acc.lb;
        GCG_DEV_ASSERT (acc.tid < n_ctg && acc.n_fwd + acc.n_bwd > 0 && (r + 1 == n_read || run_base[r] + n_runs < run_base[r + 1]));
        dst[n_runs] = g;
      }
      ++n_runs;
    };
    for (long long i0 = beg; i0 < end; i0 += 32) {
      const long long i = i0 + lane;
      const bool valid = i < end;
      const unsigned long long a = valid ? anchors[i] : 0ULL;
      const int32_t tid = valid ? contig_of (cbase, n_ctg, (int64_t) ((a >> 2) & 0x3FFFFFFFFULL)) : -2;
      // FORW: the read's strand flag and the contig occurrence's strand flag agree (ctg_graph.c:617-618)
      const bool fwd = (((a >> 1) ^ a) & 1ULL) == 0ULL;
      int32_t prev = __shfl_up_sync (0xffffffffu, tid, 1);
      if (lane == 0) prev = acc.tid;
      const uint32_t smask = __ballot_sync (0xffffffffu, valid && tid != prev);
      const uint32_t vmask = __ballot_sync (0xffffffffu, valid);
      const uint32_t fmask = __ballot_sync (0xffffffffu, valid && fwd);
      const int nvalid = __popc (vmask);
      int pos = 0;
      while (pos < nvalid) {
        const uint32_t after = pos >= 31 ? 0u : (smask & ~((2u << pos) - 1u));      // run starts strictly behind pos
        const int nxt = after ? __ffs (after) - 1 : nvalid;
        const uint32_t seg = (nxt >= 32 ? 0xffffffffu : ((1u << nxt) - 1u)) & ~((1u << pos) - 1u);
        if ((smask >> pos) & 1u) {
          close_run ();
          acc.tid = __shfl_sync (0xffffffffu, tid, pos);
          acc.n_fwd = acc.n_bwd = 0; acc.ff = acc.lf = acc.fb = acc.lb = 0;
        }
        const uint32_t F = fmask & seg, B = vmask & seg & ~fmask;
        if (F) {
          const unsigned long long first = __shfl_sync (0xffffffffu, a, __ffs (F) - 1), last = __shfl_sync (0xffffffffu, a, 31 - __clz (F));
          if (acc.n_fwd == 0) acc.ff = first;
          acc.lf = last;
          acc.n_fwd += __popc (F);
        }
        if (B) {
          const unsigned long long first = __shfl_sync (0xffffffffu, a, __ffs (B) - 1), last = __shfl_sync (0xffffffffu, a, 31 - __clz (B));
          if (acc.n_bwd == 0) acc.fb = first;
          acc.lb = last;
          acc.n_bwd += __popc (B);
        }
        pos = nxt;
      }
    }
    close_run ();
    if (!EMIT && lane == 0) run_count[r] = n_runs;
  }
}

// exclusive prefix sum of n 32-bit counts in place, one block (n is a number of reads); *total receives the sum
__global__ void __launch_bounds__ (1024)
runs_scan_kernel (uint32_t * __restrict__ v, int64_t n, unsigned long long * __restrict__ total)
{
  __shared__ uint32_t s_warp[32];
  __shared__ unsigned long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads ();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t idx = base + threadIdx.x;
    const uint32_t c = idx < n ? v[idx] : 0;
    uint32_t x = c;
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads ();
    if (threadIdx.x < 32) {
      uint32_t wv = s_warp[threadIdx.x], wx = wv;
      for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync (0xffffffffu, wx, o); if (threadIdx.x >= o) wx += y; }
      s_warp[threadIdx.x] = wx - wv;
    }
    __syncthreads ();
    const unsigned long long excl = (unsigned long long) (x - c) + s_warp[threadIdx.x >> 5] + s_carry;
    if (idx < n) v[idx] = (uint32_t) excl;
    __syncthreads ();
    if (threadIdx.x == 1023) s_carry = excl + c;
    __syncthreads ();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

// d_anchors: compact anchors in (read,pos) order; d_read_off: n_read + 1 offsets, every entry filled.
// Returns pinned arrays (gcg_free): runs_out[n_run] and run_off_out[n_read + 1].
int gcg_runs_from_anchors (gcg_ctx * ctx, const gcg_table * t, const void * d_anchors, const long long * d_read_off, int64_t n_read,
                           gcg_run ** runs_out, int64_t ** run_off_out, int64_t * n_run)
{
  GCG_CHECK (ctx && t && t->d_cbase && d_read_off && runs_out && run_off_out && n_run && n_read >= 0, GCG_EINVAL, "gcg_runs: bad argument");
  GCG_CHECK (t->n_contig < 0x7FFFFFFF, GCG_ERANGE, "gcg_runs: too many contigs");
  *runs_out = nullptr; *n_run = 0;
  int64_t * run_off = (int64_t *) gcg_pinned_alloc ((size_t) (n_read + 1) * 8);
  GCG_CHECK (run_off != nullptr, GCG_ENOMEM, "gcg_runs: pinned alloc of %lld offsets failed", (long long) n_read + 1);
  *run_off_out = run_off;
  if (n_read == 0) { run_off[0] = 0; return GCG_OK; }
  uint32_t * d_cnt = nullptr;
  gcg_run * d_runs = nullptr;
  cudaError_t e = gcg_dmalloc (ctx, &d_cnt, (size_t) n_read * 4);
  if (e != cudaSuccess) { gcg_set_error ("gcg_runs: cudaMalloc failed: %s", cudaGetErrorString (e)); return GCG_ENOMEM; }
  const int grid = (int) std::min<int64_t> ((n_read + 7) / 8, (int64_t) ctx->sm_count * 8);
  int rc = GCG_OK;
  { gcg_kscope ks (ctx, "runs_count");
    runs_kernel<false><<<grid, 256, 0, ctx->stream>>> ((const unsigned long long *) d_anchors, d_read_off, n_read, t->d_cbase, (int32_t) t->n_contig, d_cnt, nullptr, nullptr); }
  { gcg_kscope ks (ctx, "runs_scan");
    runs_scan_kernel<<<1, 1024, 0, ctx->stream>>> (d_cnt, n_read, ctx->d_counters + 6); }
  if (cudaMemcpyAsync (ctx->h_counters + 6, ctx->d_counters + 6, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize (ctx->stream) != cudaSuccess) { gcg_set_error ("gcg_runs: count pass failed: %s", cudaGetErrorString (cudaGetLastError ())); rc = GCG_ECUDA; }
  const int64_t total = rc ? 0 : (int64_t) ctx->h_counters[6];
  if (!rc && total >= 0xFFFFFFFFLL) { gcg_set_error ("gcg_runs: %lld runs exceed the 32-bit run index", (long long) total); rc = GCG_ERANGE; }
  gcg_run * runs = nullptr;
  std::vector<uint32_t> h_base;
  if (!rc) {
    h_base.resize ((size_t) n_read);
    if (total > 0) {
      runs = (gcg_run *) gcg_pinned_alloc ((size_t) total * sizeof (gcg_run));
      if (!runs) { gcg_set_error ("gcg_runs: pinned alloc of %lld runs failed", (long long) total); rc = GCG_ENOMEM; }
      if (!rc && (e = gcg_dmalloc (ctx, &d_runs, (size_t) total * sizeof (gcg_run))) != cudaSuccess) { gcg_set_error ("gcg_runs: cudaMalloc failed: %s", cudaGetErrorString (e)); rc = GCG_ENOMEM; }
      if (!rc) {
        gcg_kscope ks (ctx, "runs_emit");
        runs_kernel<true><<<grid, 256, 0, ctx->stream>>> ((const unsigned long long *) d_anchors, d_read_off, n_read, t->d_cbase, (int32_t) t->n_contig, nullptr, d_cnt, d_runs);
      }
      if (!rc && cudaMemcpyAsync (runs, d_runs, (size_t) total * sizeof (gcg_run), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) { gcg_set_error ("gcg_runs: download failed"); rc = GCG_ECUDA; }
    }
    if (!rc && (cudaMemcpyAsync (h_base.data (), d_cnt, (size_t) n_read * 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                cudaStreamSynchronize (ctx->stream) != cudaSuccess)) { gcg_set_error ("gcg_runs: emit pass failed: %s", cudaGetErrorString (cudaGetLastError ())); rc = GCG_ECUDA; }
  }
  gcg_dfree (ctx, d_cnt); gcg_dfree (ctx, d_runs);
  if (rc) { gcg_free (runs); gcg_free (run_off); *run_off_out = nullptr; return rc; }
  for (int64_t r = 0; r < n_read; ++r) run_off[r] = (int64_t) h_base[(size_t) r];
  run_off[n_read] = total;
  *runs_out = runs;
  *n_run = total;
  return GCG_OK;
}
