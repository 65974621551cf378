// K7..K9: batched Smith-Waterman fill (anti-diagonal wavefront inside a warp), end-cell selection
// with the reference's tie rules, trace spill (4 bits/cell) to HBM and CIGAR emission.
//
// Reference semantics restated (paths relative to /root/reference/gap_closer):
//   DP cell         sw.c:203-252   ms = H[i-1][j-1] + mat[q*type_c+t]
//                                  ds = max(H[i-1][j]-del_o  (only if strictly greater), ds[i-1][j]-del_e)
//                                  is = max(H[i][j-1]-ins_o  (only if strictly greater), is[i][j-1]-ins_e)
//                                  H  = ms if ms>=ds && ms>=is ; else is if is>ds ; else ds
//   borders         sw.c:61-110    row 0 / column 0: is = ds = -inf, run lengths 0, scores 0 or affine
//   end cell        sw.c:254-280   last column (last maximum wins), then last row (strictly greater,
//                                  or equal and closer to the diagonal)
//   traceback       sw.c:282-335   as shipped the cell is never re-fetched (GCG_SW_ASIS); GCG_SW_FIXED re-fetches
//   cigar           cigar.c:82-125 BAM encoding (len<<4 | op), "MIDNSHP=XB", reversed at the end
//
// Arithmetic.  Every DP value v is carried as 16*v + tag.  The four low bits make the reference's
// tie rules fall out of a plain integer max and leave the traceback flags in place:
//   I candidates : open 0b0000, extend 0b0001      (extend wins a tie  <=> open needs strictly greater)
//   D candidates : open 0b0100, extend 0b0110      (D beats I on a tie <=> "is > ds" needed for I)
//   M candidate  :      0b1000                     (M wins every tie)
// so  trace nibble = ((D|I) & 3) | (H & 12):  bit0 I-extended, bit1 D-extended, bit2 H==D, bit3 H==M.
// Two kernels share this scheme: a packed one (two alignments per warp in the halves of
// s16x2 registers: VIADD.16x2 / VIMNMX.S16x2 / VIADDMNMX.S16x2, scores by PRMT from byte
// profiles) and a generic s32 one (any alphabet <= 8, any lengths).  Which pairs may use the packed
// kernel is decided on the host from provable value bounds — never from the data.
//
// Wavefront.  Lane l of a warp owns 8 consecutive query columns of a 256-column band and walks
// down the target rows; at step s it is on row s-l+1 and receives H and I of its left neighbour's
// last column by __shfl_up_sync.  The last column of a band is parked (shared memory in the packed
// kernel, an L2-resident scratch in the generic one) for lane 0 of the next band.
//
// Trace layout per alignment (what hits HBM, 0.5 byte per cell, fully coalesced 128-byte rows):
//   word[(band * (tlen+31) + step) * 32 + lane], nibble of column c at bit shift(c)
#include <algorithm>
#include <type_traits>
#include <string>
#include <thread>
#include <vector>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gcg_internal.cuh"

#define SW_BAND 256
#define SW_COLS 8

struct sw_task {
  long long qoff, toff;             // offsets into the batch symbol buffers
  int qlen, tlen;
  unsigned long long trace_off;     // in u32 words, into the wave's trace buffer
  long long edge_off;               // in ints, into the wave's edge buffer: lastcol[tlen] then lastrow[qlen]
  int pair;                         // index in the batch
  int pad;
};

struct sw_end { int score, bt_tidx, bt_qidx, seg_len; };

struct sw_consts {
  int type_c, del_o, del_e, ins_o, ins_e, strategy;
  int border_kind, b_del_o, b_del_e, b_ins_o, b_ins_e;
  int mat16[64];                    // 16*mat + 8
  unsigned prof4[4];                // packed kernel: byte t of prof4[q] = (signed char)(16*mat[q][t]+8)
  // packed kernel: the four gap constants as s16x2 words, built on the host so that the step loop
  // reads them as constant-bank operands instead of rebuilding them (IMAD/LOP3 per step)
  unsigned pk_do, pk_de, pk_io, pk_ie;
  unsigned pk_de32, pk_ie32;        // -16*e * 65537: both halves in one 32-bit add (values are biased, see SW_BIAS2)
};

__constant__ sw_consts c_sw;

__host__ __device__ __forceinline__ int sw_shift (int c) { return c < 4 ? 4 * (3 - c) : 16 + 4 * (7 - c); }

__device__ __forceinline__ int sw_brow (int j)   // border score of row 0, column j (sw.c:96-102)
{
  return (c_sw.border_kind && j >= 1) ? -c_sw.b_ins_o - (j - 1) * c_sw.b_ins_e : 0;
}
__device__ __forceinline__ int sw_bcol (int i)   // border score of column 0, row i (sw.c:104-109)
{
  return (c_sw.border_kind && i >= 1) ? -c_sw.b_del_o - (i - 1) * c_sw.b_del_e : 0;
}

// ---------------------------------------------------------------------------------------------
// end-cell selection by one warp over the parked last column / last row (sw.c:254-280)
// ---------------------------------------------------------------------------------------------
__device__ sw_end sw_pick_end (const int * edges, int qlen, int tlen, int lane, sw_end * out)
{
  const int * lastcol = edges, * lastrow = edges + tlen;
  // last column: the LAST maximum wins (>=)
  int best = INT_MIN, bi = tlen;
  for (int idx = lane; idx < tlen; idx += 32) {
    int v = __ldcg (lastcol + idx);
    if (v >= best) { best = v; bi = idx + 1; }
  }
  for (int o = 16; o; o >>= 1) {
    int ob = __shfl_xor_sync (0xffffffffu, best, o), oi = __shfl_xor_sync (0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi > bi && ob != INT_MIN)) { best = ob; bi = oi; }
  }
  if (tlen == 0) bi = 0;
  int bt_t = bi, bt_q = qlen, score = best, seg = 0;
  if (tlen == 0) score = INT_MIN;
  if (c_sw.strategy != GCG_SWOS_LEADING_INDEL && qlen > 0) {
    int m2 = INT_MIN;
    for (int idx = lane; idx < qlen; idx += 32) m2 = max (m2, __ldcg (lastrow + idx));
    for (int o = 16; o; o >>= 1) m2 = max (m2, __shfl_xor_sync (0xffffffffu, m2, o));
    if (m2 >= score) {
      // among columns holding m2: smallest |tlen - j|, first such j
      int bd = INT_MAX, bj = INT_MAX;
      for (int idx = lane; idx < qlen; idx += 32) {
        if (__ldcg (lastrow + idx) == m2) {
          int j = idx + 1, d = abs (tlen - j);
          if (d < bd || (d == bd && j < bj)) { bd = d; bj = j; }
        }
      }
      for (int o = 16; o; o >>= 1) {
        int od = __shfl_xor_sync (0xffffffffu, bd, o), oj = __shfl_xor_sync (0xffffffffu, bj, o);
        if (od < bd || (od == bd && oj < bj)) { bd = od; bj = oj; }
      }
      bool take = m2 > score || bd < abs (bt_t - bt_q);
      if (take) { bt_t = tlen; bt_q = bj; score = m2; seg = qlen - bj; }
    }
  }
  if (score == INT_MIN) score = sw_brow (bt_q);      // only reachable with tlen == 0: cell (0, bt_q)
  sw_end e;
  e.score = score; e.bt_tidx = bt_t; e.bt_qidx = bt_q; e.seg_len = seg;     // the same in every lane
  if (lane == 0) *out = e;
  return e;
}

// degenerate alignments (an empty side): the "last column / last row" are border cells
__device__ void sw_fill_degenerate (int * edges, int qlen, int tlen, int lane)
{
  if (qlen == 0) for (int i = lane; i < tlen; i += 32) __stcg (edges + i, sw_bcol (i + 1));
  if (tlen == 0) for (int j = lane; j < qlen; j += 32) __stcg (edges + tlen + j, sw_brow (j + 1));
}

// ---------------------------------------------------------------------------------------------
// K7 generic: one alignment per warp, s32 values
// ---------------------------------------------------------------------------------------------
#define SW_NEG32 (-(1 << 29))

__global__ void __launch_bounds__ (128)
sw_fill_generic_kernel (const uint8_t * __restrict__ qry, const uint8_t * __restrict__ tgt,
                        const sw_task * __restrict__ tasks, const int * __restrict__ items, int n_items,
                        int * __restrict__ counter, uint32_t * __restrict__ trace, int * __restrict__ edges,
                        int2 * __restrict__ bound, long long bound_stride, sw_end * __restrict__ ends)
{
  __shared__ int s_mat[64];
  if (threadIdx.x < 64) s_mat[threadIdx.x] = c_sw.mat16[threadIdx.x];
  __syncthreads ();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int2 * mybound = bound + (long long) (blockIdx.x * 4 + wid) * bound_stride;
  const int type_c = c_sw.type_c;
  const int cDO = -16 * c_sw.del_o + 4, cDE = -16 * c_sw.del_e, cIO = -16 * c_sw.ins_o, cIE = -16 * c_sw.ins_e;
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd (counter, 1);
    item = __shfl_sync (0xffffffffu, item, 0);
    if (item >= n_items) break;
    const sw_task tk = tasks[items[item]];
    const int qlen = tk.qlen, tlen = tk.tlen;
    int * my_edges = edges + tk.edge_off;
    if (qlen == 0 || tlen == 0) {
      sw_fill_degenerate (my_edges, qlen, tlen, lane);
    } else {
      const uint8_t * q = qry + tk.qoff, * t = tgt + tk.toff;
      uint32_t * tr = trace + tk.trace_off;
      const int nsteps = tlen + 31, nbands = (qlen + SW_BAND - 1) / SW_BAND;
      const int lastband = (qlen - 1) / SW_BAND, lastlane = ((qlen - 1) % SW_BAND) / SW_COLS, cstar = (qlen - 1) % SW_COLS;
      for (int band = 0; band < nbands; ++band) {
        const int j0 = band * SW_BAND + lane * SW_COLS;
        int qs[SW_COLS], Hup[SW_COLS], Dup[SW_COLS];
#pragma unroll
        for (int c = 0; c < SW_COLS; ++c) {
          qs[c] = (j0 + c < qlen ? (int) q[j0 + c] : 0) * type_c;
          Hup[c] = 16 * sw_brow (j0 + c + 1);
          Dup[c] = SW_NEG32;
        }
        int prevInH = 16 * sw_brow (j0), outH = 0, outI = SW_NEG32;
        for (int step = 0; step < nsteps; ++step) {
          int inH = __shfl_up_sync (0xffffffffu, outH, 1), inI = __shfl_up_sync (0xffffffffu, outI, 1);
          const int i = step - lane + 1;
          const bool active = i >= 1 && i <= tlen;
          if (lane == 0 && active) {
            if (band == 0) { inH = 16 * sw_bcol (i); inI = SW_NEG32; }
            else { int2 b = __ldcg (mybound + i); inH = b.x; inI = b.y; }
          }
          if (active) {
            const int ts = t[i - 1];
            int Hd = prevInH, Hl = inH, Il = inI;
            uint32_t acc = 0;
#pragma unroll
            for (int c = 0; c < SW_COLS; ++c) {
              int s = s_mat[qs[c] + ts];
              int dE = (Dup[c] | 2) + cDE;
              int D = __viaddmax_s32 (Hup[c], cDO, dE);
              int iE = (Il | 1) + cIE;
              int I = __viaddmax_s32 (Hl, cIO, iE);
              int g = max (D, I);
              int H = __viaddmax_s32 (Hd, s, g);
              int Hc = H & ~15;
              acc |= ((uint32_t) (((D | I) & 3) | (H & 12))) << sw_shift (c);
              Hd = Hup[c]; Hup[c] = Hc; Dup[c] = D; Hl = Hc; Il = I;
            }
            prevInH = inH;
            outH = Hl; outI = Il;
            tr[((size_t) band * nsteps + step) * 32 + lane] = acc;
            if (lane == 31 && band + 1 < nbands) __stcg (mybound + i, make_int2 (outH, outI));
            if (band == lastband && lane == lastlane) {
              int v = 0;
#pragma unroll
              for (int c = 0; c < SW_COLS; ++c) if (c == cstar) v = Hup[c];
              __stcg (my_edges + (i - 1), v >> 4);
            }
            if (i == tlen) {
#pragma unroll
              for (int c = 0; c < SW_COLS; ++c) if (j0 + c < qlen) __stcg (my_edges + tlen + j0 + c, Hup[c] >> 4);
            }
          }
        }
        __syncwarp ();
      }
    }
    __syncwarp ();
    sw_pick_end (my_edges, qlen, tlen, lane, ends + tk.pair);
    __syncwarp ();
  }
}

// ---------------------------------------------------------------------------------------------
// K9 traceback -> CIGAR, one thread per alignment (sw.c:282-335, cigar.c:82-125).  The generic fill
// kernel is followed by sw_cigar_kernel; the packed fill kernel calls sw_emit_cigar itself, item by item.
// ---------------------------------------------------------------------------------------------
// COHERENT: the trace was written by this very kernel (the packed fill kernel walks its own items):
// loads go to the L2 (ld.global.cg), never through the non-coherent path, which may still hold
// lines of the item that used the trace slot before.
// build-time switches of the traceback (A/B variants, scripts/build_variants.sh), measured on 47 360 cfg3 pairs,
// GCUPS in FIXED / ASIS mode (profiles/r02_sw_walk_variants.log):
//   round-1 form (two walks, one load per cell, inlined)            2481 / 2641
//   SW_ONE_WALK      one walk into a scratch list, reversed into the pool   2528 / 2658   <- default
//   SW_BATCH_FETCH   eight cells of a run per round trip            1964 / 2276   (with ONE_WALK 2185 / 2289)
//   SW_WALK_NOINLINE the walk as a real call                        2451 / 2531
// The packed kernel runs at its 96-register budget: anything that adds live registers or a call frame to the walk
// changes the register allocation of the fill loop it is inlined into and costs more than the walk takes (7 %).
#ifndef SW_BATCH_FETCH
#define SW_BATCH_FETCH 0
#endif
#ifndef SW_ONE_WALK
#define SW_ONE_WALK 1
#endif
#ifndef SW_WALK_NOINLINE
#define SW_WALK_NOINLINE 0
#endif
#if SW_WALK_NOINLINE
#define SW_WALK_ATTR __noinline__
#else
#define SW_WALK_ATTR
#endif
template <bool COHERENT>
struct sw_cell_view {
  const uint32_t * tr;
  int nsteps;
  __device__ __forceinline__ uint32_t nib (int i, int j) const   // interior cell, 1-based
  {
    int col = j - 1, band = col >> 8, lane = (col & 255) >> 3, c = col & 7;
    const uint32_t * p = tr + ((size_t) band * nsteps + (i - 1 + lane)) * 32 + lane;
    uint32_t w = COHERENT ? __ldcg (p) : __ldg (p);
    return (w >> sw_shift (c)) & 15u;
  }
  __device__ __forceinline__ uint32_t nib_or0 (int i, int j) const { return (i >= 1 && j >= 1) ? nib (i, j) : 0u; }
  // status as the CIGAR op the reference would take: 0 = M, 2 = D, 1 = I (the else branch, also for border cells)
  // len = ml / dl / il of that cell (sw.c:221-249)
  // The run lengths are walked EIGHT cells per round trip: the cells of a run lie on different 128-byte lines of the
  // trace (one line per DP row), so a walk that loads a nibble, looks at it and only then knows the next address pays
  // one L2 / HBM latency per cell — 12 000 of them per cfg3 alignment, 7 % of the whole fill + traceback time in the
  // FIXED mode.  The next seven cells of the same kind of run are loaded together with the first (independent loads,
  // wasted when the run ends earlier: runs between two sequencing errors are about ten cells long).
#if !SW_BATCH_FETCH
  __device__ void fetch (int i, int j, uint32_t * op, uint32_t * len) const
  {
    if (i < 1 || j < 1) { *op = 1; *len = 0; return; }
    uint32_t n = nib (i, j);
    if (n & 8u) {                              // M: ml = consecutive M cells up the diagonal
      uint32_t l = 0;
      while (i >= 1 && j >= 1 && (nib (i, j) & 8u)) { ++l; --i; --j; }
      *op = 0; *len = l;
    } else if (n & 4u) {                       // D: dl = 1 + D-extended flags walking up
      uint32_t l = 0;
      while (i >= 1) { ++l; if (!(nib (i, j) & 2u)) break; --i; }
      *op = 2; *len = l;
    } else {                                   // I: il = 1 + I-extended flags walking left
      uint32_t l = 0;
      while (j >= 1) { ++l; if (!(nib (i, j) & 1u)) break; --j; }
      *op = 1; *len = l;
    }
  }
#else
  __device__ void fetch (int i, int j, uint32_t * op, uint32_t * len) const
  {
    if (i < 1 || j < 1) { *op = 1; *len = 0; return; }
    uint32_t w[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) w[u] = nib_or0 (i - u, j - u);          // the cell and the diagonal behind it
    const uint32_t n = w[0];
    if (n & 8u) {                              // M: ml = consecutive M cells up the diagonal
      uint32_t l = 0;
      for (;;) {
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) if (cnt == u && (w[u] & 8u)) ++cnt;
        l += (uint32_t) cnt;
        if (cnt < 8) break;
        i -= 8; j -= 8;
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = nib_or0 (i - u, j - u);
      }
      *op = 0; *len = l;
    } else if (n & 4u) {                       // D: dl = 1 + D-extended flags walking up
      uint32_t l = 0;
      bool first = true;
      for (;;) {
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = (first && u == 0) ? n : ((i - u >= 1) ? nib (i - u, j) : 0u);
        first = false;
        int cnt = 0; bool stop = false;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (!stop && i - u >= 1) { ++cnt; if (!(w[u] & 2u)) stop = true; }
        l += (uint32_t) cnt;
        if (stop || cnt < 8) break;
        i -= 8;
      }
      *op = 2; *len = l;
    } else {                                   // I: il = 1 + I-extended flags walking left
      uint32_t l = 0;
      bool first = true;
      for (;;) {
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = (first && u == 0) ? n : ((j - u >= 1) ? nib (i, j - u) : 0u);
        first = false;
        int cnt = 0; bool stop = false;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (!stop && j - u >= 1) { ++cnt; if (!(w[u] & 1u)) stop = true; }
        l += (uint32_t) cnt;
        if (stop || cnt < 8) break;
        j -= 8;
      }
      *op = 1; *len = l;
    }
  }
#endif
};

// WRITE 0: count only; 1: write reversed into out[0 .. n_total) (cigar_reverse); 2: write in walking order into out
// (a scratch list the caller reverses into the pool once the count is known: ONE walk instead of two)
template <int WRITE, class VIEW>
__device__ int sw_walk (const VIEW & cv, int mode, int qlen, int tlen, const sw_end & e,
                        uint32_t * out, int n_total, int * align_off, int * softclip)
{
  int n = 0;
  auto emit = [&] (uint32_t v) { if (WRITE == 1) out[n_total - 1 - n] = v; else if (WRITE == 2) out[n] = v; ++n; };
  int bt_t = e.bt_tidx, bt_q = e.bt_qidx;
  uint32_t seg = (uint32_t) e.seg_len, pre = 0, op = 1, len = 0;
  const int strategy = c_sw.strategy;
  *softclip = 0;
  if (seg > 0 && strategy == GCG_SWOS_SOFTCLIP) { emit ((seg << 4) | 4u); seg = 0; *softclip = 1; }
  cv.fetch (bt_t, bt_q, &op, &len);
  bool inited = false;
  do {
    if (mode == GCG_SW_FIXED && inited) cv.fetch (bt_t, bt_q, &op, &len);
    if (op == 0) { bt_t -= (int) len; bt_q -= (int) len; }
    else if (op == 2) bt_t -= (int) len;
    else bt_q -= (int) len;
    if (inited && op != pre) { emit ((seg << 4) | pre); seg = 0; }
    seg += len;
    pre = op;
    inited = true;
    if (len == 0) break;      // border end cell: an index is already 0, the reference leaves the loop as well
  } while (bt_t > 0 && bt_q > 0);
  emit ((seg << 4) | pre);
  if (strategy == GCG_SWOS_SOFTCLIP) {
    if (bt_q > 0) emit ((((uint32_t) bt_q) << 4) | 4u);
    *align_off = bt_t;
  } else {
    if (bt_t > 0) emit ((((uint32_t) bt_t) << 4) | 2u);
    if (bt_q > 0) emit ((((uint32_t) bt_q) << 4) | 1u);
    *align_off = 0;
  }
  return n;
}

// one alignment: count the CIGAR operations, reserve them in the pool, write them (a reservation that
// ends beyond pool_cap is left unwritten: the host grows the pool and has the alignment done again)
// scratch: room for qlen + tlen + 4 operations (every step of the walk consumes at least one row or column), or NULL
template <bool COHERENT>
__device__ SW_WALK_ATTR void sw_emit_cigar (const uint32_t * tr, int qlen, int tlen, int pair, const sw_end & e, int mode,
                               uint32_t * __restrict__ pool, unsigned long long pool_cap,
                               unsigned long long * __restrict__ cursor, gcg_sw_result * __restrict__ results,
                               uint32_t * __restrict__ scratch)
{
  sw_cell_view<COHERENT> cv;
  cv.tr = tr;
  cv.nsteps = tlen + 31;
  int off = 0, sc = 0;
#if !SW_ONE_WALK
  scratch = nullptr;
#endif
  const int n = scratch ? sw_walk<2> (cv, mode, qlen, tlen, e, scratch, 0, &off, &sc)
                        : sw_walk<0> (cv, mode, qlen, tlen, e, nullptr, 0, &off, &sc);
  unsigned long long at = atomicAdd (cursor, (unsigned long long) n);
  gcg_sw_result r;
  r.score = e.score; r.alignment_offset = off; r.has_softclip = sc;
  r.bt_tidx = e.bt_tidx; r.bt_qidx = e.bt_qidx; r.n_cigar = n; r.cigar_off = (long long) at;
  results[pair] = r;
  if (at + (unsigned long long) n <= pool_cap) {
    if (scratch) {                                  // the list in walking order -> the pool in CIGAR order
      for (int i = 0; i < n; ++i) pool[at + (unsigned long long) (n - 1 - i)] = COHERENT ? __ldcg (scratch + i) : scratch[i];
    } else sw_walk<1> (cv, mode, qlen, tlen, e, pool + at, n, &off, &sc);
  }
}

__global__ void __launch_bounds__ (128)
sw_cigar_kernel (const sw_task * __restrict__ tasks, int n_tasks, const uint32_t * __restrict__ trace,
                 const sw_end * __restrict__ ends, int mode, uint32_t * __restrict__ pool, unsigned long long pool_cap,
                 unsigned long long * __restrict__ cursor, gcg_sw_result * __restrict__ results, uint32_t * __restrict__ ops)
{
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_tasks) return;
  const sw_task tk = tasks[idx];
  // (scratch list of a pair: behind its edge scores' offset, four more words per pair)
  sw_emit_cigar<false> (trace + tk.trace_off, tk.qlen, tk.tlen, tk.pair, ends[tk.pair], mode, pool, pool_cap, cursor, results,
                        ops ? ops + tk.edge_off + 4LL * tk.pair : nullptr);
}


// ---------------------------------------------------------------------------------------------
// K7 packed: two alignments per warp in s16x2 halves.  Shared memory per warp (rows = tmax):
//   bH[rows], bI[rows]  last column of the previous band (both alignments packed)
//   tsel[rows]          PRMT selector for the two target symbols of that row
// ---------------------------------------------------------------------------------------------
#define SW_NEG16V (-2040)                     // "minus infinity" in score units for the packed kernel
#define SW_PACKED_MAXROWS 2040

// PRMT in its default mode: selector nibble bit 3 replicates the sign of the selected byte
// (__byte_perm masks that bit away, so the instruction is issued directly)
__device__ __forceinline__ uint32_t prmt (uint32_t a, uint32_t b, uint32_t sel)
{
  uint32_t d;
  asm ("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

__device__ __forceinline__ uint32_t pack16 (int v) { return ((uint32_t) v & 0xFFFFu) * 0x10001u; }
__device__ __forceinline__ int half_lo (uint32_t x) { return (int) (short) (x & 0xFFFFu); }
__device__ __forceinline__ int half_hi (uint32_t x) { return ((int) x) >> 16; }
// The packed kernel carries its values biased by 0x8000 per half (unsigned order == signed order of
// the scores).  With both halves non-negative a packed add of two small negative constants is one
// ordinary 32-bit add — the borrow out of the low half is always the same and is folded into the
// constant (c * 65537) — which lets the two standalone adds per cell run as IMAD on the FMA pipe
// instead of VIADD.16x2 on the integer pipe the VIADDMNMX / VIMNMX / PRMT / LOP3 work saturates.
#define SW_BIAS2 0x80008000u
__device__ __forceinline__ uint32_t pack16b (int v) { return pack16 (v) ^ SW_BIAS2; }
__device__ __forceinline__ int bhalf_lo (uint32_t x) { return (int) (x & 0xFFFFu) - 32768; }
__device__ __forceinline__ int bhalf_hi (uint32_t x) { return (int) (x >> 16) - 32768; }

// A task record re-read from global memory at every use (asm volatile: the compiler may neither
// hoist the load out of the band loop nor keep the fields in registers across the step loop).  The
// packed kernel runs at its 96-register budget; with the records loaded once per item, their
// pointers and lengths (about twenty registers) stayed live through the hot loop and the edge code
// alone cost it 12 % (measured by deleting the block).  48 bytes = three 16-byte loads, L2 hits.
__device__ __forceinline__ sw_task ld_task (const sw_task * p)
{
  static_assert (sizeof (sw_task) == 48, "sw_task is read as three 16-byte words");
  unsigned long long w[6];
  asm volatile ("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(w[0]), "=l"(w[1]) : "l"(p));
  asm volatile ("ld.global.cg.v2.u64 {%0,%1}, [%2+16];" : "=l"(w[2]), "=l"(w[3]) : "l"(p));
  asm volatile ("ld.global.cg.v2.u64 {%0,%1}, [%2+32];" : "=l"(w[4]), "=l"(w[5]) : "l"(p));
  sw_task t;
  t.qoff = (long long) w[0]; t.toff = (long long) w[1];
  t.qlen = (int) (uint32_t) w[2]; t.tlen = (int) (uint32_t) (w[2] >> 32);
  t.trace_off = w[3];
  t.edge_off = (long long) w[4];
  t.pair = (int) (uint32_t) w[5]; t.pad = 0;
  return t;
}

struct sw_packed_consts { uint32_t pk_do, pk_io, pk_de32, pk_ie32, one; };
struct sw_cigar_args { int mode; uint32_t * pool; unsigned long long pool_cap; unsigned long long * cursor; gcg_sw_result * results;
                       uint32_t * ops; unsigned long long ops_stride; };     // packed kernel: one scratch list per (block, walking lane)

template <int MINB>
__global__ void __launch_bounds__ (32, MINB)
sw_fill_packed_kernel (const uint8_t * __restrict__ qry, const uint8_t * __restrict__ tgt,
                       const sw_task * __restrict__ tasks, const int2 * __restrict__ items, int n_items,
                       int * __restrict__ counter, uint32_t * __restrict__ trace_slots, unsigned long long slot_words,
                       int * __restrict__ edges, uint2 * __restrict__ bound, int rows_cap, sw_end * __restrict__ ends,
                       const sw_packed_consts gk, const sw_cigar_args cg)
{
  // Persistent warps: every warp owns one trace slot (room for its largest item) and pulls items until
  // the list is empty; it walks the traceback of an item (sw_emit_cigar) right after filling it, so the
  // slot is free again and a batch of any size runs as ONE launch — no wave boundary at which the
  // whole GPU drains (measured on cfg3 shapes: the drain of a one-item-per-warp launch costs 14 %).
  // gk: the gap constants and `one` == 1 as kernel arguments: they reach the step loop as uniform-register
  // operands (no LDC per step, no registers), and x * one + y stays an IMAD (FMA pipe).
  // shared: one PRMT selector per target row.  The last column of the previous band (H, I of both
  // alignments, 8 bytes per row) lives in an L2-resident scratch: lane 31 parks it row by row, all
  // 32 lanes fetch the next 32 rows in one coalesced load a full 32 steps before lane 0 needs them.
  extern __shared__ unsigned short tsel[];
  const int lane = threadIdx.x;
  uint2 * bnd = bound + (size_t) blockIdx.x * rows_cap;
  uint32_t * const trace = trace_slots + (size_t) blockIdx.x * slot_words;      // task.trace_off is relative to the slot
#define cDO gk.pk_do
#define cIO gk.pk_io
  const uint32_t NEG2 = pack16b (16 * SW_NEG16V);
  const uint32_t cDE32 = gk.pk_de32, cIE32 = gk.pk_ie32, one = gk.one, sixteen = gk.one << 4;
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd (counter, 1);
    item = __shfl_sync (0xffffffffu, item, 0);
    if (item >= n_items) break;
    const int2 it = items[item];
    const bool has_b = it.y >= 0;
    const sw_task * pta = tasks + it.x, * ptb = tasks + (has_b ? it.y : it.x);
    int TL, nbands;
    {
      // per-row PRMT selectors: byte0 <- profA[tA], byte1 <- its sign, byte2 <- profB[tB], byte3 <- its sign
      const sw_task ta = ld_task (pta), tb = ld_task (ptb);
      TL = max (ta.tlen, tb.tlen);                                        // both > 0 (host guarantees)
      nbands = (max (ta.qlen, tb.qlen) + SW_BAND - 1) / SW_BAND;
      const uint8_t * tga = tgt + ta.toff, * tgb = tgt + tb.toff;
      __syncwarp ();
      for (int r = lane; r < TL; r += 32) {
        uint32_t sa = r < ta.tlen ? tga[r] : 0, sb = r < tb.tlen ? tgb[r] : 0;
        tsel[r] = (unsigned short) (sa | ((sa | 8u) << 4) | ((4u + sb) << 8) | ((12u + sb) << 12));
      }
      __syncwarp ();
    }
    const int nsteps = TL + 31;
    for (int band = 0; band < nbands; ++band) {
      const int j0 = band * SW_BAND + lane * SW_COLS;
      uint32_t PA[SW_COLS], PB[SW_COLS], Hup[SW_COLS], Dup[SW_COLS];
      // everything that does not change inside the step loop, from the task records read afresh
      uint32_t * pa, * pb;
      bool wrA, wrB, no_owner;
      int edgeA, edgeB, tlenA, tlenB;
      {
        const sw_task ta = ld_task (pta), tb = ld_task (ptb);
        const uint8_t * qa = qry + ta.qoff, * qb = qry + tb.qoff;
#pragma unroll
        for (int c = 0; c < SW_COLS; ++c) {
          PA[c] = c_sw.prof4[j0 + c < ta.qlen ? qa[j0 + c] : 0];
          PB[c] = c_sw.prof4[j0 + c < tb.qlen ? qb[j0 + c] : 0];
          Hup[c] = pack16b (16 * sw_brow (j0 + c + 1));
          Dup[c] = NEG2;
        }
        const int nba = (ta.qlen + SW_BAND - 1) / SW_BAND, nbb = (tb.qlen + SW_BAND - 1) / SW_BAND;
        wrA = band < nba; wrB = has_b && band < nbb;
        pa = trace + ta.trace_off + ((size_t) band * (ta.tlen + 31)) * 32 + lane;
        pb = trace + tb.trace_off + ((size_t) band * (tb.tlen + 31)) * 32 + lane;
        const bool ownA = j0 < ta.qlen && ta.qlen <= j0 + SW_COLS, ownB = has_b && j0 < tb.qlen && tb.qlen <= j0 + SW_COLS;
        edgeA = ownA ? 1 : ta.tlen; edgeB = ownB ? 1 : (has_b ? tb.tlen : 0x7fffffff);   // rows >= edge need the slow path
        tlenA = ta.tlen; tlenB = tb.tlen;
        no_owner = band != nba - 1 && (!has_b || band != nbb - 1);
      }
      uint32_t prevInH = pack16b (16 * sw_brow (j0)), outH = SW_BIAS2, outI = NEG2;
      // boundary prefetch registers: rows [32k, 32k+32) of the previous band, one row per lane
      uint2 pf_cur = make_uint2 (SW_BIAS2, NEG2), pf_nxt = make_uint2 (SW_BIAS2, NEG2);
      if (band > 0) {
        if (lane < TL) pf_cur = __ldcg (bnd + lane);
        if (32 + lane < TL) pf_nxt = __ldcg (bnd + 32 + lane);
      }
      const bool park = lane == 31 && band + 1 < nbands;
      const uint32_t col0step = pack16 (-16 * c_sw.b_del_e * (c_sw.border_kind ? 1 : 0));
      uint32_t col0 = pack16b (16 * sw_bcol (1));                  // 16 * border score of column 0, row `step + 1`
      // The eight cells of this lane on one row.  Integer-pipe work per cell pair: PRMT (score),
      // 3 x VIADDMNMX.U16x2, VIMNMX.U16x2, 4 x LOP3; the two gap-extension adds, the tag of H
      // (H - Hc) and the two nibble accumulators are IMADs (x * one + y) on the FMA pipe.
      auto row_cells = [&] (const uint32_t inH, const uint32_t inI, const uint32_t sel, uint32_t & wA, uint32_t & wB) {
        uint32_t Hd = prevInH, Hl = inH, Il = inI, aH = 0, aT = 0, bH = 0, bT = 0;
#pragma unroll
        for (int c = 0; c < SW_COLS; ++c) {
          const uint32_t s = prmt (PA[c], PB[c], sel);
          const uint32_t dE = (Dup[c] | 0x00020002u) * one + cDE32;
          const uint32_t D = __viaddmax_u16x2 (Hup[c], cDO, dE);
          const uint32_t iE = (Il | 0x00010001u) * one + cIE32;
          const uint32_t I = __viaddmax_u16x2 (Hl, cIO, iE);
          const uint32_t g = __vmaxu2 (D, I);
          const uint32_t H = __viaddmax_u16x2 (Hd, s, g);
          const uint32_t Hc = H & 0xFFF0FFF0u;
          const uint32_t tH = H * one - Hc;                        // tag of H: 8 = M, 4|2 = D, 0|1 = I
          const uint32_t tT = (D | I) & 0x00030003u;               // bit 1 D extended, bit 0 I extended (covers tH & 3)
          if (c < 4) { aH = aH * sixteen + tH; aT = aT * sixteen + tT; }
          else       { bH = bH * sixteen + tH; bT = bT * sixteen + tT; }
          Hd = Hup[c]; Hup[c] = Hc; Dup[c] = D; Hl = Hc; Il = I;
        }
        prevInH = inH;
        outH = Hl; outI = Il;
        wA = aH | aT; wB = bH | bT;
      };
      auto rotate_prefetch = [&] (const int step) {                // after step == 31 (mod 32): rows of the chunk after the next one
        pf_cur = pf_nxt;
        const int r = step + 33 + lane;
        if (r < TL) pf_nxt = __ldcg (bnd + r);
      };
      // A step with every test in place: ramps, last rows, last columns.
      auto edge_step = [&] (const int step) {
        uint32_t inH = __shfl_up_sync (0xffffffffu, outH, 1), inI = __shfl_up_sync (0xffffffffu, outI, 1);
        const int i = step - lane + 1;
        if (band > 0) {
          // lane 0 is on row step+1: its left neighbour is row `step` of the parked column
          uint32_t bh = __shfl_sync (0xffffffffu, pf_cur.x, step & 31), bi = __shfl_sync (0xffffffffu, pf_cur.y, step & 31);
          if (lane == 0) { inH = bh; inI = bi; }
          if ((step & 31) == 31) rotate_prefetch (step);
        } else {
          if (lane == 0) { inH = col0; inI = NEG2; }
          col0 = __vadd2 (col0, col0step);
        }
        if (i >= 1 && i <= TL) {
          uint32_t wA, wB;
          row_cells (inH, inI, tsel[i - 1], wA, wB);
          if (wrA && i <= tlenA) pa[(uint32_t) step * 32u] = __byte_perm (wA, wB, 0x5410);
          if (wrB && i <= tlenB) pb[(uint32_t) step * 32u] = __byte_perm (wA, wB, 0x7632);
          // park the last column for the next band.  Row r is written at step r+31 and was fetched
          // (for this band) no later than step r-1, so reusing the buffer in place is safe.
          if (park) __stcg (bnd + (i - 1), make_uint2 (outH, outI));
          if (i >= edgeA || i >= edgeB) {
            // rare: this lane holds the last column of an alignment, or is on its last row
            const sw_task ta = ld_task (pta), tb = ld_task (ptb);
            int * ea = edges + ta.edge_off, * eb = edges + tb.edge_off;
            const bool ownA = j0 < ta.qlen && ta.qlen <= j0 + SW_COLS, ownB = has_b && j0 < tb.qlen && tb.qlen <= j0 + SW_COLS;
            if (ownA && i <= ta.tlen) {
              uint32_t v = 0; const int cs = (ta.qlen - 1) % SW_COLS;
#pragma unroll
              for (int c = 0; c < SW_COLS; ++c) if (c == cs) v = Hup[c];
              __stcg (ea + (i - 1), bhalf_lo (v) >> 4);
            }
            if (ownB && i <= tb.tlen) {
              uint32_t v = 0; const int cs = (tb.qlen - 1) % SW_COLS;
#pragma unroll
              for (int c = 0; c < SW_COLS; ++c) if (c == cs) v = Hup[c];
              __stcg (eb + (i - 1), bhalf_hi (v) >> 4);
            }
            if (i == ta.tlen) {
#pragma unroll
              for (int c = 0; c < SW_COLS; ++c) if (j0 + c < ta.qlen) __stcg (ea + ta.tlen + j0 + c, bhalf_lo (Hup[c]) >> 4);
            }
            if (has_b && i == tb.tlen) {
#pragma unroll
              for (int c = 0; c < SW_COLS; ++c) if (j0 + c < tb.qlen) __stcg (eb + tb.tlen + j0 + c, bhalf_hi (Hup[c]) >> 4);
            }
          }
        }
      };
      // Steady steps — all 32 lanes on rows 1 .. min(tlen)-1 of a band in which no lane owns a last
      // column — run in chunks of 32 (one prefetch register set per chunk) without row-range tests,
      // store predicates on the row or edge code: the eight column recurrences of a step form one
      // basic block that the scheduler interleaves freely.
      const int steady_end = no_owner ? min (nsteps, min (tlenA, tlenB) - 1) : 0;      // steady steps: [32, steady_end)
      int step = 0;
      while (step < nsteps) {
        if (step >= 32 && step + 32 <= steady_end) {
          const unsigned short * ts = tsel + (step - lane);                                  // row i-1 = step - lane
          uint32_t * qa_ = pa + (size_t) step * 32u, * qb_ = pb + (size_t) step * 32u;
          uint2 * pk_ = bnd + (step - lane);
          if (band > 0) {
#pragma unroll 1
            for (int k = 0; k < 32; ++k) {
              uint32_t inH = __shfl_up_sync (0xffffffffu, outH, 1), inI = __shfl_up_sync (0xffffffffu, outI, 1);
              const uint32_t bh = __shfl_sync (0xffffffffu, pf_cur.x, k), bi = __shfl_sync (0xffffffffu, pf_cur.y, k);
              if (lane == 0) { inH = bh; inI = bi; }
              uint32_t wA, wB;
              row_cells (inH, inI, ts[k], wA, wB);
              if (wrA) qa_[k * 32] = __byte_perm (wA, wB, 0x5410);
              if (wrB) qb_[k * 32] = __byte_perm (wA, wB, 0x7632);
              if (park) __stcg (pk_ + k, make_uint2 (outH, outI));
            }
            rotate_prefetch (step + 31);
          } else {
#pragma unroll 1
            for (int k = 0; k < 32; ++k) {
              uint32_t inH = __shfl_up_sync (0xffffffffu, outH, 1), inI = __shfl_up_sync (0xffffffffu, outI, 1);
              if (lane == 0) { inH = col0; inI = NEG2; }
              col0 = __vadd2 (col0, col0step);
              uint32_t wA, wB;
              row_cells (inH, inI, ts[k], wA, wB);
              if (wrA) qa_[k * 32] = __byte_perm (wA, wB, 0x5410);
              if (wrB) qb_[k * 32] = __byte_perm (wA, wB, 0x7632);
              if (park) __stcg (pk_ + k, make_uint2 (outH, outI));
            }
          }
          step += 32;
        } else {
          edge_step (step);
          ++step;
        }
      }
      __syncwarp ();
    }
    __syncwarp ();
    {
      // end cells (all lanes), then lane 0 walks alignment A and lane 1 alignment B
      const sw_task ta = ld_task (pta), tb = ld_task (ptb);
      const sw_end ea = sw_pick_end (edges + ta.edge_off, ta.qlen, ta.tlen, lane, ends + ta.pair);
      sw_end eb = ea;
      if (has_b) eb = sw_pick_end (edges + tb.edge_off, tb.qlen, tb.tlen, lane, ends + tb.pair);
      __syncwarp ();                                 // the trace stores of all lanes are visible to the walkers
      uint32_t * const ops = cg.ops ? cg.ops + ((size_t) blockIdx.x * 2 + (size_t) lane) * cg.ops_stride : nullptr;
      if (lane == 0) sw_emit_cigar<true> (trace + ta.trace_off, ta.qlen, ta.tlen, ta.pair, ea, cg.mode, cg.pool, cg.pool_cap, cg.cursor, cg.results, ops);
      if (lane == 1 && has_b) sw_emit_cigar<true> (trace + tb.trace_off, tb.qlen, tb.tlen, tb.pair, eb, cg.mode, cg.pool, cg.pool_cap, cg.cursor, cg.results, ops);
    }
    __syncwarp ();                                   // the slot is rewritten by the next item only after the walks
  }
#undef cDO
#undef cIO
}

// per-pair maximum symbol (decides whether a pair may use the 4-letter packed kernel)
__global__ void __launch_bounds__ (128)
sw_maxsym_kernel (const uint8_t * __restrict__ qry, const long long * __restrict__ qoff,
                  const uint8_t * __restrict__ tgt, const long long * __restrict__ toff, long long n, int * __restrict__ maxsym)
{
  __shared__ int s[4];
  for (long long p = blockIdx.x; p < n; p += gridDim.x) {
    int m = 0;
    for (long long i = qoff[p] + threadIdx.x; i < qoff[p + 1]; i += blockDim.x) m = max (m, (int) qry[i]);
    for (long long i = toff[p] + threadIdx.x; i < toff[p + 1]; i += blockDim.x) m = max (m, (int) tgt[i]);
    m = __reduce_max_sync (0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads ();
    if (threadIdx.x == 0) maxsym[p] = max (max (s[0], s[1]), max (s[2], s[3]));
    __syncthreads ();
  }
}

// =============================================================================================
// host side
// =============================================================================================
// resident warps per SM of the packed kernel = its register budget (GCG_SW_OCC: 20 | 24).  Measured on
// cfg3-shaped waves: 20 warps at 96 registers 62.9 ms; 24 warps at 80 registers (76 bytes spilled) 69.8 ms
// for the same 2960-warp wave — and a wave that fills 24 warps per SM needs 74 GB of trace.
typedef void (* sw_packed_fn_t) (const uint8_t *, const uint8_t *, const sw_task *, const int2 *, int, int *, uint32_t *, unsigned long long, int *, uint2 *,
                                 int, sw_end *, sw_packed_consts, sw_cigar_args);
static sw_packed_fn_t sw_packed_fn ()
{
  static int occ = -1;
  if (occ < 0) { const char * e = getenv ("GCG_SW_OCC"); occ = e ? atoi (e) : 20; }
  switch (occ) {
    case 24: return sw_fill_packed_kernel<24>;
    default: return sw_fill_packed_kernel<20>;
  }
}

// device scratch that survives between align calls on the same batch (cudaMalloc of tens of GB of
// trace is far more expensive than the kernels that fill it)
// (all of it comes from the context's stream-ordered pool, which never trims: a batch that is
// freed hands its trace buffer to the next batch without a trip to the driver)
struct sw_devbuf {
  void * p = nullptr;
  size_t cap = 0;
  cudaError_t reserve (gcg_ctx * ctx, size_t bytes)
  {
    if (bytes <= cap) return cudaSuccess;
    gcg_dfree (ctx, p);
    p = nullptr; cap = 0;
    cudaError_t e = gcg_dmalloc (ctx, &p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release (gcg_ctx * ctx) { gcg_dfree (ctx, p); p = nullptr; cap = 0; }
};

struct gcg_swbatch {
  gcg_ctx * ctx = nullptr;
  sw_devbuf s_trace, s_edges, s_tasks, s_wave_tasks, s_items, s_counter, s_bound, s_pbound, s_ops, s_gops;
  int64_t n = 0;
  std::vector<long long> qoff, toff;
  uint8_t * d_qry = nullptr, * d_tgt = nullptr;
  long long * d_qoff = nullptr, * d_toff = nullptr;
  int * d_maxsym = nullptr;
  std::vector<int> maxsym;
  int64_t cells = 0;
  // results of the last align
  gcg_sw_result * d_results = nullptr;
  sw_end * d_ends = nullptr;
  uint32_t * d_pool = nullptr;
  unsigned long long pool_cap = 0, pool_used = 0;
  int64_t n_packed = 0, n_generic = 0;
  bool aligned = false;
};

extern "C" void gcg_swbatch_free (gcg_swbatch * b)
{
  if (!b) return;
  gcg_ctx * ctx = b->ctx;
  gcg_dfree (ctx, b->d_qry); gcg_dfree (ctx, b->d_tgt); gcg_dfree (ctx, b->d_qoff); gcg_dfree (ctx, b->d_toff); gcg_dfree (ctx, b->d_maxsym);
  gcg_dfree (ctx, b->d_results); gcg_dfree (ctx, b->d_ends); gcg_dfree (ctx, b->d_pool);
  b->s_trace.release (ctx); b->s_edges.release (ctx); b->s_tasks.release (ctx); b->s_wave_tasks.release (ctx);
  b->s_items.release (ctx); b->s_counter.release (ctx); b->s_bound.release (ctx); b->s_pbound.release (ctx);
  b->s_ops.release (ctx); b->s_gops.release (ctx);
  delete b;
}

extern "C" int64_t gcg_swbatch_cells (const gcg_swbatch * b) { return b ? b->cells : 0; }

extern "C" int gcg_swbatch_path_counts (const gcg_swbatch * b, int64_t counts[2])
{
  GCG_CHECK (b && counts, GCG_EINVAL, "gcg_swbatch_path_counts: bad argument");
  counts[0] = b->n_packed; counts[1] = b->n_generic;
  return GCG_OK;
}

static int h2d_chunked (gcg_ctx * ctx, void * dst, const void * src, size_t bytes)
{
  // a source in pinned memory (gcg_host_alloc, cudaHostAlloc, cudaHostRegister) goes over PCIe as it is
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes (&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost) {
    GCG_CUDA (cudaMemcpyAsync (dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return GCG_OK;
  }
  cudaGetLastError ();
  // pageable source -> pinned ring -> device, double buffered
  int rc = gcg_stage_reserve (ctx);
  if (rc) return rc;
  size_t off = 0;
  while (off < bytes) {
    int slot = ctx->stage.next;
    ctx->stage.next ^= 1;
    size_t nb = std::min (ctx->stage.cap, bytes - off);
    if (ctx->stage.busy[slot]) { GCG_CUDA (cudaEventSynchronize (ctx->stage.ev[slot])); ctx->stage.busy[slot] = false; }
    gcg_par_memcpy (ctx, ctx->stage.h[slot], (const char *) src + off, nb);
    GCG_CUDA (cudaMemcpyAsync ((char *) dst + off, ctx->stage.h[slot], nb, cudaMemcpyHostToDevice, ctx->stream));
    GCG_CUDA (cudaEventRecord (ctx->stage.ev[slot], ctx->stream));
    ctx->stage.busy[slot] = true;
    off += nb;
  }
  return GCG_OK;
}

extern "C" int gcg_swbatch_upload (gcg_ctx * ctx, const char * qry, const int64_t * qoff, const char * tgt,
                                   const int64_t * toff, int64_t n, gcg_swbatch ** out)
{
  GCG_CHECK (ctx && out && qoff && toff && n >= 0, GCG_EINVAL, "gcg_swbatch_upload: bad argument");
  GCG_CHECK (n < 0x7FFFFFFF, GCG_ERANGE, "gcg_swbatch_upload: too many pairs");
  GCG_CUDA (cudaSetDevice (ctx->device));
  gcg_swbatch * b = new gcg_swbatch ();
  b->ctx = ctx; b->n = n;
  b->qoff.assign (qoff, qoff + n + 1);
  b->toff.assign (toff, toff + n + 1);
  for (int64_t p = 0; p < n; ++p) {
    long long ql = qoff[p + 1] - qoff[p], tl = toff[p + 1] - toff[p];
    if (ql < 0 || tl < 0 || ql > (1 << 24) || tl > (1 << 24)) {
      gcg_set_error ("gcg_swbatch_upload: pair %lld has lengths %lld x %lld outside [0, 2^24]", (long long) p, ql, tl);
      delete b; return GCG_ERANGE;
    }
    b->cells += ql * tl;
  }
  size_t qb = (size_t) qoff[n], tb = (size_t) toff[n];
  int rc = GCG_OK;
  cudaError_t e;
  if ((e = gcg_dmalloc (ctx, &b->d_qry, std::max<size_t> (qb, 16))) != cudaSuccess || (e = gcg_dmalloc (ctx, &b->d_tgt, std::max<size_t> (tb, 16))) != cudaSuccess ||
      (e = gcg_dmalloc (ctx, &b->d_qoff, (size_t) (n + 1) * 8)) != cudaSuccess || (e = gcg_dmalloc (ctx, &b->d_toff, (size_t) (n + 1) * 8)) != cudaSuccess ||
      (e = gcg_dmalloc (ctx, &b->d_maxsym, (size_t) std::max<int64_t> (n, 1) * 4)) != cudaSuccess ||
      (e = gcg_dmalloc (ctx, &b->d_results, (size_t) std::max<int64_t> (n, 1) * sizeof (gcg_sw_result))) != cudaSuccess ||
      (e = gcg_dmalloc (ctx, &b->d_ends, (size_t) std::max<int64_t> (n, 1) * sizeof (sw_end))) != cudaSuccess) {
    gcg_set_error ("gcg_swbatch_upload: cudaMalloc failed: %s", cudaGetErrorString (e));
    gcg_swbatch_free (b);
    return GCG_ENOMEM;
  }
  if (qb) rc = h2d_chunked (ctx, b->d_qry, qry, qb);
  if (!rc && tb) rc = h2d_chunked (ctx, b->d_tgt, tgt, tb);
  if (!rc) {
    cudaMemcpyAsync (b->d_qoff, b->qoff.data (), (size_t) (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync (b->d_toff, b->toff.data (), (size_t) (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream);
    b->maxsym.assign ((size_t) n, 0);
    if (n > 0) {
      { gcg_kscope ks (ctx, "sw_maxsym");
        int grid = (int) std::min<int64_t> (n, (int64_t) ctx->sm_count * 16);
        sw_maxsym_kernel<<<grid, 128, 0, ctx->stream>>> (b->d_qry, b->d_qoff, b->d_tgt, b->d_toff, n, b->d_maxsym); }
      cudaMemcpyAsync (b->maxsym.data (), b->d_maxsym, (size_t) n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (cudaStreamSynchronize (ctx->stream) != cudaSuccess || cudaGetLastError () != cudaSuccess) {
      gcg_set_error ("gcg_swbatch_upload: %s", cudaGetErrorString (cudaGetLastError ()));
      rc = GCG_ECUDA;
    }
  }
  if (rc) { gcg_swbatch_free (b); return rc; }
  *out = b;
  return GCG_OK;
}

// value bounds (score units) of everything the DP can produce for a qlen x tlen alignment
static void sw_bounds (const gcg_sw_params * P, long long qlen, long long tlen, long long * lo, long long * hi)
{
  long long mx = 0, mn = 0;
  for (int i = 0; i < P->type_c * P->type_c; ++i) { mx = std::max<long long> (mx, P->mat[i]); mn = std::min<long long> (mn, P->mat[i]); }
  long long row_min = 0, col_min = 0;     // most negative border scores
  if (P->border_kind) {
    row_min = std::min<long long> (0, -(long long) P->b_ins_o - (qlen - 1) * (long long) P->b_ins_e);
    col_min = std::min<long long> (0, -(long long) P->b_del_o - (tlen - 1) * (long long) P->b_del_e);
    if (P->b_ins_e < 0) row_min = std::min<long long> (row_min, -(long long) P->b_ins_o);
    if (P->b_del_e < 0) col_min = std::min<long long> (col_min, -(long long) P->b_del_o);
  }
  // H(i,j) >= border_row(j) - del_o - (i-1) del_e  and  >= border_col(i) - ins_o - (j-1) ins_e
  long long l1 = row_min - P->del_o - std::max<long long> (0, tlen - 1) * P->del_e;
  long long l2 = col_min - P->ins_o - std::max<long long> (0, qlen - 1) * P->ins_e;
  long long hmin = std::max (l1, l2);
  // every candidate the kernels form from an H >= hmin: H - open, (H - open) - extend, H + mat
  long long omax = std::max (P->del_o, P->ins_o), emax = std::max (P->del_e, P->ins_e);
  *lo = hmin + std::min (mn, -(omax + emax)) - 1;
  // H <= best border (<= 0 when penalties are non-negative) + min(q,t) * max score
  long long bmax = 0;
  if (P->border_kind) {
    if (P->b_ins_o < 0 || P->b_ins_e < 0) bmax = std::max (bmax, -(long long) P->b_ins_o + std::max<long long> (0, -(long long) P->b_ins_e) * qlen);
    if (P->b_del_o < 0 || P->b_del_e < 0) bmax = std::max (bmax, -(long long) P->b_del_o + std::max<long long> (0, -(long long) P->b_del_e) * tlen);
  }
  *hi = bmax + std::min (qlen, tlen) * mx + 1;
}

static bool sw_params_packable (const gcg_sw_params * P)
{
  if (P->type_c < 1) return false;
  if (P->del_o < 0 || P->del_e < 0 || P->ins_o < 0 || P->ins_e < 0) return false;   // lower bound derivation needs penalties >= 0
  if (P->del_e > 8 || P->ins_e > 8 || P->del_o > 64 || P->ins_o > 64) return false;
  int n = std::min (P->type_c, 4);
  for (int q = 0; q < n; ++q)
    for (int t = 0; t < n; ++t) {
      int v = 16 * P->mat[q * P->type_c + t] + 8;
      if (v < -128 || v > 127) return false;
    }
  return true;
}

extern "C" int gcg_swbatch_align (gcg_ctx * ctx, gcg_swbatch * b, const gcg_sw_params * P, int mode)
{
  GCG_CHECK (ctx && b && P, GCG_EINVAL, "gcg_swbatch_align: bad argument");
  GCG_CHECK (mode == GCG_SW_ASIS || mode == GCG_SW_FIXED, GCG_EINVAL, "gcg_swbatch_align: mode %d", mode);
  GCG_CHECK (P->type_c >= 1 && P->type_c <= 8, GCG_ERANGE, "gcg_swbatch_align: type_c=%d outside [1,8]", P->type_c);
  GCG_CHECK (P->strategy >= 0 && P->strategy <= 3, GCG_EINVAL, "gcg_swbatch_align: overhang strategy %d", P->strategy);
  GCG_CUDA (cudaSetDevice (ctx->device));
  const int64_t n = b->n;
  b->n_packed = b->n_generic = 0;
  b->pool_used = 0;
  b->aligned = false;
  if (n == 0) { b->aligned = true; return GCG_OK; }
  for (int64_t p = 0; p < n; ++p)
    GCG_CHECK (b->maxsym[(size_t) p] < P->type_c, GCG_EINVAL, "gcg_swbatch_align: pair %lld holds symbol %d >= type_c %d (sw.c:216 would index outside mat)",
               (long long) p, b->maxsym[(size_t) p], P->type_c);

  // ---- constants
  sw_consts hc;
  memset (&hc, 0, sizeof hc);
  hc.type_c = P->type_c; hc.del_o = P->del_o; hc.del_e = P->del_e; hc.ins_o = P->ins_o; hc.ins_e = P->ins_e;
  hc.strategy = P->strategy; hc.border_kind = P->border_kind;
  hc.b_del_o = P->b_del_o; hc.b_del_e = P->b_del_e; hc.b_ins_o = P->b_ins_o; hc.b_ins_e = P->b_ins_e;
  for (int i = 0; i < P->type_c * P->type_c; ++i) hc.mat16[i] = 16 * P->mat[i] + 8;
  const bool packable = sw_params_packable (P) && getenv ("GCG_SW_FORCE_GENERIC") == nullptr;
  if (packable) {
    int nn = std::min (P->type_c, 4);
    for (int q = 0; q < 4; ++q) {
      unsigned w = 0;
      for (int t = 0; t < 4; ++t) {
        int v = (q < nn && t < nn) ? 16 * P->mat[q * P->type_c + t] + 8 : 8;
        w |= ((unsigned) (v & 0xFF)) << (8 * t);
      }
      hc.prof4[q] = w;
    }
  }
  {
    auto pk = [] (int v) { return ((unsigned) v & 0xFFFFu) * 0x10001u; };
    hc.pk_do = pk (-16 * P->del_o + 4); hc.pk_de = pk (-16 * P->del_e);
    hc.pk_io = pk (-16 * P->ins_o); hc.pk_ie = pk (-16 * P->ins_e);
    hc.pk_de32 = (unsigned) (-16 * P->del_e * 65537); hc.pk_ie32 = (unsigned) (-16 * P->ins_e * 65537);
  }
  GCG_CUDA (cudaMemcpyToSymbolAsync (c_sw, &hc, sizeof hc, 0, cudaMemcpyHostToDevice, ctx->stream));

  gcg_trace_mark (ctx, "sw.align: constants");
  // ---- classify
  std::vector<int> packed_ids, generic_ids;
  std::vector<sw_task> tasks ((size_t) n);
  for (int64_t p = 0; p < n; ++p) {
    sw_task & t = tasks[(size_t) p];
    t.qoff = b->qoff[(size_t) p]; t.toff = b->toff[(size_t) p];
    t.qlen = (int) (b->qoff[(size_t) p + 1] - t.qoff); t.tlen = (int) (b->toff[(size_t) p + 1] - t.toff);
    t.pair = (int) p; t.pad = 0; t.trace_off = 0; t.edge_off = 0;
    long long lo, hi;
    sw_bounds (P, t.qlen, t.tlen, &lo, &hi);
    GCG_CHECK (lo > -(1LL << 24) && hi < (1LL << 24), GCG_ERANGE, "gcg_swbatch_align: pair %lld: score range [%lld,%lld] exceeds the s32 kernel", (long long) p, lo, hi);
    bool pk = packable && t.qlen > 0 && t.tlen > 0 && t.tlen <= SW_PACKED_MAXROWS && b->maxsym[(size_t) p] < 4 &&
              lo >= -2030 && hi <= 2040;
    (pk ? packed_ids : generic_ids).push_back ((int) p);
  }
  std::sort (packed_ids.begin (), packed_ids.end (), [&] (int a, int c) {
    const sw_task & x = tasks[(size_t) a], & y = tasks[(size_t) c];
    if (x.tlen != y.tlen) return x.tlen < y.tlen;
    if (x.qlen != y.qlen) return x.qlen < y.qlen;
    return a < c;
  });
  b->n_packed = (int64_t) packed_ids.size ();
  b->n_generic = (int64_t) generic_ids.size ();

  gcg_trace_mark (ctx, "sw.align: classify");
  // ---- memory plan: trace + edge buffers for one wave
  // (the driver is only asked for the free memory when no trace buffer is at hand: cudaMemGetInfo
  // was measured at 0.1 .. 11 ms per call)
  auto query_budget_words = [&] () -> unsigned long long {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo (&free_b, &total_b) != cudaSuccess) { cudaGetLastError (); return (unsigned long long) ((size_t) 8 << 30) / 4; }
    // memory parked in the stream-ordered pool is as good as free: the next allocation reuses it
    cudaMemPool_t pool;
    unsigned long long reserved = 0, used = 0;
    if (cudaDeviceGetDefaultMemPool (&pool, ctx->device) == cudaSuccess &&
        cudaMemPoolGetAttribute (pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
        cudaMemPoolGetAttribute (pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
      free_b += (size_t) (reserved - used);
    free_b += ctx->dparked_bytes + ctx->dbig_cls;   // so are the blocks parked in the context's own cache (released on demand)
    free_b += b->s_trace.cap;                       // this batch's own trace buffer is reused or replaced
    return (unsigned long long) (std::min<size_t> (free_b / 4 * 3, (size_t) 64 << 30) / 4);
  };
  auto trace_words = [] (const sw_task & t) -> unsigned long long {
    if (t.qlen == 0 || t.tlen == 0) return 0;
    return (unsigned long long) ((t.qlen + SW_BAND - 1) / SW_BAND) * (unsigned long long) (t.tlen + 31) * 32ULL;
  };
  unsigned long long max_single = 0;
  for (auto & t : tasks) max_single = std::max (max_single, trace_words (t));
  unsigned long long total_trace = 0; long long total_edges = 0;
  for (auto & t : tasks) { total_trace += trace_words (t); total_edges += (long long) t.qlen + t.tlen; }
  // A trace buffer that is already there — this batch's from an earlier align, or the context's spare
  // from a batch that was freed — sets the budget as long as it holds whole alignments: asking the
  // driver again would give a slightly different answer every time (free memory moves), and a
  // different wave size means freeing and re-allocating tens of GB (20-40 ms per align).
  const unsigned long long need_min = std::max<unsigned long long> (2 * max_single, 32);
  const unsigned long long own = b->s_trace.cap / 4, spare = ctx->dbig_cls / 4;      // (a parked block has exactly its class size)
  unsigned long long budget_words;
  if (own >= need_min) budget_words = own;
  else if (spare >= need_min) budget_words = spare;
  else budget_words = query_budget_words ();
  if (const char * e = getenv ("GCG_SW_TRACE_BUDGET_MB")) budget_words = (unsigned long long) atoll (e) * (1 << 20) / 4;
  GCG_CHECK (2 * max_single <= budget_words || max_single == 0, GCG_ENOMEM, "gcg_swbatch_align: one alignment needs %llu MB of trace, more than the budget",
             (unsigned long long) (max_single * 4 >> 20));

  // work units: packed items (two alignments per warp where they may share one) and generic alignments
  struct unit { int a, b; bool packed; };
  // Two alignments share a warp only if neither half can leave its 16 bits anywhere in the common
  // (band-padded) rectangle: the packed kernel adds both halves with one 32-bit add, so a borrow out
  // of the low half — harmless garbage past the end of the shorter alignment when the halves were
  // added separately — would reach the other alignment.  An alignment alone in its warp sits in the
  // low half, which nothing can reach.
  auto pair_ok = [&] (int x, int y) -> bool {
    const sw_task & a = tasks[(size_t) x], & c = tasks[(size_t) y];
    long long lo, hi, qpad = ((long long) std::max (a.qlen, c.qlen) + SW_BAND - 1) / SW_BAND * SW_BAND;
    sw_bounds (P, qpad, std::max (a.tlen, c.tlen), &lo, &hi);
    return lo >= -2030 && hi <= 2040;
  };
  auto pair_up = [&] (const std::vector<int> & ids) {          // ids sorted by (tlen, qlen): neighbours are alike
    std::vector<unit> u;
    for (size_t i = 0; i < ids.size (); ) {
      if (i + 1 < ids.size () && pair_ok (ids[i], ids[i + 1])) { u.push_back ({ids[i], ids[i + 1], true}); i += 2; }
      else { u.push_back ({ids[i], -1, true}); i += 1; }
    }
    return u;
  };
  std::vector<unit> punits = pair_up (packed_ids), units;
  for (int id : generic_ids) units.push_back ({id, -1, false});
  auto unit_words = [&] (const unit & u) { return trace_words (tasks[(size_t) u.a]) + (u.b >= 0 ? trace_words (tasks[(size_t) u.b]) : 0ULL); };
  // The packed kernel is persistent: every resident warp owns one trace slot that holds its largest item.
  unsigned long long slot_words = 0, generic_trace = 0;
  for (auto & u : punits) slot_words = std::max (slot_words, unit_words (u));
  for (auto & u : units) generic_trace += unit_words (u);
  int packed_slots = 0, packed_rows_cap = 32;
  if (!punits.empty ()) {
    int per_sm = 0;
    for (int id : packed_ids) packed_rows_cap = std::max (packed_rows_cap, tasks[(size_t) id].tlen);
    packed_rows_cap = (packed_rows_cap + 31) & ~31;
    GCG_CUDA (cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, sw_packed_fn (), 32, (size_t) packed_rows_cap * 2));
    packed_slots = (int) std::min<size_t> (punits.size (), (size_t) ctx->sm_count * std::max (per_sm, 1));
  }
  const unsigned long long need_trace = std::max<unsigned long long> (std::max ((unsigned long long) packed_slots * slot_words, generic_trace), 32);

  uint32_t * d_trace = nullptr; int * d_edges = nullptr; sw_task * d_tasks = nullptr; int * d_items = nullptr; int * d_counter = nullptr;
  int2 * d_bound = nullptr;
  uint2 * d_pbound = nullptr;
  sw_task * d_wave_tasks = nullptr;
  unsigned long long trace_cap = std::min (budget_words, need_trace);
  int rc = GCG_OK;
  cudaError_t ce;
  // the per-launch work lists are small; allocate for the whole batch once
  long long max_tlen = 1;
  for (auto & t : tasks) max_tlen = std::max<long long> (max_tlen, t.tlen);
  const int gen_blocks = ctx->sm_count * 4;          // 4 warps per block, 4 blocks per SM resident
  long long bound_stride = max_tlen + 1;
  // keep the trace buffer of an earlier align on this batch whenever it can hold whole alignments
  // (the memory it occupies is no longer "free", so the budget above shrinks on later calls)
  if (b->s_trace.cap / 4 >= std::max<unsigned long long> (2 * max_single, 32))
    trace_cap = std::min<unsigned long long> (need_trace, b->s_trace.cap / 4);
  if ((ce = b->s_trace.reserve (ctx, (size_t) trace_cap * 4)) != cudaSuccess ||
      (ce = b->s_edges.reserve (ctx, (size_t) std::max<long long> (total_edges, 1) * 4)) != cudaSuccess ||
      (ce = b->s_tasks.reserve (ctx, (size_t) n * sizeof (sw_task))) != cudaSuccess ||
      (ce = b->s_wave_tasks.reserve (ctx, (size_t) n * sizeof (sw_task))) != cudaSuccess ||
      (ce = b->s_items.reserve (ctx, (size_t) n * 8)) != cudaSuccess ||
      (ce = b->s_counter.reserve (ctx, 2 * sizeof (int))) != cudaSuccess ||
      (generic_ids.size () && (ce = b->s_bound.reserve (ctx, (size_t) gen_blocks * 4 * (size_t) bound_stride * sizeof (int2))) != cudaSuccess) ||
      (generic_ids.size () && (ce = b->s_gops.reserve (ctx, (size_t) (total_edges + 4 * n + 8) * 4)) != cudaSuccess)) {
    gcg_set_error ("gcg_swbatch_align: cudaMalloc failed: %s", cudaGetErrorString (ce));
    rc = GCG_ENOMEM;
  }
  d_trace = (uint32_t *) b->s_trace.p; d_edges = (int *) b->s_edges.p; d_tasks = (sw_task *) b->s_tasks.p;
  d_wave_tasks = (sw_task *) b->s_wave_tasks.p; d_items = (int *) b->s_items.p; d_counter = (int *) b->s_counter.p;
  d_bound = (int2 *) b->s_bound.p;
  // edge offsets are global for the batch (small: 4 bytes per symbol)
  if (!rc) {
    long long eo = 0;
    for (auto & t : tasks) { t.edge_off = eo; eo += (long long) t.qlen + t.tlen; }
  }
  // CIGAR pool: start from a typical size, grow on demand
  if (!rc) {
    // as-is CIGARs are a handful of operations; a re-fetching traceback of a noisy read changes operation
    // every few bases.  (A pool that turns out too small costs a second pass over the alignments it failed.)
    unsigned long long want = (unsigned long long) n * 64;
    if (mode == GCG_SW_FIXED) for (auto & t : tasks) want += (unsigned long long) std::min (t.qlen, t.tlen) / 2;
    want = std::max<unsigned long long> (1 << 16, want);
    if (const char * e = getenv ("GCG_SW_POOL_INIT")) { want = (unsigned long long) atoll (e); gcg_dfree (ctx, b->d_pool); b->d_pool = nullptr; b->pool_cap = 0; }
    if (b->pool_cap < want) {
      gcg_dfree (ctx, b->d_pool); b->d_pool = nullptr; b->pool_cap = 0;
      if ((ce = gcg_dmalloc (ctx, &b->d_pool, (size_t) want * 4)) != cudaSuccess) { gcg_set_error ("gcg_swbatch_align: cigar pool: %s", cudaGetErrorString (ce)); rc = GCG_ENOMEM; }
      else b->pool_cap = want;
    }
  }
  unsigned long long * d_cursor = ctx->d_counters + 10;
  if (!rc) GCG_CUDA (cudaMemsetAsync (d_cursor, 0, 8, ctx->stream));
  gcg_trace_mark (ctx, "sw.align: scratch");

  // ---- packed items: ONE persistent launch; alignments whose CIGAR did not fit the pool are done again
  for (int attempt = 0; !rc && !punits.empty (); ++attempt) {
    const int n_slots = (int) std::max<unsigned long long> (1, std::min<unsigned long long> ((unsigned long long) packed_slots, trace_cap / std::max<unsigned long long> (slot_words, 1)));
    // largest items first: the launch ends with the short ones
    std::vector<int2> pitems (punits.size ());
    for (size_t i = 0; i < punits.size (); ++i) {
      const unit & u = punits[punits.size () - 1 - i];
      pitems[i] = make_int2 (u.a, u.b);
      tasks[(size_t) u.a].trace_off = 0;                                     // relative to the warp's slot
      if (u.b >= 0) tasks[(size_t) u.b].trace_off = trace_words (tasks[(size_t) u.a]);
    }
    const int grid = (int) std::min<size_t> (pitems.size (), (size_t) n_slots);
    if ((ce = b->s_pbound.reserve (ctx, (size_t) grid * packed_rows_cap * sizeof (uint2))) != cudaSuccess) {
      gcg_set_error ("gcg_swbatch_align: boundary scratch: %s", cudaGetErrorString (ce)); rc = GCG_ENOMEM; break; }
    d_pbound = (uint2 *) b->s_pbound.p;
    // one scratch list of CIGAR operations per walking lane (two per warp): the traceback is walked once
    unsigned long long ops_stride = 8;
    for (int id : packed_ids) ops_stride = std::max<unsigned long long> (ops_stride, (unsigned long long) tasks[(size_t) id].qlen + tasks[(size_t) id].tlen + 4);
    if ((ce = b->s_ops.reserve (ctx, (size_t) grid * 2 * ops_stride * 4)) != cudaSuccess) {
      gcg_set_error ("gcg_swbatch_align: traceback scratch: %s", cudaGetErrorString (ce)); rc = GCG_ENOMEM; break; }
    GCG_CUDA (cudaMemcpyAsync (d_tasks, tasks.data (), (size_t) n * sizeof (sw_task), cudaMemcpyHostToDevice, ctx->stream));
    GCG_CUDA (cudaMemcpyAsync (d_items, pitems.data (), pitems.size () * sizeof (int2), cudaMemcpyHostToDevice, ctx->stream));
    GCG_CUDA (cudaMemsetAsync (d_counter, 0, 2 * sizeof (int), ctx->stream));
    { gcg_kscope ks (ctx, "k7_sw_fill_packed");
      sw_packed_fn ()<<<grid, 32, (size_t) packed_rows_cap * 2, ctx->stream>>> (
          b->d_qry, b->d_tgt, d_tasks, (const int2 *) d_items, (int) pitems.size (), d_counter, d_trace, slot_words, d_edges, d_pbound,
          packed_rows_cap, b->d_ends, sw_packed_consts {hc.pk_do, hc.pk_io, hc.pk_de32, hc.pk_ie32, 1u},
          sw_cigar_args {mode, b->d_pool, b->pool_cap, d_cursor, b->d_results, (uint32_t *) b->s_ops.p, ops_stride});
      GCG_CUDA (cudaGetLastError ()); }
    GCG_CUDA (cudaMemcpyAsync (ctx->h_counters + 10, d_cursor, 8, cudaMemcpyDeviceToHost, ctx->stream));
    GCG_CUDA (cudaStreamSynchronize (ctx->stream));
    gcg_trace_mark (ctx, "sw.align: packed items done");
    const unsigned long long cur = ctx->h_counters[10];
    if (cur <= b->pool_cap) { b->pool_used = cur; break; }
    if (attempt == 2) { gcg_set_error ("gcg_swbatch_align: cigar pool overflow persists"); rc = GCG_ERANGE; break; }
    // Reservations are handed out in increasing order, so everything from the first one that ended
    // beyond the pool is unwritten: grow the pool (now the size is known), rewind the cursor to that
    // reservation and run those alignments again.
    std::vector<gcg_sw_result> hres ((size_t) n);
    GCG_CUDA (cudaMemcpy (hres.data (), b->d_results, (size_t) n * sizeof (gcg_sw_result), cudaMemcpyDeviceToHost));
    std::vector<int> failed;
    unsigned long long base = cur;
    for (int id : packed_ids) {                      // (sorted order is kept: the failed ones pair up as before)
      const gcg_sw_result & r = hres[(size_t) id];
      if ((unsigned long long) r.cigar_off + (unsigned long long) r.n_cigar > b->pool_cap) { failed.push_back (id); base = std::min (base, (unsigned long long) r.cigar_off); }
    }
    if (attempt > 0) {                               // only the alignments of the previous attempt can have failed
      std::vector<char> in_prev ((size_t) n, 0);
      for (auto & u : punits) { in_prev[(size_t) u.a] = 1; if (u.b >= 0) in_prev[(size_t) u.b] = 1; }
      std::vector<int> f2; base = cur;
      for (int id : failed) if (in_prev[(size_t) id]) { f2.push_back (id); base = std::min (base, (unsigned long long) hres[(size_t) id].cigar_off); }
      failed.swap (f2);
    }
    const unsigned long long ncap = std::max (b->pool_cap * 2, cur + (cur - base) / 8 + 1024);
    uint32_t * np = nullptr;
    if ((ce = gcg_dmalloc (ctx, &np, (size_t) ncap * 4)) != cudaSuccess) { gcg_set_error ("gcg_swbatch_align: cigar pool growth to %llu ops failed: %s", ncap, cudaGetErrorString (ce)); rc = GCG_ENOMEM; break; }
    if (base) GCG_CUDA (cudaMemcpyAsync (np, b->d_pool, (size_t) base * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->h_counters[10] = base;
    GCG_CUDA (cudaMemcpyAsync (d_cursor, ctx->h_counters + 10, 8, cudaMemcpyHostToDevice, ctx->stream));
    GCG_CUDA (cudaStreamSynchronize (ctx->stream));
    gcg_dfree (ctx, b->d_pool);
    b->d_pool = np; b->pool_cap = ncap; b->pool_used = base;
    punits = pair_up (failed);
  }

  size_t u0 = 0;
  while (!rc && u0 < units.size ()) {
    // ---- cut a wave
    unsigned long long used = 0;
    size_t u1 = u0;
    while (u1 < units.size ()) {
      unsigned long long need = trace_words (tasks[(size_t) units[u1].a]) + (units[u1].b >= 0 ? trace_words (tasks[(size_t) units[u1].b]) : 0);
      if (used + need > trace_cap && u1 > u0) break;
      used += need;
      ++u1;
    }
    // assign trace offsets, build launch lists
    std::vector<int> gitems; std::vector<sw_task> wave_tasks;
    unsigned long long off = 0;
    for (size_t u = u0; u < u1; ++u) {
      for (int id : {units[u].a, units[u].b}) {
        if (id < 0) continue;
        sw_task & t = tasks[(size_t) id];
        t.trace_off = off; off += trace_words (t);
        wave_tasks.push_back (t);
      }
      gitems.push_back (units[u].a);
    }
    GCG_CUDA (cudaMemcpyAsync (d_tasks, tasks.data (), (size_t) n * sizeof (sw_task), cudaMemcpyHostToDevice, ctx->stream));
    GCG_CUDA (cudaMemcpyAsync (d_wave_tasks, wave_tasks.data (), wave_tasks.size () * sizeof (sw_task), cudaMemcpyHostToDevice, ctx->stream));
    GCG_CUDA (cudaMemsetAsync (d_counter, 0, 2 * sizeof (int), ctx->stream));
    if (!gitems.empty ()) {
      int * d_gitems = d_items;
      GCG_CUDA (cudaMemcpyAsync (d_gitems, gitems.data (), gitems.size () * sizeof (int), cudaMemcpyHostToDevice, ctx->stream));
      int grid = (int) std::min<size_t> ((gitems.size () + 3) / 4, (size_t) gen_blocks);
      gcg_kscope ks (ctx, "k7_sw_fill_generic");
      sw_fill_generic_kernel<<<grid, 128, 0, ctx->stream>>> (b->d_qry, b->d_tgt, d_tasks, d_gitems, (int) gitems.size (), d_counter + 1,
                                                             d_trace, d_edges, d_bound, bound_stride, b->d_ends);
      GCG_CUDA (cudaGetLastError ());
    }
    gcg_trace_mark (ctx, "sw.align: wave launched");
    // ---- CIGAR for the wave; grow the pool and redo the wave's CIGARs if it overflowed
    unsigned long long wave_start = b->pool_used;
    for (int attempt = 0; attempt < 3 && !rc; ++attempt) {
      { gcg_kscope ks (ctx, "k9_sw_cigar");
        int nt = (int) wave_tasks.size ();
        sw_cigar_kernel<<<(nt + 127) / 128, 128, 0, ctx->stream>>> (d_wave_tasks, nt, d_trace, b->d_ends, mode, b->d_pool, b->pool_cap, d_cursor, b->d_results, (uint32_t *) b->s_gops.p); }
      GCG_CUDA (cudaGetLastError ());
      GCG_CUDA (cudaMemcpyAsync (ctx->h_counters + 10, d_cursor, 8, cudaMemcpyDeviceToHost, ctx->stream));
      GCG_CUDA (cudaStreamSynchronize (ctx->stream));
      unsigned long long cur = ctx->h_counters[10];
      if (cur <= b->pool_cap) { b->pool_used = cur; break; }
      // grow: keep what earlier waves wrote
      unsigned long long ncap = std::max (cur + (cur - wave_start) * (unsigned long long) (units.size () - u1) / std::max<size_t> (u1 - u0, 1), b->pool_cap * 2);
      uint32_t * np = nullptr;
      if ((ce = gcg_dmalloc (ctx, &np, (size_t) ncap * 4)) != cudaSuccess) { gcg_set_error ("gcg_swbatch_align: cigar pool growth to %llu ops failed: %s", ncap, cudaGetErrorString (ce)); rc = GCG_ENOMEM; break; }
      if (wave_start) GCG_CUDA (cudaMemcpyAsync (np, b->d_pool, (size_t) wave_start * 4, cudaMemcpyDeviceToDevice, ctx->stream));
      GCG_CUDA (cudaStreamSynchronize (ctx->stream));
      gcg_dfree (ctx, b->d_pool);
      b->d_pool = np; b->pool_cap = ncap;
      ctx->h_counters[10] = wave_start;
      GCG_CUDA (cudaMemcpyAsync (d_cursor, ctx->h_counters + 10, 8, cudaMemcpyHostToDevice, ctx->stream));
      GCG_CUDA (cudaStreamSynchronize (ctx->stream));
      if (attempt == 2) { gcg_set_error ("gcg_swbatch_align: cigar pool overflow persists"); rc = GCG_ERANGE; }
    }
    gcg_trace_mark (ctx, "sw.align: wave done");
    u0 = u1;
  }
  if (!rc && cudaStreamSynchronize (ctx->stream) != cudaSuccess) { gcg_set_error ("gcg_swbatch_align: %s", cudaGetErrorString (cudaGetLastError ())); rc = GCG_ECUDA; }
  if (!rc) b->aligned = true;
  return rc;
}

extern "C" int gcg_swbatch_download (gcg_ctx * ctx, gcg_swbatch * b, gcg_sw_result * results,
                                     uint32_t ** cigar_pool, int64_t * n_cigar_pool)
{
  GCG_CHECK (ctx && b && results && cigar_pool && n_cigar_pool, GCG_EINVAL, "gcg_swbatch_download: bad argument");
  GCG_CHECK (b->aligned, GCG_EINVAL, "gcg_swbatch_download: no alignment has been run on this batch");
  GCG_CUDA (cudaSetDevice (ctx->device));
  *cigar_pool = nullptr;
  *n_cigar_pool = (int64_t) b->pool_used;
  if (b->n) GCG_CUDA (cudaMemcpyAsync (results, b->d_results, (size_t) b->n * sizeof (gcg_sw_result), cudaMemcpyDeviceToHost, ctx->stream));
  if (b->pool_used) {
    *cigar_pool = (uint32_t *) gcg_pinned_alloc ((size_t) b->pool_used * 4);
    if (!*cigar_pool) { gcg_set_error ("gcg_swbatch_download: pinned alloc of %llu CIGAR ops failed", b->pool_used); return GCG_ENOMEM; }
    GCG_CUDA (cudaMemcpyAsync (*cigar_pool, b->d_pool, (size_t) b->pool_used * 4, cudaMemcpyDeviceToHost, ctx->stream));
  }
  GCG_CUDA (cudaStreamSynchronize (ctx->stream));
  return GCG_OK;
}

extern "C" int gcg_sw_batch (gcg_ctx * ctx, const gcg_sw_params * P, int mode,
                             const char * qry, const int64_t * qoff, const char * tgt, const int64_t * toff, int64_t n,
                             gcg_sw_result * results, uint32_t ** cigar_pool, int64_t * n_cigar_pool)
{
  gcg_swbatch * b = nullptr;
  gcg_trace_mark (ctx, nullptr);
  int rc = gcg_swbatch_upload (ctx, qry, qoff, tgt, toff, n, &b);
  if (rc) return rc;
  gcg_trace_mark (ctx, "sw: pairs to HBM");
  rc = gcg_swbatch_align (ctx, b, P, mode);
  gcg_trace_mark (ctx, "sw: fill + cigar");
  if (!rc) rc = gcg_swbatch_download (ctx, b, results, cigar_pool, n_cigar_pool);
  gcg_trace_mark (ctx, "sw: results to host");
  gcg_swbatch_free (b);
  return rc;
}

// The same batch sharded over several contexts, one per GPU (SURVEY 8e, row SW: pairs are independent,
// contiguous ranges of about equal cells, no collective, results in input order).  One host thread per
// context runs the ordinary gcg_sw_batch on its range; the CIGAR pools are joined into one pinned
// array and the offsets of the later ranges moved behind the earlier ones.
extern "C" int gcg_sw_batch_multi (gcg_ctx * const * ctxs, int n_ctx, const gcg_sw_params * P, int mode,
                                   const char * qry, const int64_t * qoff, const char * tgt, const int64_t * toff, int64_t n,
                                   gcg_sw_result * results, uint32_t ** cigar_pool, int64_t * n_cigar_pool)
{
  GCG_CHECK (ctxs && n_ctx >= 1 && n_ctx <= 64 && P && qoff && toff && n >= 0 && results && cigar_pool && n_cigar_pool, GCG_EINVAL, "gcg_sw_batch_multi: bad argument");
  for (int i = 0; i < n_ctx; ++i) GCG_CHECK (ctxs[i] != nullptr, GCG_EINVAL, "gcg_sw_batch_multi: context %d is NULL", i);
  if (n_ctx == 1) return gcg_sw_batch (ctxs[0], P, mode, qry, qoff, tgt, toff, n, results, cigar_pool, n_cigar_pool);
  // share d ends at the first pair where the running cell count reaches (d + 1) / n_ctx of the total
  std::vector<long double> cum ((size_t) n + 1, 0.0L);
  for (int64_t p = 0; p < n; ++p) cum[(size_t) p + 1] = cum[(size_t) p] + (long double) (qoff[p + 1] - qoff[p]) * (long double) (toff[p + 1] - toff[p]);
  std::vector<int64_t> bound ((size_t) n_ctx + 1, n);
  bound[0] = 0;
  for (int d = 1; d < n_ctx; ++d) {
    const long double goal = cum[(size_t) n] * d / n_ctx;
    int64_t r = bound[(size_t) d - 1];
    while (r < n && cum[(size_t) r] < goal) ++r;
    bound[(size_t) d] = r;
  }
  struct share { int rc = GCG_OK; uint32_t * pool = nullptr; int64_t n_pool = 0; std::string err; };
  std::vector<share> sh ((size_t) n_ctx);
  std::vector<std::thread> th;
  for (int d = 0; d < n_ctx; ++d) {
    th.emplace_back ([&, d] () {
      const int64_t b = bound[(size_t) d], m = bound[(size_t) d + 1] - b;
      share & s = sh[(size_t) d];
      if (m == 0) return;
      std::vector<int64_t> qo ((size_t) m + 1), to ((size_t) m + 1);      // offsets counted from the share's first pair
      for (int64_t p = 0; p <= m; ++p) { qo[(size_t) p] = qoff[b + p] - qoff[b]; to[(size_t) p] = toff[b + p] - toff[b]; }
      s.rc = gcg_sw_batch (ctxs[d], P, mode, qry + qoff[b], qo.data (), tgt + toff[b], to.data (), m, results + b, &s.pool, &s.n_pool);
      if (s.rc) s.err = gcg_last_error ();                                  // (the message is thread local)
    });
  }
  for (auto & t : th) t.join ();
  int rc = GCG_OK;
  int64_t total = 0;
  for (int d = 0; d < n_ctx; ++d) {
    if (sh[(size_t) d].rc && !rc) { rc = sh[(size_t) d].rc; gcg_set_error ("gcg_sw_batch_multi: share %d (device %d): %s", d, ctxs[d]->device, sh[(size_t) d].err.c_str ()); }
    total += sh[(size_t) d].n_pool;
  }
  uint32_t * pool = nullptr;
  if (!rc) {
    pool = (uint32_t *) gcg_pinned_alloc ((size_t) std::max<int64_t> (total, 1) * 4);
    if (!pool) { gcg_set_error ("gcg_sw_batch_multi: pinned allocation of %lld CIGAR words failed", (long long) total); rc = GCG_ENOMEM; }
  }
  int64_t at = 0;
  for (int d = 0; d < n_ctx; ++d) {
    share & s = sh[(size_t) d];
    if (!rc && s.n_pool > 0) {
      memcpy (pool + at, s.pool, (size_t) s.n_pool * 4);
      for (int64_t p = bound[(size_t) d]; p < bound[(size_t) d + 1]; ++p) results[p].cigar_off += at;
      at += s.n_pool;
    }
    if (s.pool) gcg_free (s.pool);
  }
  if (rc) return rc;
  *cigar_pool = pool;
  *n_cigar_pool = total;
  return GCG_OK;
}
