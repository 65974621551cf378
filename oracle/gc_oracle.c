/* TEST INFRASTRUCTURE — CPU restatement of the reference algorithm for the gap-closing hot path.
 *
 * This file is the *checker*, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load it.  The shipped path is the CUDA library
 * (superplus_b200/csrc) and fails loudly when that is missing.
 *
 * Parity status: PINNED.  Every function below is checked (tests/test_oracle_vs_ref.py, golden
 * fixtures under tests/golden/) against the reference's own sources compiled unmodified into
 * oracle/_ref/ (gc, ref_kmer, libref_sw_{asis,fixed}.so) — the reference ships no test vectors
 * for this path, so outputs of the reference itself are the pin.
 *
 * Plain C, written for obviousness, not speed.  Each function cites the reference lines it
 * restates (paths relative to /root/reference/gap_closer).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * 1. 2-bit code, rolling canonical k-mer
 * ---------------------------------------------------------------------------------------- */

/* bio.h:24  base2int(b) = (b>>1)&3  -> A/a=0 C/c=1 T/t=2 G/g=3, N/n=3, any byte by formula.
 * `char` is signed on x86-64 gcc; >> on a negative int is arithmetic there and &3 keeps the
 * same two bits as the unsigned shift, so taking the byte as unsigned is equivalent. */
static inline uint64_t gco_base2int (char b) { return (uint64_t) ((((unsigned char) b) >> 1) & 3); }

/* bio.h:22  int_comp(i) = i ^ 2 */
static inline uint64_t gco_comp (uint64_t i) { return i ^ 2; }

/* kseq1.h:37-46 followed literally */
static uint64_t gco_revcomp (uint64_t x, int k)
{
  uint64_t r = x ^ 0xAAAAAAAAAAAAAAAAULL;
  r = ((r & 0x3333333333333333ULL) << 2)  | ((r & 0xCCCCCCCCCCCCCCCCULL) >> 2);
  r = ((r & 0x0F0F0F0F0F0F0F0FULL) << 4)  | ((r & 0xF0F0F0F0F0F0F0F0ULL) >> 4);
  r = ((r & 0x00FF00FF00FF00FFULL) << 8)  | ((r & 0xFF00FF00FF00FF00ULL) >> 8);
  r = ((r & 0x0000FFFF0000FFFFULL) << 16) | ((r & 0xFFFF0000FFFF0000ULL) >> 16);
  r = ((r & 0x00000000FFFFFFFFULL) << 32) | ((r & 0xFFFFFFFF00000000ULL) >> 32);
  return r >> (64 - (k << 1));
}

/* kmer.c:73-117 (contigs) and ont.c:155-203 (reads) share this loop: first k-mer built with
 * kseq1_new_b (kseq1.h:28-35), its reverse complement with kseq1_fast_reverse_comp, then
 * kseq1_next (kseq1.h:54-59) / kseq1_prev (kseq1.h:48-52) per base.  Canonical = forward iff
 * fwd < rc (kmer.c:86, ont.c:161), else rc with the REV flag.
 * Writes n = max(0, len-k+1) entries; returns n. */
int64_t gco_chop (const char * s, int64_t len, int k, uint64_t * kseq_out, uint8_t * rev_out)
{
  int64_t n = len - k + 1, i, j;
  uint64_t mask, fwd = 0, rc, b;
  if (n < 1) return 0;
  mask = (((uint64_t) 1) << (k << 1)) - 1;            /* kseq1.h:71-74 */
  for (i = 0; i < k; ++i) fwd = (fwd << 2) | gco_base2int (s[i]);
  rc = gco_revcomp (fwd, k);
  for (i = 0, j = k; ; ++i, ++j) {
    if (fwd < rc) { kseq_out[i] = fwd; rev_out[i] = 0; }
    else          { kseq_out[i] = rc;  rev_out[i] = 1; }
    if (i + 1 >= n) break;
    b = gco_base2int (s[j]);
    fwd = ((fwd << 2) & mask) | b;
    rc = (rc >> 2) | (gco_comp (b) << ((k - 1) << 1));
  }
  return n;
}

/* ------------------------------------------------------------------------------------------
 * 2. contig k-mer table: distinct canonical k-mers with multiplicity and first occurrence
 *    (kmer.c:124-152 -> hash.c:113-152: a duplicate key only does ++multi and keeps the key
 *    pointer of the first inserted occurrence; threads scan contigs in index order and
 *    positions in order, so "first" = lowest (tid,pos)).  Partition id crc32 % n_thread
 *    (kmer.c:88) is unobservable and is not modelled.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint64_t n_slot;          /* power of two */
  uint64_t n_item;
  uint64_t * key;           /* UINT64_MAX = empty */
  int32_t * multi;
  int32_t * tid;
  int32_t * pos;
  uint8_t * rev;
  int32_t * ont_multi;      /* ONT-side multiplicity, filled by gco_search (ont.c:230-254) */
  int k;
} gco_table;

static inline uint64_t gco_mix (uint64_t x)
{
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

static uint64_t gco_find (const gco_table * t, uint64_t key)
{
  uint64_t h = gco_mix (key) & (t->n_slot - 1);
  while (t->key[h] != UINT64_MAX && t->key[h] != key) h = (h + 1) & (t->n_slot - 1);
  return h;
}

gco_table * gco_table_build (const char * buf, const int64_t * off, int32_t n_contig, int k)
{
  int64_t total = 0, maxlen = 0, i, c;
  gco_table * t = (gco_table *) calloc (1, sizeof (gco_table));
  for (c = 0; c < n_contig; ++c) {
    int64_t l = off[c + 1] - off[c];
    if (l > maxlen) maxlen = l;
    if (l >= k) total += l - k + 1;
  }
  t->k = k;
  t->n_slot = 1024;
  while (t->n_slot < (uint64_t) total * 2 + 2) t->n_slot <<= 1;
  t->key = (uint64_t *) malloc (t->n_slot * 8);
  memset (t->key, 0xFF, t->n_slot * 8);
  t->multi = (int32_t *) calloc (t->n_slot, 4);
  t->tid = (int32_t *) calloc (t->n_slot, 4);
  t->pos = (int32_t *) calloc (t->n_slot, 4);
  t->rev = (uint8_t *) calloc (t->n_slot, 1);
  t->ont_multi = (int32_t *) calloc (t->n_slot, 4);
  uint64_t * ks = (uint64_t *) malloc ((maxlen + 1) * 8);
  uint8_t * rv = (uint8_t *) malloc (maxlen + 1);
  for (c = 0; c < n_contig; ++c) {
    int64_t n = gco_chop (buf + off[c], off[c + 1] - off[c], k, ks, rv);
    for (i = 0; i < n; ++i) {
      uint64_t h = gco_find (t, ks[i]);
      if (t->key[h] == UINT64_MAX) {
        t->key[h] = ks[i]; t->multi[h] = 1; t->tid[h] = (int32_t) c; t->pos[h] = (int32_t) i; t->rev[h] = rv[i];
        ++t->n_item;
      } else
        ++t->multi[h];
    }
  }
  free (ks); free (rv);
  return t;
}

void gco_table_free (gco_table * t)
{
  if (!t) return;
  free (t->key); free (t->multi); free (t->tid); free (t->pos); free (t->rev); free (t->ont_multi); free (t);
}

int64_t gco_table_size (const gco_table * t) { return (int64_t) t->n_item; }

/* kmer.c:265-287: total = #items, unique = #items with multi==1 */
void gco_table_stats (const gco_table * t, int64_t out[2])
{
  uint64_t h;
  out[0] = (int64_t) t->n_item; out[1] = 0;
  for (h = 0; h < t->n_slot; ++h) if (t->key[h] != UINT64_MAX && t->multi[h] == 1) ++out[1];
}

/* unordered dump of every item (caller sorts) */
int64_t gco_table_dump (const gco_table * t, uint64_t * key, int32_t * multi, int32_t * tid, int32_t * pos, uint8_t * rev)
{
  uint64_t h; int64_t n = 0;
  for (h = 0; h < t->n_slot; ++h) if (t->key[h] != UINT64_MAX) {
    key[n] = t->key[h]; multi[n] = t->multi[h]; tid[n] = t->tid[h]; pos[n] = t->pos[h]; rev[n] = t->rev[h]; ++n;
  }
  return n;
}

/* ------------------------------------------------------------------------------------------
 * 3. ONT search (ont.c:141-204) + ONT-side multiplicity (ont.c:230-254) + stats (kmer.c:290-312)
 *    For every read r and position p in [0, l-k]: canonical key; if it is in the contig table
 *    with multi==1 it is an anchor: (r, p, tid, cpos, kmer flag, ont flag).  Reads shorter
 *    than k contribute nothing (ont.c:155).  ont_stats[0] = #distinct anchored k-mers,
 *    ont_stats[1] = those anchored exactly once over all reads.
 *    Hits are produced in (read,pos) order.  Returns the number of hits (may exceed cap; only
 *    the first cap are stored).
 * ---------------------------------------------------------------------------------------- */
int64_t gco_search (gco_table * t, const char * buf, const int64_t * off, int64_t n_read, int k,
                    int64_t cap, int64_t * h_read, int32_t * h_pos, int32_t * h_tid, int32_t * h_cpos,
                    uint8_t * h_krev, uint8_t * h_orev, int64_t ont_stats[2])
{
  int64_t r, i, n_hit = 0, maxlen = 0;
  uint64_t h;
  for (r = 0; r < n_read; ++r) if (off[r + 1] - off[r] > maxlen) maxlen = off[r + 1] - off[r];
  uint64_t * ks = (uint64_t *) malloc ((maxlen + 1) * 8);
  uint8_t * rv = (uint8_t *) malloc (maxlen + 1);
  memset (t->ont_multi, 0, t->n_slot * 4);
  for (r = 0; r < n_read; ++r) {
    int64_t n = gco_chop (buf + off[r], off[r + 1] - off[r], k, ks, rv);
    for (i = 0; i < n; ++i) {
      h = gco_find (t, ks[i]);
      if (t->key[h] == UINT64_MAX || t->multi[h] != 1) continue;      /* ont.c:171,195 */
      ++t->ont_multi[h];
      if (n_hit < cap) {
        h_read[n_hit] = r; h_pos[n_hit] = (int32_t) i; h_tid[n_hit] = t->tid[h]; h_cpos[n_hit] = t->pos[h];
        h_krev[n_hit] = t->rev[h]; h_orev[n_hit] = rv[i];
      }
      ++n_hit;
    }
  }
  ont_stats[0] = ont_stats[1] = 0;
  for (h = 0; h < t->n_slot; ++h) if (t->key[h] != UINT64_MAX && t->ont_multi[h] > 0) {
    ++ont_stats[0];
    if (t->ont_multi[h] == 1) ++ont_stats[1];
  }
  free (ks); free (rv);
  return n_hit;
}

/* ------------------------------------------------------------------------------------------
 * 3b. partition id of a k-mer and un-anchored segments of a read
 * ---------------------------------------------------------------------------------------- */

/* crc32.h:15-81 restated: the table there is the reflected CRC-32 of polynomial 0xEDB88320 (zlib's);
 * crc32(0, buf, len) starts from ~0, folds one byte per step through the table, returns ~crc.
 * kmer.h:46-49: kseq_crc32 runs it over the 8 bytes of the kseq1_t as they lie in memory (little
 * endian on the x86-64 the reference runs on).  kmer.c:88,110: hs_id = that % n_thread. */
static uint32_t gco_crc_table[256];
static int gco_crc_ready = 0;
static void gco_crc_init (void)
{
  uint32_t i, c;
  int j;
  for (i = 0; i < 256; ++i) {
    c = i;
    for (j = 0; j < 8; ++j) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
    gco_crc_table[i] = c;
  }
  gco_crc_ready = 1;
}

uint32_t gco_kseq_crc32 (uint64_t kseq)
{
  uint32_t crc = 0 ^ 0xffffffffu;
  int b;
  if (!gco_crc_ready) gco_crc_init ();
  for (b = 0; b < 8; ++b) {
    /* crc32.h:77: the byte is read through a `char *` (signed), the xor is masked with 0xff afterwards */
    char byte = (char) (kseq >> (8 * b));
    crc = gco_crc_table[((uint32_t) crc ^ (uint32_t) byte) & 0xff] ^ (crc >> 8);
  }
  return crc ^ 0xffffffffu;
}

void gco_hs_id (const uint64_t * kseq, int64_t n, int32_t n_thread, int32_t * hs_id_out)
{
  int64_t i;
  for (i = 0; i < n; ++i) hs_id_out[i] = (int32_t) (gco_kseq_crc32 (kseq[i]) % (uint32_t) n_thread);
}

/* ont.c:264-309 find_unankor_segs: okmers[] holds one entry per BASE of the read (ont.c:489-491
 * resizes to r->l), anchored[i] != 0 where okmers[i].kmer != NULL.  A segment opens at the first
 * un-anchored index after an anchored one (or at 0) and closes at the next anchored index (or at
 * n).  Writes {beg,end} pairs; returns the number of segments (only the first cap are stored). */
int64_t gco_unanchored_segs (const uint8_t * anchored, int64_t n, int32_t * beg_end_out, int64_t cap)
{
  int64_t i, n_seg = 0;
  int prev_anchored = 1, open = 0;
  for (i = 0; i < n; ++i) {
    int cur = anchored[i] != 0;
    if (!prev_anchored && cur) { if (n_seg - 1 < cap) beg_end_out[2 * (n_seg - 1) + 1] = (int32_t) i; open = 0; }
    else if (prev_anchored && !cur) { if (n_seg < cap) beg_end_out[2 * n_seg] = (int32_t) i; ++n_seg; open = 1; }
    prev_anchored = cur;
  }
  if (open && n_seg - 1 < cap) beg_end_out[2 * (n_seg - 1) + 1] = (int32_t) n;
  return n_seg;
}

/* ------------------------------------------------------------------------------------------
 * 4. Smith-Waterman (sw.c) + CIGAR (cigar.c)
 * ---------------------------------------------------------------------------------------- */
#define GCO_SW_M 1   /* sw.c:41-43 */
#define GCO_SW_I 2
#define GCO_SW_D 4
#define GCO_SOFTCLIP      0   /* sw.h:20-23 */
#define GCO_LEADING_INDEL 1
#define GCO_INDEL         2
#define GCO_IGNORE        3

typedef struct {
  int32_t type_c;
  int32_t del_o, del_e, ins_o, ins_e;
  int32_t strategy;
  /* Border scores the aligner currently holds.  In the reference they are written by
   * init_matrix_values (sw.c:61-110) at sw_set_parameter / matrix growth time and are NOT
   * rewritten when a later sw_set_parameter selects SOFTCLIP (early return, sw.c:93-94), so
   * they are state of the aligner, not a pure function of the current parameters:
   * border_kind 0: all zero (calloc, sw.c:157 / SOFTCLIP); 1: row0[j] = -b_ins_o-(j-1)*b_ins_e,
   * col0[i] = -b_del_o-(i-1)*b_del_e for i,j >= 1 (sw.c:96-109). */
  int32_t border_kind;
  int32_t b_del_o, b_del_e, b_ins_o, b_ins_e;
  int32_t mat[64];           /* type_c x type_c, type_c <= 8 here */
} gco_sw_params;

typedef struct { int32_t ms, is, ds, ml, dl, il, score, status; } gco_cell;   /* sw.h:31-36 */

static void gco_cigar_add (uint32_t * c, int32_t * n, int32_t cap, uint32_t e) { if (*n < cap) c[*n] = e; ++*n; }

/* mode 0 = as shipped (traceback never re-fetches the cell, sw.c:289-319)
 * mode 1 = fixed (cell re-fetched at the top of every do-iteration)
 * out: [0] score [1] alignment_offset [2] has_softclip [3] n_cigar [4] end-cell target index
 *      [5] end-cell query index [6] trailing seg_len chosen by the end-cell search
 * trace (optional, (tlen+1)*(qlen+1) bytes): bit0-1 status (1=M 2=I 3=D, 0 border),
 *      bit2 = D was extended, bit3 = I was extended — the 4 bits/cell the CUDA path spills. */
int gco_sw_align (const gco_sw_params * P, int32_t qlen, const char * qry, int32_t tlen, const char * tgt,
                  int mode, int64_t * out, uint32_t * cigar, int32_t cigar_cap, uint8_t * trace)
{
  const int64_t W = (int64_t) qlen + 1;
  int64_t i, j;
  gco_cell * sm = (gco_cell *) calloc ((size_t) ((int64_t) (tlen + 1) * W), sizeof (gco_cell));
  if (!sm) return -1;
  /* borders: sw.c:61-110 */
  for (j = 0; j <= qlen; ++j) { sm[j].is = INT32_MIN / 2; sm[j].ds = INT32_MIN / 2; }
  for (i = 1; i <= tlen; ++i) { sm[i * W].is = INT32_MIN / 2; sm[i * W].ds = INT32_MIN / 2; }
  if (P->border_kind == 1) {
    for (j = 1; j <= qlen; ++j) sm[j].score = -P->b_ins_o - (int32_t) (j - 1) * P->b_ins_e;
    for (i = 1; i <= tlen; ++i) sm[i * W].score = -P->b_del_o - (int32_t) (i - 1) * P->b_del_e;
  }
  /* fill: sw.c:203-252 */
  for (i = 1; i <= tlen; ++i) {
    char t = tgt[i - 1];
    for (j = 1; j <= qlen; ++j) {
      char q = qry[j - 1];
      gco_cell * c = sm + i * W + j, * lc = c - 1, * uc = c - W, * ulc = c - W - 1;
      int32_t pre_gap;
      c->ms = ulc->score + P->mat[q * P->type_c + t];
      pre_gap = uc->score - P->del_o;
      c->ds = uc->ds - P->del_e;
      if (pre_gap > c->ds) { c->ds = pre_gap; c->dl = 1; } else c->dl = uc->dl + 1;
      pre_gap = lc->score - P->ins_o;
      c->is = lc->is - P->ins_e;
      if (pre_gap > c->is) { c->is = pre_gap; c->il = 1; } else c->il = lc->il + 1;
      if (c->ms >= c->ds && c->ms >= c->is) { c->status = GCO_SW_M; c->score = c->ms; c->ml = ulc->ml + 1; }
      else if (c->is > c->ds)               { c->status = GCO_SW_I; c->score = c->is; c->ml = 0; }
      else                                  { c->status = GCO_SW_D; c->score = c->ds; c->ml = 0; }
      if (trace) {
        uint8_t st = c->status == GCO_SW_M ? 1 : (c->status == GCO_SW_I ? 2 : 3);
        trace[i * W + j] = (uint8_t) (st | ((c->dl > 1) << 2) | ((c->il > 1) << 3));
      }
    }
  }
  /* end cell: sw.c:254-280 */
  int32_t bt_tidx = tlen, bt_qidx = qlen, max_score = INT32_MIN, cur;
  uint32_t seg_len = 0, step_len, pre_opr = 0, cur_opr;
  for (i = 1; i <= tlen; ++i) {
    cur = sm[i * W + qlen].score;
    if (cur >= max_score) { bt_tidx = (int32_t) i; max_score = cur; }
  }
  if (P->strategy != GCO_LEADING_INDEL) {
    for (i = 1; i <= qlen; ++i) {
      cur = sm[(int64_t) tlen * W + i].score;
      if (cur > max_score || (cur == max_score && abs (tlen - (int32_t) i) < abs (bt_tidx - bt_qidx))) {
        bt_tidx = tlen; bt_qidx = (int32_t) i; max_score = cur; seg_len = (uint32_t) (qlen - i);
      }
    }
  }
  out[4] = bt_tidx; out[5] = bt_qidx; out[6] = seg_len;
  int32_t n = 0;
  out[2] = 0;
  if (seg_len > 0 && P->strategy == GCO_SOFTCLIP) {            /* sw.c:282-287 */
    gco_cigar_add (cigar, &n, cigar_cap, (seg_len << 4) | 4);
    seg_len = 0;
    out[2] = 1;
  }
  gco_cell * c = sm + (int64_t) bt_tidx * W + bt_qidx;
  out[0] = c->score;
  int inited = 0;
  do {                                                          /* sw.c:292-319 */
    if (mode == 1) c = sm + (int64_t) bt_tidx * W + bt_qidx;
    if (c->status == GCO_SW_M)      { step_len = (uint32_t) c->ml; cur_opr = 0; bt_tidx -= step_len; bt_qidx -= step_len; }
    else if (c->status == GCO_SW_D) { step_len = (uint32_t) c->dl; cur_opr = 2; bt_tidx -= step_len; }
    else                            { step_len = (uint32_t) c->il; cur_opr = 1; bt_qidx -= step_len; }
    if (inited && cur_opr != pre_opr) { gco_cigar_add (cigar, &n, cigar_cap, (seg_len << 4) | pre_opr); seg_len = 0; }
    seg_len += step_len;
    pre_opr = cur_opr;
    inited = 1;
    if (step_len == 0) break;   /* guard: the reference would spin forever here (border end cell with both indices > 0 cannot happen; kept for safety) */
  } while (bt_tidx > 0 && bt_qidx > 0);
  gco_cigar_add (cigar, &n, cigar_cap, (seg_len << 4) | pre_opr);
  if (P->strategy == GCO_SOFTCLIP) {                            /* sw.c:323-332 */
    if (bt_qidx > 0) gco_cigar_add (cigar, &n, cigar_cap, (((uint32_t) bt_qidx) << 4) | 4);
    out[1] = bt_tidx;
  } else {
    if (bt_tidx > 0) gco_cigar_add (cigar, &n, cigar_cap, (((uint32_t) bt_tidx) << 4) | 2);
    if (bt_qidx > 0) gco_cigar_add (cigar, &n, cigar_cap, (((uint32_t) bt_qidx) << 4) | 1);
    out[1] = 0;
  }
  /* cigar_reverse: cigar.c:108-125 */
  { int32_t m = n < cigar_cap ? n : cigar_cap, a; for (a = 0; a < m / 2; ++a) { uint32_t x = cigar[a]; cigar[a] = cigar[m - 1 - a]; cigar[m - 1 - a] = x; } }
  out[3] = n;
  free (sm);
  return 0;
}

/* last row / last column of the score plane, for end-cell parity debugging */
int gco_sw_edges (const gco_sw_params * P, int32_t qlen, const char * qry, int32_t tlen, const char * tgt,
                  int32_t * last_col /* tlen+1 */, int32_t * last_row /* qlen+1 */)
{
  /* two-row rolling version of the same recurrence (sw.c:203-252) */
  const int64_t W = (int64_t) qlen + 1;
  int64_t i, j;
  int32_t * H0 = (int32_t *) calloc (W, 4), * H1 = (int32_t *) calloc (W, 4), * D = (int32_t *) malloc (W * 4), * tmp;
  for (j = 0; j <= qlen; ++j) { D[j] = INT32_MIN / 2; H0[j] = (P->border_kind == 1 && j >= 1) ? -P->b_ins_o - (int32_t) (j - 1) * P->b_ins_e : 0; }
  last_col[0] = H0[qlen];
  for (i = 1; i <= tlen; ++i) {
    int32_t I = INT32_MIN / 2;
    H1[0] = (P->border_kind == 1) ? -P->b_del_o - (int32_t) (i - 1) * P->b_del_e : 0;
    for (j = 1; j <= qlen; ++j) {
      int32_t ms = H0[j - 1] + P->mat[qry[j - 1] * P->type_c + tgt[i - 1]];
      int32_t g = H0[j] - P->del_o, d = D[j] - P->del_e; if (g > d) d = g; D[j] = d;
      g = H1[j - 1] - P->ins_o; I = I - P->ins_e; if (g > I) I = g;
      H1[j] = (ms >= d && ms >= I) ? ms : (I > d ? I : d);
    }
    last_col[i] = H1[qlen];
    tmp = H0; H0 = H1; H1 = tmp;
  }
  for (j = 0; j <= qlen; ++j) last_row[j] = H0[j];
  free (H0); free (H1); free (D);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * 5. CIGAR helpers (cigar.c:196-265) on a plain uint32 array
 * ---------------------------------------------------------------------------------------- */
#define GCO_CIGAR_TYPE 0x3C1A7
int32_t gco_cigar2ref_len (const uint32_t * c, int32_t n)
{ int32_t i, l = 0; for (i = 0; i < n; ++i) if ((GCO_CIGAR_TYPE >> ((c[i] & 0xf) << 1) & 3) & 2) l += c[i] >> 4; return l; }
int32_t gco_cigar2qry_len (const uint32_t * c, int32_t n)
{ int32_t i, l = 0; for (i = 0; i < n; ++i) if ((GCO_CIGAR_TYPE >> ((c[i] & 0xf) << 1) & 3) & 1) l += c[i] >> 4; return l; }

/* ------------------------------------------------------------------------------------------
 * 6. Blizzard string hash (hash_func.c:22-69)
 * ---------------------------------------------------------------------------------------- */
static uint64_t gco_crypt[0x500];
static int gco_crypt_ready = 0;
static void gco_crypt_init (void)
{
  uint64_t seed = 0x00100001, i1, i2, idx, t1, t2;
  for (i1 = 0; i1 < 0x100; ++i1)
    for (idx = i1, i2 = 0; i2 < 5; ++i2, idx += 0x100) {
      seed = (seed * 125 + 3) % 0x2AAAAB; t1 = (seed & 0xFFFF) << 0x10;
      seed = (seed * 125 + 3) % 0x2AAAAB; t2 = (seed & 0xFFFF);
      gco_crypt[idx] = t1 | t2;
    }
  gco_crypt_ready = 1;
}
uint64_t gco_blizzard (const char * s, int32_t len, int hash_type)
{
  uint64_t seed1 = 0x7FED7FED, seed2 = 0xEEEEEEEE, ch;
  int32_t i;
  if (!gco_crypt_ready) gco_crypt_init ();
  for (i = 0; i < len; ++i) {
    ch = (uint64_t) (unsigned char) s[i];
    if (ch >= 'a' && ch <= 'z') ch -= 32;
    seed1 = gco_crypt[(hash_type << 8) + ch] ^ (seed1 + seed2);
    seed2 = ch + seed1 + seed2 + (seed2 << 5) + 3;
  }
  return seed1;
}
