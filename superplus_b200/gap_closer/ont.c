/* ont.c — replacement for gap_closer/ont.c: same prototypes (ont.h:50-59).  The per-base
 * crc32 + chained-hash probe loop of the reference (ont.c:141-204) and its thread pool
 * (ont.c:316-400) are gone; search_kmers_on_ont_reads is now a batcher over libgcgpu:
 *
 *   SEARCH   (reference ont.c:361-368 -> chop1ont2kmer/chop1segm ont.c:141-223)
 *            + REHASH (reference ont.c:371-377 -> rehash1okseq ont.c:230-254)
 *                 -> gcg_search: reads to HBM, canonical k-mers, table probes, ONT-side
 *                    multiplicity, anchors back in (read,pos) order
 *   back-fill of the dense okmers[] the unchanged consumers index by read position
 *            (ctg_graph.c:603-651): host threads over whole reads, one 16-byte store per anchor,
 *            from the compact 8-byte anchors the device sends (gcg_search_compact)
 *   UNANKOR  (reference ont.c:381-388 -> find_unankor_segs ont.c:264-309): derived from the
 *            sorted anchor list instead of re-scanning 16 bytes per ONT base
 *
 * With GC_DEVICES=0,1,... the reads are sharded by batch over several GPUs of the box (SURVEY 8e,
 * replicated table): contiguous read ranges of about equal bases, one host thread and one table
 * replica per device, anchors back-filled per share, the replicas' ONT-side multiplicity folded into
 * the primary table afterwards (gcg_table_merge_ont) so that kmer_stat2 prints the counts over all reads.
 *
 * okmer->kmer points at &ctgs[tid].kmers[cpos] (what the reference's search phase stores,
 * ont.c:174); the reference later re-points it to a byte-identical copy inside anchored_ksets
 * (ont.c:245,253).  The anchored sets stay empty host objects (main.c:68-72 creates and frees
 * them); their only reader, kmer_stat2, prints the device-side counters (kmer.c).
 */
#include <time.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <pthread.h>

#include "mp.h"
#include "ont.h"
#include "rseq.h"
#include "utils.h"
#include "kmer.h"
#include "gcg_bridge.h"

/* ------------------------------------------------------------------ back-fill ---------- */
/* The device returns COMPACT anchors (include/gcgpu.h, gcg_search_compact): one 64-bit word per anchor,
 * pos << 36 | gpos << 2 | flags, grouped by read through read_off[]; gpos counts contig bases in the
 * order the contigs were uploaded (chop_contig_seqs2kmers), so cbase[] below turns it back into
 * (tid, cpos).  Inputs beyond the bit budget of that form (a read of 2^28 bases, 2^34 contig bases) come
 * back as 16-byte gcg_hit records instead (gcg_search); both forms fill the same okmers[]. */
typedef struct {
  const uint64_t * anchors;   /* compact form, or NULL */
  const int64_t * read_off;
  const gcg_hit * hits;       /* 16-byte form, or NULL */
  int64_t beg, end;           /* compact: reads [beg, end) of the share; 16-byte: hits [beg, end) */
  int64_t read_off0;          /* index of the share's first read (both forms count reads from their share's start) */
  const int64_t * cbase;      /* n_ctg + 1 */
  int64_t n_ctg;
  mp_t(okseq) * okseqs;
  mp_t(ctg) * ctgs;
  int sparse, kmer_len, n_thread;   /* GC_SPARSE_KMERS: ctg->kmers[] was not filled by the chop; write the records anchors point at */
} fill_arg_t;

static uint64_t canon_at (const char * s, int k, int * rev);

/* the kmer_t record of contig position cpos, as chop_kmer_core writes it (kmer.c:73-117); kmer_len is written last and
 * doubles as the "filled" mark (two threads that meet on one record write the same bytes) */
static inline void
sparse_kmer_fill (kmer_t * km, const ctg_t * c, int64_t tid, int64_t cpos, int kmer_len, int n_thread)
{
  int rev;
  if (km->kmer_len != 0) return;
  km->kseq = canon_at (c->seq->s + cpos, kmer_len, &rev);
  km->hs_id = (int32_t) (kseq_crc32 (&km->kseq) % (uint32_t) n_thread);     /* kmer.c:88 */
  km->tid = (int32_t) tid;
  km->pos = (int32_t) cpos;
  km->flag = rev ? KMER_REV : 0;
  __atomic_store_n (&km->kmer_len, (int16_t) kmer_len, __ATOMIC_RELEASE);
}

static void *
fill_core (void * data)
{
  fill_arg_t * a = (fill_arg_t *) data;
  int64_t i, r, tid = 0;
  if (a->anchors == NULL) {
    for (i = a->beg; i < a->end; ++i) {
      const gcg_hit * h = a->hits + i;
      okseq_t * okseq = a->okseqs->pool + a->read_off0 + h->read;
      ont_kmer_t * ok = okseq->okmers->pool + h->pos;
      kmer_t * km = a->ctgs->pool[h->tid].kmers + (h->cpos_flags >> 2);
      if (a->sparse) sparse_kmer_fill (km, a->ctgs->pool + h->tid, h->tid, h->cpos_flags >> 2, a->kmer_len, a->n_thread);
      ok->kmer = km;
      ok->ont_pos = h->pos;
      ok->hs_id = (int16_t) km->hs_id;
      ok->flag = (h->cpos_flags & 2u) ? ONT_KMER_REV : 0;
    }
    return NULL;
  }
  for (r = a->beg; r < a->end; ++r) {
    ont_kmer_t * okmers = a->okseqs->pool[a->read_off0 + r].okmers->pool;
    for (i = a->read_off[r]; i < a->read_off[r + 1]; ++i) {
      const uint64_t w = a->anchors[i];
      const int64_t gpos = GCG_ANCHOR_GPOS (w);
      const int32_t pos = GCG_ANCHOR_POS (w);
      ont_kmer_t * ok = okmers + pos;
      kmer_t * km;
      if (gpos < a->cbase[tid] || gpos >= a->cbase[tid + 1]) {
        /* neighbouring anchors of a read mostly sit on one contig; otherwise bisect (empty contigs share a base: take the last) */
        int64_t lo = 0, hi = a->n_ctg;
        while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (a->cbase[mid] <= gpos) lo = mid; else hi = mid; }
        tid = lo;
      }
      km = a->ctgs->pool[tid].kmers + (gpos - a->cbase[tid]);
      if (a->sparse) sparse_kmer_fill (km, a->ctgs->pool + tid, tid, gpos - a->cbase[tid], a->kmer_len, a->n_thread);
      ok->kmer = km;
      ok->ont_pos = pos;
      ok->hs_id = (int16_t) km->hs_id;
      ok->flag = (w & 2u) ? ONT_KMER_REV : 0;
    }
  }
  return NULL;
}

/* maximal runs of un-anchored positions of the reads [r0, r1) (reference find_unankor_segs, ont.c:264-309),
 * from their share's sorted anchors instead of a scan over 16 bytes per ONT base */
static void
seg_add (okseq_t * okseq, int64_t beg, int64_t end)
{
  ont_seg_t * seg = mp_alloc (oseg, okseq->segs);
  seg->beg = (int32_t) beg;
  seg->end = (int32_t) end;
}

static void
rebuild_segs (mp_t(okseq) * okseqs, int64_t r0, int64_t r1, const gcg_hit * hits, int64_t n_hit,
    const uint64_t * anchors, const int64_t * read_off)
{
  int64_t r, n_reads = r1 - r0, h = 0;
  for (r = 0; r < n_reads; ++r) {
    okseq_t * okseq = mp_at (okseq, okseqs, r0 + r);
    int64_t n = mp_cnt (okseq->okmers), cur = 0, pos;
    mp_clear (oseg, okseq->segs, NULL);
    if (anchors != NULL) {
      for (h = read_off[r]; h < read_off[r + 1]; ++h) {
        pos = GCG_ANCHOR_POS (anchors[h]);
        if (pos > cur) seg_add (okseq, cur, pos);
        cur = pos + 1;
      }
    } else {
      while (h < n_hit && hits[h].read == r) {
        pos = hits[h].pos;
        if (pos > cur) seg_add (okseq, cur, pos);
        cur = pos + 1;
        ++h;
      }
    }
    if (cur < n) seg_add (okseq, cur, n);
  }
}

/* ------------------------------------------------------------------ GC_RUNS (N2 + N3) -- */
/* canonical k-mer of the contig occurrence an anchor points at (kmer.c:73-117 restated for ONE position): forward
 * word from the k bases, its reverse complement by kseq1_fast_reverse_comp's definition (kseq1.h:37-46) */
static uint64_t
canon_at (const char * s, int k, int * rev)
{
  uint64_t fwd = 0, rc = 0;
  int i;
  for (i = 0; i < k; ++i) {
    uint64_t b = ((unsigned char) s[i] >> 1) & 3;                  /* bio.h:24 */
    fwd = (fwd << 2) | b;
    rc |= (b ^ 2) << (2 * i);                                       /* bio.h:22: complement = code ^ 2 */
  }
  *rev = !(fwd < rc);                                               /* kmer.c:86 */
  return *rev ? rc : fwd;
}

static void
okmer_from_anchor (ont_kmer_t * ok, kmer_t * km, uint64_t w, mp_t(ctg) * ctgs, const int64_t * cbase, int64_t n_ctg, int kmer_len, int n_thread)
{
  const int64_t gpos = GCG_ANCHOR_GPOS (w);
  int64_t lo = 0, hi = n_ctg;
  int rev;
  ctg_t * c;
  while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (cbase[mid] <= gpos) lo = mid; else hi = mid; }
  c = mp_at (ctg, ctgs, lo);
  km->kseq = canon_at (c->seq->s + (gpos - cbase[lo]), kmer_len, &rev);
  km->hs_id = (int32_t) (kseq_crc32 (&km->kseq) % (uint32_t) n_thread);     /* kmer.c:88 */
  km->tid = (int32_t) lo;
  km->pos = (int32_t) (gpos - cbase[lo]);
  km->flag = (w & 1u) ? KMER_REV : 0;
  km->kmer_len = (int16_t) kmer_len;
  ok->kmer = km;
  ok->ont_pos = GCG_ANCHOR_POS (w);
  ok->hs_id = (int16_t) km->hs_id;
  ok->flag = (w & 2u) ? ONT_KMER_REV : 0;
  if (((w >> 1) ^ w) & 1u) ok->flag |= ONT_SCAF_REV;                /* ctg_graph.c:619-623: set on BACK anchors */
}

struct share_s;
static void runs_to_okmers (gcg_bridge_t * br, struct share_s * share, int n_dev, mp_t(okseq) * okseqs, mp_t(ctg) * ctgs,
    const int64_t * cbase, int64_t n_ctg, int kmer_len);

/* one share of the read batch on one device */
typedef struct share_s {
  gcg_ctx * ctx;
  gcg_table * table;
  const char ** ptrs;
  const int32_t * lens;
  int64_t r0, r1;
  int kmer_len;
  uint64_t * anchors;       /* compact form ... */
  int64_t * read_off;
  gcg_hit * hits;           /* ... or the 16-byte form */
  int64_t n_hit;
  const uint64_t * packed;  /* the loader's 2-bit words of ALL reads (pack on ingest) and their word offsets, or NULL */
  const int64_t * pwoff;
  gcg_run * runs;           /* GC_RUNS: run records instead of anchors */
  int64_t * run_off;
  int64_t n_run;
  int rc;
  char err[600];
} share_t;

static void *
share_core (void * data)
{
  share_t * s = (share_t *) data;
  if (gcg_bridge_runs_mode ()) {
    s->rc = s->packed
          ? gcg_search_runs_packed (s->ctx, s->table, s->packed, s->pwoff + s->r0, s->lens + s->r0, s->r1 - s->r0, s->kmer_len, &s->runs, &s->run_off, &s->n_run, &s->n_hit)
          : gcg_search_runs (s->ctx, s->table, s->ptrs + s->r0, s->lens + s->r0, s->r1 - s->r0, s->kmer_len, &s->runs, &s->run_off, &s->n_run, &s->n_hit);
    if (s->rc != 0) snprintf (s->err, sizeof s->err, "%s", gcg_last_error ());
    return NULL;
  }
  s->rc = getenv ("GC_ANCHORS16") ? GCG_ERANGE
        : s->packed ? gcg_search_compact_packed (s->ctx, s->table, s->packed, s->pwoff + s->r0, s->lens + s->r0, s->r1 - s->r0, s->kmer_len, &s->anchors, &s->read_off, &s->n_hit)
        : gcg_search_compact (s->ctx, s->table, s->ptrs + s->r0, s->lens + s->r0, s->r1 - s->r0, s->kmer_len, &s->anchors, &s->read_off, &s->n_hit);
  if (s->rc == GCG_ERANGE) {        /* beyond the compact form's bit budget (or GC_ANCHORS16 set: A/B of the two forms) */
    s->anchors = NULL; s->read_off = NULL;
    s->rc = gcg_search (s->ctx, s->table, s->ptrs + s->r0, s->lens + s->r0, s->r1 - s->r0, s->kmer_len, &s->hits, &s->n_hit);
  }
  if (s->rc != 0) snprintf (s->err, sizeof s->err, "%s", gcg_last_error ());   /* the message is thread local */
  return NULL;
}

/* joins the shares' run records in the bridge and gives every read its 2 x (number of runs) boundary anchors:
 * okmers[2q] / okmers[2q+1] = first / last anchor of run q's majority direction (what ont_node_init keeps of a run,
 * ctg_graph.c:93-181); a run without a 20-fold majority keeps zeroed entries (its node is deleted) */
static void
runs_to_okmers (gcg_bridge_t * br, share_t * share, int n_dev, mp_t(okseq) * okseqs, mp_t(ctg) * ctgs,
    const int64_t * cbase, int64_t n_ctg, int kmer_len)
{
  int d;
  int64_t r, q, n_reads = mp_cnt (okseqs), n_run = 0, at = 0;
  const int n_thread = br->n_thread > 0 ? br->n_thread : 1;
  for (d = 0; d < n_dev; ++d) n_run += share[d].n_run;
  free (br->runs); free (br->run_off); free (br->run_kmers);
  br->runs = (gcg_run *) ckalloc (n_run + 1, sizeof (gcg_run));
  br->run_off = (int64_t *) ckalloc (n_reads + 1, sizeof (int64_t));
  br->run_kmers = (kmer_t *) ckalloc (2 * n_run + 1, sizeof (kmer_t));
  br->n_run = n_run;
  for (d = 0; d < n_dev; ++d) {
    if (share[d].n_run > 0) memcpy (br->runs + at, share[d].runs, share[d].n_run * sizeof (gcg_run));
    for (r = share[d].r0; r < share[d].r1; ++r) br->run_off[r] = at + share[d].run_off[r - share[d].r0];
    at += share[d].n_run;
  }
  br->run_off[n_reads] = n_run;
  for (r = 0; r < n_reads; ++r) {
    okseq_t * okseq = mp_at (okseq, okseqs, r);
    const int64_t q0 = br->run_off[r], nq = br->run_off[r + 1] - q0;
    mp_resize (okmer, okseq->okmers, 2 * nq);
    okseq->okmers->n = 2 * nq;
    for (q = 0; q < nq; ++q) {
      const gcg_run * run = br->runs + q0 + q;
      ont_kmer_t * ok = okseq->okmers->pool + 2 * q;
      kmer_t * km = br->run_kmers + 2 * (q0 + q);
      memset (ok, 0, 2 * sizeof (ont_kmer_t));
      if (run->n_fwd > 20 * run->n_bwd) {
        okmer_from_anchor (ok, km, run->first_fwd, ctgs, cbase, n_ctg, kmer_len, n_thread);
        okmer_from_anchor (ok + 1, km + 1, run->last_fwd, ctgs, cbase, n_ctg, kmer_len, n_thread);
      } else if (run->n_bwd > 20 * run->n_fwd) {
        okmer_from_anchor (ok, km, run->first_bwd, ctgs, cbase, n_ctg, kmer_len, n_thread);
        okmer_from_anchor (ok + 1, km + 1, run->last_bwd, ctgs, cbase, n_ctg, kmer_len, n_thread);
      }
    }
  }
  br->runs_mode = 1;
}

int
search_kmers_on_ont_reads (mp_t(rs) * ont_seqs, mp_t(ctg) * ctg_seqs,
    xh_t ** ctg_khashs, mp_t(okseq) * okseqs, xh_set_t(kmer) ** anchored_ksets,
    const char * prefix, int n_thread, int kmer_len)
{
  int i, d, nt, n_dev, n_fill = 0;
  time_t time_beg, mod_tbeg;
  int64_t r, n_reads, n_hit = 0, total_bases = 0, run = 0, n_ctg;
  int64_t * cbase;
  const char ** ptrs;
  int32_t * lens;
  pthread_t * pids;
  pthread_t dev_pid[GCG_BRIDGE_MAX_DEV];
  fill_arg_t * args;
  share_t share[GCG_BRIDGE_MAX_DEV];
  gcg_bridge_t * br = gcg_bridge ();

  time (&time_beg);
  if (br->table == NULL)
    err_mesg ("[%s] put_contig_kmers2hashs has not been called", __func__);
  if (kmer_len != br->kmer_len)
    err_mesg ("[%s] kmer_len %d differs from the table's %d", __func__, kmer_len, br->kmer_len);
  /* The reference searches only okseq->segs from the second k-mer length on and keeps the anchors of the
   * earlier lengths (ont.c:207-223); this batcher searches whole reads and points okmer->kmer into
   * ctg->kmers[], which the next chop overwrites.  The shipped main.c runs one length (n_kmers == 1,
   * main.c:29); a second one must fail loudly instead of corrupting the earlier anchors. */
  if (br->searched_k != 0 && br->searched_k != kmer_len)
    err_mesg ("[%s] a second k-mer length (%d after %d) is not supported by the B200 path (main.c runs n_kmers == 1)", __func__, kmer_len, br->searched_k);
  br->searched_k = kmer_len;
  nt = n_thread > 0 ? n_thread : 1;
  n_dev = br->n_dev;
  for (d = 0; d < n_dev; ++d)
    GCG_CK (gcg_set_host_threads (br->ctxs[d], nt / n_dev > 0 ? nt / n_dev : 1));

  /* search + ONT-side multiplicity on the device(s) */
  time (&mod_tbeg);
  n_reads = mp_cnt (ont_seqs);
  ptrs = (const char **) ckalloc (n_reads + 1, sizeof (char *));
  lens = (int32_t *) ckalloc (n_reads + 1, sizeof (int32_t));
  for (r = 0; r < n_reads; ++r) {
    /* the reference searches okseq->seq over its single segment [0, l) (ont.c:207-223, 505-507) */
    okseq_t * okseq = mp_at (okseq, okseqs, r);
    ptrs[r] = okseq->seq->b;
    lens[r] = okseq->seq->l;
    total_bases += lens[r];
  }
  /* scaffold coordinate of the compact anchors: contigs in upload order (kmer.c of this directory) */
  n_ctg = mp_cnt (ctg_seqs);
  cbase = (int64_t *) ckalloc (n_ctg + 2, sizeof (int64_t));
  for (r = 0; r < n_ctg; ++r)
    cbase[r + 1] = cbase[r] + mp_at (ctg, ctg_seqs, r)->seq->l;
  /* contiguous shares of about equal bases: share d ends at the first read where the running base
   * count reaches (d + 1) / n_dev of the total (superplus_b200/api.py split_reads_by_bases) */
  memset (share, 0, sizeof share);
  /* reads the loader packed on ingest (rseq_fast.c) are searched from their 2-bit words; okseq->seq must still be
   * the loader's records, in order (ont_kseqs_init sets them so) */
  {
    const uint64_t * pk = NULL; const int64_t * pw = NULL;
    if (rseq_packed_lookup ((const void *) ont_seqs, n_reads, &pk, &pw)) {
      for (r = 0; r < n_reads; ++r)
        if (mp_at (okseq, okseqs, r)->seq != mp_at (rs, ont_seqs, r)) { pk = NULL; break; }
      for (d = 0; d < n_dev && pk != NULL; ++d) { share[d].packed = pk; share[d].pwoff = pw; }
      if (pk != NULL && getenv ("GCG_TRACE")) fprintf (stderr, "[gcg] searching the reads from the loader's 2-bit words (pack on ingest)\n");
    }
  }
  for (d = 0, r = 0; d < n_dev; ++d) {
    int64_t goal = (total_bases * (d + 1) + n_dev - 1) / n_dev;
    share[d].ctx = br->ctxs[d];
    share[d].table = d == 0 ? br->table : br->replicas[d];
    share[d].ptrs = ptrs; share[d].lens = lens; share[d].kmer_len = kmer_len;
    share[d].r0 = r;
    if (d == n_dev - 1) r = n_reads;
    else while (r < n_reads && run < goal) run += lens[r++];
    share[d].r1 = r;
    if (share[d].table == NULL)
      err_mesg ("[%s] device %d holds no replica of the table", __func__, br->devs[d]);
  }
  if (n_dev == 1)
    share_core (share);
  else {
    for (d = 0; d < n_dev; ++d) ckpthread_create (dev_pid + d, NULL, share_core, (void *) (share + d));
    for (d = 0; d < n_dev; ++d) ckpthread_join (dev_pid[d]);
  }
  for (d = 0; d < n_dev; ++d) {
    if (share[d].rc != 0)
      err_mesg ("[%s] search on device %d failed (%d): %s", __func__, br->devs[d], share[d].rc, share[d].err);
    n_hit += share[d].n_hit;
  }
  for (d = 1; d < n_dev; ++d)
    GCG_CK (gcg_table_merge_ont (br->ctx, br->table, br->replicas[d]));
  if (n_dev > 1 && getenv ("GCG_TRACE"))
    for (d = 0; d < n_dev; ++d)
      fprintf (stderr, "[gcg] device %d: reads [%ld, %ld), %ld anchors\n", br->devs[d], (long) share[d].r0, (long) share[d].r1, (long) share[d].n_hit);
  printf ("\n  chop and search ont kmers cost: %lds\n", time (NULL) - mod_tbeg);

  if (gcg_bridge_runs_mode ()) {
    /* GC_RUNS: the device has reduced the anchors to run records; materialise the two boundary anchors of every run */
    time (&mod_tbeg);
    runs_to_okmers (br, share, n_dev, okseqs, ctg_seqs, cbase, n_ctg, kmer_len);
    printf ("\n  re-hash ont kmers cost: %lds\n", time (NULL) - mod_tbeg);
    printf ("\n  find un-ankored positions on onts costs: %lds\n", 0L);      /* (segments are not kept in this mode: one k-mer length) */
    for (d = 0; d < n_dev; ++d) { gcg_free (share[d].runs); gcg_free (share[d].run_off); }
    free (ptrs); free (lens); free (cbase);
    printf ("\n  search ONT kmers total cost: %lds\n", time (NULL) - time_beg);
    return 0;
  }

  /* back-fill okmers[] (the reference's REHASH phase re-pointed the same entries) */
  time (&mod_tbeg);
  if (nt > n_hit / 4096 + 1) nt = (int) (n_hit / 4096 + 1);
  pids = (pthread_t *) ckalloc (nt + n_dev, sizeof (pthread_t));
  args = (fill_arg_t *) ckalloc (nt + n_dev, sizeof (fill_arg_t));
  for (d = 0; d < n_dev; ++d) {
    /* host threads in proportion to the share's anchors, at least one */
    int t, nt_d = n_hit > 0 ? (int) ((share[d].n_hit * nt + n_hit - 1) / n_hit) : 1;
    int64_t prev = 0, n_rd = share[d].r1 - share[d].r0;
    if (nt_d < 1) nt_d = 1;
    if (n_fill + nt_d > nt + n_dev) nt_d = nt + n_dev - n_fill;
    for (t = 0; t < nt_d; ++t, ++n_fill) {
      fill_arg_t * fa = args + n_fill;
      fa->anchors = share[d].anchors; fa->read_off = share[d].read_off; fa->hits = share[d].hits;
      fa->read_off0 = share[d].r0;
      fa->cbase = cbase; fa->n_ctg = n_ctg;
      fa->okseqs = okseqs;
      fa->ctgs = ctg_seqs;
      fa->sparse = gcg_bridge_sparse_kmers (); fa->kmer_len = kmer_len; fa->n_thread = nt;
      if (share[d].anchors != NULL) {
        /* whole reads per thread, cut where the anchor count reaches (t + 1) / nt_d of the share's */
        int64_t goal = share[d].n_hit * (t + 1) / nt_d, lo = prev, hi = n_rd;
        if (t == nt_d - 1) lo = n_rd;
        else while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (share[d].read_off[mid] < goal) lo = mid + 1; else hi = mid; }
        fa->beg = prev; fa->end = lo;
        prev = lo;
      } else {
        fa->beg = share[d].n_hit * t / nt_d;
        fa->end = share[d].n_hit * (t + 1) / nt_d;
      }
      ckpthread_create (pids + n_fill, NULL, fill_core, (void *) fa);
    }
  }
  for (i = 0; i < n_fill; ++i)
    ckpthread_join (pids[i]);
  printf ("\n  re-hash ont kmers cost: %lds\n", time (NULL) - mod_tbeg);

  /* un-anchored segments */
  time (&mod_tbeg);
  for (d = 0; d < n_dev; ++d)
    rebuild_segs (okseqs, share[d].r0, share[d].r1, share[d].hits, share[d].n_hit, share[d].anchors, share[d].read_off);
  printf ("\n  find un-ankored positions on onts costs: %lds\n", time (NULL) - mod_tbeg);

  for (d = 0; d < n_dev; ++d) { gcg_free (share[d].hits); gcg_free (share[d].anchors); gcg_free (share[d].read_off); }
  free (pids); free (args); free (ptrs); free (lens); free (cbase);

  printf ("\n  search ONT kmers total cost: %lds\n", time (NULL) - time_beg);
  return 0;
}

/* ------------------------------------------------------------------ okseq set-up -------- */
typedef struct {
  int64_t * cursor;
  mp_t(rs) * ont_seqs;
  mp_t(okseq) * okseqs;
  int runs_mode;
} init_arg_t;

static void *
init_core (void * data)
{
  init_arg_t * a = (init_arg_t *) data;
  int64_t n = mp_cnt (a->ont_seqs), i;
  for (;;) {
    rseq_t * r;
    okseq_t * okseq;
    ont_seg_t * seg;
    i = __sync_fetch_and_add (a->cursor, 1);
    if (i >= n) break;
    r = mp_at (rs, a->ont_seqs, i);
    okseq = a->okseqs->pool + i;
    okseq->seq = r;
    okseq->ont_id = (int) i;
    if (a->runs_mode) {                          /* GC_RUNS: the read gets two entries per run after the search */
      okseq->okmers->n = 0;
    } else {
      mp_resize (okmer, okseq->okmers, r->l);    /* zero-filled: okmer->kmer == NULL means no anchor */
      okseq->okmers->n = r->l;
    }
    mp_clear (oseg, okseq->segs, NULL);
    seg = mp_alloc (oseg, okseq->segs);
    seg->beg = 0;
    seg->end = r->l;
  }
  return NULL;
}

/* same post-conditions as reference ont.c:483-512; the per-read zeroed allocations (16 bytes per
 * ONT base, the largest single phase of the reference run) are spread over the host cores */
int
ont_kseqs_init (mp_t(rs) * ont_seqs, mp_t(okseq) * okseqs)
{
  int i, nt;
  int64_t cursor = 0, n = mp_cnt (ont_seqs);
  time_t t_beg = time (NULL);
  long ncpu = sysconf (_SC_NPROCESSORS_ONLN);
  pthread_t * pids;
  init_arg_t arg;

  mp_resize (okseq, okseqs, n);
  okseqs->n = n;

  nt = ncpu > 16 ? 16 : (ncpu < 1 ? 1 : (int) ncpu);
  if (nt > n) nt = n > 0 ? (int) n : 1;
  arg.cursor = &cursor; arg.ont_seqs = ont_seqs; arg.okseqs = okseqs; arg.runs_mode = gcg_bridge_runs_mode ();
  pids = (pthread_t *) ckalloc (nt, sizeof (pthread_t));
  for (i = 0; i < nt; ++i)
    ckpthread_create (pids + i, NULL, init_core, (void *) &arg);
  for (i = 0; i < nt; ++i)
    ckpthread_join (pids[i]);
  free (pids);
  if (getenv ("GCG_TRACE")) fprintf (stderr, "[gcg] ont_kseqs_init: %ld reads on %d threads, %ld s\n", (long) n, nt, (long) (time (NULL) - t_beg));
  return 0;
}

int
ont_kseqs_dump (mp_t(okseq) * okseqs)
{
  int64_t i, j;
  for (i = 0; i < mp_cnt (okseqs); ++i) {
    okseq_t * okseq = mp_at (okseq, okseqs, i);
    printf ("> ont %ld\n", i);
    for (j = 0; j < mp_cnt (okseq->okmers); ++j) {
      ont_kmer_t * ok = mp_at (okmer, okseq->okmers, j);
      if (ok->kmer != NULL)
        printf ("%d:%d:%d\n", ok->ont_pos, ok->kmer->kmer_len, ok->kmer->tid);
    }
  }
  return 0;
}
