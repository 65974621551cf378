/* TEST INFRASTRUCTURE — driver linked against the *reference's own* objects (contig.c, rseq.c,
 * kmer.c, hash.c, ont.c ... compiled from /root/reference/gap_closer by oracle/build_ref.sh).
 * It performs exactly the call sequence of gap_closer/main.c:147-187 for one k and dumps what
 * the reference computed, so that the CPU restatement (oracle/gc_oracle.c) and the CUDA path can
 * be compared with it bit for bit.  It contains no k-mer arithmetic of its own.
 *
 *   ref_kmer <scaff.fa> <ont.fq> <n_thread> <k> <out_prefix> <dump_level>
 *     dump_level 0: timings + the four stat integers only  (-> <prefix>.json)
 *     dump_level 1: + <prefix>.hits.bin   one 20-byte record per anchored ONT position, in
 *                     (read,pos) order: i32 read, i32 ont_pos, i32 tid, i32 ctg_pos, u16 kmer flag, u16 ont flag
 *     dump_level 2: + <prefix>.table.bin  one 24-byte record per distinct contig k-mer:
 *                     u64 kseq, i32 multi, i32 tid, i32 pos, u16 flag, i16 kmer_len   (first occurrence)
 *                   + <prefix>.ctgk.bin   one 20-byte record per contig position:
 *                     u64 kseq, i32 tid, i32 pos, u16 flag, i16 kmer_len
 *                   + <prefix>.hsid.bin   i32 hs_id per contig position (= crc32(kseq) % n_thread, kmer.c:88)
 *     dump_level >= 1 also writes <prefix>.segs.bin: per read i32 n_seg, then n_seg x {i32 beg, i32 end}
 *                     (okseq->segs after find_unankor_segs, ont.c:264-309)
 *
 * The same source, compiled with -DGCG_SHIM_HARNESS and linked against the replacement files of
 * superplus_b200/gap_closer/ instead of the reference's kmer.c / hash.c / ont.c (target shim_kmer of
 * that directory's Makefile), dumps what the B200 path leaves in the same host structures; the GPU
 * tests compare the two sets of files byte for byte.  In that build the host tables are empty handles
 * (the table lives in HBM), so table.bin and the recomputed statistics are skipped — the statistics
 * are compared through the lines kmer_stat / kmer_stat2 print.
 */
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mp.h"
#include "def.h"
#include "str.h"
#include "hash.h"
#include "kseq1.h"
#include "utils.h"
#include "contig.h"
#include "rseq.h"
#include "kmer.h"
#include "ont.h"
#include "ctg_graph.h"

static double now_s (void)
{
  struct timespec t;
  clock_gettime (CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

int main (int argc, char * argv[])
{
  if (argc < 7) {
    fprintf (stderr, "usage: ref_kmer <scaff> <ont> <n_thread> <k> <out_prefix> <dump_level>\n");
    return 1;
  }
  const char * scaff = argv[1];
  const char * ont = argv[2];
  int n_thread = atoi (argv[3]);
  int k = atoi (argv[4]);
  const char * prefix = argv[5];
  int dump = atoi (argv[6]);
  char path[4096];
  int i;
  int64_t r, p;
  double t0, t_load_ctg, t_load_ont, t_okinit, t_chop, t_put, t_search;

  hash_func_init ();

  xh_set_t(kmer) ** anchored = (xh_set_t(kmer) **) ckmalloc (n_thread * sizeof (xh_set_t(kmer) *));
  for (i = 0; i < n_thread; ++i)
    anchored[i] = xh_set_init (kmer, 65536, 0.75, NULL, NULL, kmer_copy_func, kmer_hash_func, kmer_is_equal);
  xh_t ** khashs = kmer_hash_init (n_thread);
  ctg_graph_t * cg = ctg_graph_init ();
  mp_t(okseq) * okseqs = mp_init (okseq, okseq_init2, NULL);
  mp_t(sf) * scafs = mp_init (sf, scaf_init2, NULL);

  t0 = now_s ();
  mp_t(ctg) * ctgs = contig_seqs_load (scaff, cg, scafs);
  t_load_ctg = now_s () - t0;

  t0 = now_s ();
  mp_t(rs) * reads = sefq_load (ont);
  t_load_ont = now_s () - t0;

  t0 = now_s ();
  ont_kseqs_init (reads, okseqs);
  t_okinit = now_s () - t0;

  contig_seqs_info_clear (ctgs);
  t0 = now_s ();
  chop_contig_seqs2kmers (ctgs, n_thread, k);
  t_chop = now_s () - t0;

  kmer_hash_clear (khashs, n_thread);
  t0 = now_s ();
  put_contig_kmers2hashs (khashs, ctgs, n_thread);
  t_put = now_s () - t0;

  t0 = now_s ();
  search_kmers_on_ont_reads (reads, ctgs, khashs, okseqs, anchored, prefix, n_thread, k);
  t_search = now_s () - t0;

  kmer_stat (khashs, n_thread, k, "Scaffold");
  kmer_stat2 (anchored, n_thread, k, "ONT");

  /* the four integers, recomputed exactly as kmer.c:265-312 does */
  int64_t st[4] = {0, 0, 0, 0};
#ifndef GCG_SHIM_HARNESS
  for (i = 0; i < n_thread; ++i) {
    xh_t * h = khashs[i];
    st[0] += h->cnt;
    for (p = 0; p < (int64_t) h->cnt; ++p) if (h->pool[p].multi == 1) ++st[1];
    h = anchored[i]->hash;
    st[2] += h->cnt;
    for (p = 0; p < (int64_t) h->cnt; ++p) if (h->pool[p].multi == 1) ++st[3];
  }
#endif

  int64_t n_ctg_kmers = 0, n_ctg_bases = 0, n_ont_kmers = 0, n_ont_bases = 0, n_hits = 0;
  for (r = 0; r < mp_cnt (ctgs); ++r) {
    ctg_t * c = mp_at (ctg, ctgs, r);
    n_ctg_kmers += c->n_kmer;
    n_ctg_bases += c->seq->l;
  }
  for (r = 0; r < mp_cnt (reads); ++r) {
    rseq_t * rs = mp_at (rs, reads, r);
    n_ont_bases += rs->l;
    if (rs->l >= k) n_ont_kmers += rs->l - k + 1;
  }

  if (dump >= 1) {
    snprintf (path, sizeof path, "%s.hits.bin", prefix);
    FILE * fp = ckopen (path, "wb");
    for (r = 0; r < mp_cnt (okseqs); ++r) {
      okseq_t * ok = mp_at (okseq, okseqs, r);
      int64_t n = mp_cnt (ok->okmers);
      for (p = 0; p < n; ++p) {
        ont_kmer_t * o = mp_at (okmer, ok->okmers, p);
        if (o->kmer == NULL) continue;
        int32_t rec[4] = { (int32_t) r, o->ont_pos, o->kmer->tid, o->kmer->pos };
        uint16_t fl[2] = { o->kmer->flag, o->flag };
        fwrite (rec, 4, 4, fp);
        fwrite (fl, 2, 2, fp);
        ++n_hits;
      }
    }
    fclose (fp);
    snprintf (path, sizeof path, "%s.segs.bin", prefix);
    fp = ckopen (path, "wb");
    for (r = 0; r < mp_cnt (okseqs); ++r) {
      okseq_t * ok = mp_at (okseq, okseqs, r);
      int32_t ns = (int32_t) mp_cnt (ok->segs);
      fwrite (&ns, 4, 1, fp);
      for (p = 0; p < ns; ++p) {
        ont_seg_t * sg = mp_at (oseg, ok->segs, p);
        int32_t be[2] = { sg->beg, sg->end };
        fwrite (be, 4, 2, fp);
      }
    }
    fclose (fp);
  } else {
    for (r = 0; r < mp_cnt (okseqs); ++r) {
      okseq_t * ok = mp_at (okseq, okseqs, r);
      int64_t n = mp_cnt (ok->okmers);
      for (p = 0; p < n; ++p)
        if (mp_at (okmer, ok->okmers, p)->kmer != NULL) ++n_hits;
    }
  }

  if (dump >= 2) {
    snprintf (path, sizeof path, "%s.table.bin", prefix);
    FILE * fp = ckopen (path, "wb");
#ifndef GCG_SHIM_HARNESS
    for (i = 0; i < n_thread; ++i) {
      xh_t * h = khashs[i];
      for (p = 0; p < (int64_t) h->cnt; ++p) {
        kmer_t * km = (kmer_t *) h->pool[p].key;
        int32_t multi = h->pool[p].multi;
        fwrite (&km->kseq, 8, 1, fp);
        fwrite (&multi, 4, 1, fp);
        fwrite (&km->tid, 4, 1, fp);
        fwrite (&km->pos, 4, 1, fp);
        fwrite (&km->flag, 2, 1, fp);
        fwrite (&km->kmer_len, 2, 1, fp);
      }
    }
#endif
    fclose (fp);
    snprintf (path, sizeof path, "%s.hsid.bin", prefix);
    fp = ckopen (path, "wb");
    for (r = 0; r < mp_cnt (ctgs); ++r) {
      ctg_t * c = mp_at (ctg, ctgs, r);
      for (p = 0; p < c->n_kmer; ++p) fwrite (&c->kmers[p].hs_id, 4, 1, fp);
    }
    fclose (fp);
    snprintf (path, sizeof path, "%s.ctgk.bin", prefix);
    fp = ckopen (path, "wb");
    for (r = 0; r < mp_cnt (ctgs); ++r) {
      ctg_t * c = mp_at (ctg, ctgs, r);
      for (p = 0; p < c->n_kmer; ++p) {
        kmer_t * km = c->kmers + p;
        fwrite (&km->kseq, 8, 1, fp);
        fwrite (&km->tid, 4, 1, fp);
        fwrite (&km->pos, 4, 1, fp);
        fwrite (&km->flag, 2, 1, fp);
        fwrite (&km->kmer_len, 2, 1, fp);
      }
    }
    fclose (fp);
  }

  snprintf (path, sizeof path, "%s.json", prefix);
  FILE * fj = ckopen (path, "w");
  fprintf (fj,
      "{\"k\": %d, \"n_thread\": %d, \"n_contigs\": %ld, \"n_reads\": %ld, "
      "\"n_ctg_bases\": %ld, \"n_ctg_kmers\": %ld, \"n_ont_bases\": %ld, \"n_ont_kmers\": %ld, \"n_hits\": %ld, "
      "\"scaf_total\": %ld, \"scaf_unique\": %ld, \"ont_total\": %ld, \"ont_unique\": %ld, "
      "\"t_load_ctg\": %.6f, \"t_load_ont\": %.6f, \"t_okseq_init\": %.6f, "
      "\"t_chop\": %.6f, \"t_put\": %.6f, \"t_search\": %.6f}\n",
      k, n_thread, (long) mp_cnt (ctgs), (long) mp_cnt (reads),
      (long) n_ctg_bases, (long) n_ctg_kmers, (long) n_ont_bases, (long) n_ont_kmers, (long) n_hits,
      (long) st[0], (long) st[1], (long) st[2], (long) st[3],
      t_load_ctg, t_load_ont, t_okinit, t_chop, t_put, t_search);
  fclose (fj);

  hash_func_free ();
  return 0;
}
