/* kmer.c — replacement for gap_closer/kmer.c: same prototypes (kmer.h:51-63), the work is done
 * by libgcgpu.so on the B200.
 *
 *   chop_contig_seqs2kmers   (reference kmer.c:155-184 / 37-121)  -> gcg_seqs_upload + gcg_chop_contigs
 *   put_contig_kmers2hashs   (reference kmer.c:187-213 / 124-152) -> gcg_table_build_seqs
 *   kmer_stat / kmer_stat2   (reference kmer.c:265-312)           -> gcg_table_stats
 *   kmer_hash_init/clear/free (reference kmer.c:232-262)          -> host handles + device table lifetime
 *
 * `xh_t ** khashs` keeps its type: it is created here, passed around by main.c and consumed only
 * by functions of this directory, so the n_thread xh_t objects stay empty host tables (they keep
 * lfr.c / gc_graph.c linkable) while the real table lives in HBM.  n_thread no longer partitions
 * anything (the partition id is unobservable); it sets the number of host threads
 * used for staging and back-fill.
 */
#include <time.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "mp.h"
#include "str.h"
#include "hash.h"
#include "utils.h"
#include "contig.h"
#include "kmer.h"
#include "gcg_bridge.h"

int
chop_contig_seqs2kmers (mp_t(ctg) * seqs, int n_thread, int kmer_len)
{
  time_t time_beg;
  int64_t i, n;
  const char ** ptrs;
  int32_t * lens;
  int32_t * n_kmer;
  void ** outs;
  ctg_t * c;
  gcg_bridge_t * br;

  time (&time_beg);
  br = gcg_bridge ();
  br->kmer_len = kmer_len;
  br->n_thread = n_thread;
  GCG_CK (gcg_set_host_threads (br->ctx, n_thread > 0 ? n_thread : 1));

  n = mp_cnt (seqs);
  ptrs = (const char **) ckalloc (n + 1, sizeof (char *));
  lens = (int32_t *) ckalloc (n + 1, sizeof (int32_t));
  n_kmer = (int32_t *) ckalloc (n + 1, sizeof (int32_t));
  outs = (void **) ckalloc (n + 1, sizeof (void *));
  for (i = 0; i < n; ++i) {
    c = mp_at (ctg, seqs, i);
    ptrs[i] = c->seq->s;
    lens[i] = c->seq->l;
    outs[i] = (void *) c->kmers;       /* room for seq.l records (contig.c:203-208) */
  }

  gcg_bridge_drop_table ();
  gcg_bridge_drop_contigs ();
  GCG_CK (gcg_seqs_upload (br->ctx, ptrs, lens, n, &br->contigs));
  if (gcg_bridge_runs_mode () || gcg_bridge_sparse_kmers ()) {
    /* GC_SPARSE_KMERS (ctg_graph.c untouched): the back-fill of the search writes the kmer_t record of
     * every contig position an anchor points at (ont.c of this directory), nothing else ever reads ctg->kmers[].
     * GC_RUNS (N2 + N3): nothing reads ctg->kmers[] — the anchors the consumers look at carry their own kmer_t
     * records (ont.c of this directory) — so the 24 bytes per contig base stay untouched zero pages */
    for (i = 0; i < n; ++i) n_kmer[i] = lens[i] >= kmer_len ? lens[i] - kmer_len + 1 : 0;
  } else
    GCG_CK (gcg_chop_contigs (br->ctx, br->contigs, kmer_len, n_thread > 0 ? n_thread : 1, outs, n_kmer));
  for (i = 0; i < n; ++i)
    mp_at (ctg, seqs, i)->n_kmer += n_kmer[i];   /* the reference appends from the cleared n_kmer (kmer.c:79,103) */

  fprintf (stdout, "  chop kmers cost: %lds\n", time (NULL) - time_beg);

  free (ptrs); free (lens); free (n_kmer); free (outs);
  return 0;
}

int
put_contig_kmers2hashs (xh_t ** khashs, mp_t(ctg) * seqs, int n_thread)
{
  time_t time_beg;
  gcg_bridge_t * br;

  time (&time_beg);
  br = gcg_bridge ();
  if (br->contigs == NULL)
    err_mesg ("[%s] chop_contig_seqs2kmers has not been called", __func__);
  gcg_bridge_drop_table ();
  GCG_CK (gcg_table_build_seqs (br->ctx, br->contigs, br->kmer_len, &br->table));
  GCG_CK (gcg_sync (br->ctx));
  gcg_bridge_replicate_table ();       /* GC_DEVICES: one replica per further GPU */

  fprintf (stdout, "  hash kmers cost: %lds\n", time (NULL) - time_beg);
  return 0;
}

void
kmer_set_dump (xh_t * khash, int kmer_len)
{
  /* debugging helper of the reference (kmer.c:216-229), no caller: dumps the device table */
  gcg_bridge_t * br = gcg_bridge ();
  int64_t i, n;
  uint64_t * key; int32_t * multi, * tid, * pos; uint8_t * rev;
  char * kmer_seq;

  if (br->table == NULL) return;
  n = gcg_table_size (br->ctx, br->table);
  key = (uint64_t *) ckalloc (n + 1, 8); multi = (int32_t *) ckalloc (n + 1, 4);
  tid = (int32_t *) ckalloc (n + 1, 4); pos = (int32_t *) ckalloc (n + 1, 4); rev = (uint8_t *) ckalloc (n + 1, 1);
  GCG_CK (gcg_table_dump (br->ctx, br->table, n, key, multi, tid, pos, rev));
  kmer_seq = ALLOC_LINE;
  for (i = 0; i < n; ++i) {
    kseq1_t k = key[i];
    kseq12seq (&k, kmer_seq, kmer_len);
    printf (">%lx\n%s\t%c\n", k, kmer_seq, rev[i] ? 'R' : 'F');
  }
  free (kmer_seq); free (key); free (multi); free (tid); free (pos); free (rev);
}

/* The handles returned by kmer_hash_init are empty host tables: the contig k-mers live in HBM.  Unchanged reference
 * code that probes them on the host (lfr.c ankor_lfr_reads -> _xh_set_search3; disabled in the reference's main.c) would
 * silently find nothing — so their hash function aborts instead. */
static uint64_t
device_table_hash_func (const void * key)
{
  err_mesg ("[kmer] the contig k-mer tables of this build live in GPU memory: host-side probes of ctg_khashs (lfr.c) are not supported");
  return 0;
}

xh_t **
kmer_hash_init (int n_thread)
{
  int i;
  xh_t ** khashs;

  khashs = (xh_t **) ckalloc (n_thread, sizeof (xh_t *));
  for (i = 0; i < n_thread; ++i)
    khashs[i] = _xh_init (256, 0.75, device_table_hash_func, kmer_is_equal);
  /* the tables live in HBM; start opening the device now, in the background, while the unchanged
   * loaders read the scaffolds and the ONT reads (main.c:149-156) */
  gcg_bridge_warmup ();
  return khashs;
}

void
kmer_hash_clear (xh_t ** khashs, int n_thread)
{
  int i;
  for (i = 0; i < n_thread; ++i)
    _xh_clear (khashs[i]);
  gcg_bridge_drop_table ();
}

void
kmer_hash_free (xh_t ** khashs, int n_thread)
{
  int i;
  for (i = 0; i < n_thread; ++i)
    _xh_free (khashs[i]);
  free (khashs);
  gcg_bridge_shutdown ();
}

static void
print_stat (int64_t total, int64_t uniq, int kmer_len, const char * mesg)
{
  printf ("\n  >>> %s Kmer%d Stat <<<\n", mesg, kmer_len);
  printf ("  total kmer count: %ld\n", total);
  printf ("  unique kmer count: %ld\n", uniq);
}

void
kmer_stat (xh_t ** khashs, int n_thread, int kmer_len, const char * mesg)
{
  int64_t st[4] = {0, 0, 0, 0};
  gcg_bridge_t * br = gcg_bridge ();
  if (br->table) GCG_CK (gcg_table_stats (br->ctx, br->table, st));
  print_stat (st[0], st[1], kmer_len, mesg);
}

void
kmer_stat2 (xh_set_t(kmer) ** khashs, int n_thread, int kmer_len, const char * mesg)
{
  int64_t st[4] = {0, 0, 0, 0};
  gcg_bridge_t * br = gcg_bridge ();
  if (br->table) GCG_CK (gcg_table_stats (br->ctx, br->table, st));
  print_stat (st[2], st[3], kmer_len, mesg);
}
