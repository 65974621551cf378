/* gcg_bridge.h — process-wide handle on libgcgpu.so shared by the replacement files of this
 * directory (kmer.c, ont.c, sw.c).  Not part of the reference's API surface. */
#ifndef GCG_BRIDGE_H
#define GCG_BRIDGE_H

#include "gcgpu.h"

struct ctg_s;

#define GCG_BRIDGE_MAX_DEV 16

typedef struct {
  gcg_ctx * ctx;          /* primary context: device $GC_DEVICE (default 0), or the first of $GC_DEVICES */
  gcg_seqs * contigs;     /* 2-bit contigs of the last chop_contig_seqs2kmers */
  gcg_table * table;      /* table of the last put_contig_kmers2hashs */
  int kmer_len;
  int n_thread;
  int searched_k;         /* k-mer length of the first search_kmers_on_ont_reads (0 = none yet) */
  /* GC_RUNS=1 (opt-in, SURVEY 8f rows N2 + N3): the device reduces every read's anchors to the run records
   * map_ont2contigs needs (ctg_graph.c:600-656); runs[run_off[r] .. run_off[r+1]) belong to read r.  The two boundary
   * anchors of a run are materialised as okmers[2q], okmers[2q+1] of the read (q = index of the run in the read) with
   * their kmer_t records in run_kmers[]; neither the dense okmers[] nor ctg->kmers[] is ever written. */
  int runs_mode;
  gcg_run * runs;
  int64_t * run_off;
  int64_t n_run;
  struct kmer_s * run_kmers;
  /* GC_DEVICES=0,1,... (or "all"): the ONT reads are sharded by batch over these GPUs, every one
   * holding a replica of the table (replicated layout).  ctxs[0] == ctx, replicas[0] == NULL. */
  int n_dev;
  int devs[GCG_BRIDGE_MAX_DEV];
  gcg_ctx * ctxs[GCG_BRIDGE_MAX_DEV];
  gcg_table * replicas[GCG_BRIDGE_MAX_DEV];
} gcg_bridge_t;

/* rseq_fast.c: the 2-bit words the FASTQ loader packed while loading `set` (0 = none: not that set, or switched off) */
int rseq_packed_lookup (const void * set, int64_t n_reads, const uint64_t ** words, const int64_t ** woff);
int gcg_bridge_sparse_kmers (void);          /* GC_SPARSE_KMERS set and not 0: ctg->kmers[] is written only where an anchor points */
int gcg_bridge_runs_mode (void);             /* GC_RUNS set and not 0 (no device is opened by asking) */
gcg_bridge_t * gcg_bridge_peek (void);       /* the handle as it is: no context is created */
gcg_bridge_t * gcg_bridge (void);            /* lazily creates the context; aborts via err_mesg on failure */
void gcg_bridge_warmup (void);               /* start opening the device on a helper thread (joined by gcg_bridge) */
void gcg_bridge_drop_table (void);             /* the table and its replicas */
void gcg_bridge_replicate_table (void);        /* clone the table onto the other devices (no-op with one) */
void gcg_bridge_drop_contigs (void);
void gcg_bridge_shutdown (void);
#define GCG_CK(call) do { int rc_ = (call); if (rc_ != 0) err_mesg ("[%s] %s failed (%d): %s", __func__, #call, rc_, gcg_last_error ()); } while (0)

#endif
