/* ctg_graph_runs.c — the reference's ctg_graph.c, UNMODIFIED and compiled where it lies (this file only
 * includes it), with one entry point redirected: map_ont2contigs (ctg_graph.c:575-667).
 *
 * Default: map_ont2contigs is the reference's own function, byte for byte (map_ont2contigs_reference below).
 *
 * GC_RUNS=1 (opt-in): the reference's map_ont2contigs walks a dense array of one 16-byte
 * ont_kmer_t per ONT BASE (80 GB at BASELINE configs[4]) only to cut every read's anchors into runs of
 * consecutive anchors on one contig, count the two strand directions per run and remember the first and last
 * anchor of the majority direction (ont_node_init, ctg_graph.c:93-181).  That reduction now happens on the
 * GPU (superplus_b200/csrc/runs.cu, gcg_search_runs): the bridge holds one run record per run, and every read's
 * okmers[] holds just the two boundary anchors of each run, okmers[2q] and okmers[2q+1].  The loop below turns
 * the run records into the same ont_node_t sequence — same pool, same allocation order, the same fields written
 * in the same cases, because add_ont_link2graph (ctg_graph.c:232-360) also reads the STALE fields of nodes that
 * ont_node_init marked deleted — and hands them to the reference's own, unmodified
 * add_ont_link2graph.  Everything downstream (ont_edge_create, combine_ont_info, fix_ont1) dereferences okmers[]
 * only through the node's beg_okid / end_okid, which here index the compact array.
 */
#define map_ont2contigs map_ont2contigs_reference
#include "ctg_graph.c"
#undef map_ont2contigs

#include "gcg_bridge.h"

/* what ont_node_init (ctg_graph.c:93-181) leaves in a node, from a run record; `base` = index of the run's first
 * boundary anchor in the read's compact okmers[] */
static void
ont_node_from_run (ont_node_t * ont_node, int32_t node_id, const gcg_run * run, int64_t ont_id, int64_t base)
{
  if (run->n_fwd > 20 * run->n_bwd) {
    ont_node->direct = FORW;
    ont_node->n_okmers = run->n_fwd;
  } else if (run->n_bwd > 20 * run->n_fwd) {
    ont_node->direct = BACK;
    ont_node->n_okmers = run->n_bwd;
  } else {
    ont_node->flag = CTG_NODE_DEL;        /* every other field keeps what the pool slot held before (ctg_graph.c:175-177) */
    return;
  }
  ont_node->tid = run->tid;
  ont_node->nid = node_id;
  ont_node->flag = 0;
  ont_node->oid = ont_id;
  ont_node->beg_okid = base;
  ont_node->end_okid = base + 1;
}

int
map_ont2contigs (mp_t(okseq) * okseqs, ctg_graph_t * g, mp_t(ctg) * ctg_seqs)
{
  int64_t i, q, n_okseqs;
  FILE * fp;
  FILE * vfp;
  ctg_t * seq;
  okseq_t * okseq;
  ont_node_t * ont_node;
  mp_t(ont_node) * ont_nodes;
  gcg_bridge_t * br = gcg_bridge_peek ();

  if (!br->runs_mode)
    return map_ont2contigs_reference (okseqs, g, ctg_seqs);

  fp = ckopen ("ont_link.txt", "w");
  vfp = ckopen ("valid_ont_link.txt", "w");
  n_okseqs = mp_cnt (okseqs);
  ont_nodes = mp_init (ont_node, NULL, NULL);
  for (i = 0; i < n_okseqs; ++i) {
    const int64_t q0 = br->run_off[i], q1 = br->run_off[i + 1];
    okseq = mp_at (okseq, okseqs, i);
    if (q0 == q1) {
      okseq->flag |= ONT_NO_ANK;
      continue;
    }
    mp_clear (ont_node, ont_nodes, NULL);
    for (q = q0; q < q1; ++q) {
      const gcg_run * run = br->runs + q;
      seq = mp_at (ctg, ctg_seqs, run->tid);
      ont_node = mp_alloc (ont_node, ont_nodes);
      ont_node_from_run (ont_node, seq->node_id, run, i, 2 * (q - q0));
    }
    fprintf (vfp, "> ONT %ld\n", i);
    fprintf (fp, "> ONT %ld\n", i);
    okseq->flag |= add_ont_link2graph (g, ont_nodes, vfp, fp);
  }
  fclose (fp);
  fclose (vfp);
  return 0;
}
