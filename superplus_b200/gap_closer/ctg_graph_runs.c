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
 *
 * GC_SW_FILL=1 (opt-in): fix_ont1 (ctg_graph.c:669-755) fills a gap with the raw ONT bases between
 * two EXTRAPOLATED junctions — "the contig ends (from_ctg->seq->l - kmer->pos) read bases behind the left anchor",
 * which is only true when the read has no insertion or deletion between the anchor and the contig end
 * (combine_ont_info, ctg_graph.c:455-531; its reverse-strand branch also ignores the k-mer length, see the TODO at
 * ctg_graph.c:516).  This mode keeps the reference's choice of the ONT read per gap and replaces the extrapolation by
 * an alignment: the contig flank from the anchor to the contig end is aligned, end to end, against the read behind the
 * anchor (global start at the anchor, flank fully consumed, read end free: overhang strategy SWOS_INDEL), and the read
 * position where the flank ends is the junction; same on the right side, mirrored.  All gaps' alignments go to the GPU
 * as ONE batch (sw_align_batch -> gcg_sw_batch).  The FASTA differs from the reference's by design; every decision is
 * written to sw_fill.tsv so that the CPU oracle can recompute it with the reference's own DP (tests/test_gc_e2e_gpu.py).
 */
#define map_ont2contigs map_ont2contigs_reference
#define fix_ont1 fix_ont1_reference
#include "ctg_graph.c"
#undef map_ont2contigs
#undef fix_ont1

#include "gcg_bridge.h"
#include "sw.h"
#include "sw_batch.h"

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


/* ------------------------------------------------------------------ GC_SW_FILL (N4) ------------------------- */
#define SWF_MAX_FLANK 4000            /* longer flanks keep the reference's extrapolated junction */

static int
sw_fill_mode (void)
{
  const char * e = getenv ("GC_SW_FILL");
  return e != NULL && atoi (e) != 0;
}

typedef struct {
  ctg_t * from_ctg, * to_ctg;
  okseq_t * okseq;
  int32_t ont_id, fwd, k;
  int32_t pL, cL, pR, cR;             /* the two anchors as the okmers hold them (read coordinates of the read as stored) */
  int32_t L, a, c;                    /* read length, left flank length (contig end - cL), right head length (cR + k) */
  int32_t pLs, pRs;                   /* anchor starts in the strand the scaffold runs in (reverse reads: L - p - k) */
  int32_t jL_ref, jR_ref, jL, jR;     /* junctions in that strand: extrapolated, and refined by the alignment */
  int32_t wL0, wL1, wR0, wR1;         /* read windows of the two alignments */
  int32_t job_a, job_b;               /* indices into the batch, or -1 */
  int32_t score_a, score_b;
} swf_gap_t;

/* base code of the read at position x of the scaffold's strand (bio.h:22-24) */
static inline char
swf_read_code (const swf_gap_t * gp, int32_t x)
{
  const char * b = gp->okseq->seq->b;
  if (gp->fwd) return (char) (((unsigned char) b[x] >> 1) & 3);
  return (char) ((((unsigned char) b[gp->L - 1 - x] >> 1) & 3) ^ 2);
}

static inline char
swf_read_base (const swf_gap_t * gp, int32_t x)
{
  if (gp->fwd) return gp->okseq->seq->b[x];
  return base_rc_tbl[(int) gp->okseq->seq->b[gp->L - 1 - x]];
}

/* the anchor really is the same k bases on both sides when the read is taken in the scaffold's strand (a link whose
 * strands disagree keeps the reference's junction) */
static int
swf_anchor_ok (const swf_gap_t * gp, const ctg_t * ctg, int32_t cpos, int32_t ps)
{
  int32_t x;
  if (ps < 0 || ps + gp->k > gp->L || cpos < 0 || cpos + gp->k > ctg->seq->l) return 0;
  for (x = 0; x < gp->k; ++x)
    if (swf_read_code (gp, ps + x) != (char) (((unsigned char) ctg->seq->s[cpos + x] >> 1) & 3)) return 0;
  return 1;
}

/* the reference's choice of the read that fills the gap (combine_ont_info, ctg_graph.c:476-503): the median of the
 * extrapolated gap lengths, sorted with the reference's own comparator */
static ont2gap_t *
swf_pick_read (ctg_einfo_t * einfo, mp_t(okseq) * okseqs, mp_t(ont_gap) * ont_gaps, ctg_t * from_ctg)
{
  int32_t i, n = mp_cnt (einfo->ont_infos);
  mp_clear (ont_gap, ont_gaps, NULL);
  for (i = 0; i < n; ++i) {
    ont2gap_t * info = mp_at (o2g, einfo->ont_infos, i);
    okseq_t * okseq = mp_at (okseq, okseqs, info->ont_id);
    ont_kmer_t * lk = mp_at (okmer, okseq->okmers, info->left_okid), * rk = mp_at (okmer, okseq->okmers, info->right_okid);
    int32_t span = info->left_okid < info->right_okid ? rk->ont_pos - lk->ont_pos : lk->ont_pos - rk->ont_pos;
    ont_gap_t * og = mp_alloc (ont_gap, ont_gaps);
    og->info_idx = i;
    og->len = span - (from_ctg->seq->l - lk->kmer->pos) - rk->kmer->pos;
  }
  qsort (ont_gaps->pool, n, sizeof (ont_gap_t), cmp_ont_gap_len);
  return mp_at (o2g, einfo->ont_infos, mp_at (ont_gap, ont_gaps, n >> 1)->info_idx);
}

int
fix_ont1 (mp_t(okseq) * okseqs, ctg_graph_t * g, mp_t(ctg) * ctg_seqs, mp_t(sf) * scafs)
{
  int32_t s, j, n_sfs, n_gap = 0, m_gap = 0, acu_len, n_scaf = 0, x;
  int64_t n_job = 0, qtot = 0, ttot = 0, p;
  swf_gap_t * gaps = NULL;
  mp_t(ont_gap) * ont_gaps;
  int32_t mat[25];
  sw_t * sw;
  char * qbuf, * tbuf;
  int64_t * qoff, * toff;
  gcg_sw_result * res;
  uint32_t * pool = NULL;
  int64_t n_pool = 0;
  FILE * fp, * tsv;

  if (!sw_fill_mode ())
    return fix_ont1_reference (okseqs, g, ctg_seqs, scafs);

  /* ---- pass 1: every gap an ONT read spans: the reference's read, its two anchors, the two alignments to run */
  ont_gaps = mp_init (ont_gap, NULL, NULL);
  n_sfs = mp_cnt (scafs);
  for (s = 0; s < n_sfs; ++s) {
    scaf_t * sf = mp_at (sf, scafs, s);
    for (j = 1; j < mp_cnt (sf->ctg_ids); ++j) {
      ctg_t * f_ctg = mp_at (ctg, ctg_seqs, sf->ctg_ids->arr[j - 1]), * t_ctg = mp_at (ctg, ctg_seqs, sf->ctg_ids->arr[j]);
      dg_edge_t * edge = digraph_edge_between (g->g, f_ctg->node_id, t_ctg->node_id);
      ctg_einfo_t * einfo;
      ont2gap_t * info;
      ont_kmer_t * lk, * rk;
      swf_gap_t * gp;
      if (edge == NULL)
        err_mesg ("no edge between %dth contig and %dth contig!", f_ctg->id, t_ctg->id);
      einfo = edge->info;
      if (mp_cnt (einfo->ont_infos) <= 0) continue;
      mp_resize (ont_gap, ont_gaps, mp_cnt (einfo->ont_infos));
      info = swf_pick_read (einfo, okseqs, ont_gaps, f_ctg);
      if (n_gap == m_gap) { m_gap = m_gap ? 2 * m_gap : 1024; gaps = (swf_gap_t *) realloc (gaps, (size_t) m_gap * sizeof (swf_gap_t)); if (gaps == NULL) err_mesg ("out of memory"); }
      gp = gaps + n_gap++;
      memset (gp, 0, sizeof *gp);
      gp->from_ctg = f_ctg; gp->to_ctg = t_ctg;
      gp->ont_id = info->ont_id;
      gp->okseq = mp_at (okseq, okseqs, info->ont_id);
      lk = mp_at (okmer, gp->okseq->okmers, info->left_okid); rk = mp_at (okmer, gp->okseq->okmers, info->right_okid);
      gp->fwd = info->left_okid < info->right_okid;
      gp->k = lk->kmer->kmer_len;
      gp->pL = lk->ont_pos; gp->cL = lk->kmer->pos; gp->pR = rk->ont_pos; gp->cR = rk->kmer->pos;
      gp->L = gp->okseq->seq->l;
      gp->a = f_ctg->seq->l - gp->cL;
      gp->c = gp->cR + gp->k;
      gp->pLs = gp->fwd ? gp->pL : gp->L - gp->pL - gp->k;
      gp->pRs = gp->fwd ? gp->pR : gp->L - gp->pR - gp->k;
      gp->jL_ref = gp->pLs + gp->a;                      /* "the contig ends a read bases behind the anchor" */
      gp->jR_ref = gp->pRs - gp->cR;                     /* "the next contig starts cR read bases before the anchor" */
      gp->jL = gp->jL_ref; gp->jR = gp->jR_ref;
      gp->job_a = gp->job_b = -1;
      gp->score_a = gp->score_b = 0;
      if (gp->a >= 1 && gp->a <= SWF_MAX_FLANK && swf_anchor_ok (gp, f_ctg, gp->cL, gp->pLs)) {
        gp->wL0 = gp->pLs;
        gp->wL1 = gp->pLs + gp->a + gp->a / 4 + 32;
        if (gp->wL1 > gp->L) gp->wL1 = gp->L;
        gp->job_a = (int32_t) n_job++;
        qtot += gp->a; ttot += gp->wL1 - gp->wL0;
      }
      if (gp->c >= 1 && gp->c <= SWF_MAX_FLANK && swf_anchor_ok (gp, t_ctg, gp->cR, gp->pRs)) {
        gp->wR1 = gp->pRs + gp->k;
        gp->wR0 = gp->wR1 - (gp->c + gp->c / 4 + 32);
        if (gp->wR0 < 0) gp->wR0 = 0;
        gp->job_b = (int32_t) n_job++;
        qtot += gp->c; ttot += gp->wR1 - gp->wR0;
      }
    }
  }

  /* ---- pass 2: one batch.  Job a: flank (query) against the read behind the anchor; job b: the same thing seen
   *      from the right anchor's END backwards (both sequences reversed), so that in both the alignment starts at the
   *      anchor, consumes the whole flank and ends wherever in the read the flank ends (SWOS_INDEL: affine borders = a
   *      global start; end cell = best of the last query column, sw.c:259-264) */
  qbuf = (char *) ckalloc (qtot + 1, 1); tbuf = (char *) ckalloc (ttot + 1, 1);
  qoff = (int64_t *) ckalloc (n_job + 1, sizeof (int64_t)); toff = (int64_t *) ckalloc (n_job + 1, sizeof (int64_t));
  res = (gcg_sw_result *) ckalloc (n_job + 1, sizeof (gcg_sw_result));
  for (s = 0, qtot = ttot = 0; s < n_gap; ++s) {
    swf_gap_t * gp = gaps + s;
    if (gp->job_a >= 0) {
      qoff[gp->job_a] = qtot; toff[gp->job_a] = ttot;
      for (x = 0; x < gp->a; ++x) qbuf[qtot++] = (char) (((unsigned char) gp->from_ctg->seq->s[gp->cL + x] >> 1) & 3);
      for (x = gp->wL0; x < gp->wL1; ++x) tbuf[ttot++] = swf_read_code (gp, x);
    }
    if (gp->job_b >= 0) {
      qoff[gp->job_b] = qtot; toff[gp->job_b] = ttot;
      for (x = gp->c - 1; x >= 0; --x) qbuf[qtot++] = (char) (((unsigned char) gp->to_ctg->seq->s[x] >> 1) & 3);
      for (x = gp->wR1 - 1; x >= gp->wR0; --x) tbuf[ttot++] = swf_read_code (gp, x);
    }
  }
  qoff[n_job] = qtot; toff[n_job] = ttot;
  for (s = 0; s < 25; ++s) mat[s] = (s / 5 == s % 5) ? 1 : -5;                 /* gc_graph.c:74-77,87-107 */
  sw = sw_init ();
  sw_set_parameter (sw, 5, mat, 2, 1, 2, 1, SWOS_INDEL);
  if (n_job > 0)
    sw_align_batch (sw, n_job, qbuf, qoff, tbuf, toff, res, &pool, &n_pool);
  for (s = 0; s < n_gap; ++s) {
    swf_gap_t * gp = gaps + s;
    if (gp->job_a >= 0) { gp->jL = gp->wL0 + res[gp->job_a].bt_tidx; gp->score_a = res[gp->job_a].score; }
    if (gp->job_b >= 0) { gp->jR = gp->wR1 - res[gp->job_b].bt_tidx; gp->score_b = res[gp->job_b].score; }
  }

  /* ---- pass 3: the scaffolds (fix_ont1's own loop, ctg_graph.c:703-735), gaps filled between the refined junctions */
  fp = ckopen ("gc_fix1.fa", "w");
  tsv = ckopen ("sw_fill.tsv", "w");
  fprintf (tsv, "#S\tscaffold\tfirst_ctg | N\tfrom_ctg\tto_ctg\tn_len | G\t<the columns below>\n");
  fprintf (tsv, "#G\tgap\tfrom_ctg\tto_ctg\tont\tfwd\tk\tpL\tcL\tpR\tcR\tread_len\tflank_a\thead_c\tjL_ref\tjR_ref\tjL\tjR\twL0\twL1\twR0\twR1\tscore_a\tscore_b\n");
  for (s = 0, p = 0; s < n_sfs; ++s) {
    scaf_t * sf = mp_at (sf, scafs, s);
    fprintf (tsv, "S\t%d\t%d\n", n_scaf, mp_at (ctg, ctg_seqs, sf->ctg_ids->arr[0])->id);
    fprintf (fp, ">%d\n", n_scaf++);
    acu_len = 0;
    contig_seq_dump (fp, mp_at (ctg, ctg_seqs, sf->ctg_ids->arr[0]), NULL, &acu_len);
    for (j = 1; j < mp_cnt (sf->ctg_ids); ++j) {
      ctg_t * f_ctg = mp_at (ctg, ctg_seqs, sf->ctg_ids->arr[j - 1]), * t_ctg = mp_at (ctg, ctg_seqs, sf->ctg_ids->arr[j]);
      ctg_einfo_t * einfo = digraph_edge_between (g->g, f_ctg->node_id, t_ctg->node_id)->info;
      ctg_vinfo_t * vinfo;
      if (mp_cnt (einfo->ont_infos) <= 0) {
        vinfo = create_origin_N_gap_info (g, t_ctg->l_pre_gap);
        fprintf (tsv, "N\t%d\t%d\t%d\n", f_ctg->id, t_ctg->id, t_ctg->l_pre_gap);
      } else {
        swf_gap_t * gp = gaps + p;
        int32_t len = gp->jR - gp->jL;
        vinfo = bmp_alloc (ctg_vinfo, g->vinfos);
        vinfo->flag = CTG_GAP;
        vinfo->l_gap = len;
        if (len <= 0)
          vinfo->flag |= CTG_NEGA_GAP;
        else {
          str_resize (vinfo->gap_seq, len);
          vinfo->gap_seq->l = len;
          for (x = 0; x < len; ++x) vinfo->gap_seq->s[x] = swf_read_base (gp, gp->jL + x);
          vinfo->gap_seq->s[len] = '\0';
        }
        fprintf (tsv, "G\t%ld\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\n", (long) p, f_ctg->id, t_ctg->id, gp->ont_id, gp->fwd, gp->k,
                 gp->pL, gp->cL, gp->pR, gp->cR, gp->L, gp->a, gp->c, gp->jL_ref, gp->jR_ref, gp->jL, gp->jR, gp->wL0, gp->wL1, gp->wR0, gp->wR1, gp->score_a, gp->score_b);
        ++p;
      }
      contig_seq_dump (fp, t_ctg, vinfo, &acu_len);
    }
    if (acu_len % 60 != 0)
      fprintf (fp, "\n");
  }
  fclose (fp);
  fclose (tsv);
  mp_free (ont_gap, ont_gaps, NULL);
  if (pool) gcg_free (pool);
  sw_free (sw);
  free (gaps); free (qbuf); free (tbuf); free (qoff); free (toff); free (res);
  return 0;
}
