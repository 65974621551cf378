/* sw.c — replacement for gap_closer/sw.c: same prototypes (sw.h:66-71) and the same sw_t result
 * fields (sw.h:44-59); the DP fill, end-cell choice, traceback and CIGAR run on the B200
 * (libgcgpu: sw_fill_*_kernel, sw_cigar_kernel).
 *
 *   sw_init            (reference sw.c:345-366)
 *   sw_set_parameter   (reference sw.c:369-397 + init_matrix_values sw.c:61-110)
 *   sw_align           (reference sw.c:400-414 -> score_matrix_init sw.c:134-162, align_core sw.c:165-335)
 *
 * The reference aligner is stateful in one observable way: the border scores of its matrix are
 * written when parameters are set or the matrix grows and are NOT rewritten when a later
 * sw_set_parameter picks SWOS_SOFTCLIP (sw.c:93-94).  That state (and the growth rule that resets
 * it, sw.c:140-158) is tracked here and handed to the kernels as gcg_sw_params.border_*.
 * No score matrix is allocated on the host: sw_t.sm points to this file's private state.
 *
 * Traceback flavour: GCG_SW_ASIS reproduces the shipped traceback bit for bit (the current cell
 * is never re-fetched, sw.c:289-319); sw_set_traceback_mode(GCG_SW_FIXED) or GC_SW_MODE=fixed
 * selects the re-fetching variant.  sw_align_batch is the throughput entry point (one launch for
 * many pairs); sw_align is the reference's one-pair call on top of it.
 */
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "sw.h"
#include "utils.h"
#include "cigar.h"
#include "gcg_bridge.h"
#include "sw_batch.h"

#define SW_UNDO 0
#define SW_DONE 1
#define FIRST_QRY_CAP 128      /* reference growth state: matrix starts 512 x 128 (sw.c:45-48,352-360) */
#define FIRST_TGT_CAP 512
#define FIRST_QRY_NBITS 7

typedef struct {
  int border_kind;                       /* 0 zeros, 1 affine */
  int32_t b_del_o, b_del_e, b_ins_o, b_ins_e;
} sw_private_t;

static int g_traceback_mode = -1;

static int
traceback_mode (void)
{
  if (g_traceback_mode < 0) {
    const char * e = getenv ("GC_SW_MODE");
    g_traceback_mode = (e && strcmp (e, "fixed") == 0) ? GCG_SW_FIXED : GCG_SW_ASIS;
  }
  return g_traceback_mode;
}

void
sw_set_traceback_mode (int mode)
{
  g_traceback_mode = (mode == GCG_SW_FIXED) ? GCG_SW_FIXED : GCG_SW_ASIS;
}

static sw_private_t * priv (sw_t * sw) { return (sw_private_t *) sw->sm; }

/* what init_matrix_values leaves behind for the current strategy / penalties */
static void
refresh_borders (sw_t * sw, int after_calloc)
{
  sw_private_t * p = priv (sw);
  if (sw->overhang_strategy == SWOS_SOFTCLIP) {
    if (after_calloc) p->border_kind = 0;       /* fresh zeroed matrix, early return keeps it zero */
    return;                                     /* otherwise: stale borders survive */
  }
  p->border_kind = 1;
  p->b_del_o = sw->del_o; p->b_del_e = sw->del_e;
  p->b_ins_o = sw->ins_o; p->b_ins_e = sw->ins_e;
}

sw_t *
sw_init (void)
{
  sw_t * sw = (sw_t *) ckalloc (1, sizeof (sw_t));
  sw->sm = (sw_cell_t *) ckalloc (1, sizeof (sw_private_t));
  sw->m_qry = FIRST_QRY_CAP;
  sw->m_tgt = FIRST_TGT_CAP;
  sw->qry_nbits = FIRST_QRY_NBITS;
  sw->sm_vol = FIRST_TGT_CAP * FIRST_QRY_CAP;
  sw->cigar = cigar_init ();
  sw->type_c = -1;
  sw->mat = NULL;
  sw->overhang_strategy = SWOS_SOFTCLIP;
  return sw;
}

void
sw_set_parameter (sw_t * sw, int type_c, int32_t * mat,
    int32_t del_o, int32_t del_e, int32_t ins_o, int32_t ins_e, int overhang_strategy)
{
  if (type_c <= 0)
    err_mesg ("[%s] invalid 'type_c'!", __func__);
  if (mat == NULL)
    err_mesg ("[%s] mat==NULL!", __func__);
  if (type_c > 8)
    err_mesg ("[%s] type_c=%d: the device aligner supports alphabets of at most 8 symbols", __func__, type_c);

  if (sw->type_c != type_c) {
    if (sw->type_c > 0)
      free (sw->mat);
    sw->type_c = type_c;
    sw->mat = (int32_t *) ckmalloc (type_c * type_c * sizeof (int32_t));
  }
  memcpy (sw->mat, mat, type_c * type_c * sizeof (int32_t));
  sw->del_o = del_o; sw->del_e = del_e;
  sw->ins_o = ins_o; sw->ins_e = ins_e;
  sw->overhang_strategy = overhang_strategy;
  refresh_borders (sw, 0);
}

/* the reference grows its matrix by doubling and then re-initialises the borders (sw.c:134-162) */
static void
track_growth (sw_t * sw, int32_t qry_len, int32_t tgt_len)
{
  int grown = 0;
  while (sw->m_qry < qry_len + 2) { sw->m_qry <<= 1; sw->qry_nbits += 1; grown = 1; }
  while (sw->m_tgt < tgt_len + 2) { sw->m_tgt <<= 1; grown = 1; }
  if (grown) {
    sw->sm_vol = sw->m_tgt << sw->qry_nbits;
    refresh_borders (sw, 1);
  }
}

static void
fill_params (sw_t * sw, gcg_sw_params * P)
{
  sw_private_t * p = priv (sw);
  int i;
  if (sw->type_c <= 0 || sw->mat == NULL)
    err_mesg ("[sw_align] sw_set_parameter has not been called");
  memset (P, 0, sizeof (*P));
  P->type_c = sw->type_c;
  P->del_o = sw->del_o; P->del_e = sw->del_e; P->ins_o = sw->ins_o; P->ins_e = sw->ins_e;
  P->strategy = sw->overhang_strategy;
  P->border_kind = p->border_kind;
  P->b_del_o = p->b_del_o; P->b_del_e = p->b_del_e; P->b_ins_o = p->b_ins_o; P->b_ins_e = p->b_ins_e;
  for (i = 0; i < sw->type_c * sw->type_c; ++i) P->mat[i] = sw->mat[i];
}

int
sw_align_batch (sw_t * sw, int64_t n, const char * qry, const int64_t * qoff, const char * tgt, const int64_t * toff,
    gcg_sw_result * results, uint32_t ** cigar_pool, int64_t * n_cigar_pool)
{
  gcg_sw_params P;
  int64_t p;
  int32_t max_q = 0, max_t = 0;
  gcg_bridge_t * br = gcg_bridge ();

  /* the batch shares one border state, as n consecutive sw_align calls would once the matrix has
   * reached its final size: grow first for the largest pair */
  for (p = 0; p < n; ++p) {
    if (qoff[p + 1] - qoff[p] > max_q) max_q = (int32_t) (qoff[p + 1] - qoff[p]);
    if (toff[p + 1] - toff[p] > max_t) max_t = (int32_t) (toff[p + 1] - toff[p]);
  }
  track_growth (sw, max_q, max_t);
  fill_params (sw, &P);
  /* GC_DEVICES: the pairs are sharded over the devices of the bridge (contiguous ranges of equal cells) */
  if (br->n_dev > 1)
    GCG_CK (gcg_sw_batch_multi (br->ctxs, br->n_dev, &P, traceback_mode (), qry, qoff, tgt, toff, n, results, cigar_pool, n_cigar_pool));
  else
    GCG_CK (gcg_sw_batch (br->ctx, &P, traceback_mode (), qry, qoff, tgt, toff, n, results, cigar_pool, n_cigar_pool));
  return 0;
}

int
sw_align (sw_t * sw, int32_t qry_len, char * qry, int32_t tgt_len, char * tgt)
{
  gcg_sw_params P;
  gcg_sw_result r;
  uint32_t * pool = NULL;
  int64_t n_pool = 0, qoff[2], toff[2];
  int32_t i;
  gcg_bridge_t * br = gcg_bridge ();

  sw->status = SW_UNDO;
  track_growth (sw, qry_len, tgt_len);
  sw->score = -1;
  sw->alignment_offset = -1;
  cigar_clear (sw->cigar);

  fill_params (sw, &P);
  qoff[0] = 0; qoff[1] = qry_len;
  toff[0] = 0; toff[1] = tgt_len;
  GCG_CK (gcg_sw_batch (br->ctx, &P, traceback_mode (), qry, qoff, tgt, toff, 1, &r, &pool, &n_pool));

  sw->score = r.score;
  sw->alignment_offset = r.alignment_offset;
  sw->has_softclip = r.has_softclip;
  for (i = 0; i < r.n_cigar; ++i)
    cigar_add (sw->cigar, pool[r.cigar_off + i]);
  gcg_free (pool);

  sw->status = SW_DONE;
  return 0;
}

void
sw_free (sw_t * sw)
{
  if (sw == NULL) return;
  if (sw->type_c > 0) free (sw->mat);
  free (sw->sm);
  cigar_free (sw->cigar);
  free (sw);
}
