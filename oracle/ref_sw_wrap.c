/* TEST INFRASTRUCTURE — thin wrapper linked against the *reference's own* sw.c / cigar.c
 * (compiled from /root/reference/gap_closer by oracle/build_ref.sh).  It adds no algorithm:
 * every call forwards to sw_init / sw_set_parameter / sw_align (sw.h:66-71) and reads back the
 * result fields of sw_t (sw.h:44-59).  Loaded through ctypes by tests/ and by bench.py's
 * cpu_baseline / --impl reference legs only.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>

#include "sw.h"
#include "cigar.h"

void * refsw_new (void) { return (void *) sw_init (); }

void refsw_set (void * h, int type_c, int32_t * mat, int32_t del_o, int32_t del_e,
                int32_t ins_o, int32_t ins_e, int strategy)
{
  sw_set_parameter ((sw_t *) h, type_c, mat, del_o, del_e, ins_o, ins_e, strategy);
}

/* out[0]=score out[1]=alignment_offset out[2]=has_softclip out[3]=n_cigar ; returns n_cigar */
int refsw_align (void * h, int32_t qlen, char * q, int32_t tlen, char * t,
                 int64_t * out, uint32_t * cigar_buf, int32_t cigar_cap)
{
  sw_t * sw = (sw_t *) h;
  int32_t n;
  sw_align (sw, qlen, q, tlen, t);
  out[0] = sw->score;
  out[1] = sw->alignment_offset;
  out[2] = sw->has_softclip;
  out[3] = sw->cigar->n;
  n = sw->cigar->n < cigar_cap ? sw->cigar->n : cigar_cap;
  if (n > 0) memcpy (cigar_buf, sw->cigar->c, (size_t) n * 4);
  return sw->cigar->n;
}

/* raw DP cell (i = target row, j = query column) after the last refsw_align: 8 x int32 */
void refsw_cell (void * h, int32_t i, int32_t j, int32_t * out8)
{
  sw_t * sw = (sw_t *) h;
  memcpy (out8, sw->sm + (((int64_t) i) << sw->qry_nbits) + j, sizeof (sw_cell_t));
}

/* row i of the score plane, columns 0..n-1 */
void refsw_score_row (void * h, int32_t i, int32_t n, int32_t * out)
{
  sw_t * sw = (sw_t *) h;
  int32_t j;
  for (j = 0; j < n; ++j) out[j] = sw->sm[(((int64_t) i) << sw->qry_nbits) + j].score;
}

void refsw_free (void * h)
{
  sw_t * sw = (sw_t *) h;
  if (sw->type_c > 0) free (sw->mat);
  free (sw->sm);
  cigar_free (sw->cigar);
  free (sw);
}

/* ---- timing helper: n_pairs fixed-size pairs spread over n_thread aligners (one sw_t each).
 * Returns wall seconds; scores[] receives sw->score per pair so the caller can cross-check. */
typedef struct {
  int tid, n_thread;
  int64_t n_pairs;
  int32_t qlen, tlen;
  char * q_all, * t_all;
  int type_c; int32_t * mat; int32_t del_o, del_e, ins_o, ins_e; int strategy;
  int32_t * scores;
} bench_arg_t;

static void * bench_core (void * data)
{
  bench_arg_t * a = (bench_arg_t *) data;
  int64_t p;
  sw_t * sw = sw_init ();
  sw_set_parameter (sw, a->type_c, a->mat, a->del_o, a->del_e, a->ins_o, a->ins_e, a->strategy);
  for (p = a->tid; p < a->n_pairs; p += a->n_thread) {
    sw_align (sw, a->qlen, a->q_all + p * a->qlen, a->tlen, a->t_all + p * a->tlen);
    a->scores[p] = sw->score;
  }
  refsw_free (sw);
  return NULL;
}

double refsw_bench (int n_thread, int64_t n_pairs, int32_t qlen, char * q_all, int32_t tlen, char * t_all,
                    int type_c, int32_t * mat, int32_t del_o, int32_t del_e, int32_t ins_o, int32_t ins_e,
                    int strategy, int32_t * scores)
{
  struct timespec t0, t1;
  pthread_t * pids = (pthread_t *) calloc (n_thread, sizeof (pthread_t));
  bench_arg_t * args = (bench_arg_t *) calloc (n_thread, sizeof (bench_arg_t));
  int i;
  clock_gettime (CLOCK_MONOTONIC, &t0);
  for (i = 0; i < n_thread; ++i) {
    bench_arg_t a = { i, n_thread, n_pairs, qlen, tlen, q_all, t_all, type_c, mat, del_o, del_e, ins_o, ins_e, strategy, scores };
    args[i] = a;
    pthread_create (pids + i, NULL, bench_core, args + i);
  }
  for (i = 0; i < n_thread; ++i) pthread_join (pids[i], NULL);
  clock_gettime (CLOCK_MONOTONIC, &t1);
  free (pids); free (args);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
