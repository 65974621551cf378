/* Lazily created libgcgpu contexts shared by the replacement kmer.c / ont.c / sw.c.
 *
 * Opening a CUDA device costs 0.3 .. 5 s per process (driver start-up without a persistence
 * daemon) and pinning the staging buffers another ~0.2 s.  gcg_bridge_warmup () does both on
 * helper threads (one per device) while the unchanged host code is still reading its inputs
 * (contig_seqs_load, sefq_load: main.c:152-156); gcg_bridge () joins them at the first device
 * call.  Without a warm-up the contexts are created on first use, as before.
 *
 * Devices: $GC_DEVICES = "0,1,2,3" or "all" shards the ONT read batch over several GPUs of the
 * box (table replicated, ont.c of this directory); otherwise the single device $GC_DEVICE (0). */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "utils.h"
#include "gcg_bridge.h"

static gcg_bridge_t g_bridge = { .n_thread = 1 };
static pthread_t g_warm_thread[GCG_BRIDGE_MAX_DEV];
static int g_warm_started = 0, g_warm_rc[GCG_BRIDGE_MAX_DEV];
static gcg_ctx * g_warm_ctx[GCG_BRIDGE_MAX_DEV];
static char g_warm_err[GCG_BRIDGE_MAX_DEV][600];

/* fills g_bridge.devs / n_dev once */
static void
bridge_devices (void)
{
  const char * list = getenv ("GC_DEVICES"), * one = getenv ("GC_DEVICE");
  if (g_bridge.n_dev > 0) return;
  if (list != NULL && strcmp (list, "all") == 0) {
    int i, n = gcg_device_count ();
    if (n < 1) err_mesg ("[gcg_bridge] GC_DEVICES=all but no CUDA device is visible (there is no CPU fallback)");
    if (n > GCG_BRIDGE_MAX_DEV) n = GCG_BRIDGE_MAX_DEV;
    for (i = 0; i < n; ++i) g_bridge.devs[i] = i;
    g_bridge.n_dev = n;
  } else if (list != NULL && *list) {
    const char * p = list;
    while (*p) {
      char * end;
      long d = strtol (p, &end, 10);
      if (end == p || d < 0) err_mesg ("[gcg_bridge] cannot parse GC_DEVICES='%s' (expected e.g. 0,1,2,3 or all)", list);
      if (g_bridge.n_dev == GCG_BRIDGE_MAX_DEV) err_mesg ("[gcg_bridge] GC_DEVICES names more than %d devices", GCG_BRIDGE_MAX_DEV);
      g_bridge.devs[g_bridge.n_dev++] = (int) d;
      p = end;
      if (*p == ',') ++p;
      else if (*p) err_mesg ("[gcg_bridge] cannot parse GC_DEVICES='%s' (expected e.g. 0,1,2,3 or all)", list);
    }
    if (g_bridge.n_dev == 0) err_mesg ("[gcg_bridge] GC_DEVICES is empty");
  } else {
    g_bridge.devs[0] = one ? atoi (one) : 0;
    g_bridge.n_dev = 1;
  }
}

static double
now_ms (void)
{
  struct timespec ts;
  clock_gettime (CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void *
warm_main (void * arg)
{
  const int i = (int) (long) arg;
  double t0 = now_ms (), t1;
  g_warm_rc[i] = gcg_init (g_bridge.devs[i], &g_warm_ctx[i]);
  t1 = now_ms ();
  if (g_warm_rc[i] == 0) g_warm_rc[i] = gcg_warmup (g_warm_ctx[i]);
  if (getenv ("GCG_TRACE")) fprintf (stderr, "[gcg] warm-up thread: device %d open %.0f ms, staging buffers %.0f ms\n", g_bridge.devs[i], t1 - t0, now_ms () - t1);
  if (g_warm_rc[i] != 0) snprintf (g_warm_err[i], sizeof g_warm_err[i], "%s", gcg_last_error ());
  return NULL;
}

void
gcg_bridge_warmup (void)
{
  int i;
  if (g_bridge.ctx != NULL || g_warm_started) return;
  if (getenv ("GC_NO_WARMUP") != NULL) return;
  bridge_devices ();
  for (i = 0; i < g_bridge.n_dev; ++i) {
    if (pthread_create (&g_warm_thread[i], NULL, warm_main, (void *) (long) i) != 0) break;
    ++g_warm_started;
  }
}

int
gcg_bridge_runs_mode (void)
{
  const char * e = getenv ("GC_RUNS");
  return e != NULL && atoi (e) != 0;
}

int
gcg_bridge_sparse_kmers (void)
{
  const char * e = getenv ("GC_SPARSE_KMERS");
  return e != NULL && atoi (e) != 0;
}

gcg_bridge_t *
gcg_bridge_peek (void)
{
  return &g_bridge;
}

gcg_bridge_t *
gcg_bridge (void)
{
  int i;
  if (g_bridge.ctx == NULL) {
    double t0 = now_ms ();
    bridge_devices ();
    for (i = 0; i < g_warm_started; ++i) {
      pthread_join (g_warm_thread[i], NULL);
      if (g_warm_rc[i] != 0)
        err_mesg ("[gcg_bridge] cannot open CUDA device %d: %s (there is no CPU fallback)", g_bridge.devs[i], g_warm_err[i]);
      g_bridge.ctxs[i] = g_warm_ctx[i];
    }
    if (g_warm_started && getenv ("GCG_TRACE")) fprintf (stderr, "[gcg] first device call waited %.0f ms for the warm-up thread\n", now_ms () - t0);
    for (i = g_warm_started; i < g_bridge.n_dev; ++i) {
      int rc = gcg_init (g_bridge.devs[i], &g_bridge.ctxs[i]);
      if (rc != 0)
        err_mesg ("[gcg_bridge] cannot open CUDA device %d: %s (there is no CPU fallback)", g_bridge.devs[i], gcg_last_error ());
    }
    g_warm_started = 0;
    g_bridge.ctx = g_bridge.ctxs[0];
  }
  return &g_bridge;
}

void
gcg_bridge_drop_table (void)
{
  int i;
  for (i = 1; i < g_bridge.n_dev; ++i)
    if (g_bridge.replicas[i]) { gcg_table_free (g_bridge.replicas[i]); g_bridge.replicas[i] = NULL; }
  if (g_bridge.table) { gcg_table_free (g_bridge.table); g_bridge.table = NULL; }
}

void
gcg_bridge_replicate_table (void)
{
  int i;
  if (g_bridge.table == NULL) return;
  for (i = 1; i < g_bridge.n_dev; ++i) {
    if (g_bridge.replicas[i]) { gcg_table_free (g_bridge.replicas[i]); g_bridge.replicas[i] = NULL; }
    GCG_CK (gcg_table_clone (g_bridge.ctxs[i], g_bridge.table, &g_bridge.replicas[i]));
  }
}

void
gcg_bridge_drop_contigs (void)
{
  if (g_bridge.contigs) { gcg_seqs_free (g_bridge.contigs); g_bridge.contigs = NULL; }
}

void
gcg_bridge_shutdown (void)
{
  int i;
  for (i = 0; i < g_warm_started; ++i) {         /* warmed up but never used */
    pthread_join (g_warm_thread[i], NULL);
    if (g_warm_rc[i] == 0) g_bridge.ctxs[i] = g_warm_ctx[i];
  }
  g_warm_started = 0;
  gcg_bridge_drop_table ();
  gcg_bridge_drop_contigs ();
  free (g_bridge.runs); free (g_bridge.run_off); free (g_bridge.run_kmers);
  g_bridge.runs = NULL; g_bridge.run_off = NULL; g_bridge.run_kmers = NULL; g_bridge.n_run = 0;
  for (i = 0; i < GCG_BRIDGE_MAX_DEV; ++i)
    if (g_bridge.ctxs[i]) { gcg_destroy (g_bridge.ctxs[i]); g_bridge.ctxs[i] = NULL; }
  g_bridge.ctx = NULL;
}
