/* Lazily created libgcgpu context shared by the replacement kmer.c / ont.c / sw.c.
 *
 * Opening a CUDA device costs 0.3 .. 5 s per process (driver start-up without a persistence
 * daemon) and pinning the staging buffers another ~0.2 s.  gcg_bridge_warmup () does both on a
 * helper thread while the unchanged host code is still reading its inputs (contig_seqs_load,
 * sefq_load: main.c:152-156); gcg_bridge () joins it at the first device call.  Without a
 * warm-up the context is created on first use, as before. */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "utils.h"
#include "gcg_bridge.h"

static gcg_bridge_t g_bridge = { NULL, NULL, NULL, 0, 1 };
static pthread_t g_warm_thread;
static int g_warm_started = 0, g_warm_rc = 0;
static gcg_ctx * g_warm_ctx = NULL;
static char g_warm_err[600] = "";

static int
bridge_device (void)
{
  const char * dev = getenv ("GC_DEVICE");
  return dev ? atoi (dev) : 0;
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
  double t0 = now_ms (), t1;
  (void) arg;
  g_warm_rc = gcg_init (bridge_device (), &g_warm_ctx);
  t1 = now_ms ();
  if (g_warm_rc == 0) g_warm_rc = gcg_warmup (g_warm_ctx);
  if (getenv ("GCG_TRACE")) fprintf (stderr, "[gcg] warm-up thread: device open %.0f ms, staging buffers %.0f ms\n", t1 - t0, now_ms () - t1);
  if (g_warm_rc != 0) snprintf (g_warm_err, sizeof g_warm_err, "%s", gcg_last_error ());
  return NULL;
}

void
gcg_bridge_warmup (void)
{
  if (g_bridge.ctx != NULL || g_warm_started) return;
  if (getenv ("GC_NO_WARMUP") != NULL) return;
  if (pthread_create (&g_warm_thread, NULL, warm_main, NULL) == 0) g_warm_started = 1;
}

gcg_bridge_t *
gcg_bridge (void)
{
  if (g_bridge.ctx == NULL) {
    if (g_warm_started) {
      double t0 = now_ms ();
      pthread_join (g_warm_thread, NULL);
      if (getenv ("GCG_TRACE")) fprintf (stderr, "[gcg] first device call waited %.0f ms for the warm-up thread\n", now_ms () - t0);
      g_warm_started = 0;
      if (g_warm_rc != 0)
        err_mesg ("[gcg_bridge] cannot open the CUDA device: %s (there is no CPU fallback)", g_warm_err);
      g_bridge.ctx = g_warm_ctx;
    } else {
      int rc = gcg_init (bridge_device (), &g_bridge.ctx);
      if (rc != 0)
        err_mesg ("[gcg_bridge] cannot open the CUDA device: %s (there is no CPU fallback)", gcg_last_error ());
    }
  }
  return &g_bridge;
}

void
gcg_bridge_drop_table (void)
{
  if (g_bridge.table) { gcg_table_free (g_bridge.table); g_bridge.table = NULL; }
}

void
gcg_bridge_drop_contigs (void)
{
  if (g_bridge.contigs) { gcg_seqs_free (g_bridge.contigs); g_bridge.contigs = NULL; }
}

void
gcg_bridge_shutdown (void)
{
  if (g_warm_started) {              /* warmed up but never used */
    pthread_join (g_warm_thread, NULL);
    g_warm_started = 0;
    if (g_warm_rc == 0) g_bridge.ctx = g_warm_ctx;
  }
  gcg_bridge_drop_table ();
  gcg_bridge_drop_contigs ();
  if (g_bridge.ctx) { gcg_destroy (g_bridge.ctx); g_bridge.ctx = NULL; }
}
