/* Lazily created libgcgpu context shared by the replacement kmer.c / ont.c / sw.c. */
#include <stdlib.h>

#include "utils.h"
#include "gcg_bridge.h"

static gcg_bridge_t g_bridge = { NULL, NULL, NULL, 0, 1 };

gcg_bridge_t *
gcg_bridge (void)
{
  if (g_bridge.ctx == NULL) {
    const char * dev = getenv ("GC_DEVICE");
    int rc = gcg_init (dev ? atoi (dev) : 0, &g_bridge.ctx);
    if (rc != 0)
      err_mesg ("[gcg_bridge] cannot open the CUDA device: %s (there is no CPU fallback)", gcg_last_error ());
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
  gcg_bridge_drop_table ();
  gcg_bridge_drop_contigs ();
  if (g_bridge.ctx) { gcg_destroy (g_bridge.ctx); g_bridge.ctx = NULL; }
}
