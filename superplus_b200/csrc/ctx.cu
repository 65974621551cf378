// Context, error reporting, per-kernel CUDA-event profiling, pinned staging ring.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "gcg_internal.cuh"

static thread_local char g_err[1024] = "";

void gcg_set_error (const char * fmt, ...)
{
  va_list ap;
  va_start (ap, fmt);
  vsnprintf (g_err, sizeof g_err, fmt, ap);
  va_end (ap);
}

extern "C" const char * gcg_last_error (void) { return g_err; }

extern "C" int gcg_device_count (void)
{
  int n = 0;
  if (cudaGetDeviceCount (&n) != cudaSuccess) { cudaGetLastError (); return 0; }
  return n;
}

extern "C" int gcg_init (int device, gcg_ctx ** out)
{
  GCG_CHECK (out != nullptr, GCG_EINVAL, "gcg_init: out == NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount (&n);
  if (e != cudaSuccess || n == 0) {
    gcg_set_error ("gcg_init: no CUDA device (%s); this library has no CPU fallback",
                   e != cudaSuccess ? cudaGetErrorString (e) : "device count 0");
    cudaGetLastError ();
    return GCG_ECUDA;
  }
  GCG_CHECK (device >= 0 && device < n, GCG_EINVAL, "gcg_init: device %d out of range [0,%d)", device, n);
  GCG_CUDA (cudaSetDevice (device));
  cudaDeviceProp prop;
  GCG_CUDA (cudaGetDeviceProperties (&prop, device));
  GCG_CHECK (prop.major == 10, GCG_ECUDA,
             "gcg_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
             device, prop.major, prop.minor);
  gcg_ctx * ctx = new gcg_ctx ();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  GCG_CUDA (cudaStreamCreateWithFlags (&ctx->stream, cudaStreamNonBlocking));
  {
    // keep freed blocks in the stream-ordered pool instead of returning them to the driver
    cudaMemPool_t pool;
    GCG_CUDA (cudaDeviceGetDefaultMemPool (&pool, device));
    unsigned long long keep = ~0ULL;
    GCG_CUDA (cudaMemPoolSetAttribute (pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  GCG_CUDA (cudaMalloc (&ctx->d_counters, 16 * sizeof (unsigned long long)));
  GCG_CUDA (cudaHostAlloc (&ctx->h_counters, 16 * sizeof (unsigned long long), cudaHostAllocDefault));
  *out = ctx;
  return GCG_OK;
}

extern "C" void gcg_destroy (gcg_ctx * ctx)
{
  if (!ctx) return;
  cudaSetDevice (ctx->device);
  cudaStreamSynchronize (ctx->stream);
  for (auto & kv : ctx->prof_map)
    for (auto & pr : kv.second.pending) { cudaEventDestroy (pr.first); cudaEventDestroy (pr.second); }
  for (int i = 0; i < 2; ++i) {
    if (ctx->stage.h[i]) cudaFreeHost (ctx->stage.h[i]);
    if (ctx->stage.d[i]) cudaFree (ctx->stage.d[i]);
    if (ctx->stage.ev[i]) cudaEventDestroy (ctx->stage.ev[i]);
  }
  cudaFree (ctx->d_counters);
  cudaFreeHost (ctx->h_counters);
  cudaStreamDestroy (ctx->stream);
  delete ctx;
}

extern "C" int gcg_set_host_threads (gcg_ctx * ctx, int n_thread)
{
  GCG_CHECK (ctx && n_thread >= 1, GCG_EINVAL, "gcg_set_host_threads: bad argument");
  ctx->host_threads = n_thread > 64 ? 64 : n_thread;
  return GCG_OK;
}

extern "C" void * gcg_stream (gcg_ctx * ctx) { return ctx ? (void *) ctx->stream : nullptr; }

extern "C" int gcg_sync (gcg_ctx * ctx)
{
  GCG_CHECK (ctx, GCG_EINVAL, "gcg_sync: ctx == NULL");
  GCG_CUDA (cudaStreamSynchronize (ctx->stream));
  return GCG_OK;
}

extern "C" int64_t gcg_launch_count (gcg_ctx * ctx) { return ctx ? ctx->launches : 0; }

extern "C" int gcg_prof_enable (gcg_ctx * ctx, int on)
{
  GCG_CHECK (ctx, GCG_EINVAL, "gcg_prof_enable: ctx == NULL");
  ctx->prof = on != 0;
  return GCG_OK;
}

static int prof_resolve (gcg_ctx * ctx)
{
  GCG_CUDA (cudaStreamSynchronize (ctx->stream));
  for (auto & kv : ctx->prof_map) {
    for (auto & pr : kv.second.pending) {
      float ms = 0.f;
      cudaEventElapsedTime (&ms, pr.first, pr.second);
      kv.second.ms += ms;
      cudaEventDestroy (pr.first);
      cudaEventDestroy (pr.second);
    }
    kv.second.pending.clear ();
  }
  return GCG_OK;
}

extern "C" int gcg_prof_reset (gcg_ctx * ctx)
{
  GCG_CHECK (ctx, GCG_EINVAL, "gcg_prof_reset: ctx == NULL");
  int rc = prof_resolve (ctx);
  ctx->prof_map.clear ();
  return rc;
}

extern "C" int gcg_prof_report (gcg_ctx * ctx, char * buf, int64_t cap)
{
  GCG_CHECK (ctx && buf && cap > 0, GCG_EINVAL, "gcg_prof_report: bad argument");
  int rc = prof_resolve (ctx);
  if (rc) return rc;
  int64_t off = 0;
  buf[0] = '\0';
  for (auto & kv : ctx->prof_map) {
    int w = snprintf (buf + off, (size_t) (cap - off), "%s %.6f %lld\n", kv.first.c_str (), kv.second.ms,
                      (long long) kv.second.launches);
    if (w < 0 || off + w >= cap) break;
    off += w;
  }
  return GCG_OK;
}

// Pinned staging ring: 2 x 64 MiB host + 2 x 64 MiB device, allocated on first use.
int gcg_stage_reserve (gcg_ctx * ctx)
{
  if (ctx->stage.cap) return GCG_OK;
  const size_t cap = (size_t) 64 << 20;
  for (int i = 0; i < 2; ++i) {
    GCG_CUDA (cudaHostAlloc (&ctx->stage.h[i], cap, cudaHostAllocDefault));
    GCG_CUDA (cudaMalloc (&ctx->stage.d[i], cap));
    GCG_CUDA (cudaEventCreateWithFlags (&ctx->stage.ev[i], cudaEventDisableTiming));
  }
  ctx->stage.cap = cap;
  return GCG_OK;
}

extern "C" void gcg_free (void * p)
{
  if (p) cudaFreeHost (p);
}
