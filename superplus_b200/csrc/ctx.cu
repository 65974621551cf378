// Context, error reporting, per-kernel CUDA-event profiling, pinned staging ring.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <mutex>
#include <thread>

#include "gcg_internal.cuh"
#include "host_par.h"

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
  ctx->trace = getenv ("GCG_TRACE") != nullptr;
  if (const char * e = getenv ("GCG_DEVICE_CACHE_MB")) ctx->dparked_limit = (size_t) atoll (e) << 20;
  GCG_CUDA (cudaStreamCreateWithFlags (&ctx->stream, cudaStreamNonBlocking));
  {
    // keep freed blocks in the stream-ordered pool instead of returning them to the driver
    cudaMemPool_t pool;
    GCG_CUDA (cudaDeviceGetDefaultMemPool (&pool, device));
    unsigned long long keep = ~0ULL;
    GCG_CUDA (cudaMemPoolSetAttribute (pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  GCG_CUDA (cudaMalloc (&ctx->d_counters, 64 * sizeof (unsigned long long)));
  GCG_CUDA (cudaHostAlloc (&ctx->h_counters, 64 * sizeof (unsigned long long), cudaHostAllocDefault));
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
  gcg_pipe_free (ctx);
  gcg_workers_destroy (ctx->workers);
  gcg_dcache_release (ctx);
  cudaFree (ctx->d_counters);
  cudaFreeHost (ctx->h_counters);
  cudaStreamDestroy (ctx->stream);
  gcg_pinned_trim ();
  delete ctx;
}

extern "C" int gcg_set_host_threads (gcg_ctx * ctx, int n_thread)
{
  GCG_CHECK (ctx && n_thread >= 1, GCG_EINVAL, "gcg_set_host_threads: bad argument");
  n_thread = n_thread > 64 ? 64 : n_thread;
  if (n_thread != ctx->host_threads && ctx->workers) { gcg_workers_destroy (ctx->workers); ctx->workers = nullptr; }
  ctx->host_threads = n_thread;
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

// Pinned staging ring: 2 x 16 MiB host + 2 x 16 MiB device, allocated on first use (page pinning
// costs roughly a millisecond per MiB and sits on the start-up path of the command line tool;
// 16 MiB chunks already run the PCIe link at full rate).
int gcg_stage_reserve (gcg_ctx * ctx)
{
  if (ctx->stage.cap) return GCG_OK;
  const size_t cap = (size_t) 16 << 20;
  for (int i = 0; i < 2; ++i) {
    GCG_CUDA (cudaHostAlloc (&ctx->stage.h[i], cap, cudaHostAllocDefault));
    GCG_CUDA (cudaMalloc (&ctx->stage.d[i], cap));
    GCG_CUDA (cudaEventCreateWithFlags (&ctx->stage.ev[i], cudaEventDisableTiming));
  }
  ctx->stage.cap = cap;
  return GCG_OK;
}

// ---- device block cache ------------------------------------------------------------------------
static size_t size_class (size_t bytes)
{
  if (bytes < 4096) return 4096;
  int lg = 63 - __builtin_clzll ((unsigned long long) bytes);
  size_t step = (size_t) 1 << (lg - 3);                 // eight classes per octave: at most 12.5 % slack
  return (bytes + step - 1) & ~(step - 1);
}

void gcg_dcache_release (gcg_ctx * ctx)
{
  for (auto & kv : ctx->dparked)
    for (void * p : kv.second) { ctx->dclass.erase (p); cudaFreeAsync (p, ctx->stream); }
  ctx->dparked.clear ();
  ctx->dparked_bytes = 0;
  if (ctx->dbig) { ctx->dclass.erase (ctx->dbig); cudaFreeAsync (ctx->dbig, ctx->stream); ctx->dbig = nullptr; ctx->dbig_cls = 0; }
}

cudaError_t gcg_dmalloc_bytes (gcg_ctx * ctx, void ** p, size_t bytes)
{
  const size_t cls = size_class (bytes);
  if (ctx->dbig && cls > ctx->dparked_limit / 2 && ctx->dbig_cls >= cls && ctx->dbig_cls / 2 <= cls) {
    *p = ctx->dbig;
    ctx->dbig = nullptr; ctx->dbig_cls = 0;
    return cudaSuccess;
  }
  auto it = ctx->dparked.find (cls);
  if (it != ctx->dparked.end () && !it->second.empty ()) {
    *p = it->second.back ();
    it->second.pop_back ();
    ctx->dparked_bytes -= cls;
    return cudaSuccess;
  }
  cudaError_t e = cudaMallocAsync (p, cls, ctx->stream);
  if (e != cudaSuccess) {
    // out of memory with blocks parked: give them back and try once more
    cudaGetLastError ();
    gcg_dcache_release (ctx);
    cudaStreamSynchronize (ctx->stream);
    e = cudaMallocAsync (p, cls, ctx->stream);
  }
  if (e == cudaSuccess) ctx->dclass[*p] = cls;
  return e;
}

void gcg_dfree (gcg_ctx * ctx, void * p)
{
  if (!p) return;
  auto it = ctx->dclass.find (p);
  if (it == ctx->dclass.end ()) { cudaFreeAsync (p, ctx->stream); return; }
  const size_t cls = it->second;
  if (cls > ctx->dparked_limit / 2) {                    // e.g. the 64 GB trace of an SW batch: one spare, not the cache
    void * drop = p;
    if (cls > ctx->dbig_cls) { drop = ctx->dbig; ctx->dbig = p; ctx->dbig_cls = cls; }     // keep the larger of the two
    if (drop) { ctx->dclass.erase (drop); cudaFreeAsync (drop, ctx->stream); }
    return;
  }
  ctx->dparked[cls].push_back (p);
  ctx->dparked_bytes += cls;
  while (ctx->dparked_bytes > ctx->dparked_limit && !ctx->dparked.empty ()) {
    auto big = std::prev (ctx->dparked.end ());             // drop the largest parked blocks first
    if (big->second.empty ()) { ctx->dparked.erase (big); continue; }
    void * q = big->second.back ();
    big->second.pop_back ();
    ctx->dparked_bytes -= big->first;
    ctx->dclass.erase (q);
    cudaFreeAsync (q, ctx->stream);
  }
}

// ---- pinned result buffers ---------------------------------------------------------------------
// Results handed to the caller (anchor lists, CIGAR pools) live in page-locked host memory.
// Pinning costs far more than the copy it serves (tens of ms per 100 MB), so blocks released with
// gcg_free are parked and handed out again by the next call that fits; gcg_destroy returns the
// parked blocks to the driver.
struct pinned_block { void * p; size_t cap; bool in_use; };
static std::mutex g_pin_mu;
static std::vector<pinned_block> g_pin;

void * gcg_pinned_alloc (size_t bytes)
{
  if (bytes == 0) bytes = 16;
  std::lock_guard<std::mutex> lk (g_pin_mu);
  int best = -1;
  for (size_t i = 0; i < g_pin.size (); ++i)
    if (!g_pin[i].in_use && g_pin[i].cap >= bytes && (best < 0 || g_pin[i].cap < g_pin[(size_t) best].cap)) best = (int) i;
  if (best >= 0) { g_pin[(size_t) best].in_use = true; return g_pin[(size_t) best].p; }
  // nothing parked fits: drop the parked blocks (bounds what the cache holds), allocate with headroom
  for (size_t i = 0; i < g_pin.size ();) {
    if (!g_pin[i].in_use) { cudaFreeHost (g_pin[i].p); g_pin.erase (g_pin.begin () + (long) i); } else ++i;
  }
  size_t cap = bytes + bytes / 8;
  void * p = nullptr;
  // mapped + portable: the search kernels of any context store anchors straight into these blocks (zero-copy results)
  const unsigned flags = cudaHostAllocMapped | cudaHostAllocPortable;
  if (cudaHostAlloc (&p, cap, flags) != cudaSuccess) {
    cudaGetLastError ();
    cap = bytes;
    if (cudaHostAlloc (&p, cap, flags) != cudaSuccess) { cudaGetLastError (); return nullptr; }
  }
  g_pin.push_back ({p, cap, true});
  return p;
}

void gcg_pinned_trim (void)
{
  std::lock_guard<std::mutex> lk (g_pin_mu);
  for (size_t i = 0; i < g_pin.size ();) {
    if (!g_pin[i].in_use) { cudaFreeHost (g_pin[i].p); g_pin.erase (g_pin.begin () + (long) i); } else ++i;
  }
}

// pinned (page-locked, mapped, portable) host memory for a caller's own input / output buffers: copies
// from and to it go straight over PCIe, without the staging copy pageable memory needs
extern "C" void * gcg_host_alloc (int64_t bytes)
{
  if (bytes < 0) { gcg_set_error ("gcg_host_alloc: negative size"); return nullptr; }
  void * p = gcg_pinned_alloc ((size_t) bytes);
  if (!p) gcg_set_error ("gcg_host_alloc: pinning %lld bytes failed", (long long) bytes);
  return p;
}

extern "C" void gcg_free (void * p)
{
  if (!p) return;
  std::lock_guard<std::mutex> lk (g_pin_mu);
  for (auto & b : g_pin)
    if (b.p == p) { b.in_use = false; return; }
  cudaFreeHost (p);
}

// ---- host-side phase trace (GCG_TRACE=1): wall-clock between marks, printed to stderr -----------
static std::chrono::steady_clock::time_point g_trace_t0;
void gcg_trace_mark (gcg_ctx * ctx, const char * label)
{
  if (!ctx || !ctx->trace) return;
  auto now = std::chrono::steady_clock::now ();
  if (label) fprintf (stderr, "[gcg] %-28s %9.3f ms\n", label, std::chrono::duration<double, std::milli> (now - g_trace_t0).count ());
  g_trace_t0 = now;
}

gcg_workers * gcg_ctx_workers (gcg_ctx * ctx)
{
  if (!ctx->workers) ctx->workers = gcg_workers_create (ctx->host_threads);
  return ctx->workers;
}

void gcg_par_memcpy (gcg_ctx * ctx, void * dst, const void * src, size_t bytes)
{
  int nt = ctx ? ctx->host_threads : 1;
  if (bytes < ((size_t) 4 << 20) || nt <= 1) { memcpy (dst, src, bytes); return; }
  gcg_workers_run (gcg_ctx_workers (ctx), nt, [=] (int64_t t) {
    size_t a = bytes * (size_t) t / (size_t) nt, b = bytes * (size_t) (t + 1) / (size_t) nt;
    a &= ~(size_t) 63; if (t + 1 < nt) b &= ~(size_t) 63;
    memcpy ((char *) dst + a, (const char *) src + a, b - a);
  });
}
