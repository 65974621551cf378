// Host-side helpers of the staging path: a persistent worker pool (std::thread spawn per chunk
// cost more than the copy it parallelised) and a gather copy with non-temporal stores.
//
// The gather moves every base of every read once, from the caller's separately allocated strings
// (rseq_t.b, rseq.c:307-374) into the pinned ring.  A plain memcpy of ~10 kB pieces uses ordinary
// stores: every destination line is first read for ownership, so the copy costs three memory
// transfers per byte.  Streaming stores (MOVNTDQ, baseline x86-64) write whole 64-byte lines
// without the read; destinations are 32-byte aligned by construction (sequences start on packed
// word boundaries).
#include <emmintrin.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "host_par.h"

struct gcg_workers {
  std::vector<std::thread> th;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  const std::function<void (int64_t)> * job = nullptr;
  int64_t n_task = 0, next = 0, running = 0;
  uint64_t gen = 0;
  bool stop = false;
};

static void worker_main (gcg_workers * w)
{
  uint64_t seen = 0;
  std::unique_lock<std::mutex> lk (w->mu);
  for (;;) {
    w->cv_go.wait (lk, [&] () { return w->stop || w->gen != seen; });
    if (w->stop) return;
    seen = w->gen;
    while (w->next < w->n_task) {
      int64_t i = w->next++;
      ++w->running;
      lk.unlock ();
      (*w->job) (i);
      lk.lock ();
      --w->running;
    }
    if (w->running == 0) w->cv_done.notify_all ();
  }
}

gcg_workers * gcg_workers_create (int n_thread)
{
  gcg_workers * w = new gcg_workers ();
  for (int i = 1; i < n_thread; ++i) w->th.emplace_back (worker_main, w);     // the caller is worker 0
  return w;
}

int gcg_workers_count (const gcg_workers * w) { return w ? (int) w->th.size () + 1 : 1; }

void gcg_workers_destroy (gcg_workers * w)
{
  if (!w) return;
  { std::lock_guard<std::mutex> lk (w->mu); w->stop = true; }
  w->cv_go.notify_all ();
  for (auto & t : w->th) t.join ();
  delete w;
}

void gcg_workers_run (gcg_workers * w, int64_t n_task, const std::function<void (int64_t)> & fn)
{
  if (!w || w->th.empty () || n_task <= 1) { for (int64_t i = 0; i < n_task; ++i) fn (i); return; }
  std::unique_lock<std::mutex> lk (w->mu);
  w->job = &fn; w->n_task = n_task; w->next = 0; w->running = 0;
  ++w->gen;
  w->cv_go.notify_all ();
  while (w->next < w->n_task) {            // the caller takes tasks too
    int64_t i = w->next++;
    ++w->running;
    lk.unlock ();
    fn (i);
    lk.lock ();
    --w->running;
  }
  w->cv_done.wait (lk, [&] () { return w->running == 0 && w->next >= w->n_task; });
  w->job = nullptr;
}

// asynchronous form: the pool works on fn while the caller does something else; the caller does not
// take tasks.  gcg_workers_wait() returns when all tasks are done.  fn must stay alive until then.
void gcg_workers_start (gcg_workers * w, int64_t n_task, const std::function<void (int64_t)> & fn)
{
  if (!w || w->th.empty ()) { for (int64_t i = 0; i < n_task; ++i) fn (i); return; }
  std::lock_guard<std::mutex> lk (w->mu);
  w->job = &fn; w->n_task = n_task; w->next = 0; w->running = 0;
  ++w->gen;
  w->cv_go.notify_all ();
}

void gcg_workers_wait (gcg_workers * w)
{
  if (!w || w->th.empty ()) return;
  std::unique_lock<std::mutex> lk (w->mu);
  w->cv_done.wait (lk, [&] () { return w->running == 0 && w->next >= w->n_task; });
  w->job = nullptr;
}

void gcg_copy_stream (void * dst_, const void * src_, size_t n)
{
  char * dst = (char *) dst_;
  const char * src = (const char *) src_;
  if (n < 256 || ((uintptr_t) dst & 15)) { memcpy (dst, src, n); return; }
  size_t i = 0;
  for (; i + 64 <= n; i += 64) {
    __m128i a = _mm_loadu_si128 ((const __m128i *) (src + i));
    __m128i b = _mm_loadu_si128 ((const __m128i *) (src + i + 16));
    __m128i c = _mm_loadu_si128 ((const __m128i *) (src + i + 32));
    __m128i d = _mm_loadu_si128 ((const __m128i *) (src + i + 48));
    _mm_stream_si128 ((__m128i *) (dst + i), a);
    _mm_stream_si128 ((__m128i *) (dst + i + 16), b);
    _mm_stream_si128 ((__m128i *) (dst + i + 32), c);
    _mm_stream_si128 ((__m128i *) (dst + i + 48), d);
  }
  if (i < n) memcpy (dst + i, src + i, n - i);
}

void gcg_copy_fence (void) { _mm_sfence (); }
