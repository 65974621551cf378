// Host-side helpers of the staging path: a persistent worker pool (std::thread spawn per chunk
// cost more than the copy it parallelised) and a gather copy with non-temporal stores.
//
// The gather moves every base of every read once, from the caller's separately allocated strings
// (rseq_t.b, rseq.c:307-374) into the pinned ring.  A plain memcpy of ~10 kB pieces uses ordinary
// stores: every destination line is first read for ownership, so the copy costs three memory
// transfers per byte.  Streaming stores (MOVNTDQ, baseline x86-64) write whole 64-byte lines
// without the read; destinations are 32-byte aligned by construction (sequences start on packed
// word boundaries).
#include <immintrin.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "host_par.h"

// The pool is used once per pipeline chunk — every ~100-400 us during a search — so what a job costs beyond its tasks
// matters: workers parked on a condition variable took 40-60 us to get going (16 threads re-acquiring one mutex), a
// third of an 8 MiB gather.  Workers therefore keep polling the generation counter for GCG_POOL_SPIN_US (default 300 us)
// after their last task before they go to sleep, tasks are handed out by compare-and-swap on one word that carries
// the job's generation (a late worker can never take a ticket of a job it has not read), and the waiter polls too.
struct gcg_workers {
  std::vector<std::thread> th;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  std::atomic<const std::function<void (int64_t)> *> job {nullptr};     // written before gen is advanced, read after it was seen
  std::atomic<int64_t> n_task {0};
  std::atomic<uint64_t> njob {0};                            // (generation << 32) | number of tasks: what a worker trusts
  std::atomic<uint64_t> gen {0};
  std::atomic<uint64_t> ticket {0};                          // (generation << 32) | next task
  std::atomic<int64_t> done {0};                             // tasks of the current job that have finished
  std::atomic<int> sleepers {0};
  std::atomic<bool> stop {false};
  std::atomic<bool> busy {false};                            // a job is out (the one-job rule)
  int64_t spin_us = 300;
};

static inline int64_t now_us ()
{
  return std::chrono::duration_cast<std::chrono::microseconds> (std::chrono::steady_clock::now ().time_since_epoch ()).count ();
}

// take tasks of generation g until none is left (or the pool has moved on)
static void take_tasks (gcg_workers * w, uint64_t g, const std::function<void (int64_t)> * fn, int64_t n)
{
  for (;;) {
    uint64_t cur = w->ticket.load (std::memory_order_acquire);
    if ((cur >> 32) != (g & 0xFFFFFFFFu) || (int64_t) (cur & 0xFFFFFFFFu) >= n) return;
    if (!w->ticket.compare_exchange_weak (cur, cur + 1, std::memory_order_acq_rel)) continue;
    (*fn) ((int64_t) (cur & 0xFFFFFFFFu));
    if (w->done.fetch_add (1, std::memory_order_acq_rel) + 1 == n) {
      std::lock_guard<std::mutex> lk (w->mu);              // (a waiter that went to sleep is woken; one that polls sees `done`)
      w->cv_done.notify_all ();
    }
  }
}

static void worker_main (gcg_workers * w)
{
  uint64_t seen = 0;
  for (;;) {
    uint64_t g = w->gen.load ();
    if (g == seen && !w->stop.load ()) {
      const int64_t t_end = now_us () + w->spin_us;
      for (int it = 0;; ++it) {
        g = w->gen.load ();
        if (g != seen || w->stop.load ()) break;
        _mm_pause ();
        if ((it & 63) == 63 && now_us () >= t_end) break;
      }
      if (g == seen && !w->stop.load ()) {
        std::unique_lock<std::mutex> lk (w->mu);
        w->sleepers.fetch_add (1);
        w->cv_go.wait (lk, [&] () { return w->stop.load () || w->gen.load () != seen; });
        w->sleepers.fetch_sub (1);
        g = w->gen.load ();
      }
    }
    if (w->stop.load ()) return;
    seen = g;
    // The task count is read together with its generation: a worker that slept through job g must not pair the count
    // of job g + 1 (being posted) with the tickets of job g.  The function pointer may belong to a later job only when
    // job g has finished, and then it has no tickets left.
    const uint64_t nj = w->njob.load ();
    if ((nj >> 32) != (g & 0xFFFFFFFFu)) continue;
    const std::function<void (int64_t)> * fn = w->job.load ();
    if (fn != nullptr) take_tasks (w, g, fn, (int64_t) (nj & 0xFFFFFFFFu));
  }
}

gcg_workers * gcg_workers_create (int n_thread)
{
  gcg_workers * w = new gcg_workers ();
  if (const char * e = getenv ("GCG_POOL_SPIN_US")) w->spin_us = std::max (0, atoi (e));
  for (int i = 1; i < n_thread; ++i) w->th.emplace_back (worker_main, w);     // the caller is worker 0
  return w;
}

int gcg_workers_count (const gcg_workers * w) { return w ? (int) w->th.size () + 1 : 1; }

void gcg_workers_destroy (gcg_workers * w)
{
  if (!w) return;
  { std::lock_guard<std::mutex> lk (w->mu); w->stop.store (true); }
  w->cv_go.notify_all ();
  for (auto & t : w->th) t.join ();
  delete w;
}

// publish a job; false when one is already out (the caller then runs its tasks itself)
static bool post_job (gcg_workers * w, int64_t n_task, const std::function<void (int64_t)> & fn)
{
  bool expect = false;
  if (!w->busy.compare_exchange_strong (expect, true)) return false;
  const uint64_t g = w->gen.load () + 1;
  w->job.store (&fn); w->n_task.store (n_task);
  w->njob.store (((g & 0xFFFFFFFFu) << 32) | (uint64_t) n_task);
  w->done.store (0);
  w->ticket.store ((g & 0xFFFFFFFFu) << 32, std::memory_order_release);
  w->gen.store (g);
  if (w->sleepers.load () > 0) { std::lock_guard<std::mutex> lk (w->mu); w->cv_go.notify_all (); }
  return true;
}

static void wait_job (gcg_workers * w)
{
  const int64_t n = w->n_task.load ();
  const int64_t t_end = now_us () + 2000;
  for (int it = 0; w->done.load (std::memory_order_acquire) < n; ++it) {
    _mm_pause ();
    if ((it & 63) == 63 && now_us () >= t_end) {           // a long job: sleep until the last task reports
      std::unique_lock<std::mutex> lk (w->mu);
      w->cv_done.wait (lk, [&] () { return w->done.load () >= n; });
      break;
    }
  }
  w->job.store (nullptr);
  w->busy.store (false);
}

// The pool holds ONE job at a time.  A second job offered while an asynchronous one is still out
// (gcg_workers_start without its gcg_workers_wait) must not touch the first one's state — that
// would drop the tasks nobody has taken yet — so it runs on the calling thread.
void gcg_workers_run (gcg_workers * w, int64_t n_task, const std::function<void (int64_t)> & fn)
{
  if (!w || w->th.empty () || n_task <= 1 || n_task >= 0x7FFFFFFF || !post_job (w, n_task, fn)) { for (int64_t i = 0; i < n_task; ++i) fn (i); return; }
  take_tasks (w, w->gen.load (), &fn, n_task);              // the caller takes tasks too
  wait_job (w);
}

// asynchronous form: the pool works on fn while the caller does something else; the caller does not
// take tasks.  gcg_workers_wait() returns when all tasks are done.  fn must stay alive until then.
void gcg_workers_start (gcg_workers * w, int64_t n_task, const std::function<void (int64_t)> & fn)
{
  if (!w || w->th.empty () || n_task >= 0x7FFFFFFF || !post_job (w, n_task, fn)) { for (int64_t i = 0; i < n_task; ++i) fn (i); return; }
}

// true when nothing is out or the job that is out has finished all its tasks (gcg_workers_wait then returns at once)
bool gcg_workers_idle (gcg_workers * w)
{
  if (!w || w->th.empty () || !w->busy.load ()) return true;
  return w->done.load (std::memory_order_acquire) >= w->n_task.load ();
}

// the caller of an asynchronous job lends a hand: runs ONE task of the job that is out, if one is left
bool gcg_workers_help (gcg_workers * w)
{
  if (!w || w->th.empty () || !w->busy.load ()) return false;
  const uint64_t g = w->gen.load ();
  const std::function<void (int64_t)> * fn = w->job.load ();
  const int64_t n = w->n_task.load ();               // (the caller posted this job itself: nothing can be stale)
  for (;;) {
    uint64_t cur = w->ticket.load (std::memory_order_acquire);
    if ((cur >> 32) != (g & 0xFFFFFFFFu) || (int64_t) (cur & 0xFFFFFFFFu) >= n || fn == nullptr) return false;
    if (!w->ticket.compare_exchange_weak (cur, cur + 1, std::memory_order_acq_rel)) continue;
    (*fn) ((int64_t) (cur & 0xFFFFFFFFu));
    w->done.fetch_add (1, std::memory_order_acq_rel);      // (the caller is the only waiter and it is here: nobody to wake)
    return true;
  }
}

void gcg_workers_wait (gcg_workers * w)
{
  if (!w || w->th.empty ()) return;
  if (!w->busy.load ()) return;                     // nothing out (the job ran inline, or was waited for already)
  wait_job (w);
}

// Self-test of the one-job rule (tests/test_host_logic.py): an asynchronous job of n_async slow tasks is
// started, a synchronous job of n_sync tasks is offered while it is out, then the first is waited for.
// Returns tasks executed: n_async * 1000 + n_sync when nothing was dropped.
extern "C" int64_t gcg_selftest_workers (int n_thread, int n_async, int n_sync)
{
  gcg_workers * w = gcg_workers_create (n_thread);
  std::vector<int> a ((size_t) n_async, 0), b ((size_t) n_sync, 0);
  std::function<void (int64_t)> fa = [&] (int64_t i) { std::this_thread::sleep_for (std::chrono::milliseconds (2)); a[(size_t) i] += 1; };
  std::function<void (int64_t)> fb = [&] (int64_t i) { b[(size_t) i] += 1; };
  gcg_workers_start (w, n_async, fa);
  gcg_workers_run (w, n_sync, fb);
  gcg_workers_wait (w);
  gcg_workers_run (w, n_sync, fb);                  // and the pool still works afterwards
  gcg_workers_destroy (w);
  int64_t na = 0, nb = 0;
  for (int v : a) na += v;
  for (int v : b) nb += v;
  return na * 1000 + nb / 2;
}

// Stress of the pool's hand-out (tests/test_host_logic.py): n_job jobs of 1..97 tasks back to back, alternately
// synchronous and asynchronous (some with a second job offered meanwhile); every task must run exactly once in ITS job —
// a worker that slept through a job must not pair the next job's task count or function with the old job's tickets.
// Returns 0 when all is well, else the 1-based number of the first job that went wrong.
extern "C" int64_t gcg_selftest_workers_stress (int n_thread, int n_job, int spin_us)
{
  gcg_workers * w = gcg_workers_create (n_thread);
  if (spin_us >= 0) w->spin_us = spin_us;
  int64_t bad = 0;
  std::atomic<int64_t> extra {0};
  int64_t want_extra = 0;
  for (int rep = 0; rep < n_job && !bad; ++rep) {
    const int n = 1 + rep % 97;
    std::vector<int> hit ((size_t) n, 0);
    std::function<void (int64_t)> fn = [&] (int64_t i) { hit[(size_t) i] += 1; };
    if (rep & 1) gcg_workers_run (w, n, fn);
    else {
      gcg_workers_start (w, n, fn);
      if (rep % 6 == 0) { std::function<void (int64_t)> f2 = [&] (int64_t) { extra.fetch_add (1); }; gcg_workers_run (w, 3, f2); want_extra += 3; }
      gcg_workers_wait (w);
    }
    for (int i = 0; i < n; ++i) if (hit[(size_t) i] != 1) bad = rep + 1;
  }
  gcg_workers_destroy (w);
  if (!bad && extra.load () != want_extra) bad = n_job + 1;
  return bad;
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

// ---------------------------------------------------------------------------------------------
// ASCII -> 2 bit on the host, for the host-buffer search: the gather has to read every base of
// the caller's strings anyway; writing 2 bits per base instead of 8 cuts the pinned-memory writes
// and the PCIe upload by four (host memory bandwidth is what bounds the end-to-end rate, see
// DESIGN.md 6).  Same packing as k1_pack_kernel / pack4: base code (c >> 1) & 3 (bio.h:23-24), 32
// bases per 64-bit word, first base in the top two bits, tail padded with code 0.
// ---------------------------------------------------------------------------------------------
static inline uint32_t pack4_swar (uint32_t x) { return (((x >> 1) & 0x03030303u) * 0x40100401u) >> 24; }

static inline uint64_t pack32_swar (const char * s)
{
  uint32_t v[8];
  memcpy (v, s, 32);
  uint64_t r = 0;
  for (int i = 0; i < 8; ++i) r = (r << 8) | pack4_swar (v[i]);
  return r;
}

__attribute__ ((target ("bmi2"))) static inline uint64_t pack32_pext (const char * s)
{
  // byte-swap so that the first base sits in the most significant byte, then gather bits 1-2 of
  // every byte: 8 bases -> 16 bits, first base on top
  const uint64_t M = 0x0606060606060606ULL;
  uint64_t a, b, c, d;
  memcpy (&a, s, 8); memcpy (&b, s + 8, 8); memcpy (&c, s + 16, 8); memcpy (&d, s + 24, 8);
  return (__builtin_ia32_pext_di (__builtin_bswap64 (a), M) << 48) | (__builtin_ia32_pext_di (__builtin_bswap64 (b), M) << 32) |
         (__builtin_ia32_pext_di (__builtin_bswap64 (c), M) << 16) | __builtin_ia32_pext_di (__builtin_bswap64 (d), M);
}

static inline void store_word_stream (uint64_t * dst, uint64_t v) { _mm_stream_si64 ((long long *) dst, (long long) v); }

__attribute__ ((target ("bmi2"))) static void pack_stream_pext (uint64_t * dst, const char * src, size_t n)
{
  size_t w = 0;
  for (; (w + 1) * 32 <= n; ++w) store_word_stream (dst + w, pack32_pext (src + w * 32));
  if (w * 32 < n) {
    char tail[32] = {0};
    memcpy (tail, src + w * 32, n - w * 32);
    store_word_stream (dst + w, pack32_pext (tail));
  }
}

// AVX2: 32 bases per iteration.  (x >> 1) & 3 per byte, then two multiply-adds fold four codes into
// one byte per 32-bit lane (b0*64 + b1*16 + b2*4 + b3), a byte shuffle collects the eight bytes in
// big-endian order (first base in the top bits of the word).
__attribute__ ((target ("avx2"))) static inline uint64_t pack32_avx2 (const char * s)
{
  const __m256i x = _mm256_loadu_si256 ((const __m256i *) s);
  const __m256i c = _mm256_and_si256 (_mm256_srli_epi16 (x, 1), _mm256_set1_epi8 (3));
  const __m256i p = _mm256_maddubs_epi16 (c, _mm256_set1_epi32 (0x01041040));          // bytes 64,16,4,1
  const __m256i q = _mm256_madd_epi16 (p, _mm256_set1_epi16 (1));
  const __m256i ctl = _mm256_setr_epi8 (12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                        12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
  const __m256i r = _mm256_shuffle_epi8 (q, ctl);
  const uint32_t hi = (uint32_t) _mm256_extract_epi32 (r, 0), lo = (uint32_t) _mm256_extract_epi32 (r, 4);
  return ((uint64_t) hi << 32) | lo;
}

__attribute__ ((target ("avx2"))) static void pack_stream_avx2 (uint64_t * dst, const char * src, size_t n)
{
  size_t w = 0;
  for (; (w + 1) * 32 <= n; ++w) store_word_stream (dst + w, pack32_avx2 (src + w * 32));
  if (w * 32 < n) {
    char tail[32] = {0};
    memcpy (tail, src + w * 32, n - w * 32);
    store_word_stream (dst + w, pack32_avx2 (tail));
  }
}

// AVX-512 VBMI: 64 bases per fold (same shift / mask / two multiply-adds on 512 bits), one byte
// permute collects the sixteen code bytes as two big-endian words; four folds fill a whole 64-byte
// line that leaves with ONE non-temporal store (a full write-combining buffer per store instead of
// eight partial ones).  bench_tools/host_pack_bench.cpp on the B200 box's host (profiles/
// r01_host_pack_bench.log): 13.0 against 9.9 GB/s of ASCII on one core, 25.6 against 19.5 on two (what
// a rank has when eight ranks share the box's 16 cores), 99 against 87 GB/s on all sixteen.
#define GCG_T512 __attribute__ ((target ("avx512f,avx512bw,avx512vbmi,avx512vl")))
GCG_T512 static inline __m512i fold64_avx512 (const char * s)
{
  const __m512i x = _mm512_loadu_si512 ((const void *) s);
  const __m512i c = _mm512_and_si512 (_mm512_srli_epi16 (x, 1), _mm512_set1_epi8 (3));
  const __m512i q = _mm512_madd_epi16 (_mm512_maddubs_epi16 (c, _mm512_set1_epi32 (0x01041040)), _mm512_set1_epi16 (1));
  // dword i holds bases 4i..4i+3 in its low byte; word 0 wants dword 0 in its top byte, i.e. last in memory
  const __m512i idx = _mm512_castsi128_si512 (_mm_setr_epi8 (28, 24, 20, 16, 12, 8, 4, 0, 60, 56, 52, 48, 44, 40, 36, 32));
  return _mm512_permutexvar_epi8 (idx, q);          // the two words in the low 128 bits
}

GCG_T512 static void pack_stream_avx512 (uint64_t * dst, const char * src, size_t n)
{
  size_t w = 0;
  const size_t full = n / 32;                       // whole words
  // head: single words until the destination sits on a 64-byte line
  for (; w < full && ((uintptr_t) (dst + w) & 63); ++w) store_word_stream (dst + w, pack32_avx2 (src + w * 32));
  for (; w + 8 <= full; w += 8) {
    const char * s = src + w * 32;
    __m512i r = fold64_avx512 (s);
    r = _mm512_inserti32x4 (r, _mm512_castsi512_si128 (fold64_avx512 (s + 64)), 1);
    r = _mm512_inserti32x4 (r, _mm512_castsi512_si128 (fold64_avx512 (s + 128)), 2);
    r = _mm512_inserti32x4 (r, _mm512_castsi512_si128 (fold64_avx512 (s + 192)), 3);
    _mm512_stream_si512 ((__m512i *) (dst + w), r);
  }
  for (; w + 2 <= full; w += 2) _mm_stream_si128 ((__m128i *) (dst + w), _mm512_castsi512_si128 (fold64_avx512 (src + w * 32)));
  for (; w < full; ++w) store_word_stream (dst + w, pack32_avx2 (src + w * 32));
  if (w * 32 < n) {
    char tail[32] = {0};
    memcpy (tail, src + w * 32, n - w * 32);
    store_word_stream (dst + w, pack32_avx2 (tail));
  }
}

static void pack_stream_swar (uint64_t * dst, const char * src, size_t n)
{
  size_t w = 0;
  for (; (w + 1) * 32 <= n; ++w) store_word_stream (dst + w, pack32_swar (src + w * 32));
  if (w * 32 < n) {
    char tail[32] = {0};
    memcpy (tail, src + w * 32, n - w * 32);
    store_word_stream (dst + w, pack32_swar (tail));
  }
}

void gcg_pack_stream (uint64_t * dst, const void * src, size_t n)
{
  // GCG_HOST_PACK = avx512 | avx2 | pext | swar forces a path (tests); default: the best the CPU has
  static const int path = [] () {
    const char * e = getenv ("GCG_HOST_PACK");
    const bool has512 = __builtin_cpu_supports ("avx512vbmi") && __builtin_cpu_supports ("avx512bw") && __builtin_cpu_supports ("avx512vl") && __builtin_cpu_supports ("avx2");
    if (e && !strcmp (e, "swar")) return 0;
    if (e && !strcmp (e, "pext") && __builtin_cpu_supports ("bmi2")) return 1;
    if (e && !strcmp (e, "avx2") && __builtin_cpu_supports ("avx2")) return 2;
    if (e && !strcmp (e, "avx512") && has512) return 3;
    return has512 ? 3 : __builtin_cpu_supports ("avx2") ? 2 : __builtin_cpu_supports ("bmi2") ? 1 : 0;
  } ();
  if (path == 3) pack_stream_avx512 (dst, (const char *) src, n);
  else if (path == 2) pack_stream_avx2 (dst, (const char *) src, n);
  else if (path == 1) pack_stream_pext (dst, (const char *) src, n);
  else pack_stream_swar (dst, (const char *) src, n);
}
