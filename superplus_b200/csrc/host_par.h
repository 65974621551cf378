// persistent host worker pool + streaming gather copy (host_par.cpp)
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <functional>

struct gcg_workers;
gcg_workers * gcg_workers_create (int n_thread);
int gcg_workers_count (const gcg_workers * w);
void gcg_workers_destroy (gcg_workers * w);
// runs fn(0 .. n_task-1) over the pool and the calling thread; returns when all are done
void gcg_workers_run (gcg_workers * w, int64_t n_task, const std::function<void (int64_t)> & fn);
// asynchronous form (the caller takes no tasks); fn must outlive the matching wait
void gcg_workers_start (gcg_workers * w, int64_t n_task, const std::function<void (int64_t)> & fn);
void gcg_workers_wait (gcg_workers * w);
// true when no asynchronous job is out, or all of its tasks have finished (the wait returns at once)
bool gcg_workers_idle (gcg_workers * w);
// the thread that started the asynchronous job runs one of its tasks (false: none left)
bool gcg_workers_help (gcg_workers * w);
// memcpy with non-temporal stores when dst is 16-byte aligned; gcg_copy_fence() before a DMA reads dst
void gcg_copy_stream (void * dst, const void * src, size_t n);
void gcg_copy_fence (void);
// ASCII -> 2-bit packed words (k1_pack_kernel's layout), ceil(n / 32) words with non-temporal stores;
// gcg_copy_fence() before a DMA reads dst
void gcg_pack_stream (uint64_t * dst, const void * src, size_t n);
