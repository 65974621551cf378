// Internal declarations shared by the translation units of libgcgpu.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/gcgpu.h"

void gcg_set_error (const char * fmt, ...);

#define GCG_CUDA(x)                                                                              \
  do {                                                                                           \
    cudaError_t e_ = (x);                                                                        \
    if (e_ != cudaSuccess) {                                                                     \
      gcg_set_error ("%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString (e_));           \
      return GCG_ECUDA;                                                                          \
    }                                                                                            \
  } while (0)

#define GCG_CHECK(cond, code, ...)                                                               \
  do {                                                                                           \
    if (!(cond)) { gcg_set_error (__VA_ARGS__); return (code); }                                 \
  } while (0)

struct gcg_prof_entry {
  double ms = 0.0;
  int64_t launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

// Staging: pinned host + device ring used to move ASCII sequences (host -> HBM).
struct gcg_stage {
  char * h[2] = {nullptr, nullptr};
  char * d[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool busy[2] = {false, false};      // an async copy out of h[i] has been enqueued and not yet waited for
  int next = 0;                       // slot the next staged chunk uses
  size_t cap = 0;
};

struct gcg_workers;
struct gcg_pipe;

struct gcg_ctx {
  int device = 0;
  int sm_count = 148;
  int host_threads = 4;
  cudaStream_t stream = nullptr;
  bool prof = false;
  bool trace = false;                 // GCG_TRACE: print host-side phase times
  std::map<std::string, gcg_prof_entry> prof_map;
  int64_t launches = 0;
  gcg_stage stage;
  gcg_workers * workers = nullptr;   // persistent host threads (staging gather, back copies)
  gcg_pipe * pipe = nullptr;         // slots and side streams of the streaming host-buffer search
  // small device scratch for reductions / counters
  unsigned long long * d_counters = nullptr;   // 64 x u64 ([16..48) = per-partition totals of gcg_route_plan)
  unsigned long long * h_counters = nullptr;   // pinned mirror
  int launch_async = -1;                       // streaming search: 1 when a kernel launch returns before the kernel has run (-1: not probed yet)
  double last_anchor_frac = 0.0;               // anchors per ONT k-mer of the previous device-resident search (sizes the next result buffer)
  // parked device blocks by size class (see gcg_dmalloc)
  std::map<size_t, std::vector<void *>> dparked;
  std::unordered_map<void *, size_t> dclass;   // every block handed out or parked -> its class size
  size_t dparked_bytes = 0, dparked_limit = (size_t) 24 << 30;
  // one spare for blocks too large to park (the tens of GB of SW trace of a batch): the next batch
  // of the same shape takes it over instead of paying cudaMallocAsync for 60 GB again (20-40 ms)
  void * dbig = nullptr;
  size_t dbig_cls = 0;
};

// RAII scope around one kernel launch: counts it and, when profiling is on, brackets it
// with CUDA events on the ctx stream (resolved lazily in gcg_prof_report).
struct gcg_kscope {
  gcg_ctx * ctx;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const char * name;
  gcg_kscope (gcg_ctx * c, const char * n) : ctx (c), name (n)
  {
    ++ctx->launches;
    if (ctx->prof) {
      cudaEventCreate (&e0);
      cudaEventCreate (&e1);
      cudaEventRecord (e0, ctx->stream);
    }
  }
  ~gcg_kscope ()
  {
    if (ctx->prof) {
      cudaEventRecord (e1, ctx->stream);
      gcg_prof_entry & p = ctx->prof_map[name];
      p.pending.emplace_back (e0, e1);
      ++p.launches;
    }
  }
};

// ---- device-resident sequence sets ---------------------------------------------------------
// Layout shared by ASCII staging and the 2-bit stream: sequence i starts at "word" woff[i];
// a word is 32 bases = 32 ASCII bytes = one uint64 of 2-bit codes, first base in the two most
// significant bits.  Every sequence is padded to a whole number of words; padding content is
// never interpreted (positions with pos + k > len are masked).
struct gcg_ascii {
  gcg_ctx * ctx = nullptr;
  int64_t n = 0, n_words = 0, n_bases = 0;
  char * d_ascii = nullptr;                 // n_words * 32 bytes
  int64_t * d_woff = nullptr;               // n + 1
  int32_t * d_len = nullptr;                // n
  int32_t * d_tseq = nullptr;               // per 32-word tile: index of the sequence holding its first word
  std::vector<int64_t> h_woff;
  std::vector<int32_t> h_len;
};

struct gcg_seqs {
  gcg_ctx * ctx = nullptr;
  int64_t n = 0, n_words = 0, n_bases = 0;
  uint64_t * d_packed = nullptr;            // n_words + 2 (slack, zero)
  int64_t * d_woff = nullptr;               // n + 1
  int32_t * d_len = nullptr;                // n
  int32_t * d_tseq = nullptr;               // per 32-word tile: index of the sequence holding its first word
  std::vector<int64_t> h_woff;
  std::vector<int32_t> h_len;
};

// ---- contig k-mer table ---------------------------------------------------------------------
// Open addressing, 32-byte buckets of four 8-byte key words (one DRAM sector per probe).
//   key word: bits 0..61 = canonical k-mer + 1 (0 = empty), bit 63 = seen more than once,
//   bit 62 of slot f of a bucket = some key with fingerprint f (= hash & 3) did not fit into this
//   bucket and went on to the next one (a 4-bit filter: an absent key continues its probe
//   sequence only when a key of its own fingerprint overflowed here).
//   vals[slot]: bit 0 = KMER_REV, bits 1..31 = contig position, bits 32..62 = contig index.
//   ont[slot/16]: 2 bits per slot, bit0 = anchored >= once, bit1 = anchored >= twice.
struct gcg_table {
  gcg_ctx * ctx = nullptr;
  int k = 0;
  uint32_t n_bucket = 0;
  uint64_t n_slot = 0;
  unsigned long long * d_keys = nullptr;
  unsigned long long * d_vals = nullptr;

  int64_t n_inserted = 0;                  // k-mer occurrences inserted
  // Pre-filter for tables beyond the L2 (see table_filter_ensure in kmer.cu): one 32-bit word per
  // probe, 2 or 3 bits per key, holding only the keys that can anchor (present exactly once).
  uint32_t * d_filter = nullptr;
  uint32_t filter_words = 0;
  int filter_k3 = 0;
  bool filter_valid = false;               // cleared by every insert
  // first base of every contig in the concatenated scaffold coordinate (compact anchors); NULL for an owner-side partition
  int64_t * d_cbase = nullptr;
  int64_t n_contig = 0, n_cbases = 0;
  // an owner-side partition whose arrays live in ONE cudaMalloc block (keys, then values) that peer GPUs map and probe
  // over NVLink (gcg_table_create_shared); NULL when the arrays come from the context's block cache
  void * shared_block = nullptr;
  uint64_t shared_cap_slots = 0;           // slots the block has room for (a rebuild of the same size reuses it: peers keep their mapping)
  // a VIEW of a partitioned table for the remote-probe search: no arrays of its own, n_part descriptors on the device
  const struct gcg_part_desc * d_parts = nullptr;
  int n_part = 0;
};

struct gcg_hits {
  gcg_ctx * ctx = nullptr;
  int64_t n = 0;
  gcg_hit * d_hits = nullptr;               // fmt 0: gcg_hit[n]; fmt 1: uint64_t[n] compact anchors
  int fmt = 0;
  int64_t n_seq = 0;
  long long * d_read_off = nullptr;         // fmt 1: [n_seq + 1], -1 for reads without a word
};

#define GCG_KEY_MASK 0x3FFFFFFFFFFFFFFFULL
#define GCG_KEY_OVF  0x4000000000000000ULL
#define GCG_KEY_MULTI 0x8000000000000000ULL
// value word: contig index << 32 | position << 1 | KMER_REV; contigs hold at most 2^30 bases and there
// are fewer than 2^31 of them, so bits 31 and 63 are free: they carry the ONT-side multiplicity of
// ont.c:245 (anchored at least once / more than once), set by the kernels that emit anchors
#define GCG_VAL_ONT1 0x0000000080000000ULL
#define GCG_VAL_ONT2 0x8000000000000000ULL

int gcg_stage_reserve (gcg_ctx * ctx);
gcg_workers * gcg_ctx_workers (gcg_ctx * ctx);   // created on first use with ctx->host_threads threads
void gcg_pipe_free (gcg_ctx * ctx);
int gcg_table_alloc (gcg_ctx * ctx, int64_t n_kmers, int k, gcg_table ** out);
int gcg_table_filter_ensure (gcg_ctx * ctx, gcg_table * t);     // builds / refreshes the pre-filter when the table wants one
int64_t gcg_mask_scan_blocks (int64_t n_words);
int gcg_mask_scan (gcg_ctx * ctx, const uint32_t * d_mask, int64_t n_words, uint32_t * d_prefix, uint32_t * d_bsum, int64_t * total);
void * gcg_pinned_alloc (size_t bytes);      // parked-block cache, released with gcg_free
void gcg_pinned_trim (void);
void gcg_trace_mark (gcg_ctx * ctx, const char * label);   // label == NULL restarts the clock
void gcg_par_memcpy (gcg_ctx * ctx, void * dst, const void * src, size_t bytes);   // memcpy over the ctx's host threads

// Device allocations of a context.  Blocks come from the stream-ordered pool (cudaMallocAsync on
// the ctx stream) but are PARKED in the context on gcg_dfree and handed out again to the next
// request of the same size class: the driver pool's own reuse is unpredictable for the 10-100 MB
// scratch blocks of a search (measured 0.01 .. 360 ms per cudaMallocAsync).  Reuse is safe without
// events because every kernel of a context runs on its one stream, in order.
cudaError_t gcg_dmalloc_bytes (gcg_ctx * ctx, void ** p, size_t bytes);
void gcg_dfree (gcg_ctx * ctx, void * p);
void gcg_dcache_release (gcg_ctx * ctx);      // return every parked block to the driver
static inline cudaError_t gcg_dmalloc (gcg_ctx * ctx, void ** p, size_t bytes) { return gcg_dmalloc_bytes (ctx, p, bytes); }
template <class T> static inline cudaError_t gcg_dmalloc (gcg_ctx * ctx, T ** p, size_t bytes)
{
  return gcg_dmalloc_bytes (ctx, (void **) p, bytes);
}
