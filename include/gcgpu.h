/* gcgpu.h — C ABI of libgcgpu.so: the B200 (sm_100a) implementation of SuperPlus gap_closer's
 * k-mer count/lookup and Smith-Waterman/CIGAR inner loops.
 *
 * This is the drop-in boundary.  The reference has no FFI of its own (it is one C program), so
 * the entry points below are *batched* forms of the reference functions they replace; the C
 * shims in superplus_b200/gap_closer/ keep every prototype of kmer.h / ont.h / sw.h / cigar.h /
 * hash.h and forward to these.  All paths below are relative to
 * /root/reference/gap_closer.
 *
 * Conventions
 *   - plain C types only; no CUDA or torch types in any signature
 *   - every function returns 0 on success or a negative GCG_E* code; gcg_last_error() returns a
 *     thread-local message.  (Reference behaviour is err_mesg -> abort, utils.c:217-230; the
 *     shims turn a non-zero return into err_mesg.)
 *   - one gcg_ctx per process and device; calls on one ctx are serialised by the caller
 *     (the reference API is not thread safe either, hash.h:13-15).  Contexts of DIFFERENT devices
 *     may be driven from different host threads at the same time (gcg_table_clone /
 *     gcg_sw_batch_multi do so).  Several contexts on ONE device work (the tests use them to
 *     exercise the multi-device paths on a one-GPU box) with one restriction: the SW scoring
 *     parameters live in that device's constant memory, so such contexts must not run SW batches
 *     with different gcg_sw_params concurrently.
 *   - input sequences are borrowed for the duration of the call; outputs returned through
 *     `**` are owned by the library and released with the matching *_free
 *   - there is NO CPU fallback: without a CUDA device gcg_init fails
 */
#ifndef GCGPU_H
#define GCGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCG_OK          0
#define GCG_ECUDA      -1   /* CUDA runtime error */
#define GCG_EINVAL     -2   /* bad argument */
#define GCG_ENOMEM     -3   /* host or device allocation failed */
#define GCG_ERANGE     -4   /* input exceeds a documented limit */

typedef struct gcg_ctx gcg_ctx;
typedef struct gcg_seqs gcg_seqs;     /* device-resident 2-bit packed sequence set */
typedef struct gcg_ascii gcg_ascii;   /* device-resident ASCII sequence set (bench: inputs already in HBM) */
typedef struct gcg_table gcg_table;   /* device-resident contig k-mer table */
typedef struct gcg_hits gcg_hits;     /* device-resident anchor list */
typedef struct gcg_swbatch gcg_swbatch; /* device-resident SW pair batch */

/* ------------------------------------------------------------------ context ---------- */
int  gcg_device_count (void);
int  gcg_init (int device, gcg_ctx ** out);
void gcg_destroy (gcg_ctx * ctx);
const char * gcg_last_error (void);
/* number of host threads used for staging / back-fill (the reference's n_thread, main.c:142) */
int  gcg_set_host_threads (gcg_ctx * ctx, int n_thread);
/* the CUDA stream every kernel of this ctx is launched on, as an opaque cudaStream_t */
void * gcg_stream (gcg_ctx * ctx);
int  gcg_sync (gcg_ctx * ctx);
/* optional: allocate the pinned staging ring and the slots / side streams of the streaming search
 * now (about 0.2 s of page pinning) instead of inside the first gcg_seqs_upload / gcg_search */
int  gcg_warmup (gcg_ctx * ctx);
/* per-kernel CUDA-event timing: enable, run, then read "name ms launches\n" lines */
int  gcg_prof_enable (gcg_ctx * ctx, int on);
int  gcg_prof_reset (gcg_ctx * ctx);
int  gcg_prof_report (gcg_ctx * ctx, char * buf, int64_t cap);
/* number of kernels launched by this ctx since creation */
int64_t gcg_launch_count (gcg_ctx * ctx);
/* measured roofline denominators of this device: packed s16x2 add+max lane-ops per second
 * (one lane-op = one VIADDMNMX.S16x2 on one thread = two 16-bit results), and device copy
 * bandwidth in bytes per second (read + write) */
int  gcg_ubench_int16 (gcg_ctx * ctx, double * lane_ops_per_s);
int  gcg_ubench_hbm (gcg_ctx * ctx, double * bytes_per_s);
/* independent random 32-byte bucket loads per second from a table of `table_bytes` (the probe
 * pattern of gcg_search; small tables measure the L2-resident case, large ones HBM) */
int  gcg_ubench_gather (gcg_ctx * ctx, int64_t table_bytes, double * lookups_per_s);

/* ------------------------------------------------------------------ sequences -------- */
/* K1: ASCII -> 2 bit, base2int(b) = (b>>1)&3 (bio.h:24; kseq1.h:28-35).  Host pointers. */
int  gcg_seqs_upload (gcg_ctx * ctx, const char * const * seq, const int32_t * len, int64_t n, gcg_seqs ** out);
/* same, from one concatenated buffer: sequence i is buf[off[i] .. off[i+1]) */
int  gcg_seqs_upload_concat (gcg_ctx * ctx, const char * buf, const int64_t * off, int64_t n, gcg_seqs ** out);
/* stage ASCII on the device without packing, then pack on the device only (K1 alone) */
int  gcg_ascii_upload_concat (gcg_ctx * ctx, const char * buf, const int64_t * off, int64_t n, gcg_ascii ** out);
int  gcg_seqs_pack (gcg_ctx * ctx, const gcg_ascii * a, gcg_seqs ** out);
void gcg_ascii_free (gcg_ascii * a);
void gcg_seqs_free (gcg_seqs * s);
int64_t gcg_seqs_count (const gcg_seqs * s);
int64_t gcg_seqs_bases (const gcg_seqs * s);
int64_t gcg_seqs_kmers (const gcg_seqs * s, int k);   /* sum over sequences of max(0, len-k+1) */
/* the same packing on the host (no device needed): what the gather of gcg_search does to every
 * read on its way into the pinned upload buffer.  words_out receives (len + 31) / 32 words, 32
 * bases per word, first base in the top two bits, tail padded with code 0. */
int  gcg_host_pack_2bit (const char * seq, int64_t len, uint64_t * words_out);
/* self-test of the host worker pool's one-job rule (no device needed; tests/test_host_logic.py): starts an
 * asynchronous job of n_async tasks, offers a synchronous one of n_sync tasks while it is out, waits.
 * Returns n_async * 1000 + n_sync when every task ran exactly once. */
int64_t gcg_selftest_workers (int n_thread, int n_async, int n_sync);
/* stress of the pool's task hand-out (spinning / sleeping workers, tickets by compare-and-swap): n_job jobs of 1..97
 * tasks back to back; spin_us < 0 keeps the default spin time.  Returns 0 when every task ran exactly once in its job. */
int64_t gcg_selftest_workers_stress (int n_thread, int n_job, int spin_us);

/* ------------------------------------------------------------------ contig k-mers ---- */
/* replaces chop_contig_seqs2kmers (kmer.c:155-184 / 37-121): one 24-byte record per contig
 * position p in [0, len-k], laid out exactly like kmer_t (def.h:58-66):
 *   u64 kseq @0, i32 hs_id @8 (= crc32(kseq bytes) % n_thread, kmer.c:88, crc32.h:70-81),
 *   i32 tid @12, i32 pos @16, u16 flag @20 (KMER_REV=1), i16 kmer_len @22.
 * kmers_out[i] must have room for max(0,len[i]-k+1) records; n_kmer_out[i] receives that count. */
int  gcg_chop_contigs (gcg_ctx * ctx, const gcg_seqs * contigs, int k, int n_thread,
                       void * const * kmers_out, int32_t * n_kmer_out);

/* replaces kmer_hash_init/clear + put_contig_kmers2hashs (kmer.c:187-213 / 124-152 ->
 * hash.c:113-152): distinct canonical k-mers of all contigs with multiplicity {1, >=2} and the
 * (tid,pos,flag) of the single occurrence when multiplicity is 1.  1 <= k <= 31. */
int  gcg_table_build_seqs (gcg_ctx * ctx, const gcg_seqs * contigs, int k, gcg_table ** out);
int  gcg_table_build (gcg_ctx * ctx, const char * const * contig_seq, const int32_t * contig_len,
                      int32_t n_contig, int k, gcg_table ** out);
void gcg_table_free (gcg_table * t);
/* replaces kmer_stat + kmer_stat2 (kmer.c:265-312): out[0] scaffold total, out[1] scaffold
 * unique, out[2] ONT total, out[3] ONT unique (the latter two reflect all searches since the
 * table was built, as the reference's anchored sets do, ont.c:230-254). */
int  gcg_table_stats (gcg_ctx * ctx, gcg_table * t, int64_t out[4]);
/* debugging / parity: dump every distinct key (unordered).  multi_out is 1 or 2 (2 = ">= 2").
 * For multi==1 entries tid/pos/rev describe the occurrence; for others they are unspecified
 * (the reference keeps the first occurrence, which no caller can observe). */
int64_t gcg_table_size (gcg_ctx * ctx, gcg_table * t);
int  gcg_table_dump (gcg_ctx * ctx, gcg_table * t, int64_t cap, uint64_t * key, int32_t * multi_out,
                     int32_t * tid, int32_t * pos, uint8_t * rev);
/* Replicas for read batches sharded over the GPUs of a box inside ONE process (SURVEY 8e,
 * "replicated"; what the reference does with its n_thread work-stealing loop over reads,
 * ont.c:316-400).  gcg_table_clone copies a built table slot for slot into the device of
 * `dst_ctx` (device-to-device, NVLink when the GPUs are peers); every replica is then searched
 * with its own share of the reads from its own host thread.  gcg_table_merge_ont folds the
 * ONT-side multiplicity a replica has collected (ont.c:245: multi = number of ONT hits of the
 * k-mer) into `dst`, saturating at "more than once", so that gcg_table_stats on `dst` reports
 * out[2] / out[3] over ALL reads, as the reference's anchored sets do.  The two tables must be
 * clones of each other (same slot layout).  The counts MOVE: `src`'s ONT-side state is cleared
 * once folded, so further searches on the replicas followed by further merges keep adding up. */
int  gcg_table_clone (gcg_ctx * dst_ctx, gcg_table * src, gcg_table ** out);
int  gcg_table_merge_ont (gcg_ctx * ctx, gcg_table * dst, gcg_table * src);

/* ------------------------------------------------------------------ ONT search ------- */
/* one anchored ONT position (ont.c:171-178,195-202): the canonical k-mer at `pos` of read
 * `read` occurs exactly once (both strands) over all contigs, at contig `tid` position `cpos`.
 * flags bit0 = KMER_REV of the contig occurrence (def.h:52), bit1 = ONT_KMER_REV (def.h:108). */
typedef struct {
  int32_t read;
  int32_t pos;
  int32_t tid;
  uint32_t cpos_flags;      /* (cpos << 2) | flags ; contigs are limited to 2^30 bases */
} gcg_hit;

/* replaces search_kmers_on_ont_reads' SEARCH + REHASH phases (ont.c:407-480 / 141-254).
 * Result is sorted by (read,pos). */
int  gcg_search_seqs (gcg_ctx * ctx, gcg_table * t, const gcg_seqs * reads, int k, gcg_hits ** out);
int64_t gcg_hits_count (const gcg_hits * h);
int  gcg_hits_download (gcg_ctx * ctx, const gcg_hits * h, gcg_hit * dst, int64_t cap);
void gcg_hits_free (gcg_hits * h);
/* host-buffer form: hits_out receives a library-owned pinned array, release with gcg_free */
int  gcg_search (gcg_ctx * ctx, gcg_table * t, const char * const * read_seq, const int32_t * read_len,
                 int64_t n_read, int k, gcg_hit ** hits_out, int64_t * n_hit);
void gcg_free (void * p);
/* page-locked host memory for a caller's own buffers (inputs of gcg_sw_batch / gcg_swbatch_upload are
 * copied straight from it, without the staging copy pageable memory needs); release with gcg_free */
void * gcg_host_alloc (int64_t bytes);

/* Compact anchors — what the shim asks for (superplus_b200/gap_closer/ont.c): ONE 64-bit word per
 * anchor instead of the 16-byte gcg_hit, grouped by read through an offset array, so that the
 * device -> host stream that bounds the end-to-end rate is halved (SURVEY 8d counts an 8-byte hit
 * record).  Anchor word:
 *     bits 63..36  pos    position of the k-mer in its read            (reads shorter than 2^28 bases)
 *     bits 35.. 2  gpos   contig_base[tid] + cpos, contig_base[i] = sum of the lengths of contigs 0..i-1
 *                         in the order they were given to gcg_table_build* (fewer than 2^34 contig bases)
 *     bit 1 ONT_KMER_REV (def.h:108), bit 0 KMER_REV of the contig occurrence (def.h:52)
 * read_off[r] .. read_off[r+1] are the anchors of read r (sorted by pos); read_off has n_read + 1
 * entries.  Same anchors, same order, same ONT-side multiplicity as gcg_search.  Inputs beyond the bit
 * budget return GCG_ERANGE (nothing has been searched): fall back to gcg_search.
 * Both arrays are library-owned pinned memory, released with gcg_free. */
int  gcg_search_compact (gcg_ctx * ctx, gcg_table * t, const char * const * read_seq, const int32_t * read_len,
                         int64_t n_read, int k, uint64_t ** anchors_out, int64_t ** read_off_out, int64_t * n_anchor);
/* The same searches on reads the caller has ALREADY 2-bit packed (the FASTQ loader of
 * superplus_b200/gap_closer packs every read while it copies its bases, so the search never re-reads the ASCII):
 * `packed` holds the words gcg_host_pack_2bit writes, read i at words [woff[i], woff[i] + (read_len[i] + 31) / 32).
 * Words in page-locked memory (gcg_host_alloc) whose reads follow each other without gaps go over PCIe from where
 * they lie — no host pass; anything else is copied into the pinned ring first (a quarter of the ASCII bytes). */
int  gcg_search_compact_packed (gcg_ctx * ctx, gcg_table * t, const uint64_t * packed, const int64_t * woff, const int32_t * read_len,
                                int64_t n_read, int k, uint64_t ** anchors_out, int64_t ** read_off_out, int64_t * n_anchor);
/* device-resident form (bench `value`); download with gcg_hits_download_compact (read_off: n + 1 entries) */
int  gcg_search_seqs_compact (gcg_ctx * ctx, gcg_table * t, const gcg_seqs * reads, int k, gcg_hits ** out);
int  gcg_hits_download_compact (gcg_ctx * ctx, const gcg_hits * h, uint64_t * anchors, int64_t cap, int64_t * read_off);
/* N3, opt-in: search + the anchor grouping of map_ont2contigs (ctg_graph.c:600-656) on the device.
 * The reference cuts every read's anchors, in position order, into runs of consecutive anchors on one contig and
 * keeps per run the contig, the number of anchors whose strand agrees with the scaffold (ONT_KMER_REV == KMER_REV,
 * ctg_graph.c:617-623: "FORW") and disagrees ("BACK"), and the first / last anchor of the majority direction
 * (ont_node_init, ctg_graph.c:93-181).  run_off[r] .. run_off[r+1] are the runs of read r in position order;
 * first_* / last_* are compact anchor words (0 when the count is 0).  Neither the anchors nor a per-base array reach
 * the host.  Same limits as gcg_search_compact (GCG_ERANGE beyond them).  Arrays are pinned, release with gcg_free. */
typedef struct {
  int32_t tid, n_fwd, n_bwd, pad;
  uint64_t first_fwd, last_fwd, first_bwd, last_bwd;
} gcg_run;
int  gcg_search_runs (gcg_ctx * ctx, gcg_table * t, const char * const * read_seq, const int32_t * read_len,
                      int64_t n_read, int k, gcg_run ** runs_out, int64_t ** run_off_out, int64_t * n_run, int64_t * n_anchor);
int  gcg_search_runs_packed (gcg_ctx * ctx, gcg_table * t, const uint64_t * packed, const int64_t * woff, const int32_t * read_len,
                             int64_t n_read, int k, gcg_run ** runs_out, int64_t ** run_off_out, int64_t * n_run, int64_t * n_anchor);
#define GCG_ANCHOR_POS(a)   ((int32_t) ((a) >> 36))
#define GCG_ANCHOR_GPOS(a)  ((int64_t) (((a) >> 2) & 0x3FFFFFFFFULL))
#define GCG_ANCHOR_FLAGS(a) ((uint32_t) ((a) & 3u))

/* ------------------------------------------------------------------ partitioned table  */
/* Multi-GPU form of the same tables (BASELINE configs[3]): the reference splits its
 * k-mer tables into n_thread partitions by crc32(kseq) % n_thread (kmer.c:88,124-152;
 * ont.c:169,193) and lets every thread scan all k-mers for its own share.  Here a partition is
 * one GPU: k-mers are routed to their owner, exchanged (NCCL all-to-all, done by the caller on
 * the device buffers below), inserted / looked up there, and the answers travel back.  Every
 * d_* argument is a DEVICE pointer; all work is enqueued on the ctx stream (gcg_stream), which
 * the caller must order its exchange after.  Which hash picks the owner is unobservable (the
 * partition id never leaves kmer.c/ont.c), so it is a 64-bit mix independent of the bucket hash. */
#define GCG_MAX_PART 16
typedef struct gcg_route gcg_route;   /* stable partition plan of the k-mers of one tile range */

/* owner partition of a canonical k-mer value (host side, for tests and host bookkeeping) */
int  gcg_kmer_owner (uint64_t canonical_kmer, int n_part);
/* sequences are addressed in tiles of 32 packed words (1024 bases): number of tiles of a set */
int64_t gcg_seqs_tiles (const gcg_seqs * s);
/* count the k-mers of tiles [tile_begin, tile_end) by owner: counts[n_part].  `s` is borrowed
 * and must outlive the plan.  One range holds fewer than 2^32 k-mer positions. */
int  gcg_route_plan (gcg_ctx * ctx, const gcg_seqs * s, int k, int n_part, int64_t tile_begin, int64_t tile_end,
                     gcg_route ** out, int64_t * counts);
int64_t gcg_route_kmers (const gcg_route * r);       /* k-mers the plan routes = elements of the send buffer */
int64_t gcg_route_positions (const gcg_route * r);   /* k-mer start positions of the range (= routed, without a pre-filter) */
/* Pre-filter of a routed search.  The filter is a caller-owned device array of 32-bit words (shape
 * from gcg_filter_shape for the TOTAL number of inserted k-mers, zero-initialised): every owner
 * adds the keys of its partition that can anchor (present exactly once), the partial filters are
 * OR-ed together (gcg_filter_or after an all-gather), and gcg_route_plan_filtered routes only the
 * positions whose k-mer passes — true anchors plus a few percent of false positives instead of
 * every ONT k-mer.  Positions that fail are misses by construction (no false negatives). */
int  gcg_filter_shape (int64_t n_keys, int64_t * n_words, int * k3);
int  gcg_filter_add_table (gcg_ctx * ctx, gcg_table * t, void * d_words, int64_t n_words, int k3);
int  gcg_filter_or (gcg_ctx * ctx, void * d_words, const void * d_other, int64_t n_words);
int  gcg_route_plan_filtered (gcg_ctx * ctx, const gcg_seqs * s, int k, int n_part, int64_t tile_begin, int64_t tile_end,
                              const void * d_filter, int64_t filter_words, int filter_k3, gcg_route ** out, int64_t * counts);
/* write the range's k-mers grouped by owner (segment d starts at counts[0]+..+counts[d-1]);
 * inside a segment the order is (sequence, position).
 *   keys:    8 bytes  = canonical k-mer + 1                      (search side, ont.c:161-170)
 *   records: 16 bytes = {canonical k-mer + 1, tid<<32 | pos<<1 | KMER_REV}  (build side, kmer.c:86-94) */
int  gcg_route_keys (gcg_ctx * ctx, gcg_route * r, void * d_send);
int  gcg_route_records (gcg_ctx * ctx, gcg_route * r, void * d_send);
/* owner side: an empty table sized for n_records occurrences; insert received records; answer
 * received keys.  An answer is the 8-byte value word (tid<<32 | pos<<1 | KMER_REV) of a key present
 * exactly once (ont.c:171,195), all ones otherwise.  Lookups also count the ONT-side
 * multiplicity of ont.c:245 in the owner's table, so gcg_table_stats of the partitions add up
 * to the reference's four numbers. */
int  gcg_table_create (gcg_ctx * ctx, int64_t n_records, int k, gcg_table ** out);
int  gcg_table_insert_records (gcg_ctx * ctx, gcg_table * t, const void * d_records, int64_t n);
int  gcg_table_lookup_keys (gcg_ctx * ctx, gcg_table * t, const void * d_keys, int64_t n, void * d_answers);
/* requester side: answers in the order gcg_route_keys wrote the keys -> anchors of the range in
 * (read,pos) order (read = index in `s`). */
int  gcg_route_collect (gcg_ctx * ctx, gcg_route * r, const void * d_answers, gcg_hits ** out);
void gcg_route_free (gcg_route * r);

/* Remote probes (round 2): the partitions stay where their owners built them and the ONE search kernel of the
 * single-GPU path reads them where they lie — a k-mer that passes the union filter loads its bucket from the owner's
 * key array over NVLink / NVSwitch peer memory, an anchor's atomicOr travels to the owner's value array (which thereby
 * counts the hits of every rank, ont.c:245) and returns (tid, pos, flag).  No routing, no exchange buffers, no
 * collective on the data path.
 *   gcg_table_create_shared : an owner-side partition whose arrays live in one cudaMalloc block peers can map
 *   gcg_table_reset_shared  : empty it for a rebuild of about the same size (peers keep their mapping)
 *   gcg_table_shared_info   : block pointer (same process) / 64-byte IPC handle (another process: gcg_window_open),
 *                             offset of the value array inside the block, bucket count
 *   gcg_search_seqs_remote  : d_keys[p] / d_vals[p] = partition p's arrays as this context addresses them; the filter
 *                             is the union filter of gcg_filter_add_table / gcg_filter_or.  Anchors as gcg_hit records. */
int  gcg_table_create_shared (gcg_ctx * ctx, int64_t n_records, int k, gcg_table ** out);
int  gcg_table_reset_shared (gcg_ctx * ctx, gcg_table * t, int64_t n_records);
int  gcg_table_shared_info (gcg_table * t, void ** d_block, int64_t * vals_offset_bytes, uint32_t * n_bucket, void * ipc_handle64);
int  gcg_search_seqs_remote (gcg_ctx * ctx, const gcg_seqs * reads, int k, int n_part, const void * const * d_keys, void * const * d_vals,
                             const uint32_t * n_bucket, const void * d_filter, int64_t filter_words, int filter_k3, gcg_hits ** out);

/* Direct exchange over NVLink / NVSwitch peer memory (no staging buffer, no collective on the
 * data path): every rank owns two WINDOWS — plain device allocations that the other ranks map
 * (CUDA IPC between processes, peer access inside one process).  The routing kernel stores each
 * owner's run of keys straight into that owner's key window, and the owner's lookup kernel
 * stores each answer straight into the requester's answer window, at the position the requester's
 * collect pass expects.  The caller synchronises the ranks between the phases (all stores of a
 * phase are complete when every rank has drained its stream).
 *   gcg_route_keys_direct : owner d's run goes to d_owner_base[d] + owner_off[d] (8-byte elements)
 *   gcg_table_lookup_keys_direct : the key window holds n_src runs, src_count[r] keys from
 *       requester r in rank order; the answers of run r go to d_answer_base[r] + answer_off[r] */
typedef struct gcg_window gcg_window;
int  gcg_window_create (gcg_ctx * ctx, int64_t bytes, gcg_window ** out);
void * gcg_window_ptr (gcg_window * w);
int64_t gcg_window_bytes (gcg_window * w);
int  gcg_window_export (gcg_window * w, void * handle64);                       /* 64-byte IPC handle */
int  gcg_window_open (gcg_ctx * ctx, const void * handle64, void ** d_peer);    /* in ANOTHER process */
int  gcg_window_close (gcg_ctx * ctx, void * d_peer);
void gcg_window_free (gcg_window * w);
int  gcg_peer_enable (gcg_ctx * ctx, int peer_device);                          /* same-process peers */
int  gcg_route_keys_direct (gcg_ctx * ctx, gcg_route * r, void * const * d_owner_base, const int64_t * owner_off);
int  gcg_table_lookup_keys_direct (gcg_ctx * ctx, gcg_table * t, const void * d_keys, int n_src, const int64_t * src_count,
                                   void * const * d_answer_base, const int64_t * answer_off);

/* ------------------------------------------------------------------ Smith-Waterman --- */
#define GCG_SWOS_SOFTCLIP      0   /* sw.h:20-23 */
#define GCG_SWOS_LEADING_INDEL 1
#define GCG_SWOS_INDEL         2
#define GCG_SWOS_IGNORE        3

#define GCG_SW_ASIS  0   /* traceback exactly as shipped (sw.c:289-319 never re-fetches the cell) */
#define GCG_SW_FIXED 1   /* cell re-fetched every step (the evidently intended behaviour)       */

typedef struct {
  int32_t type_c;                       /* alphabet size, symbols are 0..type_c-1 (sw.c:216) */
  int32_t del_o, del_e, ins_o, ins_e;   /* sw.h:52-55 */
  int32_t strategy;                     /* overhang strategy */
  /* border scores the aligner holds (state of init_matrix_values, sw.c:61-110):
   * kind 0 = zeros; kind 1 = row0[j] = -b_ins_o-(j-1)*b_ins_e, col0[i] = -b_del_o-(i-1)*b_del_e */
  int32_t border_kind;
  int32_t b_del_o, b_del_e, b_ins_o, b_ins_e;
  int32_t mat[64];                      /* type_c x type_c, row = query symbol (sw.c:216); type_c <= 8 */
} gcg_sw_params;

typedef struct {
  int32_t score;             /* sw_t.score            (sw.h:44) */
  int32_t alignment_offset;  /* sw_t.alignment_offset (sw.h:45) */
  int32_t has_softclip;      /* sw_t.has_softclip     (sw.h:51) */
  int32_t bt_tidx, bt_qidx;  /* end cell chosen by sw.c:259-280 */
  int32_t n_cigar;           /* number of BAM-encoded ops */
  int64_t cigar_off;         /* offset of the first op in the shared uint32 pool */
} gcg_sw_result;

/* host-buffer form of a batch of sw_align calls (sw.c:400-414).  qry/tgt are integer-coded
 * symbols, pair p is qry[qoff[p]..qoff[p+1]) vs tgt[toff[p]..toff[p+1]).  results[n];
 * *cigar_pool is a library-owned pinned uint32 array (release with gcg_free). */
int  gcg_sw_batch (gcg_ctx * ctx, const gcg_sw_params * P, int mode,
                   const char * qry, const int64_t * qoff, const char * tgt, const int64_t * toff, int64_t n,
                   gcg_sw_result * results, uint32_t ** cigar_pool, int64_t * n_cigar_pool);

/* the same batch sharded over several contexts, one per GPU of the box (pairs are
 * independent): contiguous pair ranges of about equal qry_len x tgt_len, one host thread per
 * context, no collective; results in input order, one pool (release with gcg_free).  What n
 * aligners of the reference, one per core, would do with a list of pairs (sw_t is not shared
 * between threads, sw.h:31-60). */
int  gcg_sw_batch_multi (gcg_ctx * const * ctxs, int n_ctx, const gcg_sw_params * P, int mode,
                         const char * qry, const int64_t * qoff, const char * tgt, const int64_t * toff, int64_t n,
                         gcg_sw_result * results, uint32_t ** cigar_pool, int64_t * n_cigar_pool);

/* device-resident form: upload once, align many times (bench `value`), download results */
int  gcg_swbatch_upload (gcg_ctx * ctx, const char * qry, const int64_t * qoff, const char * tgt,
                         const int64_t * toff, int64_t n, gcg_swbatch ** out);
int  gcg_swbatch_align (gcg_ctx * ctx, gcg_swbatch * b, const gcg_sw_params * P, int mode);
int  gcg_swbatch_download (gcg_ctx * ctx, gcg_swbatch * b, gcg_sw_result * results,
                           uint32_t ** cigar_pool, int64_t * n_cigar_pool);
int64_t gcg_swbatch_cells (const gcg_swbatch * b);
/* which kernel family the last gcg_swbatch_align used per pair: counts[0] packed-s16 pairs,
 * counts[1] generic-s32 pairs */
int  gcg_swbatch_path_counts (const gcg_swbatch * b, int64_t counts[2]);
void gcg_swbatch_free (gcg_swbatch * b);

#ifdef __cplusplus
}
#endif
#endif
