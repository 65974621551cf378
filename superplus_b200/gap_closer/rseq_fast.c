/* rseq_fast.c — sefq_load (rseq.h:106) for the B200 build of gap_closer.
 *
 * SURVEY 8f, row N1.  The reference reads the ONT FASTQ with one fgetc per byte and two
 * str_add per base (rseq.c:307-374): 1.4 s for the 277 MB of BASELINE configs[1], the largest
 * phase left once the k-mer path runs on the GPU.  The rest of rseq.c (rseq_t life cycle, the
 * paired reader) stays the reference's: its object is compiled from the reference tree with
 * -Dsefq_load=sefq_load_reference and this file provides sefq_load.
 *
 * Same contract, restated from rseq.c:307-374:
 *   - the file is a sequence of lines; line 4i+1 holds the bases and line 4i+3 the qualities of
 *     read i; a read exists once its fourth newline has been seen (a trailing record without
 *     the final newline is dropped);
 *   - every byte of those two lines is kept (no trimming of '\r' or blanks);
 *   - l = length, m = ((l >> INIT_SIZE_BT_WIDTH) + 1) << INIT_SIZE_BT_WIDTH, b and q are separate
 *     ckmalloc'ed, NUL terminated buffers of m bytes;
 *   - "l_base != l_qual" aborts through err_mesg;
 *   - a byte 0xFF ends the input (the reference holds fgetc's result in a char and compares it
 *     with EOF, rseq.c:311,330);
 *   - prints "Total loading %ld reads, cost %lds".
 * Here the file is mapped, the newlines are located by all host cores, and the per-read
 * allocations and copies are spread over the cores as well.
 *
 * Pack on ingest (the rest of row N1): while a read's bases are copied into its rseq_t they are also packed to
 * 2 bits per base (gcg_host_pack_2bit: the library's word layout) into one array for the whole read set, which
 * search_kmers_on_ont_reads (ont.c of this directory) hands to gcg_search_*_packed — the search then reads 8 bytes
 * per 32 bases instead of gathering and packing the ASCII again.  rseq_packed_lookup () finds the array by the
 * address of the read set; GC_NO_PACK_ON_INGEST=1 switches it off (the search packs, as before).
 */
#include <fcntl.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "rseq.h"
#include "utils.h"
#include "gcg_bridge.h"

/* the packed words of the last read set loaded (one per process is all gap_closer needs, main.c:156) */
static struct { const void * set; uint64_t * words; int64_t * woff; int64_t n; } g_packed;

int
rseq_packed_lookup (const void * set, int64_t n_reads, const uint64_t ** words, const int64_t ** woff)
{
  if (g_packed.set == NULL || g_packed.set != set || g_packed.n != n_reads) return 0;
  *words = g_packed.words; *woff = g_packed.woff;
  return 1;
}

#define INIT_SIZE_BT_WIDTH 7     /* private to the reference's rseq.c (rseq.c:22): buffer sizes are multiples of 128 */

typedef struct {
  const char * buf;
  int64_t beg, end;            /* byte range scanned by this worker */
  int64_t n_nl;                /* newlines found in it */
  int64_t * nl;                /* their offsets (second pass) */
} scan_arg_t;

static void *
count_core (void * data)
{
  scan_arg_t * a = (scan_arg_t *) data;
  const char * p = a->buf + a->beg, * e = a->buf + a->end;
  int64_t n = 0;
  while (p < e && (p = (const char *) memchr (p, '\n', (size_t) (e - p))) != NULL) { ++n; ++p; }
  a->n_nl = n;
  return NULL;
}

static void *
locate_core (void * data)
{
  scan_arg_t * a = (scan_arg_t *) data;
  const char * p = a->buf + a->beg, * e = a->buf + a->end;
  int64_t * out = a->nl;
  while (p < e && (p = (const char *) memchr (p, '\n', (size_t) (e - p))) != NULL) { *out++ = p - a->buf; ++p; }
  return NULL;
}

typedef struct {
  const char * buf;
  const int64_t * nl;          /* newline offsets */
  rseq_t * reads;
  int64_t n_reads;
  int64_t * cursor;
  int bad;                     /* a record with l_base != l_qual was seen */
  uint64_t * words;            /* pack on ingest: 2-bit words of all reads, read i at woff[i] (NULL = off) */
  const int64_t * woff;
} fill_arg_t;

static void *
fill_core (void * data)
{
  fill_arg_t * a = (fill_arg_t *) data;
  for (;;) {
    int64_t i0 = __sync_fetch_and_add (a->cursor, 64), i;
    if (i0 >= a->n_reads) break;
    for (i = i0; i < i0 + 64 && i < a->n_reads; ++i) {
      /* line 4i starts after newline 4i-1; the bases are line 4i+1, the qualities line 4i+3 */
      int64_t b0 = a->nl[4 * i] + 1, b1 = a->nl[4 * i + 1];
      int64_t q0 = a->nl[4 * i + 2] + 1, q1 = a->nl[4 * i + 3];
      rseq_t * r = a->reads + i;
      if (b1 - b0 != q1 - q0) { a->bad = 1; continue; }
      r->l = (int32_t) (b1 - b0);
      r->m = ((r->l >> INIT_SIZE_BT_WIDTH) + 1) << INIT_SIZE_BT_WIDTH;
      r->b = (char *) ckmalloc (r->m);
      r->q = (char *) ckmalloc (r->m);
      memcpy (r->b, a->buf + b0, (size_t) r->l); r->b[r->l] = '\0';
      memcpy (r->q, a->buf + q0, (size_t) r->l); r->q[r->l] = '\0';
      if (a->words != NULL && r->l > 0) gcg_host_pack_2bit (r->b, r->l, a->words + a->woff[i]);     /* (the bases are in the cache) */
    }
  }
  return NULL;
}

mp_t(rs) *
sefq_load (const char * fq_file)
{
  int fd, t, nt;
  int64_t size = 0, n_nl = 0, n_reads, cursor = 0, i;
  const char * buf = NULL, * ff;
  int64_t * nl = NULL;
  struct stat st;
  time_t time_beg;
  long ncpu = sysconf (_SC_NPROCESSORS_ONLN);
  pthread_t * pids;
  scan_arg_t * sargs;
  fill_arg_t * fargs;
  mp_t(rs) * set;

  time (&time_beg);
  if ((fd = open (fq_file, O_RDONLY)) < 0 || fstat (fd, &st) != 0)
    err_mesg ("fail to open file: '%s'!", fq_file);
  size = (int64_t) st.st_size;
  if (size > 0) {
    buf = (const char *) mmap (NULL, (size_t) size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    if (buf == MAP_FAILED)
      err_mesg ("fail to map file: '%s'!", fq_file);
    /* the reference stops at the first byte that reads as EOF through a char */
    if ((ff = (const char *) memchr (buf, 0xFF, (size_t) size)) != NULL) size = ff - buf;
  }
  nt = ncpu > 16 ? 16 : (ncpu < 1 ? 1 : (int) ncpu);
  if (size < (1 << 20)) nt = 1;
  pids = (pthread_t *) ckalloc (nt, sizeof (pthread_t));
  sargs = (scan_arg_t *) ckalloc (nt, sizeof (scan_arg_t));
  fargs = (fill_arg_t *) ckalloc (nt, sizeof (fill_arg_t));

  /* newline offsets: count per slice, prefix, locate */
  for (t = 0; t < nt; ++t) {
    sargs[t].buf = buf; sargs[t].beg = size * t / nt; sargs[t].end = size * (t + 1) / nt;
    ckpthread_create (pids + t, NULL, count_core, (void *) (sargs + t));
  }
  for (t = 0; t < nt; ++t) { ckpthread_join (pids[t]); n_nl += sargs[t].n_nl; }
  nl = (int64_t *) ckalloc (n_nl + 1, sizeof (int64_t));
  for (t = 0, i = 0; t < nt; ++t) { sargs[t].nl = nl + i; i += sargs[t].n_nl; }
  for (t = 0; t < nt; ++t) ckpthread_create (pids + t, NULL, locate_core, (void *) (sargs + t));
  for (t = 0; t < nt; ++t) ckpthread_join (pids[t]);

  /* newline k ends line k; record i is lines 4i .. 4i+3 and exists once newline 4i+3 does */
  n_reads = n_nl / 4;
  set = mp_init (rs, NULL, NULL);
  mp_resize (rs, set, n_reads);
  set->n = n_reads;
  free (g_packed.words); free (g_packed.woff);
  memset (&g_packed, 0, sizeof g_packed);
  if (n_reads > 0 && getenv ("GC_NO_PACK_ON_INGEST") == NULL) {
    /* word offset of every read: lengths are known from the newline positions alone */
    int64_t w = 0;
    g_packed.woff = (int64_t *) ckalloc (n_reads + 1, sizeof (int64_t));
    for (i = 0; i < n_reads; ++i) {
      int64_t l = nl[4 * i + 1] - nl[4 * i] - 1;
      g_packed.woff[i] = w;
      w += (l + 31) >> 5;
    }
    g_packed.woff[n_reads] = w;
    if (posix_memalign ((void **) &g_packed.words, 64, (size_t) (w + 2) * 8) != 0) err_mesg ("fail to allocate %ld packed words", (long) w);
    g_packed.words[w] = g_packed.words[w + 1] = 0;
    g_packed.set = set; g_packed.n = n_reads;
  }
  if (n_reads > 0) {
    for (t = 0; t < nt; ++t) {
      fargs[t].buf = buf; fargs[t].nl = nl; fargs[t].reads = set->pool; fargs[t].n_reads = n_reads;
      fargs[t].cursor = &cursor; fargs[t].bad = 0;
      fargs[t].words = g_packed.words; fargs[t].woff = g_packed.woff;
      ckpthread_create (pids + t, NULL, fill_core, (void *) (fargs + t));
    }
    for (t = 0; t < nt; ++t) {
      ckpthread_join (pids[t]);
      if (fargs[t].bad) err_mesg ("l_base != l_qual");
    }
  }
  if (buf != NULL && st.st_size > 0) munmap ((void *) buf, (size_t) st.st_size);
  close (fd);
  free (nl); free (pids); free (sargs); free (fargs);

  printf ("Total loading %ld reads, cost %lds\n", mp_cnt (set), time (NULL) - time_beg);
  return set;
}
