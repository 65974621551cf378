/* cigar.c — replacement for gap_closer/cigar.c (cigar.h:26-55): growable array of BAM-encoded
 * CIGAR operations, (length << 4) | op with op indexing "MIDNSHP=XB".  Host-side container; the
 * operations themselves are produced on the GPU (libgcgpu sw_cigar_kernel) and appended here by
 * sw.c.  Written from cigar.h and the BAM specification.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>

#include "str.h"
#include "utils.h"
#include "cigar.h"

#define OP_OF(e)   ((e) & 0xfu)
#define LEN_OF(e)  ((e) >> 4)
#define OP_CHARS   "MIDNSHP=XB"
/* two bits per op: bit0 consumes query, bit1 consumes reference (BAM spec table) */
#define OP_CONSUMES(op) ((0x3C1A7 >> ((op) << 1)) & 3)

enum { C_MATCH = 0, C_INS = 1, C_DEL = 2, C_SOFT = 4, C_HARD = 5, C_PAD = 6 };

#define CIGAR_FIRST_CAP 4

void
cigar_init2 (cigar_t * c)
{
  c->n = 0;
  c->m = CIGAR_FIRST_CAP;
  c->c = (uint32_t *) ckalloc (CIGAR_FIRST_CAP, sizeof (uint32_t));
}

cigar_t *
cigar_init (void)
{
  cigar_t * c = (cigar_t *) ckalloc (1, sizeof (cigar_t));
  cigar_init2 (c);
  return c;
}

void cigar_clear (cigar_t * c) { c->n = 0; }

void cigar_free2 (cigar_t * c) { free (c->c); }

void
cigar_free (cigar_t * c)
{
  cigar_free2 (c);
  free (c);
}

void
cigar_resize (cigar_t * c, int32_t cnt)
{
  int32_t cap = c->m;

  if (cnt <= cap)
    return;
  if (cap <= 0)
    cap = CIGAR_FIRST_CAP;
  while (cap < cnt)
    cap *= 2;
  c->m = cap;
  c->c = (uint32_t *) ckrealloc (c->c, (size_t) cap * sizeof (uint32_t));
}

void
cigar_add (cigar_t * c, uint32_t cigar_elem)
{
  cigar_resize (c, c->n + 1);
  c->c[c->n] = cigar_elem;
  c->n += 1;
}

void
cigar_copy (cigar_t * dst, cigar_t * src)
{
  cigar_resize (dst, src->n);
  memcpy (dst->c, src->c, (size_t) src->n * sizeof (uint32_t));
  dst->n = src->n;
}

void
cigar_reverse (cigar_t * c)
{
  int32_t lo, hi;

  if (c->n <= 0)
    warn_mesg ("empty_cigar");
  for (lo = 0, hi = c->n - 1; lo < hi; ++lo, --hi) {
    uint32_t tmp = c->c[lo];
    c->c[lo] = c->c[hi];
    c->c[hi] = tmp;
  }
}

void
cigar_dump (FILE * fp, cigar_t * c)
{
  int32_t i;
  for (i = 0; i < c->n; ++i)
    fprintf (fp, "%d%c", LEN_OF (c->c[i]), OP_CHARS[OP_OF (c->c[i])]);
}

int
cigar_write (FILE * fp, cigar_t * c)
{
  fwrite (&c->n, sizeof (int32_t), 1, fp);
  fwrite (c->c, sizeof (uint32_t), c->n, fp);
  return 0;
}

int
cigar_read (FILE * fp, cigar_t * c)
{
  if (fread (&c->n, sizeof (int32_t), 1, fp) != 1)
    return -1;
  cigar_resize (c, c->n);
  if (fread (c->c, sizeof (uint32_t), c->n, fp) != (size_t) c->n)
    return -1;
  return 0;
}

void
cigar2str (cigar_t * c, str_t * s)
{
  int32_t i, at;

  if (c->n <= 0) {
    str_resize (s, 1);
    s->s[0] = '*';
    s->s[1] = '\0';
    s->l = 1;
    return;
  }
  s->l = 0;
  str_resize (s, c->n << 2);
  for (i = 0; i < c->n; ++i) {
    uint32_t len = LEN_OF (c->c[i]);
    at = s->l;
    s->l += int2deci_nbits (len) + 1;
    str_resize (s, s->l);
    sprintf (s->s + at, "%d%c", len, OP_CHARS[OP_OF (c->c[i])]);
  }
}

static int32_t
cigar_span (cigar_t * c, int which)
{
  int32_t i, total = 0;
  for (i = 0; i < c->n; ++i)
    if (OP_CONSUMES (OP_OF (c->c[i])) & which)
      total += LEN_OF (c->c[i]);
  return total;
}

int32_t cigar2ref_len (cigar_t * c) { return cigar_span (c, 2); }
int32_t cigar2qry_len (cigar_t * c) { return cigar_span (c, 1); }

/* drop zero-length elements and deletions that would lead the alignment */
int
cigar_cleanup (cigar_t * in, cigar_t * out)
{
  int32_t i;

  cigar_clear (out);
  for (i = 0; i < in->n; ++i) {
    if (LEN_OF (in->c[i]) == 0)
      continue;
    if (out->n == 0 && OP_OF (in->c[i]) == C_DEL)
      continue;
    cigar_add (out, in->c[i]);
  }
  return 0;
}

int
cigar_unclip (cigar_t * in, cigar_t * out)
{
  int32_t i;

  cigar_clear (out);
  for (i = 0; i < in->n; ++i) {
    uint32_t op = OP_OF (in->c[i]);
    if (op == C_SOFT || op == C_HARD || op == C_PAD)
      continue;
    cigar_add (out, in->c[i]);
  }
  return 0;
}

int
cigar_has_zero_size_element (cigar_t * c)
{
  int32_t i;
  for (i = 0; i < c->n; ++i)
    if (LEN_OF (c->c[i]) == 0)
      return 1;
  return 0;
}
