/* hash_func.c — replacement for gap_closer/hash_func.c (hash_func.h:13-17): the Blizzard
 * "crypt table" string hash.  Reachable only from bio.h:41-45 (chromosome-name map), which gc
 * never exercises; main.c:147,204 only calls hash_func_init / hash_func_free.  Host C, zero
 * dynamic weight.  The table is the well-known 0x500-entry LCG table
 * (seed 0x00100001, x -> (125 x + 3) mod 0x2AAAAB, two draws per entry).
 */
#include <ctype.h>
#include <stdint.h>
#include <stdlib.h>

#include "utils.h"
#include "hash_func.h"

#define CRYPT_ENTRIES 0x500

static uint64_t * crypt_tbl = NULL;

static uint64_t
lcg_next (uint64_t * state)
{
  *state = (*state * 125 + 3) % 0x2AAAAB;
  return *state & 0xFFFF;
}

void
hash_func_init (void)
{
  uint64_t state = 0x00100001;
  uint64_t col, plane;

  crypt_tbl = (uint64_t *) ckalloc (CRYPT_ENTRIES, sizeof (uint64_t));
  for (col = 0; col < 0x100; ++col)
    for (plane = 0; plane < 5; ++plane) {
      uint64_t hi = lcg_next (&state) << 16;
      uint64_t lo = lcg_next (&state);
      crypt_tbl[plane * 0x100 + col] = hi | lo;
    }
}

void
hash_func_free (void)
{
  free (crypt_tbl);
  crypt_tbl = NULL;
}

uint64_t
blizzard_hash_func (const char * key, int key_len, int dwHashType)
{
  uint64_t a = 0x7FED7FED, b = 0xEEEEEEEE, ch;
  int i;

  for (i = 0; i < key_len; ++i) {
    ch = (uint64_t) toupper (key[i]);
    a = crypt_tbl[((uint64_t) dwHashType << 8) + ch] ^ (a + b);
    b = ch + a + b + (b << 5) + 3;
  }
  return a;
}
