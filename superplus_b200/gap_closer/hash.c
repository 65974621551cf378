/* hash.c — replacement for gap_closer/hash.c: the generic host-side hash set / map behind
 * hash.h:60-110,240-248 (chained table over a growable item pool, keys are caller-owned
 * pointers, hashing and equality through function pointers).  It stays host C: its remaining
 * users (gc_graph.c:120, lfr.c:30, bio.c:42, the anchored sets main.c:68-72 builds) are not on
 * the hot path; the k-mer tables that were its hot instantiation now live in HBM (kmer.c).
 *
 * Written from the contract in hash.h, not from the reference implementation:
 *   - items live in `pool` in insertion order, `id` = index, `cnt` = number of items
 *   - a key that is already present only bumps `multi` (or revives a deleted item)
 *   - `slots[hash % size]` heads a singly linked chain through `next`
 *   - the table grows when cnt would exceed max = size * load_factor
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hash.h"
#include "utils.h"

/* relink every pooled item into a fresh slot array */
static void
xh_relink (xh_t * h)
{
  uint64_t i, os;
  xh_item_t * it;

  memset (h->slots, 0, h->size * sizeof (xh_item_t *));
  for (i = 0; i < h->cnt; ++i) {
    it = h->pool + i;
    os = it->hash_val % h->size;
    it->next = h->slots[os];
    h->slots[os] = it;
  }
}

static void
xh_grow (xh_t * h, uint64_t need)
{
  uint64_t size = h->size;

  while ((uint64_t) ((double) size * h->load_factor) < need)
    size = next_prime (size < 0xFFFFFFFU ? size << 1 : size + 0xFFFFFFFU);

  h->size = size;
  h->max = (uint64_t) ((double) size * h->load_factor);
  h->pool = (xh_item_t *) ckrealloc (h->pool, h->max * sizeof (xh_item_t));
  free (h->slots);
  h->slots = (xh_item_t **) ckalloc (h->size, sizeof (xh_item_t *));
  xh_relink (h);
}

static xh_item_t *
xh_locate (xh_t * h, const void * key, uint64_t hash_val)
{
  xh_item_t * it;

  for (it = h->slots[hash_val % h->size]; it != NULL; it = it->next)
    if (h->is_equal_func (key, it->key))
      return it;
  return NULL;
}

/* the one insertion routine behind _xh_set_add, _xh_set_add2 and _xh_map_add */
static int
xh_put (xh_t * h, void * key, void * val, int with_val, void ** old_key)
{
  uint64_t hash_val, os;
  xh_item_t * it;

  hash_val = h->hash_func (key);
  it = xh_locate (h, key, hash_val);
  if (it != NULL) {
    if (it->deleted) {
      it->multi = 1;
      it->deleted = 0;
      --h->del_c;
    } else
      ++it->multi;
    if (old_key) *old_key = it->key;
    return XH_EXIST;
  }

  if (h->cnt + 1 > h->max)
    xh_grow (h, h->cnt + 1);
  os = hash_val % h->size;
  it = h->pool + h->cnt;
  it->key = key;
  if (with_val) it->val = val;
  it->hash_val = hash_val;
  it->id = (int64_t) h->cnt;
  it->multi = 1;
  it->deleted = 0;
  it->next = h->slots[os];
  h->slots[os] = it;
  ++h->cnt;
  if (old_key) *old_key = NULL;
  return XH_NEW;
}

xh_t *
_xh_init (int64_t size, double load_factor, HashFunc hash_f, IsEqualFunc is_equal_f)
{
  xh_t * h;

  if (hash_f == NULL) err_mesg ("hash function is not set!");
  if (is_equal_f == NULL) err_mesg ("compare function is not set!");

  h = (xh_t *) ckmalloc (sizeof (xh_t));
  h->hash_func = hash_f;
  h->is_equal_func = is_equal_f;
  h->size = size < 256 ? 256 : (uint64_t) size;
  h->load_factor = (load_factor <= 0.0 || load_factor > 1.0) ? 0.75 : load_factor;
  h->max = (uint64_t) (h->size * h->load_factor);
  h->cnt = 0;
  h->del_c = 0;
  h->slots = (xh_item_t **) ckalloc (h->size, sizeof (xh_item_t *));
  h->pool = (xh_item_t *) ckalloc (h->max, sizeof (xh_item_t));
  return h;
}

void
_xh_clear (xh_t * h)
{
  h->cnt = 0;
  h->del_c = 0;
  memset (h->slots, 0, h->size * sizeof (xh_item_t *));
}

void
_xh_free (xh_t * h)
{
  free (h->slots);
  free (h->pool);
  free (h);
}

int
_xh_set_add (xh_t * h, void * key)
{
  return xh_put (h, key, NULL, 0, NULL);
}

int
_xh_set_add2 (xh_t * h, void * new_key, void ** old_key)
{
  return xh_put (h, new_key, NULL, 0, old_key);
}

int
_xh_set_search (xh_t * h, void * key)
{
  return xh_locate (h, key, h->hash_func (key)) ? XH_EXIST : XH_FAIL;
}

void *
_xh_set_search2 (xh_t * h, void * key)
{
  xh_item_t * it = xh_locate (h, key, h->hash_func (key));
  return it ? it->key : NULL;
}

xh_item_t *
_xh_set_search3 (xh_t * h, void * key)
{
  return xh_locate (h, key, h->hash_func (key));
}

int
_xh_map_add (xh_t * h, void * key, void * val)
{
  return xh_put (h, key, val, 1, NULL);
}

void *
_xh_map_search (xh_t * h, void * key)
{
  xh_item_t * it = xh_locate (h, key, h->hash_func (key));
  return it ? it->val : NULL;
}
