/* sw_batch.h — additions next to the reference's sw.h (which stays as it is): a batched
 * entry point, the traceback-mode switch and the destructor the reference never wrote. */
#ifndef SW_BATCH_H
#define SW_BATCH_H

#include <stdint.h>
#include "sw.h"
#include "gcgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* GCG_SW_ASIS (default, bit-exact to the shipped sw.c) or GCG_SW_FIXED; also GC_SW_MODE=fixed */
void sw_set_traceback_mode (int mode);

/* n alignments in one device batch with the aligner's current parameters and border state.
 * Pair p is qry[qoff[p]..qoff[p+1]) vs tgt[toff[p]..toff[p+1]) (integer-coded symbols, sw.c:216).
 * results[n]; *cigar_pool must be released with gcg_free. */
int sw_align_batch (sw_t * sw, int64_t n, const char * qry, const int64_t * qoff, const char * tgt, const int64_t * toff,
                    gcg_sw_result * results, uint32_t ** cigar_pool, int64_t * n_cigar_pool);

void sw_free (sw_t * sw);

#ifdef __cplusplus
}
#endif
#endif
