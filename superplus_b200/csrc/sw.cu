// placeholder until the SW kernels land (next commit)
#include "gcg_internal.cuh"
extern "C" int gcg_sw_batch (gcg_ctx *, const gcg_sw_params *, int, const char *, const int64_t *, const char *, const int64_t *, int64_t, gcg_sw_result *, uint32_t **, int64_t *) { gcg_set_error ("sw: not built yet"); return GCG_EINVAL; }
extern "C" int gcg_swbatch_upload (gcg_ctx *, const char *, const int64_t *, const char *, const int64_t *, int64_t, gcg_swbatch **) { gcg_set_error ("sw: not built yet"); return GCG_EINVAL; }
extern "C" int gcg_swbatch_align (gcg_ctx *, gcg_swbatch *, const gcg_sw_params *, int) { gcg_set_error ("sw: not built yet"); return GCG_EINVAL; }
extern "C" int gcg_swbatch_download (gcg_ctx *, gcg_swbatch *, gcg_sw_result *, uint32_t **, int64_t *) { gcg_set_error ("sw: not built yet"); return GCG_EINVAL; }
extern "C" int64_t gcg_swbatch_cells (const gcg_swbatch *) { return 0; }
extern "C" int gcg_swbatch_path_counts (const gcg_swbatch *, int64_t *) { return GCG_EINVAL; }
extern "C" void gcg_swbatch_free (gcg_swbatch *) {}
