// Roofline denominators measured on the device the library runs on ("measure
// R_int16, do not assume the lane count"; B200_PROFILING: peaks are measured, not nominal).
//   gcg_ubench_int16   issue rate of VIADDMNMX.S16x2 (the packed add+max the SW kernel is built
//                      from) with 8 independent chains per thread on every SM
//   gcg_ubench_hbm     device-to-device copy bandwidth (read + write bytes) of a 1 GiB buffer
//   gcg_ubench_gather  rate of independent random 32-byte bucket loads (the probe the k-mer search
//                      issues, LDG.E.256 with one sector per lane) from a table of a given size:
//                      the ceiling of any one-probe-per-k-mer design on this part, L2- or HBM-resident
#include "gcg_internal.cuh"

#define UB_ILP 8
#define UB_ITERS 4096

__global__ void __launch_bounds__ (256)
ubench_int16_kernel (uint32_t * out, uint32_t seed)
{
  uint32_t a[UB_ILP], b = seed | 0x00010001u, c = seed * 3u;
#pragma unroll
  for (int i = 0; i < UB_ILP; ++i) a[i] = seed + i * 0x01010101u + threadIdx.x;
  for (int it = 0; it < UB_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < UB_ILP; ++i) {
      a[i] = __viaddmax_s16x2 (a[i], b, c);
      asm volatile ("" : "+r"(a[i]));
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < UB_ILP; ++i) r ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

extern "C" int gcg_ubench_int16 (gcg_ctx * ctx, double * lane_ops_per_s)
{
  GCG_CHECK (ctx && lane_ops_per_s, GCG_EINVAL, "gcg_ubench_int16: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  int blocks = ctx->sm_count * 8;
  uint32_t * d;
  GCG_CUDA (cudaMalloc (&d, (size_t) blocks * 256 * 4));
  cudaEvent_t e0, e1;
  GCG_CUDA (cudaEventCreate (&e0));
  GCG_CUDA (cudaEventCreate (&e1));
  double best = 0;
  for (int rep = 0; rep < 4; ++rep) {
    GCG_CUDA (cudaEventRecord (e0, ctx->stream));
    { gcg_kscope ks (ctx, "ubench_int16");
      ubench_int16_kernel<<<blocks, 256, 0, ctx->stream>>> (d, 12345u + rep); }
    GCG_CUDA (cudaEventRecord (e1, ctx->stream));
    GCG_CUDA (cudaEventSynchronize (e1));
    float ms = 0;
    GCG_CUDA (cudaEventElapsedTime (&ms, e0, e1));
    double rate = (double) blocks * 256 * UB_ITERS * UB_ILP / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy (e0); cudaEventDestroy (e1);
  cudaFree (d);
  *lane_ops_per_s = best;
  return GCG_OK;
}

extern "C" int gcg_ubench_hbm (gcg_ctx * ctx, double * bytes_per_s)
{
  GCG_CHECK (ctx && bytes_per_s, GCG_EINVAL, "gcg_ubench_hbm: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  const size_t n = (size_t) 1 << 30;
  char * a, * b;
  GCG_CUDA (cudaMalloc (&a, n));
  GCG_CUDA (cudaMalloc (&b, n));
  GCG_CUDA (cudaMemsetAsync (a, 1, n, ctx->stream));
  cudaEvent_t e0, e1;
  GCG_CUDA (cudaEventCreate (&e0));
  GCG_CUDA (cudaEventCreate (&e1));
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    GCG_CUDA (cudaEventRecord (e0, ctx->stream));
    GCG_CUDA (cudaMemcpyAsync (b, a, n, cudaMemcpyDeviceToDevice, ctx->stream));
    GCG_CUDA (cudaEventRecord (e1, ctx->stream));
    GCG_CUDA (cudaEventSynchronize (e1));
    float ms = 0;
    GCG_CUDA (cudaEventElapsedTime (&ms, e0, e1));
    double rate = 2.0 * (double) n / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy (e0); cudaEventDestroy (e1);
  cudaFree (a); cudaFree (b);
  *bytes_per_s = best;
  return GCG_OK;
}

template <int ILP>
__global__ void __launch_bounds__ (256)
ubench_gather_kernel (const unsigned long long * __restrict__ tab, uint32_t n_bucket, int iters, unsigned long long * out)
{
  uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  unsigned long long acc = 0;
  for (int it = 0; it < iters; ++it) {
    unsigned long long a[ILP], b[ILP], c[ILP], d[ILP];
#pragma unroll
    for (int u = 0; u < ILP; ++u) {
      x ^= x << 13; x ^= x >> 17; x ^= x << 5;                 // xorshift32: every lane its own sector
      const unsigned long long * p = tab + 4ULL * __umulhi (x, n_bucket);
      asm volatile ("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a[u]), "=l"(b[u]), "=l"(c[u]), "=l"(d[u]) : "l"(p));
    }
#pragma unroll
    for (int u = 0; u < ILP; ++u) acc += a[u] ^ b[u] ^ c[u] ^ d[u];
  }
  if (acc == 0x1234567ULL) out[0] = acc;                        // keeps the loads alive
}

extern "C" int gcg_ubench_gather (gcg_ctx * ctx, int64_t table_bytes, double * lookups_per_s)
{
  GCG_CHECK (ctx && lookups_per_s && table_bytes >= 4096 && table_bytes <= ((int64_t) 64 << 30), GCG_EINVAL, "gcg_ubench_gather: bad argument");
  GCG_CUDA (cudaSetDevice (ctx->device));
  uint32_t n_bucket = (uint32_t) (table_bytes / 32);
  unsigned long long * tab, * out;
  GCG_CUDA (gcg_dmalloc (ctx, &tab, (size_t) n_bucket * 32));
  GCG_CUDA (gcg_dmalloc (ctx, &out, 64));
  GCG_CUDA (cudaMemsetAsync (tab, 0x5a, (size_t) n_bucket * 32, ctx->stream));
  cudaEvent_t e0, e1;
  GCG_CUDA (cudaEventCreate (&e0));
  GCG_CUDA (cudaEventCreate (&e1));
  const int blocks = ctx->sm_count * 8, iters = 64;
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    const int ilp = (rep & 1) ? 8 : 4;
    GCG_CUDA (cudaEventRecord (e0, ctx->stream));
    { gcg_kscope ks (ctx, "ubench_gather");
      if (ilp == 4) ubench_gather_kernel<4><<<blocks, 256, 0, ctx->stream>>> (tab, n_bucket, iters * 2, out);
      else ubench_gather_kernel<8><<<blocks, 256, 0, ctx->stream>>> (tab, n_bucket, iters, out); }
    GCG_CUDA (cudaEventRecord (e1, ctx->stream));
    GCG_CUDA (cudaEventSynchronize (e1));
    float ms = 0;
    GCG_CUDA (cudaEventElapsedTime (&ms, e0, e1));
    double rate = (double) blocks * 256 * iters * 8 / (ms * 1e-3);
    if (rep >= 2 && rate > best) best = rate;
  }
  cudaEventDestroy (e0); cudaEventDestroy (e1);
  gcg_dfree (ctx, tab); gcg_dfree (ctx, out);
  GCG_CUDA (cudaStreamSynchronize (ctx->stream));
  *lookups_per_s = best;
  return GCG_OK;
}
