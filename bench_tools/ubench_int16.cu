// Microbenchmark: issue rate of the integer instructions the SW kernel is made of, per SM.
// Each thread runs ILP independent dependency chains of one instruction kind; the grid fills
// every SM with 2048 threads.  Output: lane-ops/clk/SM and T lane-ops/s for the whole chip.
// Used for the SW roofline denominator ("measure R_int16, do not assume it").
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ILP 8
#define ITERS 4096

template <int KIND>
__global__ void __launch_bounds__ (256) k (uint32_t * out, uint32_t seed)
{
  uint32_t a[ILP], b = seed | 0x00010001u, c = seed * 3u;
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = seed + i * 0x01010101u + threadIdx.x;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (KIND == 0) a[i] = __vadd2 (a[i], b);                       // VIADD.16x2
      if (KIND == 1) a[i] = __vmaxs2 (a[i], b ^ it);                 // VIMNMX.S16x2
      if (KIND == 2) a[i] = __viaddmax_s16x2 (a[i], b, c);           // VIADDMNMX.S16x2
      if (KIND == 3) a[i] = (a[i] & b) ^ c;                          // LOP3
      if (KIND == 4) asm volatile ("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));   // PRMT
      if (KIND == 5) a[i] = a[i] * 16u + b;                          // IMAD
      if (KIND == 6) a[i] = a[i] + b;                                // IADD3
      if (KIND == 7) a[i] = max ((int) a[i], (int) (b ^ it));        // VIMNMX (s32)
      if (KIND == 8) a[i] = __viaddmax_s32 (a[i], b, c);             // VIADDMNMX (s32)
      if (KIND == 9) { a[i] = __viaddmax_s16x2 (a[i], b, c); asm volatile ("" : "+r"(a[i])); a[i] = a[i] * 16u + b; }   // ALU + FMA pipe mix
      asm volatile ("" : "+r"(a[i]));     // keep every op: no folding or fusing across iterations
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int KIND> static void run (const char * name, int sms, double ops_per_iter)
{
  uint32_t * d;
  int blocks = sms * 8;
  cudaMalloc (&d, (size_t) blocks * 256 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate (&e0); cudaEventCreate (&e1);
  k<KIND><<<blocks, 256>>> (d, 12345u);
  cudaDeviceSynchronize ();
  cudaEventRecord (e0);
  k<KIND><<<blocks, 256>>> (d, 12345u);
  cudaEventRecord (e1);
  cudaEventSynchronize (e1);
  float ms = 0;
  cudaEventElapsedTime (&ms, e0, e1);
  double lane_ops = (double) blocks * 256 * ITERS * ILP * ops_per_iter;
  int clk_khz = 0;
  cudaDeviceGetAttribute (&clk_khz, cudaDevAttrClockRate, 0);
  printf ("%-22s %8.3f ms  %7.2f T lane-ops/s  %6.1f lane-ops/clk/SM (at %d MHz nominal)\n", name, ms, lane_ops / ms / 1e9,
          lane_ops / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1000);
  cudaFree (d);
}

int main ()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties (&p, 0);
  printf ("%s, %d SMs\n", p.name, p.multiProcessorCount);
  int s = p.multiProcessorCount;
  run<0> ("VIADD.16x2", s, 1); run<1> ("VIMNMX.S16x2", s, 1); run<2> ("VIADDMNMX.S16x2", s, 1);
  run<3> ("LOP3", s, 1); run<4> ("PRMT", s, 1); run<5> ("IMAD", s, 1); run<6> ("IADD3", s, 1);
  run<7> ("VIMNMX.S32", s, 1); run<8> ("VIADDMNMX.S32", s, 1); run<9> ("VIADDMNMX16x2+IMAD", s, 2);
  return 0;
}
