// Per-instruction issue rate of the packed int16x2 ops on sm_100a (8 independent chains per thread).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
typedef unsigned u32;
template <int OP>
__global__ void __launch_bounds__(256) k(u32* out, int iters, u32 seed, long long* cyc)
{
  long long t0 = clock64();
  u32 a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = seed * (threadIdx.x + 1) + i * 0x00010003u;
  const u32 g = seed | 0x00010001u, h = seed ^ 0x00030005u;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (OP == 0) a[i] = __vadd2(a[i], g);
        if (OP == 1) a[i] = __vmaxs2(a[i], a[(i + 1) & 7]);
        if (OP == 2) a[i] = __viaddmax_s16x2(a[i], g, a[(i + 3) & 7]);
        if (OP == 3) a[i] = __vimax3_s16x2(a[i], a[(i + 1) & 7], a[(i + 2) & 7]);
        if (OP == 4) a[i] = __vsub2(a[i], a[(i + 5) & 7]);
        if (OP == 5) a[i] = a[i] + g;                       // plain 32-bit IADD
        if (OP == 6) a[i] = (a[i] & g) | (a[(i + 1) & 7] & ~g); // LOP3
        if (OP == 7) a[i] = a[i] * h + g;                   // IMAD
        if (OP == 8) { a[i] = __vadd2(a[i], g); a[(i + 4) & 7] = a[(i + 4) & 7] * h + g; } // ALU + FMA pipes
        if (OP == 9) a[i] = __vmins2(a[i], a[(i + 1) & 7]);
      }
    }
  }
  u32 r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char* name, u32* out, int warps_per_sm, double inst_per_iter)
{
  const int blocks = 148 * warps_per_sm / 8, iters = 4096;
  static long long* dc = nullptr; if (!dc) cudaMalloc(&dc, 8);
  k<OP><<<blocks, 256>>>(out, 16, 0x1234567u, dc);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  k<OP><<<blocks, 256>>>(out, iters, 0x1234567u, dc);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double inst = (double)blocks * 8 * iters * inst_per_iter; // warp-instructions
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double cyc = ms * 1e-3 * clk * 1e3;
  long long hc; cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s warps/SM=%2d : %.3f warp-inst/clk/SMSP by wall clock @%d kHz; %.3f by clock64 (block 0: %lld cycles, %.2f ms -> %.0f MHz)\n", name, warps_per_sm,
         inst / (148.0 * 4) / cyc, clk, (double)warps_per_sm / 4 * iters * inst_per_iter / (double)hc, hc, ms, hc / (ms * 1e3));
}
int main()
{
  u32* out; cudaMalloc(&out, 148 * 64 * 256 * 4);
  for (int w : {8, 16, 64}) {
    run<0>("VIADD.16x2 (vadd2)", out, w, 32);
    run<1>("VIMNMX.S16x2 (vmaxs2)", out, w, 32);
    run<2>("VIADDMNMX.S16x2", out, w, 32);
    run<3>("VIMNMX3.S16x2", out, w, 32);
    run<4>("vsub2", out, w, 32);
    run<5>("IADD 32", out, w, 32);
    run<6>("LOP3", out, w, 32);
    run<7>("IMAD", out, w, 32);
    run<8>("vadd2 + IMAD pair", out, w, 64);
    run<9>("vmins2", out, w, 32);
  }
  return 0;
}
