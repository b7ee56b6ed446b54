// Issue rate of the 32-bit integer add / max forms the time-parallel kernels (map_scan.cuh) could use on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int OP>
__global__ void __launch_bounds__(256) k(int* out, int iters, int seed, long long* cyc)
{
  long long t0 = clock64();
  int a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = seed * (threadIdx.x + 1) + i * 65539;
  const int g = seed | 3, h = seed ^ 0x30005;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (OP == 0) a[i] = __viaddmax_s32(a[i], g, a[(i + 3) & 7]);          // VIADDMNMX (32 bit)
        if (OP == 1) a[i] = max(a[i], a[(i + 1) & 7]);                         // VIMNMX / IMNMX
        if (OP == 2) a[i] = a[i] + g;                                          // IADD / VIADD
        if (OP == 3) a[i] = __vimax3_s32(a[i], a[(i + 1) & 7], a[(i + 2) & 7]); // VIMNMX3
        if (OP == 4) { int t; asm volatile("add.s32 %0, %1, %2;" : "=r"(t) : "r"(a[i]), "r"(g)); asm volatile("max.s32 %0, %1, %2;" : "=r"(a[i]) : "r"(t), "r"(a[(i + 3) & 7])); } // un-fused add, max
        if (OP == 5) a[i] = (int)__viaddmax_s16x2((unsigned)a[i], (unsigned)g, (unsigned)a[(i + 3) & 7]);
      }
    }
  }
  int r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char* name, int* out, int warps_per_sm, double inst_per_iter)
{
  const int blocks = 148 * warps_per_sm / 8, iters = 4096;
  static long long* dc = nullptr; if (!dc) cudaMalloc(&dc, 8);
  k<OP><<<blocks, 256>>>(out, 16, 0x1234567, dc);
  k<OP><<<blocks, 256>>>(out, iters, 0x1234567, dc);
  cudaDeviceSynchronize();
  long long hc; cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s warps/SM=%2d : %.3f warp-inst/clk/SMSP (block 0: %lld cycles)\n", name, warps_per_sm, (double)warps_per_sm / 4 * iters * inst_per_iter / (double)hc, hc);
}
int main()
{
  int* out; cudaMalloc(&out, 148 * 64 * 256 * 4);
  for (int w : {8, 32}) {
    run<0>("VIADDMNMX 32", out, w, 32);
    run<1>("max 32", out, w, 32);
    run<2>("add 32", out, w, 32);
    run<3>("max3 32", out, w, 32);
    run<4>("add + max (2 instr)", out, w, 64);
    run<5>("VIADDMNMX.S16x2", out, w, 32);
  }
  return 0;
}
