// Which instruction kinds co-issue with the packed int16x2 ops on B200?  Each test interleaves two independent sets
// of dependency chains (set A: 8 chains of VIADD.16x2, set B: 8 chains of the partner op) and reports the total
// warp-instruction rate per clock and SM sub-partition (clock64 of a 1-block-per-SM launch).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
typedef unsigned u32;
template <int OP>
__global__ void __launch_bounds__(256) k(u32* out, int iters, u32 seed, long long* cyc)
{
  u32 a[8]; float f[8]; u32 b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = seed * (threadIdx.x + 1) + i * 0x00010003u; f[i] = (float)(threadIdx.x + i); b[i] = a[i] ^ 0x5a5a5a5au; }
  const u32 g = seed | 0x00010001u;
  const float fg = (float)seed * 1e-9f;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        if (OP != 100 && OP != 101 && OP != 102) a[i] = __vadd2(a[i], a[(i + 3) & 7]);
        if (OP == 1) f[i] = f[i] + f[(i + 3) & 7];                        // FADD
        if (OP == 2) f[i] = fmaxf(f[i], f[(i + 1) & 7]);       // FMNMX
        if (OP == 3) b[i] = (b[i] & b[(i + 5) & 7]) ^ b[(i + 1) & 7];       // LOP3
        if (OP == 4) b[i] = __byte_perm(b[i], b[(i + 1) & 7], 0x5410 + i); // PRMT
        if (OP == 5) b[i] = __funnelshift_r(b[i], b[(i + 1) & 7], 3);      // SHF
        if (OP == 6) f[i] = fmaf(f[i], f[(i + 3) & 7], f[(i + 1) & 7]);    // FFMA
        if (OP == 7) b[i] = __vmaxs2(b[i], b[(i + 1) & 7]);    // second int chain (control: should stay 0.5 total)
        if (OP == 8) b[i] = b[i] * b[(i + 3) & 7] + b[(i + 1) & 7];                     // IMAD
        if (OP == 100) f[i] = f[i] + f[(i + 3) & 7];                       // FADD alone
        if (OP == 101) f[i] = fmaxf(f[i], f[(i + 1) & 7]);     // FMNMX alone
        if (OP == 102) f[i] = fmaf(f[i], f[(i + 3) & 7], f[(i + 1) & 7]);  // FFMA alone
      }
    }
  }
  long long t1 = clock64();
  u32 r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= a[i] ^ b[i] ^ __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP>
void run(const char* name, u32* out, long long* dc, double inst_per_iter)
{
  const int iters = 2048;
  k<OP><<<148, 256>>>(out, 16, 0x1234567u, dc);
  k<OP><<<148, 256>>>(out, iters, 0x1234567u, dc);
  cudaDeviceSynchronize();
  long long hc; cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s : %.3f warp-inst/clk/SMSP total (8 warps/SM)\n", name, 2.0 * iters * inst_per_iter / (double)hc);
}
int main()
{
  u32* out; cudaMalloc(&out, 148 * 256 * 4); long long* dc; cudaMalloc(&dc, 8);
  run<0>("VIADD.16x2 alone", out, dc, 32);
  run<1>("VIADD.16x2 + FADD", out, dc, 64);
  run<2>("VIADD.16x2 + FMNMX", out, dc, 64);
  run<3>("VIADD.16x2 + LOP3", out, dc, 64);
  run<4>("VIADD.16x2 + PRMT", out, dc, 64);
  run<5>("VIADD.16x2 + SHF", out, dc, 64);
  run<6>("VIADD.16x2 + FFMA", out, dc, 64);
  run<7>("VIADD.16x2 + VIMNMX.S16x2", out, dc, 64);
  run<8>("VIADD.16x2 + IMAD", out, dc, 64);
  run<100>("FADD alone", out, dc, 32);
  run<101>("FMNMX alone", out, dc, 32);
  run<102>("FFMA alone", out, dc, 32);
  return 0;
}
