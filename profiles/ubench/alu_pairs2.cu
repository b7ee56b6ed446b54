// Pipe assignment of the packed int16x2 instructions on B200: pairs of independent chain sets (A, B); the total
// warp-instruction rate per clock and SM sub-partition tells whether A and B share an execution pipe (0.5) or not (~1).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
typedef unsigned u32;
enum { ADD, MAX, ADDMAX, MAX3, IMAD_, LOP, PRMT_, SHF_, NONE };
template <int O> __device__ __forceinline__ void op(u32 (&v)[8], int i)
{
  if (O == ADD) v[i] = __vadd2(v[i], v[(i + 3) & 7]);
  if (O == MAX) v[i] = __vmaxs2(v[i], v[(i + 3) & 7]);
  if (O == ADDMAX) v[i] = __viaddmax_s16x2(v[i], v[(i + 3) & 7], v[(i + 5) & 7]);
  if (O == MAX3) v[i] = __vimax3_s16x2(v[i], v[(i + 3) & 7], v[(i + 5) & 7]);
  if (O == IMAD_) v[i] = v[i] * v[(i + 3) & 7] + v[(i + 1) & 7];
  if (O == LOP) v[i] = (v[i] & v[(i + 5) & 7]) ^ v[(i + 1) & 7];
  if (O == PRMT_) v[i] = __byte_perm(v[i], v[(i + 1) & 7], 0x5410 + i);
  if (O == SHF_) v[i] = __funnelshift_r(v[i], v[(i + 1) & 7], 3);
}
template <int A, int B>
__global__ void __launch_bounds__(256) k(u32* out, int iters, u32 seed, long long* cyc)
{
  u32 a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = seed * (threadIdx.x + 1) + i * 0x00010003u; b[i] = a[i] ^ 0x5a5a5a5au; }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int i = 0; i < 8; i++) { op<A>(a, i); op<B>(b, i); }
    }
  }
  long long t1 = clock64();
  u32 r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r ^= a[i] ^ b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int A, int B>
void run(const char* name, u32* out, long long* dc)
{
  const int iters = 2048;
  k<A, B><<<148, 256>>>(out, 16, 0x1234567u, dc);
  k<A, B><<<148, 256>>>(out, iters, 0x1234567u, dc);
  cudaDeviceSynchronize();
  long long hc; cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
  const double n = B == NONE ? 32 : 64;
  printf("%-28s : %.3f warp-inst/clk/SMSP total\n", name, 2.0 * iters * n / (double)hc);
}
int main()
{
  u32* out; cudaMalloc(&out, 148 * 256 * 4); long long* dc; cudaMalloc(&dc, 8);
  run<ADD, NONE>("VIADD", out, dc); run<MAX, NONE>("VIMNMX", out, dc); run<ADDMAX, NONE>("VIADDMNMX", out, dc); run<MAX3, NONE>("VIMNMX3", out, dc);
  run<LOP, NONE>("LOP3", out, dc); run<PRMT_, NONE>("PRMT", out, dc); run<SHF_, NONE>("SHF", out, dc); run<IMAD_, NONE>("IMAD", out, dc);
  run<ADD, MAX>("VIADD + VIMNMX", out, dc);
  run<ADD, ADDMAX>("VIADD + VIADDMNMX", out, dc);
  run<ADD, MAX3>("VIADD + VIMNMX3", out, dc);
  run<MAX, ADDMAX>("VIMNMX + VIADDMNMX", out, dc);
  run<MAX, MAX3>("VIMNMX + VIMNMX3", out, dc);
  run<ADDMAX, MAX3>("VIADDMNMX + VIMNMX3", out, dc);
  run<ADDMAX, IMAD_>("VIADDMNMX + IMAD", out, dc);
  run<MAX, IMAD_>("VIMNMX + IMAD", out, dc);
  run<MAX, LOP>("VIMNMX + LOP3", out, dc);
  run<ADD, LOP>("VIADD + LOP3", out, dc);
  run<ADDMAX, LOP>("VIADDMNMX + LOP3", out, dc);
  run<ADDMAX, PRMT_>("VIADDMNMX + PRMT", out, dc);
  run<MAX, PRMT_>("VIMNMX + PRMT", out, dc);
  run<LOP, PRMT_>("LOP3 + PRMT", out, dc);
  run<LOP, SHF_>("LOP3 + SHF", out, dc);
  run<IMAD_, PRMT_>("IMAD + PRMT", out, dc);
  return 0;
}
