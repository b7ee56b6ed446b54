// Micro-benchmark: cp.async.bulk global->shared throughput per SM as a function of copy size (B200).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ void mbar_init(unsigned b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned b, unsigned n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned b, unsigned ph)
{
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk(unsigned dst, const void* src, unsigned bytes, unsigned bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// each warp: ring of D stages; per iteration lane<ncopy issues one copy of `bytes`
template <int D>
__global__ void k(const char* g, size_t gbytes, int bytes, int ncopy, int iters, unsigned long long* out, int stride_blocks)
{
  extern __shared__ __align__(128) char sm[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const unsigned per_stage = (unsigned)bytes * ncopy;
  char* base = sm + (size_t)w * (D * per_stage + 64);
  unsigned bs = (unsigned)__cvta_generic_to_shared(base);
  unsigned bars = bs + D * per_stage;
  if (lane == 0) { for (int i = 0; i < D; i++) mbar_init(bars + 8 * i, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  size_t off = ((size_t)(blockIdx.x * nw + w) * 1315423911ull) % (gbytes / 2);
  off &= ~(size_t)127;
  const size_t step = (size_t)stride_blocks * 128;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters + D; it++) {
    const int s = it % D;
    if (it >= D) { // consume
      mbar_wait(bars + 8 * s, ((it / D) - 1) & 1);
      acc += *(volatile unsigned*)(base + s * per_stage + lane * 4);
      __syncwarp();
    }
    if (it < iters) {
      if (lane == 0) mbar_expect(bars + 8 * s, per_stage);
      __syncwarp();
      if (lane < ncopy) {
        size_t o = (off + (size_t)(it * ncopy + lane) * step) % (gbytes - 8192);
        o &= ~(size_t)127;
        bulk(bs + s * per_stage + lane * bytes, g + o, bytes, bars + 8 * s);
      }
    }
  }
  long long t1 = clock64();
  if (lane == 0) { out[(blockIdx.x * nw + w) * 2] = t1 - t0; out[(blockIdx.x * nw + w) * 2 + 1] = acc; }
}
int main(int argc, char** argv)
{
  size_t gbytes = (size_t)(argc > 1 ? atoi(argv[1]) : 64) << 20; // working set MB (64 = L2 resident, 2048 = DRAM)
  char* g; cudaMalloc(&g, gbytes); cudaMemset(g, 1, gbytes);
  unsigned long long* out; cudaMalloc(&out, 148 * 32 * 16);
  int sizes[] = {128, 256, 512, 768, 1024, 2048, 3072, 4096};
  int warps_list[] = {8, 16};
  for (int nw : warps_list)
    for (int ncopy : {1, 4, 12})
      for (int bytes : sizes) {
        const int D = 3;
        size_t smem = (size_t)nw * (D * bytes * ncopy + 64);
        if (smem > 200 * 1024) continue;
        cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int iters = 2000;
        k<3><<<148, nw * 32, smem>>>(g, gbytes, bytes, ncopy, 50, out, 37);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        k<3><<<148, nw * 32, smem>>>(g, gbytes, bytes, ncopy, iters, out, 37);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        cudaError_t e = cudaGetLastError();
        unsigned long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        double cyc = (double)h[0];
        double ops_per_sm = (double)nw * ncopy * iters;
        printf("ws=%zuMB warps=%2d copies/iter=%2d bytes=%4d : %.1f cyc/op/SM  %.1f B/cyc/SM  chip %.0f GB/s  (%s)\n", gbytes >> 20, nw, ncopy, bytes, cyc / ops_per_sm,
               ops_per_sm * bytes / cyc, 148.0 * ops_per_sm * bytes / (ms * 1e-3) / 1e9, cudaGetErrorString(e));
      }
  return 0;
}
