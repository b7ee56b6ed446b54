// Micro-benchmark: cp.async.bulk.tensor.4d global->shared with the MAP kernel's box shape:
// box = (8 words, 4 blocks, L rows, P planes) out of tensor (8, nb, W, 6) with strides (., 6*ps*2, 32, ps*2) bytes.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ void mbar_init(unsigned b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned b, unsigned n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned b, unsigned ph)
{
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma4(unsigned dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, unsigned bar)
{
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst), "l"(tm), "r"(c0),
               "r"(c1), "r"(c2), "r"(c3), "r"(bar)
               : "memory");
}
template <int D>
__global__ void k(const CUtensorMap* tm, int box_bytes, int L, int W, int nb, int reps, unsigned long long* out, unsigned* sink)
{
  extern __shared__ __align__(128) char sm[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  char* base = sm + (size_t)w * (D * box_bytes + 128);
  unsigned bs = (unsigned)__cvta_generic_to_shared(base);
  unsigned bars = bs + D * box_bytes;
  if (lane == 0) { for (int i = 0; i < D; i++) mbar_init(bars + 8 * i, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  const int gw = blockIdx.x * nw + w;      // global warp = group of 4 blocks
  const int blk0 = (gw * 4) % nb;
  const int chunks = W / L;
  unsigned acc = 0;
  long long t0 = clock64();
  int it = 0;
  for (int r = 0; r < reps; r++)
    for (int c = 0; c < chunks + D; c++, it++) {
      const int s = it % D;
      if (c >= D) {
        mbar_wait(bars + 8 * s, ((it / D) - 1) & 1);
        acc += *(volatile unsigned*)(base + s * box_bytes + lane * 4);
        __syncwarp();
      } else if (r > 0) { mbar_wait(bars + 8 * s, ((it / D) - 1) & 1); }
      if (c < chunks) {
        if (lane == 0) { mbar_expect(bars + 8 * s, box_bytes); tma4(bs + s * box_bytes, tm, 0, blk0, W - (c + 1) * L, 0, bars + 8 * s); }
      } else { if (lane == 0) mbar_expect(bars + 8 * s, 0); }
      __syncwarp();
    }
  long long t1 = clock64();
  if (lane == 0) { out[gw] = t1 - t0; sink[gw] = acc; }
}
int main(int argc, char** argv)
{
  const int K = 6144, N = 16, W = K / N, ps = K;
  const int nb = argc > 1 ? atoi(argv[1]) : 9472;
  size_t elems = (size_t)nb * 6 * ps;
  int16_t* ws; cudaMalloc(&ws, elems * 2); cudaMemset(ws, 1, elems * 2);
  EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr);
  if (!enc) { printf("no encode fn\n"); return 1; }
  unsigned long long* out; cudaMalloc(&out, 148 * 32 * 8); unsigned* sink; cudaMalloc(&sink, 148 * 32 * 4);
  for (int P : {3, 2, 1})
    for (int L : {4, 8, 16})
      for (int nw : {8, 16}) {
        CUtensorMap tm;
        cuuint64_t dims[4] = {8, (cuuint64_t)nb, (cuuint64_t)W, 6};
        cuuint64_t strides[3] = {(cuuint64_t)6 * ps * 2, 32, (cuuint64_t)ps * 2};
        cuuint32_t box[4] = {8, 4, (cuuint32_t)L, (cuuint32_t)P};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, ws, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        CUtensorMap* dtm; cudaMalloc(&dtm, sizeof(tm)); cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
        const int box_bytes = 32 * 4 * L * P;
        const int D = 3;
        size_t smem = (size_t)nw * (D * box_bytes + 128);
        if (smem > 220 * 1024) { cudaFree(dtm); continue; }
        cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int reps = 8;
        k<3><<<148, nw * 32, smem>>>(dtm, box_bytes, L, W, nb, 1, out, sink);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        k<3><<<148, nw * 32, smem>>>(dtm, box_bytes, L, W, nb, reps, out, sink);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        cudaError_t e = cudaGetLastError();
        unsigned long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        double ops_per_sm = (double)nw * reps * (W / L);
        printf("planes=%d L=%2d warps=%2d box=%5d B : %.0f cyc/op/SM (%.0f cyc per warp-chunk)  %.1f B/cyc/SM  chip %.0f GB/s  (%s)\n", P, L, nw, box_bytes,
               (double)h / ops_per_sm, (double)h / (reps * (W / L)), ops_per_sm * box_bytes / (double)h, 148.0 * ops_per_sm * box_bytes / (ms * 1e-3) / 1e9,
               cudaGetErrorString(e));
        cudaFree(dtm);
      }
  return 0;
}
