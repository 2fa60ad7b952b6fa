// Micro-benchmark (exploration, not part of the product): streaming bandwidth of 1D cp.async.bulk (TMA) copies of
// 8832-B tiles into shared memory versus per-thread 16-B loads, per SM occupancy / pipeline depth.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_bw bulk_bw.cu && ./bulk_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define TILE 8832
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t par)
{
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
  return ok;
}
template <int S>
__global__ void k_bulk(const char* src, size_t ntiles, int tiles_per_item, unsigned long long* sink)
{
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * tiles_per_item * TILE);
  const int t = threadIdx.x;
  if(t == 0)
  {
    for(int s = 0; s < S; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bars[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t nitems = ntiles / tiles_per_item;
  size_t mine = (nitems - blockIdx.x + gridDim.x - 1) / gridDim.x;
  if(blockIdx.x >= nitems) mine = 0;
  unsigned long long acc = 0;
  auto issue = [&](size_t k)
  {
    const int s = (int)(k % S);
    const size_t item = blockIdx.x + k * gridDim.x;
    const uint32_t bar = s32(&bars[s]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tiles_per_item * TILE) : "memory");
    for(int q = 0; q < tiles_per_item; q++)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + ((size_t)s * tiles_per_item + q) * TILE)),
                   "l"(src + (item * tiles_per_item + q) * TILE), "r"(TILE), "r"(bar)
                   : "memory");
  };
  if(t == 0)
    for(size_t k = 0; k < (size_t)S && k < mine; k++) issue(k);
  for(size_t k = 0; k < mine; k++)
  {
    const int s = (int)(k % S);
    const uint32_t par = (uint32_t)((k / S) & 1);
    while(!try_wait(s32(&bars[s]), par)) {}
    acc += reinterpret_cast<const unsigned long long*>(smem + (size_t)s * tiles_per_item * TILE)[t];
    __syncthreads();
    if(t == 0 && k + S < mine) issue(k + S);
  }
  if(acc == 0x1234567) sink[0] = acc;
}
// per-thread loads: each thread reads 16 B pieces, 256 threads, U pieces in flight per thread
template <int U>
__global__ void k_ldg(const char* src, size_t nbytes, unsigned long long* sink)
{
  const size_t n16 = nbytes / 16;
  const uint4* p = reinterpret_cast<const uint4*>(src);
  unsigned long long acc = 0;
  for(size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x; i + (size_t)(U - 1) * blockDim.x < n16; i += (size_t)gridDim.x * blockDim.x * U)
  {
    uint4 v[U];
#pragma unroll
    for(int u = 0; u < U; u++) v[u] = __ldcs(p + i + (size_t)u * blockDim.x);
#pragma unroll
    for(int u = 0; u < U; u++) acc += v[u].x ^ v[u].w;
  }
  if(acc == 0x1234567) sink[0] = acc;
}
template <typename F>
float timeit(F f, int reps)
{
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  cudaEventRecord(a);
  for(int i = 0; i < reps; i++) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}
int main()
{
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  unsigned long long* sink; cudaMalloc(&sink, 8);
  for(size_t ntiles : {(size_t)262144 * 2, (size_t)4096})   // 4.6 GB (DRAM) and 36 MB (L2 resident)
  {
    char* src; cudaMalloc(&src, ntiles * TILE); cudaMemset(src, 1, ntiles * TILE);
    const double gb = ntiles * (double)TILE / 1e9;
    const int reps = ntiles > 100000 ? 3 : 50;
    printf("== %.1f MB source\n", gb * 1e3);
    auto run = [&](auto kern, int S, int tpi, int ctas_per_sm, const char* name)
    {
      const size_t smem = (size_t)S * tpi * TILE + 64;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      float ms = timeit([&] { kern<<<sms * ctas_per_sm, 256, smem>>>(src, ntiles, tpi, sink); }, reps);
      cudaError_t e = cudaGetLastError();
      printf("%s stages %d x %d tiles/item, %d CTAs/SM: %.3f ms  %.0f GB/s %s\n", name, S, tpi, ctas_per_sm, ms, gb / ms * 1e3, e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    run(k_bulk<2>, 2, 2, 3, "bulk");
    run(k_bulk<3>, 3, 2, 3, "bulk");
    run(k_bulk<3>, 3, 2, 4, "bulk");
    run(k_bulk<4>, 4, 2, 3, "bulk");
    run(k_bulk<6>, 6, 2, 2, "bulk");
    run(k_bulk<6>, 6, 1, 4, "bulk");
    run(k_bulk<8>, 8, 1, 3, "bulk");
    run(k_bulk<12>, 12, 1, 2, "bulk");
    for(int c : {2, 4, 8})
    {
      float ms = timeit([&] { k_ldg<4><<<sms * c, 256>>>(src, ntiles * TILE, sink); }, reps);
      printf("ldg.128 x4 per thread, %d CTAs/SM: %.3f ms %.0f GB/s\n", c, ms, gb / ms * 1e3);
      ms = timeit([&] { k_ldg<8><<<sms * c, 256>>>(src, ntiles * TILE, sink); }, reps);
      printf("ldg.128 x8 per thread, %d CTAs/SM: %.3f ms %.0f GB/s\n", c, ms, gb / ms * 1e3);
    }
    cudaFree(src);
  }
  return 0;
}
