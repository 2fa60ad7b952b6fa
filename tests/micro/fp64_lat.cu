// exploratory: FP64 latency / throughput per SM on this part (nvcc -arch=sm_100a -O3 -o fp64_lat fp64_lat.cu)
#include <cstdio>
#include <cuda_runtime.h>
template <int CH, typename T>
__global__ void k(T* out, long long* cyc, int n, T a, T b)
{
  T v[CH];
#pragma unroll
  for(int i = 0; i < CH; i++) v[i] = (T)threadIdx.x + (T)i;
  __syncthreads();
  long long t0 = clock64();
  for(int it = 0; it < n; it++)
#pragma unroll
    for(int i = 0; i < CH; i++) v[i] = v[i] * a + b;
  long long t1 = clock64();
  T s = 0;
#pragma unroll
  for(int i = 0; i < CH; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if(threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CH, typename T>
void run(const char* name, int threads)
{
  T* out; long long* cyc;
  cudaMalloc(&out, sizeof(T) * 2048); cudaMalloc(&cyc, 8);
  const int n = 2000;
  k<CH, T><<<1, threads>>>(out, cyc, n, (T)1.0000001, (T)1e-9);
  k<CH, T><<<1, threads>>>(out, cyc, n, (T)1.0000001, (T)1e-9);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%s threads %4d chains %d: %.2f cycles per FMA per thread-chain step, %.2f cycles per warp-instruction per SMSP\n", name, threads, CH,
         (double)c / n, (double)c / n / CH / ((threads + 127) / 128));
}
int main()
{
  run<1, double>("f64", 32); run<1, double>("f64", 128); run<4, double>("f64", 128); run<1, double>("f64", 1024); run<4, double>("f64", 1024);
  run<1, float>("f32", 32); run<4, float>("f32", 1024);
  return 0;
}
