// DFMA throughput vs resident warps per SM and ILP (independent chains per thread).  Development probe:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_ilp dfma_ilp.cu && ./dfma_ilp
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__(256) k(double* out, int iters, double seed) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = seed + threadIdx.x + i;
  const double m = 0.999999, c = 1e-7;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  if (s == 123.456) out[0] = s;
}

template <int ILP>
void run(int blocks_per_sm, int threads, int sms, double* d) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int iters = 20000 / ILP;
  // dynamic smem to cap residency at blocks_per_sm
  size_t smem = (227 * 1024) / blocks_per_sm - 2048;
  cudaFuncSetAttribute(k<ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<ILP><<<sms * blocks_per_sm, threads, smem>>>(d, 10, 1.0);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<ILP><<<sms * blocks_per_sm, threads, smem>>>(d, iters, 1.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double flops = 2.0 * 8 * ILP * (double)iters * sms * blocks_per_sm * threads;
  printf("warps/SM %2d  ILP %d : %6.2f TFLOP/s\n", blocks_per_sm * threads / 32, ILP, flops / (ms * 1e-3) / 1e12);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double* d;
  cudaMalloc(&d, 8);
  for (int bps : {1, 2, 4, 8}) {
    for (int th : {128, 256}) {
      run<1>(bps, th, p.multiProcessorCount, d);
      run<2>(bps, th, p.multiProcessorCount, d);
      run<4>(bps, th, p.multiProcessorCount, d);
      run<8>(bps, th, p.multiProcessorCount, d);
    }
  }
  return 0;
}
