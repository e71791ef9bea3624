// Is the FP64 tensor-core MMA (mma.sync .f64) a pipe of its own on B200, i.e. does it run beside DFMA?
//   (a) DFMA alone, (b) DMMA m8n8k4 alone, (c) DMMA m16n8k8 / m16n8k16 alone, (d) DFMA and DMMA interleaved in one warp,
//   (e) half the warps DFMA, half DMMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d[0]), "+d"(d[1])
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, "
      "{%0,%1,%2,%3};"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]),
        "d"(b[2]), "d"(b[3]));
}

// MODE 0: DFMA only (8 chains) ; 1: m8n8k4 only (8 accumulator pairs) ; 2: interleaved 8 DFMA + NM m8n8k4 per round ;
// 3: warps alternate (even warps DFMA, odd warps DMMA) ; 4: m16n8k8 only ; 5: m16n8k16 only
template <int MODE, int NM>
__global__ void __launch_bounds__(256) k(double* out, int iters, double seed) {
  double f[8], acc[8][2], acc4[4][4], a8[8], b4[4];
  const double m = 0.999999, c = 1e-7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    f[i] = seed + threadIdx.x + i;
    acc[i][0] = seed * i;
    acc[i][1] = seed + i;
    a8[i] = seed * 1e-3 * (i + 1);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    b4[i] = seed * 1e-3 * (i + 2);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc4[i][j] = seed + i + j;
  }
  const double a = seed * 1e-3, b = seed * 2e-3;
  const bool odd = (threadIdx.x >> 5) & 1;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (MODE == 0 || MODE == 2 || (MODE == 3 && !odd)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fma(f[i], m, c);
      }
      if (MODE == 1 || (MODE == 3 && odd)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(acc[i], a, b);
      }
      if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < NM; ++i) dmma884(acc[i], a, b);
      }
      if (MODE == 4) {
        const double a4[4] = {a8[0], a8[1], a8[2], a8[3]};
        const double b2[2] = {b4[0], b4[1]};
#pragma unroll
        for (int i = 0; i < 4; ++i) dmma1688(acc4[i], a4, b2);
      }
      if (MODE == 5) {
#pragma unroll
        for (int i = 0; i < 4; ++i) dmma16816(acc4[i], a8, b4);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i] + acc[i][0] + acc[i][1];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += acc4[i][j];
  if (s == 123.456) out[0] = s;
}

template <int MODE, int NM>
void run(const char* name, int sms, double* d, int ctas_per_sm) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 1000, blocks = sms * ctas_per_sm;
  k<MODE, NM><<<blocks, 256>>>(d, 10, 1.0);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE, NM><<<blocks, 256>>>(d, iters, 1.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double threads = (double)blocks * 256, warps = threads / 32;
  double dfma = 0, mac = 0;
  if (MODE == 0 || MODE == 2) dfma = 64.0 * iters * threads;
  if (MODE == 3) dfma = 64.0 * iters * threads / 2;
  if (MODE == 1) mac = 64.0 * iters * warps * 256;
  if (MODE == 2) mac = 8.0 * NM * iters * warps * 256;
  if (MODE == 3) mac = 64.0 * iters * warps / 2 * 256;
  if (MODE == 4) mac = 32.0 * iters * warps * 1024;
  if (MODE == 5) mac = 32.0 * iters * warps * 2048;
  printf("%-38s ctas/SM=%d: %8.3f ms  DFMA %6.2f TFLOP/s  DMMA %6.2f TFLOP/s  sum %6.2f  err=%s\n", name, ctas_per_sm, ms,
         2 * dfma / (ms * 1e-3) / 1e12, 2 * mac / (ms * 1e-3) / 1e12, 2 * (dfma + mac) / (ms * 1e-3) / 1e12,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double* d;
  cudaMalloc(&d, 8);
  const int sms = p.multiProcessorCount;
  for (int c = 1; c <= 2; ++c) {
    run<0, 0>("DFMA only", sms, d, c);
    run<1, 0>("DMMA m8n8k4 only", sms, d, c);
    run<4, 0>("DMMA m16n8k8 only", sms, d, c);
    run<5, 0>("DMMA m16n8k16 only", sms, d, c);
    run<2, 1>("8 DFMA + 1 m8n8k4 per warp round", sms, d, c);
    run<2, 2>("8 DFMA + 2 m8n8k4 per warp round", sms, d, c);
    run<2, 4>("8 DFMA + 4 m8n8k4 per warp round", sms, d, c);
    run<2, 8>("8 DFMA + 8 m8n8k4 per warp round", sms, d, c);
    run<3, 0>("even warps DFMA, odd warps m8n8k4", sms, d, c);
  }
  return 0;
}
