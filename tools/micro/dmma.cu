// Does B200 have a usable FP64 tensor pipe next to the FP64 FMA pipe?
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int C, int NF>  // C independent DMMA accumulators, NF extra DFMA per DMMA
__global__ void k_884(int iters, double a, double b, double* sink) {
  double c0[C], c1[C], f[8];
#pragma unroll
  for (int c = 0; c < C; ++c) { c0[c] = threadIdx.x; c1[c] = c; }
#pragma unroll
  for (int c = 0; c < 8; ++c) f[c] = c + a;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int c = 0; c < C; ++c) {
        dmma884(c0[c], c1[c], a, b);
#pragma unroll
        for (int q = 0; q < NF; ++q) f[(c * NF + q) & 7] = fma(f[(c * NF + q) & 7], a, b);
      }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) s += c0[c] + c1[c];
#pragma unroll
  for (int c = 0; c < 8; ++c) s += f[c];
  if (s == 1.2345) sink[0] = s;
}

template <int C>
__global__ void k_1688(int iters, double a, double b, double* sink) {
  double acc[C][4], av[4] = {a, a + 1, a + 2, a + 3}, bv[2] = {b, b + 1};
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[c][q] = threadIdx.x + q;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int c = 0; c < C; ++c) dmma1688(acc[c], av, bv);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
  if (s == 1.2345) sink[0] = s;
}

template <class F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  double* sink;
  cudaMalloc(&sink, 8);
  const int sms = 148, iters = 2000;
  printf("kind,acc,warps_per_sm,extra_dfma,mma_tflops,dfma_tflops,ms\n");
#define RUN884(C, NF, W)                                                                        \
  {                                                                                             \
    float ms = timeit([&] { k_884<C, NF><<<sms, W * 32>>>(iters, 1.0000001, 1e-9, sink); });    \
    double n = 8.0 * C * iters * W * sms;                                                       \
    printf("m8n8k4,%d,%d,%d,%.2f,%.2f,%.3f\n", C, W, NF, n * 512 / ms / 1e9,                    \
           n * NF * 64 / ms / 1e9, ms);                                                         \
  }
  RUN884(1, 0, 4) RUN884(2, 0, 4) RUN884(4, 0, 4) RUN884(8, 0, 4)
  RUN884(4, 0, 8) RUN884(8, 0, 8) RUN884(4, 0, 16) RUN884(8, 0, 16)
  RUN884(4, 2, 8) RUN884(4, 4, 8) RUN884(4, 8, 8) RUN884(4, 4, 16) RUN884(4, 8, 16)
#define RUN1688(C, W)                                                                           \
  {                                                                                             \
    float ms = timeit([&] { k_1688<C><<<sms, W * 32>>>(iters, 1.0000001, 1e-9, sink); });       \
    double n = 8.0 * C * iters * W * sms;                                                       \
    printf("m16n8k8,%d,%d,0,%.2f,0,%.3f\n", C, W, n * 2048 / ms / 1e9, ms);                     \
  }
  RUN1688(1, 4) RUN1688(2, 4) RUN1688(4, 4) RUN1688(4, 8) RUN1688(4, 16)
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
