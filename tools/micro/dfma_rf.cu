// DFMA throughput as a function of the operand pattern (register-file bandwidth / reuse cache):
//   pattern 0: acc[i] = fma(a, b, acc[i])               two operands in the reuse cache
//   pattern 1: acc[r][k] = fma(g[r], a[k], acc[r][k])   the P2 loop of sdb_fused.cuh (4 x 10 accumulators)
//   pattern 2: acc[i] = fma(x[i], y[i], acc[i])         three fresh 64-bit operands per instruction
//   pattern 3: pattern 1 with an LDS.64 per four DFMAs (the a[k] operand comes from shared memory)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_rf dfma_rf.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int PAT>
__global__ void __launch_bounds__(256, 1) k(int iters, double* sink, const double* in) {
  __shared__ double sm[64 * 32];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) sm[i] = 1e-9 * i;
  __syncthreads();
  double acc[40], x[40], y[40];
#pragma unroll
  for (int i = 0; i < 40; ++i) { acc[i] = in[i] * threadIdx.x; x[i] = in[40 + i] + threadIdx.x; y[i] = in[80 + i] - threadIdx.x; }
  const volatile double* p = sm + lane;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (PAT == 0) {
#pragma unroll
      for (int i = 0; i < 40; ++i) acc[i] = fma(x[0], y[0], acc[i]);
    } else if (PAT == 1) {
#pragma unroll
      for (int kk = 0; kk < 10; ++kk)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[4 * kk + r] = fma(x[r], y[kk], acc[4 * kk + r]);
    } else if (PAT == 2) {
#pragma unroll
      for (int i = 0; i < 40; ++i) acc[i] = fma(x[i], y[i], acc[i]);
    } else {
#pragma unroll
      for (int kk = 0; kk < 10; ++kk) {
        const double a = p[32 * ((kk + it) & 63)];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[4 * kk + r] = fma(x[r], a, acc[4 * kk + r]);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 40; ++i) s += acc[i];
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
  double *sink, *in;
  cudaMalloc(&sink, 8);
  cudaMalloc(&in, 120 * 8);
  double h[120];
  for (int i = 0; i < 120; ++i) h[i] = 1.0 + 1e-3 * i;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const int sms = 148, iters = 20000;
  printf("pattern,warps_per_smsp,cycles_per_dfma_per_smsp\n");
#define RUN(PAT, W)                                                                       \
  {                                                                                        \
    float ms = timeit([&] { k<PAT><<<sms, W * 128>>>(iters, sink, in); });                 \
    printf("%d,%d,%.3f\n", PAT, W, ms * 1e-3 * 1.965e9 / (40.0 * iters * W));              \
  }
  RUN(0, 1) RUN(0, 2) RUN(1, 1) RUN(1, 2) RUN(2, 1) RUN(2, 2) RUN(3, 1) RUN(3, 2)
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
