// Microbenchmarks that size the fused kernel: DFMA latency/ILP need, LDS broadcast feed, LDC feed.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; ./ubench
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void k_dfma(int iters, double a, double b, double* sink) {
  double acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = a + c + threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = fma(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) s += acc[c];
  if (s == 1.2345) sink[0] = s;
}

// DFMA fed by broadcast LDS.128 (two operands per load): acc[c] = fma(acc[c], B[k], y) pattern
template <int C>
__global__ void k_dfma_lds(int iters, double a, double* sink) {
  extern __shared__ double2 sm[];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) sm[i] = make_double2(1.0 + 1e-9 * i, 1.0 - 1e-9 * i);
  __syncthreads();
  double acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = a + c + threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int c = 0; c < C; c += 2) {
        const double2 bv = sm[(u * C + c) / 2 + (i & 7) * 16];
        acc[c] = fma(acc[c], bv.x, a);
        acc[c + 1] = fma(acc[c + 1], bv.y, a);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) s += acc[c];
  if (s == 1.2345) sink[0] = s;
}

struct Params { double v[1024]; };
// DFMA fed by kernel-parameter constants (LDC / LDCU)
template <int C>
__global__ void k_dfma_ldc(int iters, double a, double* sink, const __grid_constant__ Params p) {
  double acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = a + c + threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 64; ++u)
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = fma(acc[c], p.v[u * C + c], a);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) s += acc[c];
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
  int sms = 148;
  const int iters = 4000;
  static Params p;
  for (int i = 0; i < 1024; ++i) p.v[i] = 1.0 + 1e-9 * i;
  printf("kind,chains,warps_per_sm,tflops,dfma_per_clk_per_sm(at 1.965GHz)\n");
#define RUN(KIND, C, W, CALL, PERIT)                                                         \
  {                                                                                           \
    float ms = timeit([&] { CALL; });                                                         \
    double fl = 2.0 * PERIT * C * (double)iters * (W * 32) * sms;                             \
    printf("%s,%d,%d,%.2f,%.1f\n", KIND, C, W, fl / ms / 1e9, fl / 2 / (ms * 1e-3) / 1.965e9 / sms); \
  }
#define SWEEP(C)                                                                              \
  RUN("reg", C, 4, (k_dfma<C><<<sms, 128>>>(iters, 1.0000001, 1e-9, sink)), 16)              \
  RUN("reg", C, 8, (k_dfma<C><<<sms, 256>>>(iters, 1.0000001, 1e-9, sink)), 16)              \
  RUN("reg", C, 16, (k_dfma<C><<<sms, 512>>>(iters, 1.0000001, 1e-9, sink)), 16)
  SWEEP(1) SWEEP(2) SWEEP(3) SWEEP(4) SWEEP(6) SWEEP(8) SWEEP(12) SWEEP(16)
#define SWEEPL(C)                                                                             \
  RUN("lds128", C, 4, (k_dfma_lds<C><<<sms, 128, 8192>>>(iters, 1.0000001, sink)), 8)        \
  RUN("lds128", C, 8, (k_dfma_lds<C><<<sms, 256, 8192>>>(iters, 1.0000001, sink)), 8)        \
  RUN("lds128", C, 16, (k_dfma_lds<C><<<sms, 512, 8192>>>(iters, 1.0000001, sink)), 8)
  SWEEPL(4) SWEEPL(8) SWEEPL(16)
#define SWEEPC(C)                                                                             \
  RUN("ldc", C, 4, (k_dfma_ldc<C><<<sms, 128>>>(iters / 8, 1.0000001, sink, p)), 64)         \
  RUN("ldc", C, 8, (k_dfma_ldc<C><<<sms, 256>>>(iters / 8, 1.0000001, sink, p)), 64)         \
  RUN("ldc", C, 16, (k_dfma_ldc<C><<<sms, 512>>>(iters / 8, 1.0000001, sink, p)), 64)
  SWEEPC(4) SWEEPC(8) SWEEPC(16)
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
