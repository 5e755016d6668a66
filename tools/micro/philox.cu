// Cost of Philox4x32-10 on one SM sub-partition: 64-bit product by IMAD.WIDE vs IMAD.HI + IMAD,
// alone and next to independent DFMA chains (the normal transform / Levy work it has to share the
// SM with).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o philox philox.cu ; ./philox
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define M0 0xD2511F53u
#define M1 0xCD9E8D57u
#define W0 0x9E3779B9u
#define W1 0xBB67AE85u

template <int NB, int MODE, int NF>
__global__ void k_philox(int iters, uint32_t k0, uint32_t k1, double a, double b, uint32_t* sink,
                         uint32_t rm0, uint32_t rm1) {
  uint32_t acc = 0;
  double f[NF > 0 ? NF : 1];
#pragma unroll
  for (int i = 0; i < NF; ++i) f[i] = a + i + threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    uint32_t c0[NB], c1[NB], c2[NB], c3[NB];
#pragma unroll
    for (int q = 0; q < NB; ++q) { c0[q] = threadIdx.x + blockIdx.x * 1024; c1[q] = it; c2[q] = q; c3[q] = acc; }
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        uint32_t h0, l0, h1, l1;
        if (MODE == 0) {
          const uint64_t p0 = (uint64_t)M0 * c0[q], p1 = (uint64_t)M1 * c2[q];
          h0 = (uint32_t)(p0 >> 32); l0 = (uint32_t)p0; h1 = (uint32_t)(p1 >> 32); l1 = (uint32_t)p1;
        } else if (MODE == 2) {  // multipliers in registers (not immediates)
          const uint64_t p0 = (uint64_t)rm0 * c0[q], p1 = (uint64_t)rm1 * c2[q];
          h0 = (uint32_t)(p0 >> 32); l0 = (uint32_t)p0; h1 = (uint32_t)(p1 >> 32); l1 = (uint32_t)p1;
        } else {  // separate high / low products (inline PTX so that they are not re-fused)
          asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(h0) : "r"(c0[q]), "r"(M0));
          asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(l0) : "r"(c0[q]), "r"(M0));
          asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(h1) : "r"(c2[q]), "r"(M1));
          asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(l1) : "r"(c2[q]), "r"(M1));
        }
        const uint32_t n0 = h1 ^ c1[q] ^ (k0 + r * W0), n2 = h0 ^ c3[q] ^ (k1 + r * W1);
        c1[q] = l1; c3[q] = l0; c0[q] = n0; c2[q] = n2;
      }
      if (NF > 0) {  // 2 DFMAs per chain per round: 20 NF DFMAs per iteration
#pragma unroll
        for (int i = 0; i < NF; ++i) { f[i] = fma(f[i], a, b); f[i] = fma(f[i], a, b); }
      }
    }
#pragma unroll
    for (int q = 0; q < NB; ++q) acc ^= c0[q] ^ c1[q] ^ c2[q] ^ c3[q];
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NF; ++i) s += f[i];
  if (acc == 0x12345u || s == 1.2345) sink[0] = acc;
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
  uint32_t* sink;
  cudaMalloc(&sink, 8);
  const int iters = 4000, sms = 148;
  printf("mode(0=IMAD.WIDE imm,2=IMAD.WIDE reg),blocks_in_flight,dfma_chains,warps_per_smsp,cycles_per_philox_block_per_smsp,dfma_cycles_alone\n");
#define RUN(NB, MODE, NF, W)                                                                       \
  {                                                                                                 \
    float ms = timeit([&] { k_philox<NB, MODE, NF><<<sms, W * 128>>>(iters, 123u, 456u, 1.0000001, 1e-9, sink, M0, M1); }); \
    printf("%d,%d,%d,%d,%.1f,%.1f\n", MODE, NB, NF, W, ms * 1e-3 * 1.965e9 / iters / (NB * W), 20.0 * NF * 2 / NB); \
  }
#define SET(MODE)                                                                                   \
  RUN(5, MODE, 0, 1) RUN(5, MODE, 0, 2) RUN(5, MODE, 0, 4) RUN(10, MODE, 0, 1) RUN(10, MODE, 0, 2)  \
  RUN(5, MODE, 4, 1) RUN(5, MODE, 4, 2) RUN(5, MODE, 4, 4) RUN(5, MODE, 8, 1) RUN(5, MODE, 8, 2)
  SET(0) SET(2)
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
