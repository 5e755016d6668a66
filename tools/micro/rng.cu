// Cost of one 128-bit random block per SM sub-partition for candidate generators, alone and next to
// independent DFMA chains (the FP64 work the headline kernel has to share the SM with).
//   0 Philox4x32-10, IMAD.WIDE          1 Philox4x32-10, mul.hi + mul.lo     2 Philox4x32-7, IMAD.WIDE
//   3 MX2 keyed expansion (12 IMAD + 24 ALU)   4 Chaskey-4 permutation (56 ALU)   5 Threefry4x32-8 (ARX)
//   6 MX2 with the adds of the cross-coupling on the ALU pipe (IADD3 instead of IMAD's addend)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rng rng.cu ; ./rng
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define M0 0xD2511F53u
#define M1 0xCD9E8D57u
#define W0 0x9E3779B9u
#define W1 0xBB67AE85u

__device__ __forceinline__ uint32_t rotl(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

template <int MODE>
__device__ __forceinline__ void block(uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3, uint32_t s,
                                      uint32_t ka, uint32_t kb, uint32_t (&o)[4]) {
  if (MODE == 0 || MODE == 1 || MODE == 2) {
    uint32_t c0 = k0, c1 = k1, c2 = k2, c3 = s;
    constexpr int R = MODE == 2 ? 7 : 10;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      uint32_t h0, l0, h1, l1;
      if (MODE != 1) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        h0 = (uint32_t)(p0 >> 32); l0 = (uint32_t)p0; h1 = (uint32_t)(p1 >> 32); l1 = (uint32_t)p1;
      } else {
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(h0) : "r"(c0), "r"(M0));
        asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(l0) : "r"(c0), "r"(M0));
        asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(h1) : "r"(c2), "r"(M1));
        asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(l1) : "r"(c2), "r"(M1));
      }
      const uint32_t n0 = h1 ^ c1 ^ (ka + r * W0), n2 = h0 ^ c3 ^ (kb + r * W1);
      c1 = l1; c3 = l0; c0 = n0; c2 = n2;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
  } else if (MODE == 3 || MODE == 6) {
    uint32_t x0 = k0 + s * 0x9E3779B1u, x1 = k1 + s * 0x85EBCA77u, x2 = k2 + s * 0xC2B2AE3Du,
             x3 = k3 + s * 0x27D4EB2Fu;
    x0 ^= x0 >> 15; x1 ^= x1 >> 15; x2 ^= x2 >> 15; x3 ^= x3 >> 15;
    uint32_t y0, y1, y2, y3;
    if (MODE == 3) {
      y0 = x0 * 0xED5AD4BBu + x1; y1 = x1 * 0xAC4C1B51u + x2; y2 = x2 * 0x31848BABu + x3; y3 = x3 * 0x7FEB352Du + x0;
    } else {
      y0 = (x0 * 0xED5AD4BBu) ^ x1; y1 = (x1 * 0xAC4C1B51u) ^ x2; y2 = (x2 * 0x31848BABu) ^ x3; y3 = (x3 * 0x7FEB352Du) ^ x0;
    }
    y0 ^= y0 >> 13; y1 ^= y1 >> 13; y2 ^= y2 >> 13; y3 ^= y3 >> 13;
    if (MODE == 3) {
      x0 = y0 * 0xAC4C1B51u + y1; x1 = y1 * 0x31848BABu + y2; x2 = y2 * 0x7FEB352Du + y3; x3 = y3 * 0xED5AD4BBu + y0;
    } else {
      x0 = (y0 * 0xAC4C1B51u) ^ y1; x1 = (y1 * 0x31848BABu) ^ y2; x2 = (y2 * 0x7FEB352Du) ^ y3; x3 = (y3 * 0xED5AD4BBu) ^ y0;
    }
    o[0] = x0 ^ (x0 >> 16); o[1] = x1 ^ (x1 >> 16); o[2] = x2 ^ (x2 >> 16); o[3] = x3 ^ (x3 >> 16);
  } else if (MODE == 4) {
    uint32_t v0 = k0 ^ s, v1 = k1, v2 = k2, v3 = k3;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      v0 += v1; v1 = rotl(v1, 5); v1 ^= v0; v0 = rotl(v0, 16);
      v2 += v3; v3 = rotl(v3, 8); v3 ^= v2;
      v0 += v3; v3 = rotl(v3, 13); v3 ^= v0;
      v2 += v1; v1 = rotl(v1, 7); v1 ^= v2; v2 = rotl(v2, 16);
    }
    o[0] = v0; o[1] = v1; o[2] = v2; o[3] = v3;
  } else {
    // Threefry4x32, 8 rounds, key (k0..k3), counter (s, 0, 0, 0)
    const uint32_t ks4 = 0x1BD11BDAu ^ k0 ^ k1 ^ k2 ^ k3;
    uint32_t X0 = s + k0, X1 = k1, X2 = k2, X3 = k3;
    const int R0[8] = {10, 11, 13, 23, 6, 17, 25, 18}, R1[8] = {26, 21, 27, 5, 20, 11, 10, 20};
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r & 1) {
        X0 += X3; X3 = rotl(X3, R0[r]); X3 ^= X0;
        X2 += X1; X1 = rotl(X1, R1[r]); X1 ^= X2;
      } else {
        X0 += X1; X1 = rotl(X1, R0[r]); X1 ^= X0;
        X2 += X3; X3 = rotl(X3, R1[r]); X3 ^= X2;
      }
      if (r == 3) { X0 += k1; X1 += k2; X2 += k3; X3 += ks4 + 1; }
      if (r == 7) { X0 += k2; X1 += k3; X2 += ks4; X3 += k0 + 2; }
    }
    o[0] = X0; o[1] = X1; o[2] = X2; o[3] = X3;
  }
}

// NB blocks in lock-step (like sdb_philox_batch), then DF DFMAs per block on NF independent chains
template <int MODE, int NB, int DF>
__global__ void k_rng(int iters, uint32_t ka, uint32_t kb, double a, double b, uint32_t* sink) {
  uint32_t acc = threadIdx.x;
  constexpr int NF = 8;
  double f[NF];
#pragma unroll
  for (int i = 0; i < NF; ++i) f[i] = a + i + threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    uint32_t o[NB][4];
    const uint32_t k0 = threadIdx.x + blockIdx.x * 1024 + acc, k1 = it, k2 = acc * 3, k3 = ~acc;
#pragma unroll
    for (int q = 0; q < NB; ++q) block<MODE>(k0, k1, k2, k3, (uint32_t)(it * NB + q), ka, kb, o[q]);
#pragma unroll
    for (int j = 0; j < DF * NB / NF; ++j)
#pragma unroll
      for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
#pragma unroll
    for (int q = 0; q < NB; ++q) acc ^= o[q][0] ^ o[q][1] ^ o[q][2] ^ o[q][3];
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
  printf("mode,blocks_in_flight,dfma_per_block,warps_per_smsp,cycles_per_block_per_smsp,dfma_cycles_alone\n");
#define RUN(MODE, NB, DF, W)                                                                          \
  {                                                                                                    \
    float ms = timeit([&] { k_rng<MODE, NB, DF><<<sms, W * 128>>>(iters, 123u, 456u, 1.0000001, 1e-9, sink); }); \
    printf("%d,%d,%d,%d,%.1f,%.1f\n", MODE, NB, DF, W, ms * 1e-3 * 1.965e9 / iters / (NB * W), 2.0 * DF); \
  }
#define SET(MODE)                                                                                      \
  RUN(MODE, 5, 0, 1) RUN(MODE, 5, 0, 2) RUN(MODE, 5, 0, 4) RUN(MODE, 5, 32, 2) RUN(MODE, 5, 64, 2)    \
  RUN(MODE, 5, 112, 2) RUN(MODE, 5, 112, 1)
  RUN(0, 5, 32, 0 + 2) RUN(0, 5, 64, 2)
  SET(0) SET(1) SET(2) SET(3) SET(6) SET(4) SET(5)
  // DFMA alone for reference: MODE 3 with no blocks is not expressible; DF only via NB=... use 8 chains
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
