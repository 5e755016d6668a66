// TMEM as a per-lane scratchpad: correctness (same warp and cross-warp within an SM sub-partition)
// and tcgen05.ld / tcgen05.st throughput with the 32x32b shape.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem tmem.cu ; ./tmem
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ST16(taddr, v)                                                                         \
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),   \
                  "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory")
#define LD16(taddr, v)                                                                         \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),        \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])   \
               : "r"(taddr) : "memory")
#define WAIT_LD() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")
#define WAIT_ST() asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory")

__device__ uint32_t tmem_alloc_512(uint32_t* slot) {
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return *slot;
}
__device__ void tmem_free_512(uint32_t base) {
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(base) : "memory");
}

// mode 0: each warp writes its own columns then reads them back (check).
// mode 1: warp w+4 reads what warp w wrote (same lane quarter): cross-warp visibility.
__global__ void k_check(int* errors, int mode) {
  __shared__ uint32_t slot;
  const uint32_t base = tmem_alloc_512(&slot);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = blockDim.x >> 5;
  const uint32_t quarter = (uint32_t)(warp & 3);
  const int per = 512 / ((nw + 3) / 4);  // columns per warp of a quarter
  const int slotw = warp >> 2;
  uint32_t v[16], r[16];
  int bad = 0;
  if (mode == 0 || warp < 4) {
    for (int c = 0; c < per; c += 16) {
      for (int i = 0; i < 16; ++i) v[i] = 0xA0000000u ^ (warp << 20) ^ (lane << 12) ^ (c + i);
      const uint32_t ta = base + ((quarter * 32u) << 16) + (uint32_t)(slotw * per + c);
      ST16(ta, v);
    }
    WAIT_ST();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int src = (mode == 0) ? warp : (warp & 3);
  const int src_slot = (mode == 0) ? slotw : 0;
  for (int c = 0; c < per; c += 16) {
    const uint32_t ta = base + ((quarter * 32u) << 16) + (uint32_t)(src_slot * per + c);
    LD16(ta, r);
    WAIT_LD();
    for (int i = 0; i < 16; ++i)
      bad += r[i] != (0xA0000000u ^ (src << 20) ^ (lane << 12) ^ (c + i));
  }
  if (bad) atomicAdd(errors, bad);
  tmem_free_512(base);
}

// throughput: each warp streams `iters` x (LD16 or ST16) over its columns.
__global__ void k_bw(int iters, int do_store, uint32_t* sink) {
  __shared__ uint32_t slot;
  const uint32_t base = tmem_alloc_512(&slot);
  const int warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  const uint32_t quarter = (uint32_t)(warp & 3);
  const int per = 512 / ((nw + 3) / 4);
  const uint32_t ta0 = base + ((quarter * 32u) << 16) + (uint32_t)((warp >> 2) * per);
  uint32_t v[16], acc = 0;
  for (int i = 0; i < 16; ++i) v[i] = threadIdx.x + i;
  for (int c = 0; c < per; c += 16) ST16(ta0 + c, v);
  WAIT_ST();
  for (int it = 0; it < iters; ++it) {
    if (do_store) {
#pragma unroll 4
      for (int c = 0; c < per; c += 16) ST16(ta0 + c, v);
      WAIT_ST();
    } else {
      for (int c = 0; c < per; c += 64) {
        uint32_t r0[16], r1[16], r2[16], r3[16];
        LD16(ta0 + c, r0);
        LD16(ta0 + c + 16, r1);
        LD16(ta0 + c + 32, r2);
        LD16(ta0 + c + 48, r3);
        WAIT_LD();
        acc ^= r0[0] ^ r1[5] ^ r2[9] ^ r3[15] ^ r0[15] ^ r1[0] ^ r2[1] ^ r3[2];
      }
    }
  }
  if (acc == 0x12345u) sink[0] = acc;
  tmem_free_512(base);
}

int main() {
  int* err;
  uint32_t* sink;
  cudaMalloc(&err, 4);
  cudaMalloc(&sink, 4);
  for (int mode = 0; mode < 2; ++mode)
    for (int threads : {128, 256, 512}) {
      cudaMemset(err, 0, 4);
      k_check<<<148, threads>>>(err, mode);
      int h = -1;
      cudaMemcpy(&h, err, 4, cudaMemcpyDeviceToHost);
      printf("check mode %d threads %d: errors %d (%s)\n", mode, threads, h, cudaGetErrorString(cudaGetLastError()));
    }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 2000;
  for (int st = 0; st < 2; ++st)
    for (int threads : {128, 256, 512}) {
      k_bw<<<148, threads>>>(iters, st, sink);
      cudaEventRecord(e0);
      k_bw<<<148, threads>>>(iters, st, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const int nw = threads / 32;
      const int per = 512 / ((nw + 3) / 4);
      const double bytes_per_sm = (double)iters * nw * per * 32 * 4;
      printf("%s warps/SM %2d: %.1f B/clk/SM (at 1.965 GHz), %.3f ms, %s\n", st ? "st" : "ld", nw,
             bytes_per_sm / (ms * 1e-3) / 1.965e9, ms, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
