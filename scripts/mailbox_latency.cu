// Micro-benchmark of inter-CTA signalling through L2 on sm_100a (informs nltgv2_grid.cuh).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mailbox_latency scripts/mailbox_latency.cu
// Modes (all: 16-byte tagged mailboxes, payload and tag in one access):
//   ping-pong between CTA 0 and CTA `peer` (one thread each), cycles per ROUND TRIP:
//     0 st/ld.relaxed.gpu.b128   1 st.volatile/ld.volatile v4   2 st.release.gpu / ld.acquire.gpu b128
//     3 st.relaxed.gpu + ld.relaxed.gpu with L1 no-allocate hint (ld.global.cv v4)   4 atom.exch.b128 writer
//   lockstep neighbour exchange, N CTAs, cycles per ITERATION:
//     10 each CTA: 32 threads publish, 32 threads poll 4 neighbours (8 entries each), 2 __syncthreads
//     11 same but only ONE entry per neighbour pair (4 polls per CTA)
//     12 same as 10 with 256 threads publishing/polling (8 neighbours x 32 entries)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint4 ld_rlx(const uint4* p) {
  unsigned long long lo, hi;
  asm volatile("{\n\t.reg .b128 t;\n\tld.relaxed.gpu.global.b128 t, [%2];\n\tmov.b128 {%0, %1}, t;\n\t}" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
  return make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
}
__device__ __forceinline__ uint4 ld_acq(const uint4* p) {
  unsigned long long lo, hi;
  asm volatile("{\n\t.reg .b128 t;\n\tld.acquire.gpu.global.b128 t, [%2];\n\tmov.b128 {%0, %1}, t;\n\t}" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
  return make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
}
__device__ __forceinline__ uint4 ld_vol(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_cv(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cv.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_rlx(uint4* p, uint4 v) {
  unsigned long long lo = (unsigned long long)v.x | ((unsigned long long)v.y << 32), hi = (unsigned long long)v.z | ((unsigned long long)v.w << 32);
  asm volatile("{\n\t.reg .b128 t;\n\tmov.b128 t, {%1, %2};\n\tst.relaxed.gpu.global.b128 [%0], t;\n\t}" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ void st_rel(uint4* p, uint4 v) {
  unsigned long long lo = (unsigned long long)v.x | ((unsigned long long)v.y << 32), hi = (unsigned long long)v.z | ((unsigned long long)v.w << 32);
  asm volatile("{\n\t.reg .b128 t;\n\tmov.b128 t, {%1, %2};\n\tst.release.gpu.global.b128 [%0], t;\n\t}" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ void st_vol(uint4* p, uint4 v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_xchg(uint4* p, uint4 v) {
  unsigned long long lo = (unsigned long long)v.x | ((unsigned long long)v.y << 32), hi = (unsigned long long)v.z | ((unsigned long long)v.w << 32);
  asm volatile("{\n\t.reg .b128 t, o;\n\tmov.b128 t, {%1, %2};\n\tatom.relaxed.gpu.global.exch.b128 o, [%0], t;\n\t}" ::"l"(p), "l"(lo), "l"(hi) : "memory");
}
template <int M> __device__ __forceinline__ uint4 LD(const uint4* p) {
  if (M == 1) return ld_vol(p);
  if (M == 2) return ld_acq(p);
  if (M == 3) return ld_cv(p);
  return ld_rlx(p);
}
template <int M> __device__ __forceinline__ void ST(uint4* p, uint4 v) {
  if (M == 1) st_vol(p, v);
  else if (M == 2) st_rel(p, v);
  else if (M == 4) st_xchg(p, v);
  else st_rlx(p, v);
}

template <int M>
__global__ void pingpong(uint4* box, int iters, int peer, long long* out, unsigned* smid) {
  if (threadIdx.x != 0) return;
  const int b = blockIdx.x;
  unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
  if (b == 0) smid[0] = sm;
  if (b == peer) smid[1] = sm;
  if (b != 0 && b != peer) return;
  uint4* mine = box + (b == 0 ? 0 : 64);    // separate 128 B lines (and sectors)
  uint4* theirs = box + (b == 0 ? 64 : 0);
  long long t0 = clock64();
  for (uint32_t i = 1; i <= (uint32_t)iters; ++i) {
    if (b == 0) {
      ST<M>(theirs, make_uint4(i, i, i, i));
      uint4 v; unsigned sp = 0; do { v = LD<M>(mine); } while (v.w < i && ++sp < 4000000u);
    } else {
      uint4 v; unsigned sp = 0; do { v = LD<M>(mine); } while (v.w < i && ++sp < 4000000u);
      ST<M>(theirs, make_uint4(i, i, i, i));
    }
  }
  long long t1 = clock64();
  if (b == 0) out[0] = (t1 - t0) / iters;
}

// lockstep neighbour exchange: CTA c has neighbours c +- 1, c +- stride (mod N)
__global__ void __launch_bounds__(256, 1) lockstep(uint4* box, int iters, int N, int per_pair, int nthr, int stride, long long* out) {
  __shared__ uint4 halo[256];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int offs[8] = {1, N - 1, stride, N - stride, stride + 1, N - stride - 1, stride - 1 + N, N - stride + 1};
  const int nn = nthr / per_pair;  // neighbours polled
  long long t0 = clock64();
  for (uint32_t i = 1; i <= (uint32_t)iters; ++i) {
    uint4* bank = box + (size_t)(i & 1) * N * 256;
    if (tid < nthr) ST<0>(bank + (size_t)c * 256 + tid, make_uint4(i, tid, c, i));
    if (tid < nthr) {
      const int nb = (c + offs[(tid / per_pair) % 8]) % N;
      const uint4* p = bank + (size_t)nb * 256 + (tid % per_pair) + (tid / per_pair) * per_pair % nthr;
      uint4 v; unsigned sp = 0; do { v = ld_rlx(p); } while (v.w < i && ++sp < 4000000u);
      halo[tid] = v;
    }
    (void)nn;
    __syncthreads();
    if (tid == 0 && halo[0].x == 0xffffffffu) out[1] = 1;  // keep the reads alive
    __syncthreads();
  }
  long long t1 = clock64();
  if (c == 0 && tid == 0) out[0] = (t1 - t0) / iters;
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  uint4* box; long long* out; unsigned* smid;
  const int N = 148;
  cudaMalloc(&box, sizeof(uint4) * 2 * 512 * 256);
  cudaMallocManaged(&out, 64);
  cudaMallocManaged(&smid, 8);
  const int iters = 1000;
  const char* names[5] = {"relaxed.gpu.b128", "volatile.v4", "release/acquire.gpu.b128", "st.relaxed + ld.cv", "atom.exch.b128 + ld.relaxed"};
  for (int peer : {1, 2, 37, 74, 147}) {
    for (int m = 0; m < 5; ++m) {
      cudaMemset(box, 0, sizeof(uint4) * 2 * 512 * 256);
      out[0] = 0;
      void* args[] = {&box, (void*)&iters, &peer, &out, &smid};
      void* fn = m == 0 ? (void*)pingpong<0> : m == 1 ? (void*)pingpong<1> : m == 2 ? (void*)pingpong<2> : m == 3 ? (void*)pingpong<3> : (void*)pingpong<4>;
      cudaLaunchCooperativeKernel(fn, dim3(N), dim3(32), args, 0, 0);
      cudaError_t e = cudaDeviceSynchronize();
      printf("pingpong peer=%3d (sm %u <-> sm %u) %-28s round trip %lld cycles %s\n", peer, smid[0], smid[1], names[m], out[0], e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  for (int n : {2, 16, 148}) {
    struct { int per_pair, nthr; } cfgs[] = {{8, 32}, {1, 4}, {32, 256}, {1, 2}};
    for (auto cf : cfgs) {
      cudaMemset(box, 0, sizeof(uint4) * 2 * 512 * 256);
      int stride = n >= 16 ? 12 : 1;
      if (n == 2) stride = 1;
      void* args[] = {&box, (void*)&iters, &n, &cf.per_pair, &cf.nthr, &stride, &out};
      out[0] = 0;
      cudaLaunchCooperativeKernel((void*)lockstep, dim3(n), dim3(256), args, 0, 0);
      cudaError_t e = cudaDeviceSynchronize();
      printf("lockstep N=%3d entries/pair=%2d threads=%3d  %lld cycles/iteration %s\n", n, cf.per_pair, cf.nthr, out[0], e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("SM clock %d kHz\n", clk);
  return 0;
}
