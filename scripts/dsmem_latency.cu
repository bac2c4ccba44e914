// Micro-benchmark of cluster-level signalling primitives on sm_100a (informs nltgv2_cluster.cuh).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dsmem_latency scripts/dsmem_latency.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) {
  uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o;
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void csync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
}

// mode 0: cluster barrier (release/acquire); 1: relaxed cluster barrier; 2: st.async ping-pong (thread 0 only waits);
// 3: st.async ping-pong, all threads wait + __syncthreads each hop; 4: st.release flag + ld.acquire polling ping-pong
// 5: __syncthreads only
__global__ void __launch_bounds__(1024, 1) k(int mode, int iters, long long* out) {
  __shared__ __align__(16) float4 buf[4];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ volatile int flag;
  const uint32_t r = ctarank();
  const int tid = threadIdx.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    flag = 0;
  }
  __syncthreads();
  csync();
  const uint32_t peer = r ^ 1u;
  const uint32_t my_mbar = smem_u32(&mbar), rm = mapa(my_mbar, peer), rbuf = mapa(smem_u32(&buf[0]), peer);
  const uint32_t rflag = mapa(smem_u32((const void*)&flag), peer);
  long long t0 = clock64();
  if (mode == 0) { for (int i = 0; i < iters; ++i) csync(); }
  else if (mode == 1) { for (int i = 0; i < iters; ++i) csync_relaxed(); }
  else if (mode == 5) { for (int i = 0; i < iters; ++i) __syncthreads(); }
  else if (mode == 2 || mode == 3) {
    // rank even sends first; each hop: arm own barrier, wait, reply
    for (int i = 0; i < iters; ++i) {
      if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 16;" ::"r"(my_mbar) : "memory");
      if ((r & 1u) == 0) {
        if (tid == 0) asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%1,%1,%1}, [%2];" ::"r"(rbuf), "f"((float)i), "r"(rm) : "memory");
        if (mode == 3 || tid == 0) mbar_wait(my_mbar, i & 1);
        if (mode == 3) __syncthreads();
      } else {
        if (mode == 3 || tid == 0) mbar_wait(my_mbar, i & 1);
        if (mode == 3) __syncthreads();
        if (tid == 0) asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%1,%1,%1}, [%2];" ::"r"(rbuf), "f"((float)i), "r"(rm) : "memory");
      }
    }
  } else if (mode == 4) {
    if (tid == 0) {
      for (int i = 1; i <= iters; ++i) {
        if ((r & 1u) == 0) {
          asm volatile("st.release.cluster.shared::cluster.u32 [%0], %1;" ::"r"(rflag), "r"(i) : "memory");
          int v; do { asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void*)&flag)) : "memory"); } while (v < i);
        } else {
          int v; do { asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void*)&flag)) : "memory"); } while (v < i);
          asm volatile("st.release.cluster.shared::cluster.u32 [%0], %1;" ::"r"(rflag), "r"(i) : "memory");
        }
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  csync();
  if (tid == 0 && blockIdx.x == 0) out[mode] = (t1 - t0) / iters;
}

int main() {
  long long* d; cudaMalloc(&d, 64 * sizeof(long long)); cudaMemset(d, 0, 64 * sizeof(long long));
  const char* names[] = {"cluster barrier release/acquire", "cluster barrier relaxed", "st.async ping-pong (1 waiter) per round trip",
                         "st.async ping-pong (1024 waiters + bar.sync) per round trip", "st.release flag + ld.acquire poll per round trip", "__syncthreads"};
  for (int csize = 2; csize <= 8; csize *= 4) {
    for (int mode = 0; mode < 6; ++mode) {
      cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(csize); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 0;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaError_t e = cudaLaunchKernelEx(&cfg, k, mode, 2000, d);
      e = cudaDeviceSynchronize();
      long long v = 0; cudaMemcpy(&v, d + mode, sizeof(v), cudaMemcpyDeviceToHost);
      printf("cluster %d  %-62s %6lld cycles  (%s)\n", csize, names[mode], v, cudaGetErrorString(e));
    }
  }
  return 0;
}
