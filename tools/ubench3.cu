// ubench3.cu -- device-wide barrier variants on a resident grid of 128 CTAs (32 clusters x 4), 256 threads each: what does one
// dependent phase boundary of the cluster step kernel cost, and which part of the protocol is it?
// Each repetition = 512 B of global stores per CTA (the epilogue's output), the barrier, and a 16 KB bulk copy issued by the
// polling thread (the next phase's activation slice).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench3_bin tools/ubench3.cu && tools/ubench3_bin
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint4 ld_relaxed4(const uint4* p) { uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_relaxed_add(unsigned* p, unsigned v) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory"); } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

struct Res { long long total, bar, copy; int fail; };

// mode: 0 red.release + relaxed poll + acq_rel fence (step kernels today)   1: same without the acquire fence
//       2: four counters (one per cluster rank), poll all four with one 16 B load, no acquire fence
//       3: cluster barrier, then ONE arrival per cluster (rank 0), everyone polls, no acquire fence
//       4: like 1 but the arrival is fence.acq_rel.gpu + relaxed red (same semantics, different instruction pair)
//       5: like 1 without the 512 B of data stores (what the release waits for)      6: like 1 without fence.proxy.async.global
//       7 / 8: like 1 with 2 / 4 polling loads kept in flight, staggered (a fresh sample of the counter every RT/2, RT/4)
__global__ void __launch_bounds__(256, 1) k_bar(unsigned* ctr, uint32_t* data, const unsigned char* img, Res* res, int mode, int reps, int do_copy) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* mb = reinterpret_cast<uint64_t*>(smem);
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_rank();
  const unsigned G = gridDim.x;
  if (tid == 0) { mbar_init(mb, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  long long t_bar = 0, t_copy = 0;
  uint32_t parity = 0;
  int fail = 0;
  const long long t0 = clock64();
  for (int it = 1; it <= reps; it++) {
    if (tid < 128 && mode != 5) data[(size_t)blockIdx.x * 128 + tid] = (uint32_t)it;   // 512 B of output
    const long long b0 = clock64();
    if (mode != 6) asm volatile("fence.proxy.async.global;" ::: "memory");
    if (mode == 3) { cluster_arrive(); cluster_wait(); } else __syncthreads();
    if (tid == 0) {
      unsigned spins = 0;
      if (mode == 7 || mode == 8) {
        red_release_add(ctr, 1u);
        const unsigned want = (unsigned)it * G;
        const int N = (mode == 7) ? 2 : 4;
        const long long gap = 1400 / N;   // ~ round trip / N
        unsigned v0, v1 = 0, v2 = 0, v3 = 0;
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v0) : "l"(ctr) : "memory");
        long long c0 = clock64();
        while (clock64() - c0 < gap) {}
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v1) : "l"(ctr) : "memory");
        if (N == 4) {
          c0 = clock64(); while (clock64() - c0 < gap) {}
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v2) : "l"(ctr) : "memory");
          c0 = clock64(); while (clock64() - c0 < gap) {}
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v3) : "l"(ctr) : "memory");
        }
        while (true) {
          if (v0 >= want) break;
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v0) : "l"(ctr) : "memory");
          if (v1 >= want) break;
          asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v1) : "l"(ctr) : "memory");
          if (N == 4) {
            if (v2 >= want) break;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v2) : "l"(ctr) : "memory");
            if (v3 >= want) break;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v3) : "l"(ctr) : "memory");
          }
          if (++spins > (1u << 22)) { fail = 1; break; }
        }
      } else if (mode == 0 || mode == 1 || mode == 5 || mode == 6) {
        red_release_add(ctr, 1u);
        while (ld_relaxed(ctr) < (unsigned)it * G) if (++spins > (1u << 22)) { fail = 1; break; }
      } else if (mode == 2) {
        red_release_add(ctr + rank, 1u);
        const unsigned want = (unsigned)it * (G / 4);
        while (true) {
          const uint4 v = ld_relaxed4(reinterpret_cast<const uint4*>(ctr));
          if (v.x >= want && v.y >= want && v.z >= want && v.w >= want) break;
          if (++spins > (1u << 22)) { fail = 1; break; }
        }
      } else if (mode == 3) {
        if (rank == 0) red_release_add(ctr, 1u);
        while (ld_relaxed(ctr) < (unsigned)it * (G / 4)) if (++spins > (1u << 22)) { fail = 1; break; }
      } else {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        red_relaxed_add(ctr, 1u);
        while (ld_relaxed(ctr) < (unsigned)it * G) if (++spins > (1u << 22)) { fail = 1; break; }
      }
      if (mode == 0) asm volatile("fence.acq_rel.gpu;" ::: "memory");
      if (do_copy) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(mb, 16384);
        bulk_g2s(smem + 128, img + (size_t)(blockIdx.x & 3) * 16384, 16384, mb);
      }
    }
    __syncthreads();
    const long long b1 = clock64();
    if (do_copy) { mbar_wait(mb, parity); parity ^= 1; }
    const long long b2 = clock64();
    t_bar += b1 - b0; t_copy += b2 - b1;
  }
  const long long t1 = clock64();
  if (tid == 0) {
    if (fail) atomicExch(&res->fail, 1);
    if (blockIdx.x == 0) { res->total = t1 - t0; res->bar = t_bar; res->copy = t_copy; }
  }
}

int main() {
  unsigned* ctr; uint32_t* data; unsigned char* img; Res* res;
  CK(cudaMalloc(&ctr, 256)); CK(cudaMalloc(&data, 148 * 512)); CK(cudaMalloc(&img, 65536)); CK(cudaMalloc(&res, sizeof(Res)));
  CK(cudaMemset(img, 1, 65536));
  const int reps = 500;
  const size_t smem = 128 + 16384;
  CK(cudaFuncSetAttribute(k_bar, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const char* names[9] = {"red.release + poll + acq_rel fence", "red.release + poll (no acquire fence)", "4 counters (per cluster rank), one 16 B poll",
                          "cluster barrier + 1 arrival per cluster", "fence.acq_rel + relaxed red + poll", "red.release + poll, NO data stores", "red.release + poll, no proxy fence",
                          "red.release + 2 staggered polls", "red.release + 4 staggered polls"};
  for (int copy = 0; copy < 2; copy++)
    for (int mode = 0; mode < 9; mode++) {
      CK(cudaMemset(ctr, 0, 256)); CK(cudaMemset(res, 0, sizeof(Res)));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(128); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
      cfg.attrs = at; cfg.numAttrs = 2;
      cudaError_t e = cudaLaunchKernelEx(&cfg, k_bar, ctr, data, (const unsigned char*)img, res, mode, reps, copy);
      if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); cudaGetLastError(); continue; }
      CK(cudaDeviceSynchronize());
      Res r; CK(cudaMemcpy(&r, res, sizeof(Res), cudaMemcpyDeviceToHost));
      printf("%-46s %s | %6.0f cyc/rep (%.2f us)  barrier %6.0f  copy-wait %5.0f %s\n", names[mode], copy ? "+16 KB slice" : "            ",
             (double)r.total / reps, (double)r.total / reps / 1965.0, (double)r.bar / reps, (double)r.copy / reps, r.fail ? " ** TIMEOUT **" : "");
    }
  return 0;
}
