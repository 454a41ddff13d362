// ubench2.cu -- micro-benchmarks behind the round-2 decode-step design: how fast can a chain of
// dependent all-to-all phases run WITHOUT a device-wide barrier?
//   * cluster occupancy for the shapes we want (cluster 2 / 4 / 8, 256 threads, ~200 KB dynamic shared memory)
//   * "sentinel dataflow" phase: every CTA polls its K-slice of the previous phase's activations straight out of L2
//     (words are pre-set to a NaN pattern no producer can emit; a word is valid as soon as it differs), exchanges K-split
//     partial sums with its cluster peers through DSMEM (st.async + mbarrier complete_tx) and writes 512 B of output
//   * the same phase with the exchange or the poll removed, to separate the costs
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench2_bin tools/ubench2.cu && tools/ubench2_bin
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr uint32_t SENT = 0xFFFFFFFFu;
constexpr int ROWS = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4* p) {
  uint4 r;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t* p, uint32_t v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, float4 v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1,%2,%3,%4}, [%5];"
               ::"r"(remote_addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 20)) return false;
  } while (!ok);
  return true;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

struct Res { long long t[8]; int fail; };

// act: [4 ring buffers][ROWS][H] bf16 (as u32 words: H/2 per row).  Cluster c of C CTAs owns features [c*FC, (c+1)*FC) of the
// N = H outputs; rank r of the cluster reads K-slice r (H/C columns) and finally owns FC/C features x ROWS rows.
// mode bit 0: poll the input slice; bit 1: DSMEM exchange of [ROWS][EXF] fp32 per peer; bit 2: write outputs (+ reset)
template <int C>
__global__ void __launch_bounds__(256, 1) k_chain(uint32_t* act, int H, int EXF, int reps, int mode, Res* res) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);             // exchange mbarrier
  float* slots = reinterpret_cast<float*>(smem + 128);           // [C][ROWS][EXF] received partials
  uint4* tile = reinterpret_cast<uint4*>(smem + 128 + C * ROWS * 64 * 4);  // staged K-slice (EXF <= 64)
  const int tid = threadIdx.x;
  const uint32_t rank = cluster_rank();
  const int cluster = blockIdx.x / C;
  const int nclusters = gridDim.x / C;
  const int words_per_row = H / 2;
  const int slice_words = words_per_row / C;           // u32 words per row of this CTA's K-slice
  const int vec_per_row = slice_words / 4;
  const int nvec = ROWS * vec_per_row;                 // 16-byte vectors in the slice
  const int FC = H / nclusters;                        // features per cluster
  const int own_f = FC / C;                            // features this CTA finally owns
  const int own_words = ROWS * own_f / 2;              // output words (bf16x2)
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  cluster_arrive(); cluster_wait();
  if (mode & 2) cluster_arrive();  // pairs with the first phase's wait
  uint32_t parity = 0;
  long long t_poll = 0, t_ex = 0, t_out = 0;
  int fail = 0;
  const long long t_begin = clock64();
  for (int p = 1; p <= reps; p++) {  // (a timed-out thread keeps the structure: no barrier is ever skipped)
    const uint32_t* in = act + (size_t)((p - 1) & 3) * ROWS * words_per_row;
    uint32_t* out = act + (size_t)(p & 3) * ROWS * words_per_row;
    uint32_t* nxt = act + (size_t)((p + 1) & 3) * ROWS * words_per_row;
    long long t0 = clock64();
    uint32_t acc = 0;
    if (mode & 1) {
      for (int v = tid; v < nvec; v += 256) {
        const int r = v / vec_per_row, c = v - r * vec_per_row;
        const uint4* src = reinterpret_cast<const uint4*>(in + (size_t)r * words_per_row + rank * slice_words) + c;
        uint4 x;
        uint32_t spins = 0;
        while (true) {
          x = ld_relaxed_v4(src);
          if (fail || (x.x != SENT && x.y != SENT && x.z != SENT && x.w != SENT)) break;
          if (++spins > (1u << 20)) { fail = 1; break; }
        }
        tile[v] = x;
        acc += x.x;
      }
    }
    __syncthreads();
    long long t1 = clock64();
    if (mode & 2) {
      // slots of the previous phase have been read by every peer: (cluster barrier split around the epilogue)
      cluster_wait();
      if (tid == 0) mbar_expect_tx(bar, (uint32_t)((C - 1) * ROWS * EXF * 4));
      const int nv = ROWS * EXF / 4;  // float4 per peer
      for (int v = tid; v < nv; v += 256) {
        const float4 val = make_float4((float)p, (float)acc, 1.f, 2.f);
#pragma unroll
        for (int d = 1; d < C; d++) {
          const uint32_t peer = (rank + d) % C;
          st_async_v4(mapa(smem_u32(slots + ((size_t)rank * ROWS * EXF) + v * 4), peer), val, mapa(smem_u32(bar), peer));
        }
      }
      if (!fail && !mbar_wait_cluster(bar, parity)) fail = 2;
      parity ^= 1;
      float s = 0.f;
      for (int v = tid; v < ROWS * EXF; v += 256)
        for (int d = 0; d < C; d++) if (d != (int)rank) s += slots[(size_t)d * ROWS * EXF + v];
      acc += (uint32_t)s;
      __syncthreads();
      cluster_arrive();
    }
    long long t2 = clock64();
    if (mode & 4) {
      for (int w = tid; w < own_words; w += 256) {
        const int r = w / (own_f / 2), c = w - r * (own_f / 2);
        const size_t o = (size_t)r * words_per_row + (cluster * FC + rank * own_f) / 2 + c;
        st_relaxed_u32(nxt + o, SENT);                        // re-arm the slot three phases behind (see DESIGN)
        st_relaxed_u32(out + o, (uint32_t)p | ((acc & 1u) << 20));
      }
    }
    long long t3 = clock64();
    t_poll += t1 - t0; t_ex += t2 - t1; t_out += t3 - t2;
  }
  const long long t_end = clock64();
  if (mode & 2) cluster_wait();
  if (tid == 0) {
    if (fail) atomicExch(&res->fail, fail);
    if (blockIdx.x == 0) { res->t[0] = t_end - t_begin; res->t[1] = t_poll; res->t[2] = t_ex; res->t[3] = t_out; }
  }
}

template <int C>
static void run_chain(uint32_t* act, Res* res, int H, int EXF, int reps, int mode, int nclusters, const char* name) {
  const int grid = nclusters * C;
  const size_t smem = 128 + (size_t)C * ROWS * 64 * 4 + (size_t)ROWS * (H / C) * 2 + 1024;
  CK(cudaFuncSetAttribute(k_chain<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (C > 8) CK(cudaFuncSetAttribute(k_chain<C>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  // phase 0 data valid in ring buffer 0, everything else armed
  const size_t words = (size_t)4 * ROWS * H / 2;
  uint32_t* h = (uint32_t*)malloc(words * 4);
  for (size_t i = 0; i < words; i++) h[i] = (i < (size_t)ROWS * H / 2) ? 7u : SENT;
  CK(cudaMemcpy(act, h, words * 4, cudaMemcpyHostToDevice));
  free(h);
  CK(cudaMemset(res, 0, sizeof(Res)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int ncl = 0;
  cudaError_t oe = cudaOccupancyMaxActiveClusters(&ncl, k_chain<C>, &cfg);
  if (oe != cudaSuccess) { printf("%-46s occupancy query failed: %s\n", name, cudaGetErrorString(oe)); cudaGetLastError(); return; }
  if (ncl < nclusters) { printf("%-46s only %d clusters of %d co-resident (need %d): skipped\n", name, ncl, C, nclusters); return; }
  CK(cudaLaunchKernelEx(&cfg, k_chain<C>, act, H, EXF, reps, mode, res));
  CK(cudaDeviceSynchronize());
  Res r;
  CK(cudaMemcpy(&r, res, sizeof(Res), cudaMemcpyDeviceToHost));
  printf("%-46s max clusters %3d | %7.0f cyc/phase (%.2f us)  poll %6.0f  exchange %6.0f  out %5.0f %s\n", name, ncl, (double)r.t[0] / reps,
         (double)r.t[0] / reps / 1965.0, (double)r.t[1] / reps, (double)r.t[2] / reps, (double)r.t[3] / reps, r.fail ? "  ** TIMEOUT **" : "");
}

int main() {
  uint32_t* act; Res* res;
  CK(cudaMalloc(&act, (size_t)4 * ROWS * 4096 * 2));
  CK(cudaMalloc(&res, sizeof(Res)));
  const int reps = 400;
  // H = 1024 activations; exchange of [32][EXF] fp32 per peer (EXF = features a rank finally owns in a qkv-like phase)
  run_chain<1>(act, res, 1024, 0, reps, 1 | 4, 128, "C=1 128 CTAs: poll 64 KB + out");
  run_chain<2>(act, res, 1024, 24, reps, 1 | 4, 64, "C=2 64x2: poll 32 KB + out");
  run_chain<2>(act, res, 1024, 24, reps, 1 | 2 | 4, 64, "C=2 64x2: poll 32 KB + exch 3 KB + out");
  run_chain<4>(act, res, 1024, 24, reps, 1 | 4, 32, "C=4 32x4: poll 16 KB + out");
  run_chain<4>(act, res, 1024, 24, reps, 1 | 2 | 4, 32, "C=4 32x4: poll 16 KB + exch 3x3 KB + out");
  run_chain<4>(act, res, 1024, 64, reps, 1 | 2 | 4, 32, "C=4 32x4: poll 16 KB + exch 3x8 KB + out");
  run_chain<4>(act, res, 1024, 24, reps, 2, 32, "C=4 32x4: exchange 3x3 KB only");
  run_chain<4>(act, res, 4096, 8, reps, 1 | 2 | 4, 32, "C=4 32x4 K=4096: poll 64 KB + exch 3x1 KB + out");
  run_chain<8>(act, res, 1024, 24, reps, 1 | 4, 16, "C=8 16x8: poll 8 KB + out");
  run_chain<8>(act, res, 1024, 24, reps, 1 | 2 | 4, 16, "C=8 16x8: poll 8 KB + exch 7x3 KB + out");
  run_chain<8>(act, res, 4096, 8, reps, 1 | 2 | 4, 16, "C=8 16x8 K=4096: poll 32 KB + exch 7x1 KB + out");
  run_chain<16>(act, res, 1024, 8, reps, 1 | 2 | 4, 8, "C=16 8x16: poll 4 KB + exch 15x1 KB + out");
  return 0;
}
