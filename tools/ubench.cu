// ubench.cu -- micro-benchmarks of the fused step kernel's primitives on a resident 148-CTA grid.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench.cu && /tmp/ubench
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// mode 0: release-red + relaxed spin + acq fence; 1: threadfence + atomicAdd + acquire spin; 2: relaxed spin with nanosleep
__device__ __forceinline__ unsigned grid_sync(unsigned* ctr, unsigned target, int mode) {
  target += gridDim.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (mode == 1) { __threadfence(); atomicAdd(ctr, 1u); while (ld_acquire(ctr) < target) {} __threadfence(); }
    else { red_release_add(ctr, 1u); while (ld_relaxed(ctr) < target) { if (mode == 2) __nanosleep(20); } asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
  }
  __syncthreads();
  return target;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do { asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory"); } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct Res { long long t[64]; };

// test: tile staging.  variant 0: bulk per row; 1: bulk 4 segs staggered; 2: cp.async 16B; 3: plain LDG.128 -> STS; 4: bulk, every CTA its own tile copy
__global__ void __launch_bounds__(256, 1) k_stage(const bf16* X, bf16* Xpriv, unsigned* ctr, Res* res, int variant, int reps, int barmode) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  bf16* tile = reinterpret_cast<bf16*>(smem + 128);
  const int pitch = 1024 + 8;
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  unsigned target = 0, parity = 0;
  long long tot = 0, tbar = 0;
  const bf16* src = (variant == 4) ? Xpriv + (size_t)blockIdx.x * 32 * 1024 : X;
  for (int it = 0; it < reps; it++) {
    if (variant >= 5 && blockIdx.x < 128) {  // every CTA rewrites its 8-column slice of all 32 rows, like the GEMM epilogue
      const int r = tid >> 3, c = tid & 7;
      const_cast<bf16*>(X)[(size_t)r * 1024 + blockIdx.x * 8 + c] = __float2bfloat16((float)(it + r + c));
    }
    long long b0 = clock64();
    target = grid_sync(ctr, target, barmode);
    long long t0 = clock64();
    if (variant == 0 || variant == 1 || variant == 4 || variant == 5) {
      if (tid < 32) {
        asm volatile("fence.proxy.async;" ::: "memory");
        if (tid == 0) mbar_expect_tx(bar, 32 * 1024 * 2);
        __syncwarp();
        if (variant == 1) {
          const int r = (tid + blockIdx.x) & 31;
          const int rot = (blockIdx.x >> 5) & 3;
          for (int s = 0; s < 4; s++) { const int sg = (s + rot) & 3; bulk_g2s(tile + r * pitch + sg * 256, src + (size_t)r * 1024 + sg * 256, 512, bar); }
        } else {
          bulk_g2s(tile + tid * pitch, src + (size_t)tid * 1024, 2048, bar);
        }
      }
      mbar_wait(bar, parity); parity ^= 1;
    } else if (variant == 2 || variant == 7) {
      for (int v = tid; v < 32 * 128; v += 256) {
        const int r = v >> 7, c = v & 127;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(tile + r * pitch + c * 8)), "l"(src + (size_t)r * 1024 + c * 8));
      }
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
      __syncthreads();
    } else {
      uint4 v[16];
#pragma unroll
      for (int i = 0; i < 16; i++) { const int idx = tid + 256 * i; v[i] = *reinterpret_cast<const uint4*>(src + (size_t)(idx >> 7) * 1024 + (idx & 127) * 8); }
#pragma unroll
      for (int i = 0; i < 16; i++) { const int idx = tid + 256 * i; *reinterpret_cast<uint4*>(tile + (idx >> 7) * pitch + (idx & 127) * 8) = v[i]; }
      __syncthreads();
    }
    long long t1 = clock64();
    tot += t1 - t0; tbar += t0 - b0;
    // touch + modify X a little so nothing is optimised away and so the line is "written by someone" each round
    if (blockIdx.x == (it % gridDim.x) && tid == 0) const_cast<bf16*>(X)[it % 1024] = tile[it % 1024];
  }
  if (tid == 0) { res[blockIdx.x].t[0] = tot / reps; res[blockIdx.x].t[1] = tbar / reps; }
}

// LayerNorm on a resident tile (no global traffic): 4-row ILP version
__global__ void __launch_bounds__(256, 1) k_ln(Res* res, int reps) {
  extern __shared__ __align__(128) unsigned char smem[];
  bf16* xs = reinterpret_cast<bf16*>(smem + 128);
  float* lnp = reinterpret_cast<float*>(smem + 128 + 32 * 1032 * 2);
  const int pitch = 1032, Kc = 1024, H = 1024;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 32 * pitch; i += 256) xs[i] = __float2bfloat16((float)((i * 37) % 101) * 0.01f);
  for (int i = tid; i < 2 * H; i += 256) lnp[i] = 1.0f + (i % 7) * 0.01f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < reps; it++) {
    bf16* row[4];
#pragma unroll
    for (int i = 0; i < 4; i++) row[i] = xs + (size_t)(warp + 8 * i) * pitch;
    float s[4] = {0, 0, 0, 0};
#pragma unroll 2
    for (int c = lane * 2; c < Kc; c += 64)
#pragma unroll
      for (int i = 0; i < 4; i++) { const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(row[i] + c)); s[i] += v.x + v.y; }
    float mean[4], q[4] = {0, 0, 0, 0}, rstd[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { float v = s[i]; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); mean[i] = v / Kc; }
#pragma unroll 2
    for (int c = lane * 2; c < Kc; c += 64)
#pragma unroll
      for (int i = 0; i < 4; i++) { const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(row[i] + c)); const float a = v.x - mean[i], d = v.y - mean[i]; q[i] += a * a + d * d; }
#pragma unroll
    for (int i = 0; i < 4; i++) { float v = q[i]; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); rstd[i] = rsqrtf(v / Kc + 1e-5f); }
#pragma unroll 2
    for (int c = lane * 2; c < Kc; c += 64) {
      const float2 g = *reinterpret_cast<const float2*>(lnp + c);
      const float2 bb = *reinterpret_cast<const float2*>(lnp + H + c);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(row[i] + c));
        *reinterpret_cast<__nv_bfloat162*>(row[i] + c) = __floats2bfloat162_rn((v.x - mean[i]) * rstd[i] * g.x + bb.x, (v.y - mean[i]) * rstd[i] * g.y + bb.y);
      }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0) res[blockIdx.x].t[0] = (t1 - t0) / reps;
}

// vectorised LayerNorm (LDS.128, the version in step.cu)
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}
__global__ void __launch_bounds__(256, 1) k_ln_vec(Res* res, int reps) {
  extern __shared__ __align__(128) unsigned char smem[];
  bf16* xs = reinterpret_cast<bf16*>(smem + 128);
  float* lnp = reinterpret_cast<float*>(smem + 128 + 32 * 1032 * 2);
  const int pitch = 1032, Kc = 1024, H = 1024;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 32 * pitch; i += 256) xs[i] = __float2bfloat16((float)((i * 37) % 101) * 0.01f);
  for (int i = tid; i < 2 * H; i += 256) lnp[i] = 1.0f + (i % 7) * 0.01f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < reps; it++) {
    bf16* row[4];
#pragma unroll
    for (int i = 0; i < 4; i++) row[i] = xs + (size_t)(warp + 8 * i) * pitch;
    float s[4] = {0, 0, 0, 0};
    for (int c = lane * 8; c < Kc; c += 256)
#pragma unroll
      for (int i = 0; i < 4; i++) { float f[8]; unpack8(*reinterpret_cast<const uint4*>(row[i] + c), f); s[i] += ((f[0] + f[1]) + (f[2] + f[3])) + ((f[4] + f[5]) + (f[6] + f[7])); }
    float mean[4], q[4] = {0, 0, 0, 0}, rstd[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { float v = s[i]; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); mean[i] = v / Kc; }
    for (int c = lane * 8; c < Kc; c += 256)
#pragma unroll
      for (int i = 0; i < 4; i++) { float f[8]; unpack8(*reinterpret_cast<const uint4*>(row[i] + c), f);
#pragma unroll
        for (int e = 0; e < 8; e++) { const float d = f[e] - mean[i]; q[i] = fmaf(d, d, q[i]); } }
#pragma unroll
    for (int i = 0; i < 4; i++) { float v = q[i]; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); rstd[i] = rsqrtf(v / Kc + 1e-5f); }
    for (int c = lane * 8; c < Kc; c += 256) {
      float g[8], bb[8];
      *reinterpret_cast<float4*>(g) = *reinterpret_cast<const float4*>(lnp + c);
      *reinterpret_cast<float4*>(g + 4) = *reinterpret_cast<const float4*>(lnp + c + 4);
      *reinterpret_cast<float4*>(bb) = *reinterpret_cast<const float4*>(lnp + H + c);
      *reinterpret_cast<float4*>(bb + 4) = *reinterpret_cast<const float4*>(lnp + H + c + 4);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        float f[8]; unpack8(*reinterpret_cast<const uint4*>(row[i] + c), f);
        uint4 o; __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int e = 0; e < 4; e++) oh[e] = __floats2bfloat162_rn((f[2 * e] - mean[i]) * rstd[i] * g[2 * e] + bb[2 * e], (f[2 * e + 1] - mean[i]) * rstd[i] * g[2 * e + 1] + bb[2 * e + 1]);
        *reinterpret_cast<uint4*>(row[i] + c) = o;
      }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0) res[blockIdx.x].t[0] = (t1 - t0) / reps;
}

// pure barrier cost
__global__ void __launch_bounds__(256, 1) k_bar(unsigned* ctr, Res* res, int reps, int mode) {
  unsigned target = 0;
  long long t0 = clock64();
  for (int it = 0; it < reps; it++) target = grid_sync(ctr, target, mode);
  long long t1 = clock64();
  if (threadIdx.x == 0) res[blockIdx.x].t[0] = (t1 - t0) / reps;
}

static void report(const char* name, Res* d_res, int grid, int nslots) {
  Res* h = (Res*)malloc(sizeof(Res) * grid);
  CK(cudaMemcpy(h, d_res, sizeof(Res) * grid, cudaMemcpyDeviceToHost));
  for (int s = 0; s < nslots; s++) {
    long long mn = 1LL << 60, mx = 0, sum = 0;
    for (int i = 0; i < grid; i++) { long long v = h[i].t[s]; mn = v < mn ? v : mn; mx = v > mx ? v : mx; sum += v; }
    printf("%-46s slot%d: avg %8.0f cyc (%.2f us)  min %lld max %lld\n", name, s, (double)sum / grid, (double)sum / grid / 1965.0, mn, mx);
  }
  free(h);
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  printf("SMs %d\n", sms);
  const int grid = sms;
  bf16 *X, *Xp;
  unsigned* ctr;
  Res* res;
  CK(cudaMalloc(&X, 32 * 1024 * 2));
  CK(cudaMalloc(&Xp, (size_t)grid * 32 * 1024 * 2));
  CK(cudaMalloc(&ctr, 256));
  CK(cudaMalloc(&res, sizeof(Res) * grid));
  CK(cudaMemset(X, 0, 32 * 1024 * 2));
  CK(cudaMemset(Xp, 0, (size_t)grid * 32 * 1024 * 2));
  const int smem = 128 + 32 * 1032 * 2 + 2 * 1024 * 4;
  CK(cudaFuncSetAttribute(k_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(k_ln, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int reps = 200;
  for (int mode = 0; mode < 3; mode++) {
    CK(cudaMemset(ctr, 0, 256));
    void* args[] = {&ctr, &res, &reps, &mode};
    CK(cudaLaunchCooperativeKernel((void*)k_bar, dim3(grid), dim3(256), args, 0, 0));
    CK(cudaDeviceSynchronize());
    char nm[64]; snprintf(nm, 64, "grid barrier mode %d", mode);
    report(nm, res, grid, 1);
  }
  const char* vn[8] = {"stage: bulk row (same tile, all CTAs)", "stage: bulk 4 seg staggered", "stage: cp.async 16B", "stage: LDG.128 x16 -> STS", "stage: bulk row, private tile per CTA", "stage: bulk row, tile REWRITTEN by 128 CTAs each round", "stage: LDG.128, tile rewritten each round", "stage: cp.async, tile rewritten each round"};
  for (int v = 0; v < 8; v++) {
    CK(cudaMemset(ctr, 0, 256));
    int barmode = 0;
    void* args[] = {&X, &Xp, &ctr, &res, &v, &reps, &barmode};
    CK(cudaLaunchCooperativeKernel((void*)k_stage, dim3(grid), dim3(256), args, smem, 0));
    CK(cudaDeviceSynchronize());
    report(vn[v], res, grid, 2);
  }
  {
    void* args[] = {&res, &reps};
    CK(cudaLaunchCooperativeKernel((void*)k_ln, dim3(grid), dim3(256), args, smem, 0));
    CK(cudaDeviceSynchronize());
    report("layernorm 32x1024 tile (smem only)", res, grid, 1);
  }
  {
    CK(cudaFuncSetAttribute(k_ln_vec, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    void* args[] = {&res, &reps};
    CK(cudaLaunchCooperativeKernel((void*)k_ln_vec, dim3(grid), dim3(256), args, smem, 0));
    CK(cudaDeviceSynchronize());
    report("layernorm vectorised (LDS.128)", res, grid, 1);
  }
  printf("done\n");
  return 0;
}
