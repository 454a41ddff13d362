// gemm_tc.cu -- the prefill linear layers as tcgen05 GEMMs (default for the bf16 model dtype; PTTS_PREFILL_TC=0 switches back to
// the mma.sync kernel of gemm.cu for A/B runs).  Validated on B200 in round 2 (tests/test_gpu_parity.py::
// test_prefill_tc_matches_default_prefill and the oracle comparison of the step-0 logits at the bench shape).
//
// Why: at prefill the decoder's linear layers see M = B*(P+1) (prompt) or B*S (encoder K/V projection) rows, ~1000-2000 for
// the bench workload: 6.6 GFLOP per matrix, tensor-bound.  Today they run the decode GEMM (gemm.cu: 32-row tiles, mma.sync,
// weights re-streamed from L2 for every 32 rows) at 115-195 us per launch, 27 ms per generate (profiles/r01_launches.md).
// A tcgen05 tile of 128 rows x 256 features with TMA-fed 128-byte-swizzled operands is the same machinery as the DAC
// convolutions (dac_tc.cu: a k=1 convolution over channels-last activations IS x W^T), so this kernel is that mainloop with
// the nn.Linear epilogues of the reference (modeling_parler_tts.py:1020-1062): optional folded LayerNorm
// (y = rstd*(acc - mean*c1) + c2, ln_stats.cuh), rounding to bf16, GELU, residual add.
// The helpers are private copies of dac_tc.cu's (namespace gtc) so that the verified DAC path is untouched until this
// file has been validated; unify afterwards.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace ptts {
namespace gtc {

constexpr int M_TILE = 128;   // rows per CTA tile (UMMA M)
constexpr int K_STAGE = 64;   // K elements per pipeline stage (one 128-byte swizzle row of bf16)
constexpr int STAGES = 3;
constexpr int THREADS = 192;  // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue (TMEM lane quarters)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_addr(b)), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 24)) { printf("ptts: gemm_tc mbarrier timeout (cta %d,%d thread %d)\n", (int)blockIdx.x, (int)blockIdx.y, (int)threadIdx.x); __trap(); }
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_addr(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_addr(bar)) : "memory");
}
// UMMA shared-memory descriptor, K-major, 128-byte swizzle: 8-row groups 1024 B apart (same encoding as dac_tc.cu)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                 "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                 "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Args {
  int M, N, K, n_tile, tmem_cols;
  const float* stats;  // [M][2] (mean, rstd) per row, or nullptr (no folded LayerNorm)
  const float* c1;     // [N] folded-LayerNorm vectors (with stats)
  const float* c2;
  int epi, act;
  const bf16* R;       // residual [M][N] (EPI_RESIDUAL)
  bf16* Y;             // [M][N]
};

// One CTA: Y[m0:m0+128, n0:n0+n_tile] = epilogue( X[m0:m0+128, :] W[n0:n0+n_tile, :]^T ), X and W row-major bf16 (K contiguous).
__global__ void __launch_bounds__(THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const Args p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a_bytes = M_TILE * K_STAGE * 2, b_bytes = p.n_tile * K_STAGE * 2;
  const int stage_bytes = (a_bytes + b_bytes + 1023) & ~1023;
  unsigned char* stages = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* cvec = reinterpret_cast<float*>(smem + STAGES * stage_bytes + 128);  // [2][n_tile]: c1 | c2

  const int m0 = blockIdx.x * M_TILE, n0 = blockIdx.y * p.n_tile;
  const int n_iter = p.K / K_STAGE;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
  }
  if (warp == 2) {  // TMEM allocation (and later deallocation) by one warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer: rows beyond M / features beyond N are zero-filled by the TMA unit =====
      for (int it = 0; it < n_iter; it++) {
        const int s = it % STAGES, use = it / STAGES;
        if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
        unsigned char* a_dst = stages + (size_t)s * stage_bytes;
        mbar_expect(&full[s], (uint32_t)(a_bytes + b_bytes));
        tma_load_2d(a_dst, &map_x, it * K_STAGE, m0, &full[s]);
        tma_load_2d(a_dst + a_bytes, &map_w, it * K_STAGE, n0, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      // instruction descriptor: D = f32, A = B = bf16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(M_TILE >> 4) << 24);
      for (int it = 0; it < n_iter; it++) {
        const int s = it % STAGES, use = it / STAGES;
        mbar_wait(&full[s], use & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_addr(stages + (size_t)s * stage_bytes);
        const uint64_t da = umma_desc_sw128(a_addr), db = umma_desc_sw128(a_addr + a_bytes);
#pragma unroll
        for (int k = 0; k < K_STAGE / 16; k++)  // 32 bytes (16 bf16) further along K inside the swizzled row: +2 in the address field
          umma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty[s]);  // the stage may be refilled once these MMAs have read it
      }
      umma_commit(acc_full);
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    const int et = threadIdx.x - 64;  // 0..127
    if (p.stats != nullptr) {
      for (int c = et; c < p.n_tile; c += 128) { cvec[c] = p.c1[n0 + c]; cvec[p.n_tile + c] = p.c2[n0 + c]; }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
    const int quarter = warp & 3;
    const int m = m0 + quarter * 32 + lane;
    const bool row_ok = m < p.M;
    float mean = 0.f, rstd = 1.f;
    if (p.stats != nullptr && row_ok) { mean = p.stats[2 * m]; rstd = p.stats[2 * m + 1]; }
    mbar_wait(acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const size_t orow = (size_t)(row_ok ? m : 0) * p.N + n0;
    for (int c0 = 0; c0 < p.n_tile; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);  // warp-collective: all lanes take part
      if (row_ok) {
        __nv_bfloat162 r2[16];
        uint4 res[4];
        if (p.epi == EPI_RESIDUAL) {
#pragma unroll
          for (int i = 0; i < 4; i++) res[i] = *reinterpret_cast<const uint4*>(p.R + orow + c0 + 8 * i);
        }
#pragma unroll
        for (int i = 0; i < 16; i++) {
          float a = __uint_as_float(v[2 * i]), b = __uint_as_float(v[2 * i + 1]);
          if (p.stats != nullptr) {
            a = rstd * (a - mean * cvec[c0 + 2 * i]) + cvec[p.n_tile + c0 + 2 * i];
            b = rstd * (b - mean * cvec[c0 + 2 * i + 1]) + cvec[p.n_tile + c0 + 2 * i + 1];
          }
          a = DT<bf16>::rnd(a); b = DT<bf16>::rnd(b);  // nn.Linear output is rounded to the model dtype
          if (p.epi == EPI_ACT) { a = apply_act(a, p.act); b = apply_act(b, p.act); }
          else if (p.epi == EPI_RESIDUAL) {
            const float2 rr = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(res)[i]);
            a = rr.x + a; b = rr.y + b;
          }
          r2[i] = __floats2bfloat162_rn(a, b);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) *reinterpret_cast<uint4*>(p.Y + orow + c0 + 8 * i) = *reinterpret_cast<const uint4*>(&r2[4 * i]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// (mean, rstd) per row of a bf16 [M][K] matrix: fp32, two passes over a register-resident row (one warp per row)
__global__ void row_stats_kernel(const bf16* __restrict__ X, int64_t ldx, int M, int K, float eps, float* __restrict__ stats) {
  const int warp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= M) return;
  const bf16* row = X + (size_t)warp * ldx;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s += __bfloat162float(row[k]);
  const float mean = warp_sum(s) / (float)K;
  float q = 0.f;
  for (int k = lane; k < K; k += 32) { const float d = __bfloat162float(row[k]) - mean; q = fmaf(d, d, q); }
  const float var = warp_sum(q) / (float)K;
  if (lane == 0) { stats[2 * warp] = mean; stats[2 * warp + 1] = rsqrtf(var + eps); }
}

// mma-fragment order (gemm.cu pack_matrix_bf16_kernel) -> row-major [N][K]
__global__ void unpack_fragments_kernel(const bf16* __restrict__ frag, bf16* __restrict__ dst, int64_t N, int K) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * K) return;
  const int64_t n = i / K, k = i - n * K;
  const int64_t nt = n >> 3, g = n & 7, kt = k >> 5, kk = k & 31;
  const int j = (int)(kk >> 4), c = (int)(kk & 15), half = c >> 3, t = (c & 7) >> 1, e = c & 1;
  const int lane = (int)g * 4 + t, reg = j * 2 + half;
  dst[i] = frag[((nt * (K >> 5) + kt) * 32 + lane) * 8 + reg * 2 + e];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}
// 2-D bf16 matrix [rows][cols] (cols contiguous, row pitch = cols), box {64, box_rows}, 128-byte swizzle, zero OOB fill
static int make_map(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(PTTS_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)K_STAGE, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(PTTS_ECUDA, "cuTensorMapEncodeTiled failed (%d) dims %llu x %llu box rows %u", (int)r, (unsigned long long)rows, (unsigned long long)cols, box_rows);
  return PTTS_OK;
}

static int pick_ntile(int N) {
  if (N % 256 == 0) return 256;
  if (N % 192 == 0) return 192;
  if (N % 128 == 0) return 128;
  if (N % 96 == 0) return 96;
  if (N % 64 == 0) return 64;
  return 32;
}

}  // namespace gtc

bool linear_tc_supported(const LinearArgs& a) {
  return a.M >= gtc::M_TILE && a.K % gtc::K_STAGE == 0 && a.N % 32 == 0 && a.ldx == a.K && a.ldy == a.N && (a.R == nullptr || a.ldr == a.N) &&
         a.epi != EPI_F32;
}

// a: as for launch_linear (bf16); w_rowmajor: the SAME matrix as a.W but row-major [N][K]; stats_scratch: float[2*M] (used when a.c1 != nullptr)
int launch_linear_tc(const LinearArgs& a, const void* w_rowmajor, float* stats_scratch, cudaStream_t st) {
  using namespace gtc;
  Args p{};
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.n_tile = pick_ntile(a.N);
  p.tmem_cols = p.n_tile <= 32 ? 32 : (p.n_tile <= 64 ? 64 : (p.n_tile <= 128 ? 128 : 256));
  p.epi = a.epi; p.act = a.act; p.R = (const bf16*)a.R; p.Y = (bf16*)a.Y;
  if (a.c1 != nullptr) {
    const int warps_per_block = 8;
    row_stats_kernel<<<(a.M + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>((const bf16*)a.X, a.ldx, a.M, a.K, a.eps, stats_scratch);
    PTTS_LAUNCH_CHECK();
    p.stats = stats_scratch; p.c1 = a.c1; p.c2 = a.c2;
  }
  CUtensorMap mx, mw;
  if (int e = make_map(&mx, a.X, (uint64_t)a.K, (uint64_t)a.M, (uint32_t)M_TILE)) return e;
  if (int e = make_map(&mw, w_rowmajor, (uint64_t)a.K, (uint64_t)a.N, (uint32_t)p.n_tile)) return e;
  const int stage_bytes = (M_TILE * K_STAGE * 2 + p.n_tile * K_STAGE * 2 + 1023) & ~1023;
  const size_t smem = (size_t)STAGES * stage_bytes + 128 + 2 * 256 * sizeof(float);
  static bool attr = false;
  if (!attr) {
    PTTS_CHECK_CUDA(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  dim3 grid((a.M + M_TILE - 1) / M_TILE, a.N / p.n_tile, 1);
  linear_tc_kernel<<<grid, THREADS, smem, st>>>(mx, mw, p);
  PTTS_LAUNCH_CHECK();
  return PTTS_OK;
}

int unpack_fragments(const void* frag, void* dst_rowmajor, int64_t N, int K, cudaStream_t st) {
  const int64_t n = N * K;
  gtc::unpack_fragments_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>((const bf16*)frag, (bf16*)dst_rowmajor, N, K);
  PTTS_LAUNCH_CHECK();
  return PTTS_OK;
}

}  // namespace ptts
