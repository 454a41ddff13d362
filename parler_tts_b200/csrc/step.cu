// step.cu -- the fused decode step: ONE persistent kernel per generated token (bf16 model dtype).
//
// Replaces, per step, everything ptts_decode_forward + ptts_sample launch as 195 separate kernels:
// embedding sum, L x {LN+QKV, self-attention with KV append, out_proj+residual, LN+q, cross-attention,
// out_proj+residual, LN+fc1+GELU, fc2+residual}, final LN + K lm heads, logits processors + sampling
// (reference: ParlerTTSForCausalLM.forward with q_len==1, modeling_parler_tts.py:1865-1974 / :983-1074, and one
// iteration of GenerationMixin._sample).
//
// Why one kernel: the step is a chain of ~195 dependent phases, each moving only 0.3-8 MB.  At the HBM
// roofline the whole step lasts ~220 us (Mini, B=32), i.e. ~1.1 us per phase, so launch latency and per-kernel
// ramp dominate a multi-kernel design.  Here one CTA per SM stays resident for the whole step:
//   * phases are separated by a device-wide barrier (monotonic counter in global memory, release/acquire);
//   * the next layer's weight slices are pulled into L2 one layer ahead with cp.async.bulk.prefetch.L2
//     (the stream is static, so HBM keeps flowing while the chain waits on barriers);
//   * the 32-row activation tile is staged by the TMA engine (cp.async.bulk -> shared memory, mbarrier
//     completion), double-buffered over K for fc2;
//   * weights stream from L2 in mma B-fragment order straight into registers (gemm.cu's layout);
//   * attention processes two (row, kv head) items per CTA concurrently (128 threads each, named barriers).
// All reductions keep a fixed order: results are bit-reproducible and identical to the multi-kernel path.
#include "attn_core.cuh"
#include "common.cuh"
#include "kernels.h"
#include "sample_core.cuh"
#include "step.h"

namespace ptts {

constexpr int ST_THREADS = 256;
constexpr int ST_WARPS = 8;

// ---- PTX helpers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
  const char* c = reinterpret_cast<const char*>(p);
  while (bytes > 0) {
    const uint32_t n = bytes > 32768u ? 32768u : bytes;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(c), "r"(n) : "memory");
    c += n;
    bytes -= n;
  }
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void ldmatrix_x4s(uint32_t (&r)[4], const void* smem_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(smem_ptr)));
}
__device__ __forceinline__ void mma_bf16s(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 ldg_stream_s(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// ---- device-wide barrier ------------------------------------------------------------------------
struct GridBar {
  unsigned* ctr;
  unsigned target;
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0) {
      target += gridDim.x;
      __threadfence();
      atomicAdd(ctr, 1u);
      while (ld_acquire(ctr) < target) {}
      __threadfence();
    }
    __syncthreads();
  }
};

// ---- shared-memory context ----------------------------------------------------------------------
struct Smem {
  uint64_t* bars;   // [2] tile buffers
  float* lnp;       // [2*H] gamma | beta
  bf16* tile[2];    // activation tile buffers, row pitch = H + 8
  unsigned char* scratch;  // start of the tile region (aliased by the K-reduction buffer and by attention)
  uint32_t parity;  // bit i: parity to wait for on bars[i]
  int pitch;
  int nbuf;
};

// TMA-stage x[0:M, col0:col0+Kc] into tile buffer `buf` (called by all threads).
__device__ __forceinline__ void stage_tile(Smem& sm, int buf, const bf16* X, int64_t ldx, int col0, int Kc, int M) {
  __syncthreads();  // every generic-proxy access to the buffer (ldmatrix, LN, reduction scratch) is done
  if (threadIdx.x < 32) {
    fence_proxy_async();
    if (threadIdx.x == 0) mbar_expect_tx(&sm.bars[buf], (uint32_t)(M * Kc * 2));
    __syncwarp();
    if ((int)threadIdx.x < M) bulk_g2s(sm.tile[buf] + (size_t)threadIdx.x * sm.pitch, X + (size_t)threadIdx.x * ldx + col0, (uint32_t)(Kc * 2), &sm.bars[buf]);
  }
}
__device__ __forceinline__ void wait_tile(Smem& sm, int buf) {
  mbar_wait(&sm.bars[buf], (sm.parity >> buf) & 1u);
  sm.parity ^= (1u << buf);
}

// In-place LayerNorm of the staged tile; each warp owns rows warp, warp+8, ...; values stay in registers.
__device__ __forceinline__ void ln_tile(bf16* xs, int pitch, int Kc, int M, const float* __restrict__ lnp, int H, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nj = Kc >> 6;  // bf16x2 per lane
  for (int r = warp; r < M; r += ST_WARPS) {
    bf16* row = xs + (size_t)r * pitch;
    float2 v[32];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j++)
      if (j < nj) {
        v[j] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(row + lane * 2 + 64 * j));
        s += v[j].x + v[j].y;
      }
    const float mean = warp_sum(s) / (float)Kc;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j++)
      if (j < nj) {
        const float a = v[j].x - mean, d = v[j].y - mean;
        q += a * a + d * d;
      }
    const float rstd = rsqrtf(warp_sum(q) / (float)Kc + eps);
#pragma unroll
    for (int j = 0; j < 32; j++)
      if (j < nj) {
        const int c = lane * 2 + 64 * j;
        const float2 g = *reinterpret_cast<const float2*>(lnp + c);
        const float2 bb = *reinterpret_cast<const float2*>(lnp + H + c);
        const float y0 = (v[j].x - mean) * rstd * g.x + bb.x;
        const float y1 = (v[j].y - mean) * rstd * g.y + bb.y;
        *reinterpret_cast<__nv_bfloat162*>(row + c) = __floats2bfloat162_rn(y0, y1);
      }
  }
}

struct GemmDesc {
  const bf16* X; int64_t ldx;
  const uint4* W;
  int N, K;
  const float* lnw; const float* lnb;
  int epi;
  const bf16* R;
  void* Y; int64_t ldy;
};

// All tasks (n-blocks of 8*NT features) of one linear layer assigned to this CTA.  M = B <= 32 rows.
template <int NT, int PF>
__device__ __noinline__ void gemm_tasks(const StepParams& p, Smem& sm, const GemmDesc& d) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = p.B, H = p.H;
  const int Kc = d.K < H ? d.K : H;
  const int n_chunks = d.K / Kc;
  const int kt_per_chunk = Kc >> 5, KT = d.K >> 5;
  const int per_chunk = (kt_per_chunk > warp) ? (kt_per_chunk - warp + ST_WARPS - 1) / ST_WARPS : 0;
  const int ntasks = d.N / (8 * NT);
  constexpr int FB = 8 * NT;
  for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
    const int nt0 = task * NT;
    uint4 wr[PF][NT];
    auto load_w = [&](uint4 (&dst)[NT], int c, int i) {
      const int ktg = c * kt_per_chunk + warp + ST_WARPS * i;
#pragma unroll
      for (int j = 0; j < NT; j++) dst[j] = ldg_stream_s(d.W + ((size_t)(nt0 + j) * KT + ktg) * 32 + lane);
    };
#pragma unroll
    for (int s = 0; s < PF; s++)
      if (s < per_chunk) load_w(wr[s], 0, s);
    // activations: chunk 0 (and chunk 1 when double-buffered)
    stage_tile(sm, 0, d.X, d.ldx, 0, Kc, M);
    if (sm.nbuf > 1 && n_chunks > 1) stage_tile(sm, 1, d.X, d.ldx, Kc, Kc, M);
    if (d.lnw != nullptr) {
      for (int i = threadIdx.x; i < H; i += ST_THREADS) { sm.lnp[i] = d.lnw[i]; sm.lnp[H + i] = d.lnb[i]; }
    }
    float acc[2][NT][4];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
      for (int j = 0; j < NT; j++)
#pragma unroll
        for (int e = 0; e < 4; e++) acc[a][j][e] = 0.f;

    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8;
    const int lcol = (lane >> 4) * 8;
    for (int c = 0; c < n_chunks; c++) {
      const int buf = (sm.nbuf > 1) ? (c & 1) : 0;
      if (c > 0) {
#pragma unroll
        for (int s = 0; s < PF; s++)
          if (s < per_chunk) load_w(wr[s], c, s);
        if (sm.nbuf == 1) stage_tile(sm, 0, d.X, d.ldx, c * Kc, Kc, M);
      }
      wait_tile(sm, buf);
      if (d.lnw != nullptr) {
        __syncthreads();  // lnp visible
        ln_tile(sm.tile[buf], sm.pitch, Kc, M, sm.lnp, H, p.eps);
        __syncthreads();
      }
      const bf16* xs = sm.tile[buf];
      for (int i0 = 0; i0 < per_chunk; i0 += PF) {
#pragma unroll
        for (int s = 0; s < PF; s++) {
          const int i = i0 + s;
          if (i < per_chunk) {
            const int kt = warp + ST_WARPS * i;
            uint32_t a[2][2][4];
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
              for (int j = 0; j < 2; j++) ldmatrix_x4s(a[mt][j], xs + (size_t)(mt * 16 + lrow) * sm.pitch + kt * 32 + j * 16 + lcol);
#pragma unroll
            for (int j = 0; j < NT; j++) {
              const uint4 w = wr[s][j];
#pragma unroll
              for (int mt = 0; mt < 2; mt++) {
                mma_bf16s(acc[mt][j], a[mt][0], w.x, w.y);
                mma_bf16s(acc[mt][j], a[mt][1], w.z, w.w);
              }
            }
            if (i + PF < per_chunk) load_w(wr[s], c, i + PF);
          }
        }
      }
      if (sm.nbuf > 1 && c + 2 < n_chunks) stage_tile(sm, buf, d.X, d.ldx, (c + 2) * Kc, Kc, M);
    }
    __syncthreads();
    float* red = reinterpret_cast<float*>(sm.scratch);  // [8][32][FB], aliases the (now idle) tile buffers
    {
      const int g = lane >> 2, t = lane & 3;
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int j = 0; j < NT; j++) {
          float* base = red + ((size_t)warp * 32 + mt * 16 + g) * FB + j * 8 + 2 * t;
          base[0] = acc[mt][j][0];
          base[1] = acc[mt][j][1];
          base[8 * FB] = acc[mt][j][2];
          base[8 * FB + 1] = acc[mt][j][3];
        }
    }
    __syncthreads();
    const int n0 = nt0 * 8;
    for (int o = threadIdx.x; o < 32 * FB; o += ST_THREADS) {
      const int r = o / FB, cidx = o - r * FB;
      if (r >= M) continue;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < ST_WARPS; w++) v += red[((size_t)w * 32 + r) * FB + cidx];
      v = DT<bf16>::rnd(v);
      const size_t yo = (size_t)r * d.ldy + n0 + cidx;
      if (d.epi == EPI_ACT) v = apply_act(v, p.act);
      else if (d.epi == EPI_RESIDUAL) v = DT<bf16>::to_f(d.R[yo]) + v;
      if (d.epi == EPI_F32) reinterpret_cast<float*>(d.Y)[yo] = v;
      else reinterpret_cast<bf16*>(d.Y)[yo] = __float2bfloat16_rn(v);
    }
  }
}

__device__ __noinline__ void run_gemm(const StepParams& p, Smem& sm, const GemmDesc& d, int nt) {
  switch (nt) {
    case 1: gemm_tasks<1, 4>(p, sm, d); break;
    case 2: gemm_tasks<2, 4>(p, sm, d); break;
    case 3: gemm_tasks<3, 4>(p, sm, d); break;
    case 4: gemm_tasks<4, 4>(p, sm, d); break;
    case 6: gemm_tasks<6, 2>(p, sm, d); break;
    default: gemm_tasks<9, 2>(p, sm, d); break;
  }
}

// bytes of this CTA's weight slice for a GEMM (first task only) -> L2, one layer ahead
__device__ __forceinline__ void prefetch_slice(const char* w, int N, int K, int nt) {
  const int ntasks = N / (8 * nt);
  for (int task = blockIdx.x; task < ntasks; task += gridDim.x)
    l2_prefetch(w + (size_t)task * nt * K * 16, (uint32_t)(nt * K * 16));
}

// Two attention items per CTA (one per 128-thread half), named barriers 1 and 2.
__device__ __noinline__ void attn_phase(const StepParams& p, Smem& sm, const AttnArgs& a, int nkv) {
  const int half = threadIdx.x >> 7, tid = threadIdx.x & 127;
  float* region = reinterpret_cast<float*>(sm.scratch) + (size_t)half * p.attn_floats_per_half;
  const int items = p.B * nkv;
  for (int it = blockIdx.x * 2 + half; it < items; it += gridDim.x * 2) {
    const int b = it / nkv, kvh = it - b * nkv;
    attention_item<bf16>(a, b, kvh, region, tid, [half] { asm volatile("bar.sync %0, 128;" ::"r"(half + 1) : "memory"); });
  }
}

template <int ITEMS>
__device__ __noinline__ void sample_phase(const SampleArgs& sa, const ptts_gen_params& gp, int BK, int cur_len) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = blockIdx.x * ST_WARPS + warp; row < BK; row += gridDim.x * ST_WARPS) sample_row<ITEMS>(sa, gp, nullptr, row, cur_len, lane);
}

template <int ITEMS>
__global__ void __launch_bounds__(ST_THREADS, 1) decode_step_kernel(const __grid_constant__ StepParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Ctrl* ctrl = p.sa.ctrl;
  if (ctrl->active == 0) return;  // generation finished: the rest of the enqueued steps are no-ops
  const int cur_len = ctrl->cur_len;
  const unsigned gen = (unsigned)ctrl->launch_gen;
  const int pos = p.P + cur_len - 1;  // cache position of the token being fed
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int H = p.H;

  Smem sm;
  sm.bars = reinterpret_cast<uint64_t*>(smem_raw);
  sm.lnp = reinterpret_cast<float*>(smem_raw + 128);
  sm.scratch = smem_raw + 128 + (size_t)2 * H * sizeof(float);
  sm.pitch = H + 8;
  sm.nbuf = p.nbuf;
  sm.tile[0] = reinterpret_cast<bf16*>(sm.scratch);
  sm.tile[1] = sm.tile[0] + (size_t)32 * sm.pitch;
  sm.parity = 0;
  if (tid == 0) {
    mbar_init(&sm.bars[0], 1);
    mbar_init(&sm.bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // rows >= B of the tile buffers are never written by the TMA copies: clear them once
  for (int i = tid; i < (int)(p.tile_region_bytes / 16); i += ST_THREADS) reinterpret_cast<uint4*>(sm.scratch)[i] = make_uint4(0, 0, 0, 0);
  GridBar bar{p.bar + (gen & 1u), 0u};
  if (blockIdx.x == 0 && tid == 0) p.bar[(gen + 1u) & 1u] = 0u;  // the counter the NEXT launch will use
  __syncthreads();

  const char* blob = p.blob;
  // ---- phase 0: embeddings (one batch row per CTA) + L2 prefetch of layer 0 ----
  if (tid == 0) {
    const char* lb = blob + p.layer0;
    prefetch_slice(lb + p.wqkv, p.qkv_rows, H, p.nt_qkv);
    prefetch_slice(lb + p.wo, H, H, p.nt_h);
    prefetch_slice(lb + p.wqc, H, H, p.nt_h);
    prefetch_slice(lb + p.woc, H, H, p.nt_h);
    prefetch_slice(lb + p.fc1, p.F, H, p.nt_fc1);
    prefetch_slice(lb + p.fc2, H, p.F, p.nt_h);
  }
  for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
    const bf16* tables = reinterpret_cast<const bf16*>(blob + p.embed);
    const bf16* postab = p.rope ? nullptr : reinterpret_cast<const bf16*>(blob + p.pos);
    for (int c = tid; c < H; c += ST_THREADS) {
      float v = 0.f;
      for (int k = 0; k < p.K; k++) {
        const int id = p.sa.cur_ids[b * p.K + k];
        const float e = __bfloat162float(tables[((size_t)k * (p.V + 1) + id) * H + c]);
        v = (k == 0) ? e : DT<bf16>::rnd(v + e);
      }
      if (postab != nullptr) v = DT<bf16>::rnd(v + __bfloat162float(postab[(size_t)pos * H + c]));
      p.x[(size_t)b * H + c] = __float2bfloat16_rn(v);
    }
  }
  bar.sync();

  AttnArgs at{};
  at.ldo = H; at.out = p.attn; at.ctrl = nullptr; at.B = p.B; at.nh = p.nh; at.q_len = 1;
  at.past_from_ctrl = 0; at.past_len = pos; at.prefix = p.P;
  at.rope = p.rope; at.rope_cos = blob + p.rope_cos; at.rope_sin = blob + p.rope_sin; at.scale = p.scale;

  // One loop over all 8L+1 dependent phases (single call site per phase type keeps code size and registers sane).
#pragma unroll 1
  for (int ph = 0; ph <= 8 * p.L; ph++) {
    const int l = ph >> 3, sub = (ph == 8 * p.L) ? 8 : (ph & 7);
    const char* lb = blob + p.layer0 + p.layer_stride * (l < p.L ? l : p.L - 1);
    if (sub == 0 && tid == 0) {  // pull the NEXT layer's weight slices (or the lm heads) into L2 while this layer runs
      if (l + 1 < p.L) {
        const char* nb = lb + p.layer_stride;
        prefetch_slice(nb + p.wqkv, p.qkv_rows, H, p.nt_qkv);
        prefetch_slice(nb + p.wo, H, H, p.nt_h);
        prefetch_slice(nb + p.wqc, H, H, p.nt_h);
        prefetch_slice(nb + p.woc, H, H, p.nt_h);
        prefetch_slice(nb + p.fc1, p.F, H, p.nt_fc1);
        prefetch_slice(nb + p.fc2, H, p.F, p.nt_h);
      } else {
        prefetch_slice(blob + p.heads, p.K * p.V, H, p.nt_heads);
      }
    }
    if (sub == 1 || sub == 4) {
      AttnArgs a = at;
      if (sub == 1) {  // self-attention over the cache (+ append of the new K/V row)
        a.q = p.qkv; a.ldq = p.qkv_rows; a.q_col0 = 0;
        a.knew = p.qkv; a.vnew = p.qkv; a.ldkv = p.qkv_rows; a.k_col0 = p.nh * HD; a.v_col0 = (p.nh + p.nkv) * HD;
        char* kc = p.self_kv + p.self_layer_stride * l;
        a.kcache = kc; a.vcache = kc + (size_t)p.B * p.nkv * p.Tmax * HD * 2;
        a.kv_b_stride = (int64_t)p.nkv * p.Tmax * HD; a.kv_h_stride = (int64_t)p.Tmax * HD; a.kv_t_stride = HD;
        a.key_mask = p.prompt_mask; a.mask_len = p.P; a.mask_ld = p.P;
        a.nkv = p.nkv; a.cross = 0; a.kv_len = 0; a.kv_capacity = p.Tmax;
      } else {         // cross-attention over the cached encoder K/V
        a.q = p.qc; a.ldq = H; a.q_col0 = 0; a.knew = nullptr; a.vnew = nullptr;
        char* ck = p.cross_kv + p.cross_layer_stride * l;
        a.kcache = ck; a.vcache = ck + (size_t)p.nckv * HD * 2;
        a.kv_b_stride = (int64_t)p.S * p.ckv_rows; a.kv_h_stride = HD; a.kv_t_stride = p.ckv_rows;
        a.key_mask = p.enc_mask; a.mask_len = p.S; a.mask_ld = p.S;
        a.nkv = p.nckv; a.cross = 1; a.kv_len = p.S; a.kv_capacity = p.S;
      }
      attn_phase(p, sm, a, a.nkv);
    } else {
      GemmDesc g{};
      int nt = p.nt_h;
      switch (sub) {
        case 0:  // qkv = LN1(x) Wqkv^T
          g = GemmDesc{p.x, H, reinterpret_cast<const uint4*>(lb + p.wqkv), p.qkv_rows, H, reinterpret_cast<const float*>(lb + p.ln1_w),
                       reinterpret_cast<const float*>(lb + p.ln1_b), EPI_STORE, nullptr, p.qkv, p.qkv_rows};
          nt = p.nt_qkv;
          break;
        case 2:  // x += attn Wo^T
          g = GemmDesc{p.attn, H, reinterpret_cast<const uint4*>(lb + p.wo), H, H, nullptr, nullptr, EPI_RESIDUAL, p.x, p.x, H};
          break;
        case 3:  // q_cross = LN2(x) Wq^T
          g = GemmDesc{p.x, H, reinterpret_cast<const uint4*>(lb + p.wqc), H, H, reinterpret_cast<const float*>(lb + p.ln2_w),
                       reinterpret_cast<const float*>(lb + p.ln2_b), EPI_STORE, nullptr, p.qc, H};
          break;
        case 5:  // x += attn Wo_cross^T
          g = GemmDesc{p.attn, H, reinterpret_cast<const uint4*>(lb + p.woc), H, H, nullptr, nullptr, EPI_RESIDUAL, p.x, p.x, H};
          break;
        case 6:  // h = act(LN3(x) W1^T)
          g = GemmDesc{p.x, H, reinterpret_cast<const uint4*>(lb + p.fc1), p.F, H, reinterpret_cast<const float*>(lb + p.ln3_w),
                       reinterpret_cast<const float*>(lb + p.ln3_b), EPI_ACT, nullptr, p.hbuf, p.F};
          nt = p.nt_fc1;
          break;
        case 7:  // x += h W2^T
          g = GemmDesc{p.hbuf, p.F, reinterpret_cast<const uint4*>(lb + p.fc2), H, p.F, nullptr, nullptr, EPI_RESIDUAL, p.x, p.x, H};
          break;
        default:  // final LayerNorm + K lm heads -> f32 logits [B, K*V]
          g = GemmDesc{p.x, H, reinterpret_cast<const uint4*>(blob + p.heads), p.K * p.V, H, reinterpret_cast<const float*>(blob + p.final_ln_w),
                       reinterpret_cast<const float*>(blob + p.final_ln_b), EPI_F32, nullptr, p.logits, (int64_t)p.K * p.V};
          nt = p.nt_heads;
          break;
      }
      run_gemm(p, sm, g, nt);
    }
    if (ph < 8 * p.L) bar.sync();
  }
  bar.sync();
  if (p.do_sample_phase) {
    const ptts_gen_params gp = *p.sa.gen;
    const int BK = p.B * p.K;
    sample_phase<ITEMS>(p.sa, gp, BK, cur_len);
    bar.sync();
  }
  if (blockIdx.x == 0 && tid == 0) {
    if (p.do_sample_phase) {
      const int n = atomicAdd(&ctrl->n_unfinished, 0);
      ctrl->cur_len = cur_len + 1;
      ctrl->active = (n > 0) ? 1 : 0;
      ctrl->steps_run += 1;
      ctrl->n_unfinished = 0;
    }
    ctrl->launch_gen = (int)(gen + 1u);
  }
}

// ---- host side ----------------------------------------------------------------------------------
int step_smem_bytes(const StepParams& p) {
  return (int)(128 + (size_t)2 * p.H * sizeof(float) + p.tile_region_bytes);
}

int launch_decode_step(const StepParams& p, int grid, cudaStream_t st) {
  const int smem = step_smem_bytes(p);
  void* args[] = {(void*)&p};
  const void* fn;
  if (p.sample_items <= 4) fn = (const void*)decode_step_kernel<4>;
  else if (p.sample_items <= 36) fn = (const void*)decode_step_kernel<36>;
  else fn = (const void*)decode_step_kernel<72>;
  static int attr_done[3] = {0, 0, 0};
  const int fi = p.sample_items <= 4 ? 0 : (p.sample_items <= 36 ? 1 : 2);
  if (!attr_done[fi]) {
    PTTS_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done[fi] = 1;
  }
  PTTS_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(ST_THREADS), args, (size_t)smem, st));
  return PTTS_OK;
}

}  // namespace ptts
