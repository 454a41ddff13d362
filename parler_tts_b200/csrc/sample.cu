// sample.cu -- logits -> next token, entirely on the device, plus the integer ops around it.
//
// One warp per row (B*K rows).  Replaces, per decode step:
//   MinNewTokensLengthLogitsProcessor, ParlerTTSLogitsProcessor (logits_processors.py:44-53, stateful:
//   quirk Q11), Temperature/TopK/TopP warpers, softmax + multinomial / argmax, finished-row padding,
//   torch.cat of the history, EosTokenCriteria + MaxLengthCriteria, `unfinished.max()==0` (a host sync
//   per step in the reference) -- i.e. one iteration of transformers' GenerationMixin._sample -- and
//   apply_delay_pattern_mask on the next input (modeling_parler_tts.py:2909, :205-211).
// The step index lives in the device control block so a captured CUDA graph replays unchanged.
// RNG: Philox4x32-10 keyed by the user seed, counter = (row, column): results do not depend on how
// the batch is sharded over GPUs.  torch.multinomial's stream cannot be reproduced
// bit-for-bit (SURVEY hard part 2); sampling parity is distribution-level, greedy is exact.
#include "common.cuh"
#include "kernels.h"

namespace ptts {

__device__ __forceinline__ uint32_t fkey(float f) {  // order-preserving float -> uint
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  c[0] = hi1 ^ c[1] ^ k0; c[1] = lo1; c[2] = hi0 ^ c[3] ^ k1; c[3] = lo0;
}
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t row, uint32_t col) {
  uint32_t c[4] = {row, col, 0x5054u, 0x5453u};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; i++) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  return (float)(c[0] >> 8) * (1.0f / 16777216.0f);  // [0, 1)
}

constexpr int SAMPLE_WARPS = 4;

template <int ITEMS>
__global__ void __launch_bounds__(SAMPLE_WARPS * 32) sample_kernel(SampleArgs p, const int64_t* __restrict__ forced) {
  pdl_launch_dependents();
  pdl_wait();
  if (p.ctrl->active == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * SAMPLE_WARPS + warp;
  const int BK = p.B * p.K;
  const int cur_len = p.ctrl->cur_len;  // the new token becomes column `cur_len`
  const ptts_gen_params g = *p.gen;
  int still_unfinished = 0;
  if (row < BK) {
    const int b = row / p.K, k = row - b * p.K;
    float v[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
      const int i = lane + 32 * j;
      v[j] = (i < p.V) ? p.logits[(size_t)row * p.V + i] : -INFINITY;
    }
    const int eos_lane = p.eos & 31, eos_j = p.eos >> 5;
    bool mask_eos = false;
    // MinNewTokensLength: prompt_length_to_skip = 1 (the BOS column)
    if (cur_len - 1 < g.min_new_tokens) mask_eos = true;
    // ParlerTTSLogitsProcessor (stateful; state double-buffered on the column parity)
    {
      const int par = cur_len & 1;
      int fu = p.first_unf[par * p.B + b];
      if (p.eos_seen[fu] > 0 && fu < b * p.K + p.K - 1) fu++;
      if (k == 0 && lane == 0) p.first_unf[(par ^ 1) * p.B + b] = fu;
      if (row > fu) mask_eos = true;
    }
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
      const int i = lane + 32 * j;
      if (mask_eos && j == eos_j && lane == eos_lane) v[j] = -INFINITY;
      if (g.suppress_special && i >= g.codebook_size) v[j] = -INFINITY;
    }
    int tok = 0;
    if (g.do_sample) {
      if (g.temperature != 1.0f) {
#pragma unroll
        for (int j = 0; j < ITEMS; j++) v[j] = v[j] / g.temperature;
      }
      if (g.top_k > 0) {
        const int kk = g.top_k < p.V ? g.top_k : p.V;
        uint32_t th = 0;
        for (int bit = 31; bit >= 0; bit--) {
          const uint32_t cand = th | (1u << bit);
          int cnt = 0;
#pragma unroll
          for (int j = 0; j < ITEMS; j++) cnt += (lane + 32 * j < p.V && fkey(v[j]) >= cand) ? 1 : 0;
          cnt = __reduce_add_sync(0xffffffffu, cnt);
          if (cnt >= kk) th = cand;
        }
#pragma unroll
        for (int j = 0; j < ITEMS; j++)
          if (fkey(v[j]) < th) v[j] = -INFINITY;  // scores < kth largest
      }
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < ITEMS; j++) m = fmaxf(m, v[j]);
      m = warp_max(m);
      float e[ITEMS];
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < ITEMS; j++) { e[j] = expf(v[j] - m); s += e[j]; }
      s = warp_sum(s);
      if (g.top_p < 1.0f) {
        // remove tokens whose ascending cumulative probability is <= 1 - top_p (the max is always kept)
        const float thr = (1.0f - g.top_p) * s;
        uint32_t th = 0;
        for (int bit = 31; bit >= 0; bit--) {
          const uint32_t cand = th | (1u << bit);
          float c = 0.f;
#pragma unroll
          for (int j = 0; j < ITEMS; j++) c += (fkey(v[j]) <= cand) ? e[j] : 0.f;
          c = warp_sum(c);
          if (c <= thr) th = cand;
        }
        const uint32_t kmax = fkey(m);
        s = 0.f;
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
          const uint32_t key = fkey(v[j]);
          if (key <= th && key != kmax) { v[j] = -INFINITY; e[j] = 0.f; }
          s += e[j];
        }
        s = warp_sum(s);
      }
      // inverse-CDF draw in index order
      const float target = philox_uniform(g.seed, (uint32_t)row, (uint32_t)cur_len) * s;
      float carry = 0.f;
      int found = -1, last_nz = -1;
#pragma unroll
      for (int j = 0; j < ITEMS; j++) {
        float x = e[j];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float y = __shfl_up_sync(0xffffffffu, x, o);
          if (lane >= o) x += y;
        }
        const float cum = carry + x;
        const unsigned hit = __ballot_sync(0xffffffffu, cum > target && e[j] > 0.f);
        const unsigned nz = __ballot_sync(0xffffffffu, e[j] > 0.f);
        if (nz) last_nz = 32 * j + (31 - __clz(nz));
        if (found < 0 && hit) found = 32 * j + (__ffs(hit) - 1);
        carry = __shfl_sync(0xffffffffu, cum, 31);
      }
      tok = found >= 0 ? found : last_nz;
    } else {
      // argmax, smallest index on ties
      float m = -INFINITY;
      int mi = 0x7fffffff;
#pragma unroll
      for (int j = 0; j < ITEMS; j++) {
        const int i = lane + 32 * j;
        if (i < p.V && (v[j] > m || (v[j] == m && i < mi))) { m = v[j]; mi = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, o);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
      }
      tok = mi;
    }
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
      const int i = lane + 32 * j;
      if (i < p.V) p.scores[(size_t)row * p.V + i] = v[j];
    }
    if (lane == 0) {
      if (forced != nullptr) tok = (int)forced[row];
      const int unf = p.unfinished[row];
      if (!unf) tok = p.pad;  // next_tokens * unfinished + pad * (1 - unfinished)
      p.raw_ids[(size_t)row * p.raw_ld + cur_len] = tok;
      if (tok == p.eos) p.eos_seen[row] = 1;
      const int new_len = cur_len + 1;
      const int done = (tok == p.eos) || (new_len >= g.max_length);
      still_unfinished = unf && !done;
      p.unfinished[row] = still_unfinished;
      // delay-pattern override of the NEXT model input (column `cur_len`), build_delay_pattern_mask :252-261
      int nxt = tok;
      if (g.max_length >= 2 * p.K - 1) {
        const bool is_bos = cur_len <= k;
        const bool is_pad = (cur_len - k) >= (g.max_length - p.K + 1);
        if (is_bos || is_pad) nxt = (is_bos ? p.bos : 0) + (is_pad ? p.pad : 0);
      }
      p.cur_ids[row] = nxt;
      if (still_unfinished) atomicAdd(&p.ctrl->n_unfinished, 1);
    }
  }
  // last block advances the control block
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int ticket = atomicAdd(&p.ctrl->done_blocks, 1);
    if (ticket == (int)gridDim.x - 1) {
      __threadfence();
      const int n = atomicAdd(&p.ctrl->n_unfinished, 0);
      p.ctrl->cur_len = cur_len + 1;
      p.ctrl->active = (n > 0) ? 1 : 0;
      p.ctrl->steps_run += 1;
      p.ctrl->n_unfinished = 0;
      p.ctrl->done_blocks = 0;
      __threadfence();
    }
  }
}

int launch_sample(const SampleArgs& a, const int64_t* forced, cudaStream_t st, bool pdl) {
  const int BK = a.B * a.K;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((BK + SAMPLE_WARPS - 1) / SAMPLE_WARPS);
  cfg.blockDim = dim3(SAMPLE_WARPS * 32);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  const int items = (a.V + 31) / 32;
  PTTS_REQUIRE(items <= 72, "sample: vocab_size %d > 2304 not supported", a.V);
  if (items <= 4) PTTS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, sample_kernel<4>, a, forced));
  else if (items <= 36) PTTS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, sample_kernel<36>, a, forced));
  else PTTS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, sample_kernel<72>, a, forced));
  return PTTS_OK;
}

__global__ void generate_begin_kernel(SampleArgs p) {
  const int BK = p.B * p.K;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    p.ctrl->cur_len = 1; p.ctrl->active = 1; p.ctrl->n_unfinished = 0; p.ctrl->done_blocks = 0; p.ctrl->steps_run = 0;
  }
  if (i < BK) {
    p.raw_ids[(size_t)i * p.raw_ld] = p.bos;
    p.cur_ids[i] = p.bos;
    p.eos_seen[i] = 0;
    p.unfinished[i] = 1;
  }
  if (i < p.B) { p.first_unf[i] = i * p.K; p.first_unf[p.B + i] = i * p.K; }
}
int launch_generate_begin(const SampleArgs& a, cudaStream_t st) {
  const int n = a.B * a.K;
  generate_begin_kernel<<<(n + 127) / 128, 128, 0, st>>>(a);
  PTTS_LAUNCH_CHECK();
  return PTTS_OK;
}

// ---- stand-alone integer operators --------------------------------------------------------------
// build_delay_pattern_mask (modeling_parler_tts.py:214-276): pattern_mask only (bit-exact int64).
__global__ void delay_build_kernel(const int64_t* ids, int BK, int seq, int K, int64_t bos, int64_t pad, int L, int64_t* mask) {
  const int row = blockIdx.x;
  const int k = row % K;
  for (int c = threadIdx.x; c < L; c += blockDim.x) {
    int64_t out = -1;
    if (L >= 2 * K - 1) {
      const bool bos_pat = c <= k;
      const bool eos_pat = (c - k) >= (L - K + 1);
      const int64_t shifted = (c >= k && c < seq + k) ? ids[(size_t)row * seq + (c - k)] : -1;
      out = ((!bos_pat && !eos_pat) ? shifted : 0) + (bos_pat ? bos : 0) + (eos_pat ? pad : 0);
    }
    mask[(size_t)row * L + c] = out;
  }
}
int launch_delay_build(const int64_t* ids, int BK, int seq, int K, int64_t bos, int64_t pad, int L, int64_t* mask, cudaStream_t st) {
  delay_build_kernel<<<BK, 128, 0, st>>>(ids, BK, seq, K, bos, pad, L, mask);
  PTTS_LAUNCH_CHECK();
  return PTTS_OK;
}
__global__ void delay_apply_kernel(const int64_t* ids, int seq, int64_t ld_ids, const int64_t* mask, int64_t ld_mask, int64_t* out) {
  const int row = blockIdx.x;
  for (int c = threadIdx.x; c < seq; c += blockDim.x) {
    const int64_t m = mask[(size_t)row * ld_mask + c];
    out[(size_t)row * seq + c] = (m == -1) ? ids[(size_t)row * ld_ids + c] : m;
  }
}
int launch_delay_apply(const int64_t* ids, int BK, int seq, int64_t ld_ids, const int64_t* mask, int64_t ld_mask, int64_t* out, cudaStream_t st) {
  delay_apply_kernel<<<BK, 128, 0, st>>>(ids, seq, ld_ids, mask, ld_mask, out);
  PTTS_LAUNCH_CHECK();
  return PTTS_OK;
}

// ParlerTTSLogitsProcessor.__call__ on a full history (the HF-loop entry point), one CTA per batch row.
__global__ void logits_processor_kernel(const int64_t* ids, int seq, int64_t ld_ids, float* scores, int V, int64_t eos, int K, int64_t* first_unf) {
  __shared__ int cnt[32];
  const int b = blockIdx.x;
  if (threadIdx.x < 32) cnt[threadIdx.x] = 0;
  __syncthreads();
  for (int k = 0; k < K; k++) {
    int c = 0;
    for (int t = threadIdx.x; t < seq; t += blockDim.x) c += ids[(size_t)(b * K + k) * ld_ids + t] == eos;
    if (c) atomicAdd(&cnt[k], c);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t fu = first_unf[b];
    const int kk = (int)(fu - (int64_t)b * K);
    if (kk >= 0 && kk < K && cnt[kk] > 0 && fu < (int64_t)b * K + K - 1) fu++;
    first_unf[b] = fu;
    for (int k = 0; k < K; k++)
      if ((int64_t)b * K + k > fu) scores[(size_t)(b * K + k) * V + eos] = -INFINITY;
  }
}
int launch_logits_processor(const int64_t* ids, int BK, int seq, int64_t ld_ids, float* scores, int V, int64_t eos, int K, int64_t* first_unf, cudaStream_t st) {
  PTTS_REQUIRE(K <= 32 && BK % K == 0, "logits_processor: bad num_codebooks %d for %d rows", K, BK);
  PTTS_REQUIRE(eos >= 0 && eos < V, "`eos_token_id` has to be in [0, vocab), got %lld", (long long)eos);
  logits_processor_kernel<<<BK / K, 128, 0, st>>>(ids, seq, ld_ids, scores, V, eos, K, first_unf);
  PTTS_LAUNCH_CHECK();
  return PTTS_OK;
}

}  // namespace ptts
