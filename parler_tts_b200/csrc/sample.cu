// sample.cu -- logits -> next token, entirely on the device, plus the integer ops around it.
//
// One CTA per row (B*K rows; sample_core.cuh).  Replaces, per decode step:
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
#include "sample_core.cuh"

namespace ptts {

template <int ITEMS>
__global__ void __launch_bounds__(SMP_THREADS) sample_kernel(SampleArgs p, const int64_t* __restrict__ forced) {
  pdl_launch_dependents();
  pdl_wait();
  if (p.ctrl->active == 0) return;
  const int row = blockIdx.x;           // one CTA per (utterance, codebook) row
  const int cur_len = p.ctrl->cur_len;  // the new token becomes column `cur_len`
  const ptts_gen_params g = *p.gen;
  sample_row_cta<ITEMS>(p, g, forced, row, cur_len);
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
  cfg.gridDim = dim3(BK);
  cfg.blockDim = dim3(SMP_THREADS);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  const int items = (a.V + SMP_THREADS - 1) / SMP_THREADS;
  PTTS_REQUIRE(items <= 9, "sample: vocab_size %d > 2304 not supported", a.V);
  if (items <= 1) PTTS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, sample_kernel<1>, a, forced));
  else if (items <= 5) PTTS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, sample_kernel<5>, a, forced));
  else PTTS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, sample_kernel<9>, a, forced));
  return PTTS_OK;
}

__global__ void generate_begin_kernel(SampleArgs p) {
  const int BK = p.B * p.K;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    p.ctrl->cur_len = 1; p.ctrl->active = 1; p.ctrl->n_unfinished = 0; p.ctrl->done_blocks = 0; p.ctrl->steps_run = 0;
    p.ctrl->launch_gen = 0;
    for (int j = 0; j < 32; j++) p.ctrl->bar[j] = 0;
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
