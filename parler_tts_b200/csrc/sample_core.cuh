// sample_core.cuh -- one row of logits -> next token (shared by sample_kernel and the fused step kernel).
#pragma once
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


// Processes row `row` with ONE CTA of SMP_THREADS threads (element i of the row lives on thread i % 256, slot i / 256): the
// bitwise top-k / top-p searches, the softmax sums and the inverse-CDF draw are CTA-wide reductions / scans with a fixed order
// (lanes by butterfly, then warps 0..7), so results are bit-reproducible and identical between sample_kernel and the fused step
// kernel (both call exactly this function).  One warp per row left 75 % of the fused kernel's warps idle for 33 us per token and
// held 72 live values per lane; one CTA per row needs ITEMS = ceil(V / 256) <= 9 values per thread.
// Every thread of the CTA must call it (it contains __syncthreads()).
constexpr int SMP_WARPS = 8;
constexpr int SMP_THREADS = SMP_WARPS * 32;

template <int ITEMS>
__device__ __forceinline__ int sample_row_cta(const SampleArgs& p, const ptts_gen_params& g, const int64_t* __restrict__ forced,
                                               int row, int cur_len) {
  __shared__ float s_f[2][SMP_WARPS];
  __shared__ int s_i[2][SMP_WARPS];
  __shared__ int s_z[2][SMP_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int buf = 0;  // reductions alternate between two scratch rows: a warp may enter the next reduction while others still read this one
  auto cta_sum_i = [&](int v) -> int {
    v = __reduce_add_sync(0xffffffffu, v);
    if (lane == 0) s_i[buf][warp] = v;
    __syncthreads();
    int r = 0;
#pragma unroll
    for (int w = 0; w < SMP_WARPS; w++) r += s_i[buf][w];
    buf ^= 1;
    return r;
  };
  auto cta_sum_f = [&](float v) -> float {
    v = warp_sum(v);
    if (lane == 0) s_f[buf][warp] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < SMP_WARPS; w++) r += s_f[buf][w];
    buf ^= 1;
    return r;
  };
  auto cta_max_f = [&](float v) -> float {
    v = warp_max(v);
    if (lane == 0) s_f[buf][warp] = v;
    __syncthreads();
    float r = s_f[buf][0];
#pragma unroll
    for (int w = 1; w < SMP_WARPS; w++) r = fmaxf(r, s_f[buf][w]);
    buf ^= 1;
    return r;
  };

  int still_unfinished = 0;
  const int b = row / p.K, k = row - b * p.K;
  float v[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; j++) {
    const int i = tid + SMP_THREADS * j;
    v[j] = (i < p.V) ? __ldcg(p.logits + (size_t)row * p.V + i) : -INFINITY;  // written by other CTAs of this launch: L2, not L1
  }
  bool mask_eos = false;
  // MinNewTokensLength: prompt_length_to_skip = 1 (the BOS column)
  if (cur_len - 1 < g.min_new_tokens) mask_eos = true;
  // ParlerTTSLogitsProcessor (stateful; state double-buffered on the column parity)
  {
    const int par = cur_len & 1;
    int fu = p.first_unf[par * p.B + b];
    // eos_seen[r] = 1 + column of row r's first EOS (0 = none).  The reference counts EOS over input_ids, i.e. columns
    // < cur_len (logits_processors.py:46): an EOS written by another CTA during THIS step (column cur_len) must not count,
    // so the test is on the column, not on a flag (rows of one batch item are sampled by different CTAs).
    const int es = __ldcg(p.eos_seen + fu);
    if (es > 0 && es <= cur_len && fu < b * p.K + p.K - 1) fu++;
    if (k == 0 && tid == 0) p.first_unf[(par ^ 1) * p.B + b] = fu;
    if (row > fu) mask_eos = true;
  }
#pragma unroll
  for (int j = 0; j < ITEMS; j++) {
    const int i = tid + SMP_THREADS * j;
    if (mask_eos && i == p.eos) v[j] = -INFINITY;
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
      for (int bit = 31; bit >= 0; bit--) {  // bitwise binary search of the k-th largest key
        const uint32_t cand = th | (1u << bit);
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < ITEMS; j++) cnt += (tid + SMP_THREADS * j < p.V && fkey(v[j]) >= cand) ? 1 : 0;
        if (cta_sum_i(cnt) >= kk) th = cand;
      }
#pragma unroll
      for (int j = 0; j < ITEMS; j++)
        if (fkey(v[j]) < th) v[j] = -INFINITY;  // scores < kth largest
    }
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) m = fmaxf(m, v[j]);
    m = cta_max_f(m);
    float e[ITEMS];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) { e[j] = expf(v[j] - m); s += e[j]; }
    s = cta_sum_f(s);
    if (g.top_p < 1.0f) {
      // remove tokens whose ascending cumulative probability is <= 1 - top_p (the max is always kept)
      const float thr = (1.0f - g.top_p) * s;
      uint32_t th = 0;
      for (int bit = 31; bit >= 0; bit--) {
        const uint32_t cand = th | (1u << bit);
        float c = 0.f;
#pragma unroll
        for (int j = 0; j < ITEMS; j++) c += (fkey(v[j]) <= cand) ? e[j] : 0.f;
        if (cta_sum_f(c) <= thr) th = cand;
      }
      const uint32_t kmax = fkey(m);
      s = 0.f;
#pragma unroll
      for (int j = 0; j < ITEMS; j++) {
        const uint32_t key = fkey(v[j]);
        if (key <= th && key != kmax) { v[j] = -INFINITY; e[j] = 0.f; }
        s += e[j];
      }
      s = cta_sum_f(s);
    }
    // inverse-CDF draw in index order: CTA-wide inclusive scan per slot j (elements 256 j .. 256 j + 255)
    const float target = philox_uniform(g.seed, (uint32_t)(row + g.row_base), (uint32_t)cur_len) * s;
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
      if (lane == 31) s_f[buf][warp] = x;  // this warp's total
      __syncthreads();
      float base = carry, tot = carry;
#pragma unroll
      for (int w = 0; w < SMP_WARPS; w++) {
        const float tw = s_f[buf][w];
        if (w < warp) base += tw;
        tot += tw;
      }
      const float cum = base + x;
      const unsigned hit = __ballot_sync(0xffffffffu, cum > target && e[j] > 0.f);
      const unsigned nz = __ballot_sync(0xffffffffu, e[j] > 0.f);
      if (lane == 0) {
        s_i[buf][warp] = hit ? SMP_THREADS * j + 32 * warp + (__ffs(hit) - 1) : 0x7fffffff;
        s_z[buf][warp] = nz ? SMP_THREADS * j + 32 * warp + (31 - __clz(nz)) : -1;
      }
      __syncthreads();
      int first = 0x7fffffff, lastz = -1;
#pragma unroll
      for (int w = 0; w < SMP_WARPS; w++) { first = min(first, s_i[buf][w]); lastz = max(lastz, s_z[buf][w]); }
      if (found < 0 && first != 0x7fffffff) found = first;
      if (lastz >= 0) last_nz = lastz;
      carry = tot;
      buf ^= 1;
    }
    tok = found >= 0 ? found : last_nz;
  } else {
    // argmax, smallest index on ties
    float m = -INFINITY;
    int mi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < ITEMS; j++) {
      const int i = tid + SMP_THREADS * j;
      if (i < p.V && (v[j] > m || (v[j] == m && i < mi))) { m = v[j]; mi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
      if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
    }
    if (lane == 0) { s_f[buf][warp] = m; s_i[buf][warp] = mi; }
    __syncthreads();
    m = s_f[buf][0]; mi = s_i[buf][0];
#pragma unroll
    for (int w = 1; w < SMP_WARPS; w++) {
      const float om = s_f[buf][w];
      const int oi = s_i[buf][w];
      if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
    }
    buf ^= 1;
    tok = mi;
  }
#pragma unroll
  for (int j = 0; j < ITEMS; j++) {
    const int i = tid + SMP_THREADS * j;
    if (i < p.V) p.scores[(size_t)row * p.V + i] = v[j];
  }
  if (tid == 0) {
    if (forced != nullptr) tok = (int)forced[row];
    const int unf = p.unfinished[row];
    if (!unf) tok = p.pad;  // next_tokens * unfinished + pad * (1 - unfinished)
    p.raw_ids[(size_t)row * p.raw_ld + cur_len] = tok;
    if (tok == p.eos && p.eos_seen[row] == 0) p.eos_seen[row] = cur_len + 1;
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
  __syncthreads();  // the scratch rows may be reused by the next row of this CTA
  return still_unfinished;
}

}  // namespace ptts
