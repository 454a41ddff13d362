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


// Processes row `row` with one warp.  Returns (on lane 0) whether the row is still unfinished.
template <int ITEMS>
__device__ __forceinline__ int sample_row(const SampleArgs& p, const ptts_gen_params& g, const int64_t* __restrict__ forced,
                                          int row, int cur_len, int lane) {
  int still_unfinished = 0;
  {
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
      // eos_seen[r] = 1 + column of row r's first EOS (0 = none).  The reference counts EOS over input_ids, i.e. columns
      // < cur_len (logits_processors.py:46): an EOS written by another warp during THIS step (column cur_len) must not count,
      // so the test is on the column, not on a flag (rows of one batch item are sampled by different warps / CTAs).
      const int es = p.eos_seen[fu];
      if (es > 0 && es <= cur_len && fu < b * p.K + p.K - 1) fu++;
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
  }
  return still_unfinished;
}

}  // namespace ptts
