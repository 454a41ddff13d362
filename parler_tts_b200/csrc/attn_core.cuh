// attn_core.cuh -- one (batch row, kv head) attention item processed by a 128-thread group.
// Shared by attention_kernel (one CTA per item) and the fused step kernel (two items per CTA, named barriers).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace ptts {

constexpr int ATT_THREADS = 128;
constexpr int ATT_WARPS = 4;
constexpr int HD = PTTS_HEAD_DIM;

template <typename Sync>
__device__ __forceinline__ float block_reduce(float v, float* sh, bool is_max, int tid, Sync sync) {
  const int warp = tid >> 5, lane = tid & 31;
  v = is_max ? warp_max(v) : warp_sum(v);
  sync();
  if (lane == 0) sh[warp] = v;
  sync();
  float r = sh[0];
#pragma unroll
  for (int w = 1; w < ATT_WARPS; w++) r = is_max ? fmaxf(r, sh[w]) : r + sh[w];
  return r;
}

// RoPE on one 64-vector held as "element d per thread": out = x*cos + rotate_half(x)*sin with every
// intermediate rounded to the model dtype (apply_rotary_pos_emb, :416-436).
template <typename T>
__device__ __forceinline__ float rope_elem(float x, float x_pair, int d, const T* cos_row, const T* sin_row) {
  const float c = DT<T>::to_f(cos_row[d]), s = DT<T>::to_f(sin_row[d]);
  const float rot = (d < HD / 2) ? -x_pair : x_pair;
  return DT<T>::rnd(DT<T>::rnd(x * c) + DT<T>::rnd(rot * s));
}


// sm: [64] q | [4][64] partials | [8] scratch | [kv capacity] scores.  `tid` in [0,128); sync() = barrier of the group.
template <typename T, typename Sync>
__device__ __forceinline__ void attention_item(const AttnArgs& p, int b, int kvh, float* sm, int tid, Sync sync) {
  const T* __restrict__ rope_cos = reinterpret_cast<const T*>(p.rope_cos);
  const T* __restrict__ rope_sin = reinterpret_cast<const T*>(p.rope_sin);
  float* qs = sm;
  float* red = sm + HD;
  float* sh = red + ATT_WARPS * HD;
  float* sc = sh + 8;
  const int warp = tid >> 5, lane = tid & 31;
  const int rep = p.nh / p.nkv;
  int past = p.past_len;
  if (p.past_from_ctrl) past = p.prefix + p.ctrl->cur_len - 1;

  T* kc = reinterpret_cast<T*>(p.kcache) + (size_t)b * p.kv_b_stride + (size_t)kvh * p.kv_h_stride;
  T* vc = reinterpret_cast<T*>(p.vcache) + (size_t)b * p.kv_b_stride + (size_t)kvh * p.kv_h_stride;

  // ---- phase A: append the new K/V rows (self-attention only) ----
  if (!p.cross) {
    const int d = tid & 63;
    const bool is_v = tid >= 64;
    for (int j = 0; j < p.q_len; j++) {
      const size_t r = (size_t)b * p.q_len + j;
      const int pos = past + j;
      if (!is_v) {
        const T* src = reinterpret_cast<const T*>(p.knew) + r * p.ldkv + p.k_col0 + kvh * HD;
        float x = DT<T>::to_f(src[d]);
        if (p.rope) {
          const float xp = DT<T>::to_f(src[d < HD / 2 ? d + HD / 2 : d - HD / 2]);
          x = rope_elem<T>(x, xp, d, rope_cos + (size_t)pos * HD, rope_sin + (size_t)pos * HD);
        }
        kc[(size_t)pos * p.kv_t_stride + d] = DT<T>::from_f(x);
      } else {
        const T* src = reinterpret_cast<const T*>(p.vnew) + r * p.ldkv + p.v_col0 + kvh * HD;
        vc[(size_t)pos * p.kv_t_stride + d] = src[d];
      }
    }
    sync();  // this CTA is the only reader of the rows it just wrote
  }

  const int grp = lane >> 3;           // key group inside the warp (4 keys per warp per iteration)
  const int d0 = (lane & 7) * 8;       // this lane's 8 dims
  const int kslot = warp * 4 + grp;    // 0..15
  const int* km = p.key_mask ? p.key_mask + (size_t)b * p.mask_ld : nullptr;

  for (int j = 0; j < p.q_len; j++) {
    const size_t r = (size_t)b * p.q_len + j;
    const int pos = past + j;
    const int T_keys = p.cross ? p.kv_len : pos + 1;
    for (int rr = 0; rr < rep; rr++) {
      const int h = kvh * rep + rr;
      sync();
      if (tid < HD) {
        const T* src = reinterpret_cast<const T*>(p.q) + r * p.ldq + p.q_col0 + h * HD;
        float x = DT<T>::to_f(src[tid]);
        if (p.rope) {
          const float xp = DT<T>::to_f(src[tid < HD / 2 ? tid + HD / 2 : tid - HD / 2]);
          x = rope_elem<T>(x, xp, tid, rope_cos + (size_t)pos * HD, rope_sin + (size_t)pos * HD);
        }
        qs[tid] = x;
      }
      sync();
      float qv[8];
#pragma unroll
      for (int e = 0; e < 8; e++) qv[e] = qs[d0 + e];

      // ---- scores ----
      float lmax = -INFINITY;
      for (int t0 = 0; t0 < T_keys; t0 += 64) {
        float kv[4][8];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int t = t0 + u * 16 + kslot;
          ok[u] = t < T_keys;
          if (ok[u]) load8(kc + (size_t)t * p.kv_t_stride + d0, kv[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int t = t0 + u * 16 + kslot;
          float s = 0.f;
          if (ok[u]) {
#pragma unroll
            for (int e = 0; e < 8; e++) s = fmaf(qv[e], kv[u][e], s);
          }
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          if (ok[u] && (lane & 7) == 0) {
            s *= p.scale;
            if (km != nullptr && t < p.mask_len && km[t] == 0) s = -INFINITY;
            sc[t] = s;
            lmax = fmaxf(lmax, s);
          }
        }
      }
      const float m = block_reduce(lmax, sh, true, tid, sync);  // (contains the barrier that publishes sc[])
      // ---- softmax numerators; fully-masked rows degrade to uniform attention (never consumed) ----
      float lsum = 0.f;
      for (int t = tid; t < T_keys; t += ATT_THREADS) {
        const float e = (m == -INFINITY) ? 1.0f : expf(sc[t] - m);
        sc[t] = e;
        lsum += e;
      }
      const float l = block_reduce(lsum, sh, false, tid, sync);
      // ---- P.V ----
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; e++) acc[e] = 0.f;
      for (int t0 = 0; t0 < T_keys; t0 += 64) {
        float vv[4][8];
        float pw[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int t = t0 + u * 16 + kslot;
          pw[u] = 0.f;
          if (t < T_keys) {
            load8(vc + (size_t)t * p.kv_t_stride + d0, vv[u]);
            pw[u] = DT<T>::rnd(sc[t]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; e++) vv[u][e] = 0.f;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
          for (int e = 0; e < 8; e++) acc[e] = fmaf(pw[u], vv[u][e], acc[e]);
      }
#pragma unroll
      for (int e = 0; e < 8; e++) {
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
        acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
      }
      if (lane < 8) {
#pragma unroll
        for (int e = 0; e < 8; e++) red[warp * HD + d0 + e] = acc[e];
      }
      sync();
      if (tid < HD) {
        float o = red[tid] + red[HD + tid] + red[2 * HD + tid] + red[3 * HD + tid];
        o = o / l;
        reinterpret_cast<T*>(p.out)[r * p.ldo + h * HD + tid] = DT<T>::from_f(o);
      }
    }
  }
}

}  // namespace ptts

namespace ptts {

// ---- decode fast path (q_len == 1): ONE WARP per (batch row, kv head) item, TMA-staged K/V tiles ---------
// The cached K/V rows of the item stream through a per-warp 2-stage shared-memory ring filled by the TMA
// engine (cp.async.bulk + mbarrier; CH keys = CH*128 B of K and of V per stage), so HBM/L2 latency is hidden
// without holding loads in registers.  Scores, online softmax (running max / sum, fp32) and P.V are fused in
// one sweep: 8 lanes x 16 B cover one key row (conflict-free LDS.128), 4 keys per warp instruction; the
// probabilities are rounded to the model dtype before P.V like torch's flash kernels.  The step's own K/V row
// is taken from shared memory (it is also written to the cache for later steps), so no global read-after-write.
// No block-level barrier anywhere: used by attention_decode_kernel (8 items per CTA) and by the fused step
// kernel -- both run exactly this code, hence bit-identical results.
__device__ __forceinline__ uint32_t att_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <typename T> struct AttChunk { static constexpr int CH = 32; };
template <> struct AttChunk<float> { static constexpr int CH = 16; };

// bytes of shared memory one warp needs (CH = keys per ring stage)
template <typename T, int CH = AttChunk<T>::CH>
__host__ __device__ constexpr int attn_decode_smem_per_warp() {
  return 2 * 2 * CH * HD * (int)sizeof(T) + (3 * HD) * (int)sizeof(float);
}

template <typename T, int CH = AttChunk<T>::CH>
__device__ __forceinline__ void attention_decode_item_warp(const AttnArgs& p, int b, int kvh, int pos, unsigned char* sm_warp, uint64_t* bars,
                                                           int lane, uint32_t& parity, int part = 0, int nparts = 1, float* xch = nullptr,
                                                           int pair_bar = 0) {
  // part / nparts: the item's cached keys are split between `nparts` warps (chunk c belongs to warp c % nparts);
  // partial (max, sum, accumulator) triples are merged through `xch` with a 64-thread named barrier `pair_bar`.
  constexpr int STAGE_ELEMS = CH * HD;  // per K (or V) stage
  T* kst = reinterpret_cast<T*>(sm_warp);                   // [2][CH][64]
  T* vst = kst + 2 * STAGE_ELEMS;                            // [2][CH][64]
  float* qs = reinterpret_cast<float*>(vst + 2 * STAGE_ELEMS);  // [64] query
  float* kn = qs + HD;                                       // [64] this step's key   (self only)
  float* vn = kn + HD;                                       // [64] this step's value (self only)
  // bars[2]: this warp's mbarriers, initialised ONCE per kernel in memory that is never aliased (re-initialising
  // a live mbarrier is undefined behaviour); `parity` carries their phase across items and phases.
  const T* __restrict__ rope_cos = reinterpret_cast<const T*>(p.rope_cos) + (size_t)pos * HD;
  const T* __restrict__ rope_sin = reinterpret_cast<const T*>(p.rope_sin) + (size_t)pos * HD;
  const int rep = p.nh / p.nkv;
  T* kc = reinterpret_cast<T*>(p.kcache) + (size_t)b * p.kv_b_stride + (size_t)kvh * p.kv_h_stride;
  T* vc = reinterpret_cast<T*>(p.vcache) + (size_t)b * p.kv_b_stride + (size_t)kvh * p.kv_h_stride;
  const int lo = lane, hi = lane + HD / 2;
  const int n_cached = p.cross ? p.kv_len : pos;  // keys that come from the cache
  const int n_chunks_all = (n_cached + CH - 1) / CH;
  const int n_chunks = (n_chunks_all > part) ? (n_chunks_all - part + nparts - 1) / nparts : 0;  // chunks of THIS warp

  auto issue = [&](int i) {  // TMA: this warp's i-th chunk -> stage i&1
    const int st = i & 1;
    const int t0 = (part + nparts * i) * CH;
    const int n = (n_cached - t0 < CH) ? (n_cached - t0) : CH;
    const uint32_t bar = att_smem_u32(&bars[st]);
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(2 * n * HD * sizeof(T))) : "memory");
    __syncwarp();
    if (lane < 2) {  // one bulk copy for the K rows, one for the V rows (rows of an item are contiguous)
      const T* src = (lane == 0 ? kc : vc) + (size_t)t0 * HD;
      T* dst = (lane == 0 ? kst : vst) + st * STAGE_ELEMS;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(att_smem_u32(dst)), "l"(src), "r"((uint32_t)(n * HD * sizeof(T))), "r"(bar) : "memory");
    }
  };
  auto wait_stage = [&](int st) {
    const uint32_t bar = att_smem_u32(&bars[st]);
    const uint32_t par = (parity >> st) & 1u;
    uint32_t ok, spins = 0;
    do {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
      if (!ok && ++spins > (1u << 16)) { if (lane == 0) printf("ptts: attention KV mbarrier timeout (cta %d b %d kvh %d stage %d cross %d n_cached %d chunks %d par %u)\n", (int)blockIdx.x, b, kvh, st, p.cross, n_cached, n_chunks, parity); __trap(); }
    } while (!ok);
    parity ^= (1u << st);
  };

  // all lanes are past their shared-memory reads of the previous item: the ring may be refilled
  __syncwarp();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (n_chunks > 0) issue(0);
  if (n_chunks > 1) issue(1);

  // the first query head's elements are requested before the K/V append below (its stores would otherwise order the loads
  // behind a second L2 round trip)
  const T* q_first = reinterpret_cast<const T*>(p.q) + (size_t)b * p.ldq + p.q_col0 + (size_t)(kvh * rep) * HD;
  const float qf0 = DT<T>::to_f(q_first[lo]), qf1 = DT<T>::to_f(q_first[hi]);
  if (!p.cross && part == 0) {  // this step's K (rotary applied) and V: to the cache (for later steps) and to shared memory (for now)
    const T* ksrc = reinterpret_cast<const T*>(p.knew) + (size_t)b * p.ldkv + p.k_col0 + kvh * HD;
    const T* vsrc = reinterpret_cast<const T*>(p.vnew) + (size_t)b * p.ldkv + p.v_col0 + kvh * HD;
    float x0 = DT<T>::to_f(ksrc[lo]), x1 = DT<T>::to_f(ksrc[hi]);
    if (p.rope) {
      const float y0 = rope_elem<T>(x0, x1, lo, rope_cos, rope_sin);
      const float y1 = rope_elem<T>(x1, x0, hi, rope_cos, rope_sin);
      x0 = y0; x1 = y1;
    }
    const T v0 = vsrc[lo], v1 = vsrc[hi];
    kc[(size_t)pos * p.kv_t_stride + lo] = DT<T>::from_f(x0);
    kc[(size_t)pos * p.kv_t_stride + hi] = DT<T>::from_f(x1);
    vc[(size_t)pos * p.kv_t_stride + lo] = v0;
    vc[(size_t)pos * p.kv_t_stride + hi] = v1;
    kn[lo] = DT<T>::rnd(x0); kn[hi] = DT<T>::rnd(x1);
    vn[lo] = DT<T>::to_f(v0); vn[hi] = DT<T>::to_f(v1);
  }
  const int grp = lane >> 3, d0 = (lane & 7) * 8;
  const int* km = p.key_mask ? p.key_mask + (size_t)b * p.mask_ld : nullptr;

  for (int rr = 0; rr < rep; rr++) {
    const int h = kvh * rep + rr;
    {
      const T* qsrc = reinterpret_cast<const T*>(p.q) + (size_t)b * p.ldq + p.q_col0 + h * HD;
      float x0 = (rr == 0) ? qf0 : DT<T>::to_f(qsrc[lo]), x1 = (rr == 0) ? qf1 : DT<T>::to_f(qsrc[hi]);
      if (p.rope) {
        const float y0 = rope_elem<T>(x0, x1, lo, rope_cos, rope_sin);
        const float y1 = rope_elem<T>(x1, x0, hi, rope_cos, rope_sin);
        x0 = y0; x1 = y1;
      }
      __syncwarp();
      qs[lo] = x0; qs[hi] = x1;
      __syncwarp();
    }
    float qv[8];
#pragma unroll
    for (int e = 0; e < 8; e++) qv[e] = qs[d0 + e];
    float m_run = -INFINITY, l_run = 0.f;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; e++) acc[e] = 0.f;

    // One block of n <= CH keys whose K/V rows sit in shared memory at kb/vb (row stride 64); bit r of `mword` = key r is
    // attendable.  Branch-free: rows >= n are clamped to row n-1 (valid data) and get probability 0 through a -inf score,
    // so the 8 per-key chains of a lane group are independent and interleave (a divergent `if` per key serialised them).
    auto process = [&](const T* kb, const T* vb, uint32_t mword, int n) {
      constexpr int PER = CH / 4;  // keys per 8-lane group
      float sloc[PER];
#pragma unroll
      for (int u = 0; u < PER; u++) {
        const int r = u * 4 + grp, rc = r < n ? r : n - 1;
        float kf[8];
        load8(kb + rc * HD + d0, kf);
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 8; e++) s = fmaf(qv[e], kf[e], s);
        sloc[u] = s;
      }
#pragma unroll
      for (int u = 0; u < PER; u++) sloc[u] += __shfl_xor_sync(0xffffffffu, sloc[u], 1);
#pragma unroll
      for (int u = 0; u < PER; u++) sloc[u] += __shfl_xor_sync(0xffffffffu, sloc[u], 2);
#pragma unroll
      for (int u = 0; u < PER; u++) sloc[u] += __shfl_xor_sync(0xffffffffu, sloc[u], 4);
      float cmax = -INFINITY;
#pragma unroll
      for (int u = 0; u < PER; u++) {
        const int r = u * 4 + grp;
        sloc[u] = (r < n && ((mword >> r) & 1u)) ? sloc[u] * p.scale : -INFINITY;
        cmax = fmaxf(cmax, sloc[u]);
      }
      cmax = warp_max(cmax);
      const float m_new = fmaxf(m_run, cmax);
      if (m_new == -INFINITY) return;  // every key so far is masked: nothing to accumulate (warp-uniform)
      const float corr = (m_run == -INFINITY) ? 0.f : expf(m_run - m_new);
      l_run *= corr;
#pragma unroll
      for (int e = 0; e < 8; e++) acc[e] *= corr;
      float lsum = 0.f;
#pragma unroll
      for (int u = 0; u < PER; u++) {
        const int r = u * 4 + grp, rc = r < n ? r : n - 1;
        const float pe = expf(sloc[u] - m_new);  // exactly 0 for a masked / out-of-range key
        lsum += pe;  // identical on the 8 lanes of the group; counted once below
        const float pw = DT<T>::rnd(pe);
        float vf[8];
        load8(vb + rc * HD + d0, vf);
#pragma unroll
        for (int e = 0; e < 8; e++) acc[e] = fmaf(pw, vf[e], acc[e]);
      }
      // sum of probabilities over the 4 groups (each group's 8 lanes hold the same value)
      lsum += __shfl_xor_sync(0xffffffffu, lsum, 8);
      lsum += __shfl_xor_sync(0xffffffffu, lsum, 16);
      l_run += lsum;
      m_run = m_new;
    };
    // attendable-key bits of the chunk starting at key t0 (the loads are issued before the wait on the chunk's TMA stage)
    auto mask_bits = [&](int t0) -> int {
      const int t = t0 + lane;
      return (km != nullptr && lane < CH && t < p.mask_len && t < n_cached) ? km[t] : 1;
    };

    if (rr > 0) {  // GQA: further query heads re-stream the same cache rows
      __syncwarp();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (n_chunks > 0) issue(0);
      if (n_chunks > 1) issue(1);
    }
    for (int c = 0; c < n_chunks; c++) {
      const int st = c & 1;
      const int t0 = (part + nparts * c) * CH;
      const int n = (n_cached - t0 < CH) ? (n_cached - t0) : CH;
      const int mk = mask_bits(t0);
      wait_stage(st);
      process(kst + st * STAGE_ELEMS, vst + st * STAGE_ELEMS, __ballot_sync(0xffffffffu, mk != 0), n);
      if (c + 2 < n_chunks) {
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(c + 2);
      }
    }
    if (!p.cross && part == 0) {  // the step's own key (position `pos`), held in shared memory as fp32
      __syncwarp();
      float s = 0.f;
      if (grp == 0) {
#pragma unroll
        for (int e = 0; e < 8; e++) s = fmaf(qv[e], kn[d0 + e], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s = __shfl_sync(0xffffffffu, s, 0) * p.scale;
      if (km != nullptr && pos < p.mask_len && km[pos] == 0) s = -INFINITY;
      const float m_new = fmaxf(m_run, s);
      if (m_new != -INFINITY) {
        const float corr = (m_run == -INFINITY) ? 0.f : expf(m_run - m_new);
        const float pe = (s == -INFINITY) ? 0.f : expf(s - m_new);
        const float pw = DT<T>::rnd(pe);
        l_run = l_run * corr + pe;
#pragma unroll
        for (int e = 0; e < 8; e++) acc[e] = acc[e] * corr + ((grp == 0) ? pw * vn[d0 + e] : 0.f);
        m_run = m_new;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; e++) {
      acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
      acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
    }
    if (nparts == 2) {  // merge the two warps' partial softmax states (fixed order: part 0 then part 1)
      if (part == 1) {
        if (lane < 8) {
#pragma unroll
          for (int e = 0; e < 8; e++) xch[d0 + e] = acc[e];
        }
        if (lane == 0) { xch[HD] = m_run; xch[HD + 1] = l_run; }
      }
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      if (part == 0) {
        const float m1 = xch[HD], l1 = xch[HD + 1];
        const float mm = fmaxf(m_run, m1);
        const float c0 = (m_run == -INFINITY) ? 0.f : expf(m_run - mm);
        const float c1 = (m1 == -INFINITY) ? 0.f : expf(m1 - mm);
        l_run = l_run * c0 + l1 * c1;
        if (lane < 8) {
#pragma unroll
          for (int e = 0; e < 8; e++) acc[e] = acc[e] * c0 + xch[d0 + e] * c1;
        }
      }
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    }
    if (lane < 8 && part == 0) {
      float o[8];
      const float inv = (l_run > 0.f) ? 1.0f / l_run : 0.f;  // fully masked row -> zeros (never consumed)
#pragma unroll
      for (int e = 0; e < 8; e++) o[e] = acc[e] * inv;
      store8(reinterpret_cast<T*>(p.out) + (size_t)b * p.ldo + h * HD + d0, o);
    }
  }
}

// per-warp mbarrier setup (once per kernel); bars = this warp's two mbarriers
__device__ __forceinline__ void attention_decode_init_warp(uint64_t* bars, int lane) {
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(att_smem_u32(&bars[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(att_smem_u32(&bars[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
}

}  // namespace ptts
