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
