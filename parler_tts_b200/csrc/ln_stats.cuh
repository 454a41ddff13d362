// ln_stats.cuh -- LayerNorm folded into the following linear layer (bf16 model dtype).
//
// Reference: y = Linear(LayerNorm(x)) (modeling_parler_tts.py:1020-1023, :1040-1041, :1059-1060, :1632 + :1917-1920).
// With LN(x)_k = (x_k - mu) * r * g_k + b_k:
//     y_n = r * ( sum_k x_k W'_nk  -  mu * c1_n ) + c2_n,   W' = bf16(g_k W_nk),  c1_n = sum_k W'_nk,  c2_n = sum_k b_k W_nk
// W', c1, c2 are produced ONCE at load (ptts_decoder_finalize); at run time the GEMM consumes the RAW activation tile and
// only the per-row (mu, r) are needed -- one vectorised pass over the staged tile instead of a three-pass in-place
// normalisation that every one of the 148 CTAs would repeat (measured 5 us per LN-fused GEMM phase).
// The statistics use a shifted single pass (shift = the row's first element) in fp32: mean = x0 + S1/n,
// var = S2/n - (S1/n)^2.  Numerics: the normalised activations are no longer rounded to bf16 before the GEMM and gamma is
// rounded into the weights instead; both effects are at the bf16 resolution of the reference's own arithmetic.
// The same function is used by linear_bf16_kernel (gemm.cu) and decode_step_kernel (step.cu): results are bit-identical.
#pragma once
#include "common.cuh"

namespace ptts {

__device__ __forceinline__ void unpack8_bf16(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; i++) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

// 256 threads (8 warps); warp w owns rows w, w+8, w+16, w+24 and walks them together.  stats: shared float[64] = (mean, rstd) per row.
__device__ __forceinline__ void tile_row_stats(const bf16* xs, int pitch, int Kc, int M, float eps, float* stats) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* row[4];
  float shift[4], s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    row[i] = xs + (size_t)(warp + 8 * i) * pitch;
    shift[i] = __bfloat162float(row[i][0]);
  }
  for (int c = lane * 8; c < Kc; c += 256) {
    uint4 u[4];  // register copies (4 LDS.128 in flight); binding a reference to shared memory would re-read 32-bit words
#pragma unroll
    for (int i = 0; i < 4; i++) u[i] = *reinterpret_cast<const uint4*>(row[i] + c);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float f[8];
      unpack8_bf16(u[i], f);
#pragma unroll
      for (int e = 0; e < 8; e++) f[e] -= shift[i];
      // fixed pairwise order (short dependency chains)
      const float a = ((f[0] + f[1]) + (f[2] + f[3])) + ((f[4] + f[5]) + (f[6] + f[7]));
      const float q = (fmaf(f[1], f[1], f[0] * f[0]) + fmaf(f[3], f[3], f[2] * f[2])) + (fmaf(f[5], f[5], f[4] * f[4]) + fmaf(f[7], f[7], f[6] * f[6]));
      s1[i] += a;
      s2[i] += q;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const float S1 = warp_sum(s1[i]) / (float)Kc, S2 = warp_sum(s2[i]) / (float)Kc;
    const int r = warp + 8 * i;
    if (lane == 0 && r < M) {
      stats[2 * r] = shift[i] + S1;
      stats[2 * r + 1] = rsqrtf(fmaxf(S2 - S1 * S1, 0.f) + eps);
    }
  }
}

}  // namespace ptts
