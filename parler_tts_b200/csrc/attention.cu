// attention.cu -- self- and cross-attention over the KV cache for q_len new positions.
//
// Replaces ParlerTTSSdpaAttention.forward after the projections (modeling_parler_tts.py:858-914):
// rotary embedding of q/k (:858-859, :880-882, quirks Q2/Q3/Q7), KV-cache append (:887-889, a
// pre-allocated [B][nkv][Tmax][64] cache instead of DynamicCache's torch.cat), repeat_kv (:896-897,
// done by indexing), the 4-D additive masks (:1546-1562, :1658-1736 -> here a per-key exclude flag)
// and F.scaled_dot_product_attention (:906-914).
//
// Roofline: HBM.  Algorithmic bytes per (batch row, kv head) = 2 * T * 64 * sizeof(T) (K and V read
// once) -- SURVEY 8(d)'s kv_tok term.  One CTA per (kv head, batch row); 8 lanes cover one 128 B key
// row with 16 B loads (fully coalesced), 16 keys in flight per CTA iteration, fp32 softmax,
// probabilities rounded to the model dtype before P.V exactly like the flash kernels torch calls.
#include "common.cuh"
#include "kernels.h"
#include "attn_core.cuh"

namespace ptts {

template <typename T>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(AttnArgs p) {
  extern __shared__ __align__(16) float sm[];
  pdl_launch_dependents();
  pdl_wait();
  if (p.ctrl != nullptr && p.ctrl->active == 0) return;
  attention_item<T>(p, blockIdx.y, blockIdx.x, sm, threadIdx.x, [] { __syncthreads(); });
}

int launch_attention(const AttnArgs& a, int dtype, cudaStream_t st, bool pdl) {
  const int kv_capacity = a.kv_capacity;
  PTTS_REQUIRE(a.B > 0 && a.nkv > 0 && a.q_len > 0, "attention: empty problem");
  const size_t smem = (size_t)(HD + ATT_WARPS * HD + 8 + kv_capacity) * sizeof(float);
  PTTS_REQUIRE(smem <= 200 * 1024, "attention: kv length %d needs %zu B of shared memory (> 200 KB)", kv_capacity, smem);
  static bool attr_done = false;
  if (!attr_done) {
    PTTS_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    PTTS_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(a.nkv, a.B);
  cfg.blockDim = dim3(ATT_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  if (dtype == PTTS_BF16)
    PTTS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attention_kernel<bf16>, a));
  else
    PTTS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, attention_kernel<float>, a));
  return PTTS_OK;
}

}  // namespace ptts
