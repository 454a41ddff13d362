"""Seeded synthetic weights in the reference's state-dict naming.

There is no network in the build/bench environment, so no published checkpoint
can be loaded; every parity and bench run uses weights generated here.
Names follow the reference modules:
  decoder stack  parler_tts/modeling_parler_tts.py:1353-1373 (embed_tokens.N, layers.N.*, layer_norm)
  layer          :945-981 (self_attn.{q,k,v,out}_proj, encoder_attn.*, fc1, fc2, three LayerNorms)
  lm heads       :1835-1840 (lm_heads.N, bias-free)
  composite      :2388-2395 (embed_prompts, optional enc_to_dec_proj)
Init follows :1093-1102 / :2427-2436 (normal(0, initializer_factor)); LayerNorm
weights are perturbed away from (1, 0) so a kernel that drops them is caught.
DAC names follow transformers.models.dac.DacModel's decoder/quantizer modules.
"""
from __future__ import annotations
import math
import torch

from .config import Cfg


def sinusoidal_table(num_embeddings: int, dim: int) -> torch.Tensor:
    """Restates ParlerTTSSinusoidalPositionalEmbedding.get_embedding (modeling_parler_tts.py:346-359)."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.int64).float() * -e)
    e = torch.arange(num_embeddings, dtype=torch.int64).float().unsqueeze(1) * e.unsqueeze(0)
    e = torch.cat([torch.cos(e), torch.sin(e)], dim=1).view(num_embeddings, -1)
    if dim % 2 == 1:
        e = torch.cat([e, torch.zeros(num_embeddings, 1)], dim=1)
    return e.float()


def make_decoder_weights(cfg: Cfg, seed: int = 0, std: float | None = None, head_std: float | None = None,
                         embed_std: float | None = None) -> dict[str, torch.Tensor]:
    """fp32 tensors keyed like ParlerTTSForConditionalGeneration.state_dict() (decoder part + embed_prompts)."""
    g = torch.Generator().manual_seed(seed)
    std = cfg.initializer_factor if std is None else std
    head_std = std if head_std is None else head_std
    embed_std = std if embed_std is None else embed_std
    H, F, V, K = cfg.hidden_size, cfg.ffn_dim, cfg.vocab_size, cfg.num_codebooks
    hd = H // cfg.num_attention_heads
    kvH = cfg.num_key_value_heads * hd
    ckvH = cfg.num_cross_attention_key_value_heads * hd

    def n(*shape, s=std):
        return torch.randn(*shape, generator=g) * s

    w: dict[str, torch.Tensor] = {}
    p = "decoder.model.decoder."
    for k in range(K):
        w[f"{p}embed_tokens.{k}.weight"] = n(V + 1, H, s=embed_std)
    if not cfg.rope_embeddings:
        w[f"{p}embed_positions.weights"] = sinusoidal_table(cfg.max_position_embeddings, H)
    for i in range(cfg.num_hidden_layers):
        q = f"{p}layers.{i}."
        w[q + "self_attn.k_proj.weight"] = n(kvH, H)
        w[q + "self_attn.v_proj.weight"] = n(kvH, H)
        w[q + "self_attn.q_proj.weight"] = n(H, H)
        w[q + "self_attn.out_proj.weight"] = n(H, H)
        w[q + "encoder_attn.k_proj.weight"] = n(ckvH, H)
        w[q + "encoder_attn.v_proj.weight"] = n(ckvH, H)
        w[q + "encoder_attn.q_proj.weight"] = n(H, H)
        w[q + "encoder_attn.out_proj.weight"] = n(H, H)
        w[q + "fc1.weight"] = n(F, H)
        w[q + "fc2.weight"] = n(H, F)
        for ln in ("self_attn_layer_norm", "encoder_attn_layer_norm", "final_layer_norm"):
            w[q + ln + ".weight"] = 1.0 + n(H, s=0.1)
            w[q + ln + ".bias"] = n(H, s=0.05)
    w[p + "layer_norm.weight"] = 1.0 + n(H, s=0.1)
    w[p + "layer_norm.bias"] = n(H, s=0.05)
    for k in range(K):
        w[f"decoder.lm_heads.{k}.weight"] = n(V, H, s=head_std)
    w["embed_prompts.weight"] = n(cfg.text_vocab_size, H, s=embed_std)
    return w


def make_dac_weights(cfg: Cfg, seed: int = 0) -> dict[str, torch.Tensor]:
    """Folded (weight-norm already applied) DAC quantizer+decoder weights, fp32.

    Keys follow transformers.models.dac.DacModel: quantizer.quantizers.N.{codebook,out_proj},
    decoder.conv1, decoder.block.N.{snake1.alpha,conv_t1,res_unitM.{snake1,conv1,snake2,conv2}},
    decoder.snake1.alpha, decoder.conv2.  Scales are chosen fan-in style so activations stay O(1)
    through the ~40-layer stack (a degenerate all-saturated tanh would hide errors).
    """
    g = torch.Generator().manual_seed(seed)

    def conv(co, ci, k):
        s = 1.0 / math.sqrt(ci * k)
        return torch.randn(co, ci, k, generator=g) * s, torch.randn(co, generator=g) * 0.02

    def alpha(c):
        return (1.0 + 0.3 * torch.randn(1, c, 1, generator=g)).abs() + 0.1

    w: dict[str, torch.Tensor] = {}
    for i in range(cfg.n_codebooks):
        w[f"quantizer.quantizers.{i}.codebook.weight"] = torch.randn(cfg.codebook_size, cfg.codebook_dim, generator=g)
        ww, bb = conv(cfg.hidden_size, cfg.codebook_dim, 1)
        w[f"quantizer.quantizers.{i}.out_proj.weight"] = ww / math.sqrt(cfg.n_codebooks)
        w[f"quantizer.quantizers.{i}.out_proj.bias"] = bb
    C = cfg.decoder_hidden_size
    w["decoder.conv1.weight"], w["decoder.conv1.bias"] = conv(C, cfg.hidden_size, 7)
    for bi, s in enumerate(cfg.upsampling_ratios):
        cin, cout = C // 2 ** bi, C // 2 ** (bi + 1)
        p = f"decoder.block.{bi}."
        w[p + "snake1.alpha"] = alpha(cin)
        # ConvTranspose1d weight is [in, out, k]; each output sample sees ~2 taps x cin
        w[p + "conv_t1.weight"] = torch.randn(cin, cout, 2 * s, generator=g) / math.sqrt(2 * cin)
        w[p + "conv_t1.bias"] = torch.randn(cout, generator=g) * 0.02
        for ri in (1, 2, 3):
            r = p + f"res_unit{ri}."
            w[r + "snake1.alpha"] = alpha(cout)
            ww, bb = conv(cout, cout, 7)
            w[r + "conv1.weight"], w[r + "conv1.bias"] = ww * 0.5, bb
            w[r + "snake2.alpha"] = alpha(cout)
            ww, bb = conv(cout, cout, 1)
            w[r + "conv2.weight"], w[r + "conv2.bias"] = ww * 0.5, bb
    cl = C // 2 ** len(cfg.upsampling_ratios)
    w["decoder.snake1.alpha"] = alpha(cl)
    w["decoder.conv2.weight"], w["decoder.conv2.bias"] = conv(1, cl, 7)
    return w
