"""Shared builders for the parity tests: the product model fed with the oracle's synthetic weights."""
from __future__ import annotations
import numpy as np
import torch

from oracle.config import Cfg


def product_decoder_config(cfg: Cfg):
    from parler_tts_b200 import ParlerTTSDecoderConfig
    return ParlerTTSDecoderConfig(
        vocab_size=cfg.vocab_size, max_position_embeddings=cfg.max_position_embeddings,
        num_hidden_layers=cfg.num_hidden_layers, ffn_dim=cfg.ffn_dim, num_attention_heads=cfg.num_attention_heads,
        num_key_value_heads=cfg.num_key_value_heads, num_cross_attention_key_value_heads=cfg.num_cross_attention_key_value_heads,
        hidden_size=cfg.hidden_size, num_codebooks=cfg.num_codebooks, pad_token_id=cfg.pad_token_id,
        eos_token_id=cfg.eos_token_id, bos_token_id=cfg.bos_token_id, rope_embeddings=cfg.rope_embeddings,
        rope_theta=cfg.rope_theta, activation_function=cfg.activation_function)


def product_dac_config(dcfg: Cfg):
    from parler_tts_b200 import DACConfig
    return DACConfig(num_codebooks=dcfg.n_codebooks, codebook_size=dcfg.codebook_size, latent_dim=dcfg.hidden_size,
                     codebook_dim=dcfg.codebook_dim, decoder_dim=dcfg.decoder_hidden_size,
                     decoder_rates=tuple(dcfg.upsampling_ratios))


def build_product_model(cfg: Cfg, dcfg: Cfg, weights: dict, dac_weights: dict, dtype=torch.float32, device="cuda"):
    from parler_tts_b200 import ParlerTTSConfig, ParlerTTSForConditionalGeneration
    pc = ParlerTTSConfig(vocab_size=cfg.text_vocab_size, text_encoder={}, audio_encoder=product_dac_config(dcfg),
                         decoder=product_decoder_config(cfg))
    m = ParlerTTSForConditionalGeneration(pc, device=device, dtype=dtype)
    m.load_state_dict(weights, dac_state_dict=dac_weights)
    return m


def synth_inputs(cfg: Cfg, B: int, S: int, P: int, seed: int = 0, masks: bool = True):
    """SURVEY 8(d) synthetic inputs: left-padded description / prompt with matching masks."""
    g = torch.Generator().manual_seed(seed)
    enc = torch.randn(B, S, cfg.hidden_size, generator=g)
    prompt = torch.randn(B, P, cfg.hidden_size, generator=g) * 0.5 if P > 0 else None
    enc_mask = prompt_mask = None
    if masks:
        enc_mask = torch.ones(B, S, dtype=torch.long)
        lens = torch.randint(max(1, S // 2), S + 1, (B,), generator=g)
        for b in range(B):
            enc_mask[b, : S - int(lens[b])] = 0
        enc = enc * enc_mask[..., None]
        if P > 0:
            prompt_mask = torch.ones(B, P, dtype=torch.long)
            plens = torch.randint(max(1, P // 2), P + 1, (B,), generator=g)
            for b in range(B):
                prompt_mask[b, : P - int(plens[b])] = 0
    return enc, enc_mask, prompt, prompt_mask


def rms(a) -> float:
    a = np.asarray(a, dtype=np.float64)
    return float(np.sqrt((a ** 2).mean()))
