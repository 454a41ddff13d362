"""Shape/config records for the oracle (plain dicts -> attribute bags).

Shapes follow the reference init scripts:
  Mini : helpers/model_init_scripts/init_model_600M.py:27-44
  Large: helpers/model_init_scripts/init_large_model.py:25-43
Field names follow parler_tts/configuration_parler_tts.py:107-172.
"""
from __future__ import annotations
import copy


class Cfg(dict):
    __getattr__ = dict.__getitem__

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self, **kw):
        c = Cfg(copy.deepcopy(dict(self)))
        c.update(kw)
        return c


def decoder_cfg(**kw) -> Cfg:
    base = dict(
        vocab_size=1088, max_position_embeddings=4096, num_hidden_layers=24, ffn_dim=4096,
        num_attention_heads=16, num_key_value_heads=None, num_cross_attention_key_value_heads=None,
        hidden_size=1024, num_codebooks=9, pad_token_id=1024, eos_token_id=1024, bos_token_id=1025,
        activation_function="gelu", rope_embeddings=False, rope_theta=10000.0, layer_norm_eps=1e-5,
        codebook_size=1024, text_vocab_size=32128, initializer_factor=0.02,
    )
    base.update(kw)
    if base["num_key_value_heads"] is None:
        base["num_key_value_heads"] = base["num_attention_heads"]
    if base["num_cross_attention_key_value_heads"] is None:
        base["num_cross_attention_key_value_heads"] = base["num_key_value_heads"]
    return Cfg(base)


def mini_cfg(**kw) -> Cfg:
    return decoder_cfg(**kw)


def large_cfg(**kw) -> Cfg:
    d = dict(hidden_size=1536, num_hidden_layers=30, num_attention_heads=24, ffn_dim=6144)
    d.update(kw)
    return decoder_cfg(**d)


def tiny_cfg(**kw) -> Cfg:
    """Small shape the oracle finishes in milliseconds; head_dim stays 64 like Mini/Large."""
    d = dict(vocab_size=96, max_position_embeddings=128, num_hidden_layers=2, ffn_dim=256,
             num_attention_heads=2, hidden_size=128, num_codebooks=4, pad_token_id=64, eos_token_id=64,
             bos_token_id=65, codebook_size=64, text_vocab_size=100)
    d.update(kw)
    return decoder_cfg(**d)


def dac_cfg(**kw) -> Cfg:
    """DAC 44.1 kHz decoder shape (dac_wrapper/configuration_dac.py:12-17 + descript-audio-codec 44khz)."""
    base = dict(n_codebooks=9, codebook_size=1024, codebook_dim=8, hidden_size=1024,
                decoder_hidden_size=1536, upsampling_ratios=[8, 8, 4, 2], sampling_rate=44100)
    base.update(kw)
    return Cfg(base)


def tiny_dac_cfg(**kw) -> Cfg:
    d = dict(n_codebooks=4, codebook_size=64, codebook_dim=8, hidden_size=64, decoder_hidden_size=96,
             upsampling_ratios=[8, 8, 4, 2])
    d.update(kw)
    return dac_cfg(**d)
