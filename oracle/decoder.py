"""CPU restatement of the KV-cached ParlerTTS decoder step (test infrastructure only).

Follows, op for op and in the reference's rounding order (every op output is
rounded to the model dtype, as torch does for bf16):
  ParlerTTSDecoder.forward            parler_tts/modeling_parler_tts.py:1392-1655
  _update_causal_mask                 :1658-1736
  ParlerTTSDecoderLayer.forward       :983-1074
  ParlerTTSSdpaAttention.forward      :819-930   (q unscaled, SDPA scales: quirk Q1)
  rotary embedding / rotate_half      :373-436   (cos/sin fp32 -> model dtype: Q3; cross-attn q rotated: Q2)
  sinusoidal positions                :327-369
  K LM heads                          :1917-1920, reshape :1960
The KV cache is a pair of pre-allocated tensors per layer (what DynamicCache's
torch.cat growth produces, without the O(T) copy).
"""
from __future__ import annotations
import math
import torch
import torch.nn.functional as F

from .config import Cfg

ACT = {"gelu": F.gelu, "relu": F.relu, "silu": F.silu, "gelu_new": lambda x: F.gelu(x, approximate="tanh")}


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


class OracleDecoder:
    def __init__(self, cfg: Cfg, weights: dict[str, torch.Tensor], dtype=torch.float32):
        self.cfg, self.dtype = cfg, dtype
        self.w = {k: v.to(dtype) for k, v in weights.items()}
        self.H = cfg.hidden_size
        self.nh = cfg.num_attention_heads
        self.hd = self.H // self.nh
        self.nkv = cfg.num_key_value_heads
        self.nckv = cfg.num_cross_attention_key_value_heads
        self.K = cfg.num_codebooks
        self.L = cfg.num_hidden_layers
        self.eps = cfg.get("layer_norm_eps", 1e-5)
        self.act = ACT[cfg.activation_function]
        if cfg.rope_embeddings:
            # ParlerTTSRotaryEmbedding.__init__ (:380)
            self.inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, self.hd, 2, dtype=torch.int64).float() / self.hd))
        self.reset()

    # ---- cache -------------------------------------------------------------------------------
    def reset(self):
        self.k_cache, self.v_cache = [None] * self.L, [None] * self.L
        self.ck, self.cv = [None] * self.L, [None] * self.L
        self.cur_len = 0
        self.enc_mask4d = None
        self.prompt_mask = None

    def _p(self, name):
        return self.w["decoder.model.decoder." + name]

    # ---- pieces ------------------------------------------------------------------------------
    def embed_ids(self, ids_bkq: torch.Tensor) -> torch.Tensor:
        """sum_k embed_k(ids[:, k]) with Python-sum rounding order (:1433)."""
        return sum([F.embedding(ids_bkq[:, k], self._p(f"embed_tokens.{k}.weight")) for k in range(self.K)])

    def _rope(self, positions: torch.Tensor):
        # ParlerTTSRotaryEmbedding.forward (:394-406) then cast (:1534)
        freqs = positions.float()[:, None] * self.inv_freq[None, :]
        emb = torch.cat((freqs, freqs), dim=-1)
        return emb.cos().to(self.dtype)[None], emb.sin().to(self.dtype)[None]  # [1, q, hd]

    def _apply_rope(self, x, cos, sin):
        cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
        return (x * cos) + (rotate_half(x) * sin)

    def _self_mask(self, B, q, past, device_like):
        """4-D additive mask as _update_causal_mask builds it (:1692-1724); None when no padding mask (Q9)."""
        if self.prompt_mask is None:
            return None
        total = past + q
        am = torch.cat([self.prompt_mask, torch.ones(B, total - self.prompt_mask.shape[1], dtype=self.prompt_mask.dtype)], dim=1)
        mn = torch.finfo(self.dtype).min
        m = torch.full((q, total), mn, dtype=self.dtype)
        if q != 1:
            m = torch.triu(m, diagonal=1)
        cache_position = torch.arange(past, past + q)
        m = m * (torch.arange(total) > cache_position.reshape(-1, 1))
        m = m[None, None].expand(B, 1, -1, -1).clone()
        pad = (m + am[:, None, None, :].to(self.dtype)) == 0
        m = m.masked_fill(pad, mn)
        return m

    def _attn(self, layer, x, cross, cos, sin, mask, past, enc=None):
        p = f"layers.{layer}." + ("encoder_attn." if cross else "self_attn.")
        B, q, _ = x.shape
        nkv = self.nckv if cross else self.nkv
        qs = F.linear(x, self._p(p + "q_proj.weight")).view(B, q, self.nh, self.hd).transpose(1, 2).contiguous()
        if self.cfg.rope_embeddings:
            qs = self._apply_rope(qs, cos, sin)
        if cross:
            if self.ck[layer] is None:
                self.ck[layer] = F.linear(enc, self._p(p + "k_proj.weight")).view(B, -1, nkv, self.hd).transpose(1, 2).contiguous()
                self.cv[layer] = F.linear(enc, self._p(p + "v_proj.weight")).view(B, -1, nkv, self.hd).transpose(1, 2).contiguous()
            ks, vs = self.ck[layer], self.cv[layer]
        else:
            kn = F.linear(x, self._p(p + "k_proj.weight")).view(B, q, nkv, self.hd).transpose(1, 2).contiguous()
            vn = F.linear(x, self._p(p + "v_proj.weight")).view(B, q, nkv, self.hd).transpose(1, 2).contiguous()
            if self.cfg.rope_embeddings:
                kn = self._apply_rope(kn, cos, sin)
            if self.k_cache[layer] is None:
                self.k_cache[layer], self.v_cache[layer] = kn, vn
            else:
                self.k_cache[layer] = torch.cat([self.k_cache[layer], kn], dim=2)
                self.v_cache[layer] = torch.cat([self.v_cache[layer], vn], dim=2)
            ks, vs = self.k_cache[layer], self.v_cache[layer]
        rep = self.nh // nkv
        if rep > 1:
            ks = ks[:, :, None].expand(B, nkv, rep, ks.shape[2], self.hd).reshape(B, self.nh, -1, self.hd)
            vs = vs[:, :, None].expand(B, nkv, rep, vs.shape[2], self.hd).reshape(B, self.nh, -1, self.hd)
        is_causal = (not cross) and mask is None and q > 1
        o = F.scaled_dot_product_attention(qs, ks, vs, attn_mask=mask, is_causal=is_causal)
        o = o.transpose(1, 2).reshape(B, q, self.H)
        return F.linear(o, self._p(p + "out_proj.weight"))

    def _ln(self, x, name):
        return F.layer_norm(x, (self.H,), self._p(name + ".weight"), self._p(name + ".bias"), self.eps)

    # ---- forward over q new positions --------------------------------------------------------
    def forward(self, inputs_embeds: torch.Tensor, enc_hidden: torch.Tensor | None, return_hidden=False):
        """inputs_embeds [B, q, H] (prompt prefix already concatenated at step 0).  Returns logits [B*K, q, V]."""
        B, q, _ = inputs_embeds.shape
        past = self.cur_len
        positions = torch.arange(past, past + q)
        cos = sin = None
        if not self.cfg.rope_embeddings:
            pos = self._p("embed_positions.weights").index_select(0, positions)
            h = inputs_embeds + pos
        else:
            h = inputs_embeds
            cos, sin = self._rope(positions)
        mask = self._self_mask(B, q, past, h)
        enc_mask = None
        if self.enc_mask4d is not None:
            enc_mask = self.enc_mask4d.expand(B, 1, q, -1)
        for i in range(self.L):
            pre = f"layers.{i}."
            r = h
            h = self._ln(h, pre + "self_attn_layer_norm")
            h = r + self._attn(i, h, False, cos, sin, mask, past)
            r = h
            h = self._ln(h, pre + "encoder_attn_layer_norm")
            h = r + self._attn(i, h, True, cos, sin, enc_mask, past, enc=enc_hidden)
            r = h
            h = self._ln(h, pre + "final_layer_norm")
            h = self.act(F.linear(h, self._p(pre + "fc1.weight")))
            h = r + F.linear(h, self._p(pre + "fc2.weight"))
        h = self._ln(h, "layer_norm")
        self.cur_len = past + q
        logits = torch.stack([F.linear(h, self.w[f"decoder.lm_heads.{k}.weight"]) for k in range(self.K)], dim=1)
        logits = logits.reshape(-1, *logits.shape[2:])  # [B*K, q, V]
        return (logits, h) if return_hidden else logits

    # ---- generate()-style entry points -------------------------------------------------------
    def prefill(self, ids: torch.Tensor, enc_hidden: torch.Tensor, enc_mask: torch.Tensor | None,
                prompt_hidden: torch.Tensor | None, prompt_mask: torch.Tensor | None):
        """Step 0 of generate(): prompt prefix + decoder ids (delay mask already applied). ids [B*K, q]."""
        self.reset()
        B = enc_hidden.shape[0]
        enc_hidden = enc_hidden.to(self.dtype)
        if enc_mask is not None:
            # _prepare_4d_attention_mask_for_sdpa (:1553): 1 -> 0, 0 -> finfo.min ; all-ones mask -> None
            if bool((enc_mask == 1).all()):
                self.enc_mask4d = None
            else:
                inv = 1.0 - enc_mask[:, None, None, :].to(self.dtype)
                self.enc_mask4d = inv.masked_fill(inv.bool(), torch.finfo(self.dtype).min)
        self.prompt_mask = prompt_mask if prompt_hidden is not None else None
        self._enc = enc_hidden
        emb = self.embed_ids(ids.reshape(B, self.K, -1))
        if prompt_hidden is not None:
            emb = torch.cat([prompt_hidden.to(self.dtype), emb], dim=1)
        return self.forward(emb, enc_hidden)

    def step(self, ids: torch.Tensor):
        """ids [B*K, 1] (delay mask already applied) -> logits [B*K, 1, V]."""
        B = ids.shape[0] // self.K
        return self.forward(self.embed_ids(ids.reshape(B, self.K, -1)), self._enc)
