"""PyTorch restatement of the reference's FAST GPU path for the decode loop (comparator only -- test / bench infrastructure).

The reference's own recipe for fast generation is SDPA attention + a static KV cache + `torch.compile(mode="reduce-overhead")`
(reference INFERENCE.md:57-72; the cache is built at modeling_parler_tts.py:3254-3309 and consumed at :861-889).  The reference
package cannot be imported on the GPU box, so this module restates that path with stock torch ops, in the same
op order as oracle/decoder.py (which is pinned against the reference's own forward): pre-allocated K/V tensors written in place
with index_copy_ at `cache_position`, F.scaled_dot_product_attention over the whole static cache with an additive mask, the K LM
heads, and HF's top-k -> softmax -> multinomial sampling.  The decode step is then replayed from a CUDA graph -- what
"reduce-overhead" does -- and, when it works on the box, additionally passed through torch.compile.
bench.py reports its tokens/s as `vs_reference_gpu` ("restatement": it is not the reference's code).
"""
from __future__ import annotations
import torch
import torch.nn.functional as F

from .config import Cfg
from .decoder import ACT


class StaticCacheDecoder:
    def __init__(self, cfg: Cfg, weights: dict[str, torch.Tensor], dtype, device, B: int, S: int, P: int, Tmax: int):
        self.cfg, self.dtype, self.dev = cfg, dtype, torch.device(device)
        self.w = {k: v.to(self.dev, dtype) for k, v in weights.items()}
        self.B, self.S, self.P, self.Tmax = B, S, P, Tmax
        self.H, self.nh, self.K, self.L = cfg.hidden_size, cfg.num_attention_heads, cfg.num_codebooks, cfg.num_hidden_layers
        self.hd = self.H // self.nh
        assert cfg.num_key_value_heads == self.nh and cfg.num_cross_attention_key_value_heads == self.nh and not cfg.rope_embeddings, \
            "the comparator covers the in-repo Mini / Large shapes (MHA, sinusoidal positions)"
        self.eps = cfg.get("layer_norm_eps", 1e-5)
        self.act = ACT[cfg.activation_function]
        z = lambda *s: torch.zeros(*s, dtype=dtype, device=self.dev)
        self.kc = [z(B, self.nh, Tmax, self.hd) for _ in range(self.L)]   # StaticCache layout [B, heads, max_len, head_dim]
        self.vc = [z(B, self.nh, Tmax, self.hd) for _ in range(self.L)]
        self.ck = [None] * self.L
        self.cv = [None] * self.L
        self.pos = torch.zeros(1, dtype=torch.long, device=self.dev)       # cache_position of the token being fed
        self.ids = torch.zeros(B * self.K, 1, dtype=torch.long, device=self.dev)
        self.key_pad = torch.zeros(B, 1, 1, Tmax, dtype=dtype, device=self.dev)  # additive: finfo.min on padded prompt keys
        self.enc_mask4d = None
        self.ar = torch.arange(Tmax, device=self.dev)

    def _p(self, n):
        return self.w["decoder.model.decoder." + n]

    def _ln(self, x, n):
        return F.layer_norm(x, (self.H,), self._p(n + ".weight"), self._p(n + ".bias"), self.eps)

    def _heads(self, x):
        return x.view(x.shape[0], x.shape[1], self.nh, self.hd).transpose(1, 2)

    def _layer(self, i, h, positions, mask, enc_mask):
        pre = f"layers.{i}."
        B, q, _ = h.shape
        r = h
        x = self._ln(h, pre + "self_attn_layer_norm")
        p = pre + "self_attn."
        qs = self._heads(F.linear(x, self._p(p + "q_proj.weight")))
        kn = self._heads(F.linear(x, self._p(p + "k_proj.weight")))
        vn = self._heads(F.linear(x, self._p(p + "v_proj.weight")))
        self.kc[i].index_copy_(2, positions, kn)   # StaticCache.update
        self.vc[i].index_copy_(2, positions, vn)
        o = F.scaled_dot_product_attention(qs, self.kc[i], self.vc[i], attn_mask=mask)
        h = r + F.linear(o.transpose(1, 2).reshape(B, q, self.H), self._p(p + "out_proj.weight"))
        r = h
        x = self._ln(h, pre + "encoder_attn_layer_norm")
        p = pre + "encoder_attn."
        qs = self._heads(F.linear(x, self._p(p + "q_proj.weight")))
        o = F.scaled_dot_product_attention(qs, self.ck[i], self.cv[i], attn_mask=enc_mask)
        h = r + F.linear(o.transpose(1, 2).reshape(B, q, self.H), self._p(p + "out_proj.weight"))
        r = h
        x = self._ln(h, pre + "final_layer_norm")
        x = self.act(F.linear(x, self._p(pre + "fc1.weight")))
        return r + F.linear(x, self._p(pre + "fc2.weight"))

    def _forward(self, emb, positions):
        B, q, _ = emb.shape
        h = emb + self._p("embed_positions.weights").index_select(0, positions)
        mn = torch.finfo(self.dtype).min
        # causal + padding mask over the WHOLE static cache (what _update_causal_mask builds for a StaticCache, :1696-1724)
        causal = (self.ar[None, :] > positions[:, None]).to(self.dtype) * mn          # [q, Tmax]
        mask = torch.minimum(causal[None, None], self.key_pad.expand(B, 1, q, self.Tmax))
        enc_mask = None if self.enc_mask4d is None else self.enc_mask4d.expand(B, 1, q, -1)
        for i in range(self.L):
            h = self._layer(i, h, positions, mask, enc_mask)
        h = self._ln(h, "layer_norm")[:, -1:]
        logits = torch.stack([F.linear(h, self.w[f"decoder.lm_heads.{k}.weight"]) for k in range(self.K)], dim=1)
        return logits.reshape(B * self.K, -1).float()

    def _embed(self, ids):
        idb = ids.reshape(self.B, self.K, -1)
        return sum([F.embedding(idb[:, k], self._p(f"embed_tokens.{k}.weight")) for k in range(self.K)])

    @torch.no_grad()
    def prefill(self, ids, enc_hidden, enc_mask, prompt_hidden, prompt_mask):
        B, P = self.B, self.P
        enc_hidden = enc_hidden.to(self.dev, self.dtype)
        mn = torch.finfo(self.dtype).min
        self.key_pad.zero_()
        if prompt_mask is not None and P > 0:
            self.key_pad[:, 0, 0, :P] = (1 - prompt_mask.to(self.dev)).to(self.dtype) * mn
        self.enc_mask4d = None
        if enc_mask is not None and not bool((enc_mask == 1).all()):
            self.enc_mask4d = ((1 - enc_mask.to(self.dev))[:, None, None, :].to(self.dtype) * mn)
        for i in range(self.L):
            p = f"layers.{i}.encoder_attn."
            self.ck[i] = self._heads(F.linear(enc_hidden, self._p(p + "k_proj.weight"))).contiguous()
            self.cv[i] = self._heads(F.linear(enc_hidden, self._p(p + "v_proj.weight"))).contiguous()
        emb = self._embed(ids.to(self.dev))
        if prompt_hidden is not None:
            emb = torch.cat([prompt_hidden.to(self.dev, self.dtype), emb], dim=1)
        q = emb.shape[1]
        logits = self._forward(emb, torch.arange(q, device=self.dev))
        self.pos.fill_(q)
        return logits

    @torch.no_grad()
    def step_and_sample(self, top_k: int):
        """One decode step on the ids staged in self.ids + HF-style top-k sampling; advances pos and stages the next ids."""
        logits = self._forward(self._embed(self.ids), self.pos)
        if top_k > 0:
            kth = torch.topk(logits, top_k)[0][..., -1, None]
            logits = logits.masked_fill(logits < kth, -float("inf"))          # TopKLogitsWarper
        nxt = torch.multinomial(F.softmax(logits, dim=-1), 1)                  # _sample
        self.ids.copy_(nxt)
        self.pos.add_(1)
        return nxt
