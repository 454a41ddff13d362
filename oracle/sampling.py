"""CPU restatement of the per-step logits processing and the `_sample` loop (test infrastructure only).

  ParlerTTSLogitsProcessor            parler_tts/logits_processors.py:23-53 (stateful, quirk Q11)
  MinNewTokensLength / Temperature / TopK / TopP warpers
                                      transformers 4.46.1 `generation/logits_process.py` (not vendored in
                                      /root/reference; same arithmetic in the container's 5.5.0 at
                                      logits_process.py:225-233, 297-299, 521-533, 581-586); PINNED bit-exact
                                      against those classes by tests/golden/warpers.npz (make_golden.gen_warpers)
  `_sample` loop                      transformers 4.46.1 `generation/utils.py::_sample` (restated; the
                                      reference calls it at modeling_parler_tts.py:3564); PINNED bit-exact against the
                                      installed transformers' `_sample` on scripted logits (tests/golden/sample_loop.npz)
  processor order                     [MinNewTokens, ParlerTTS (custom), Temperature, TopK, TopP]
                                      (`_get_logits_processor`: defaults, then merged custom list, then warpers)
"""
from __future__ import annotations
import math
import numpy as np
import torch

from .delay_pattern import build_delay_pattern_mask, apply_delay_pattern_mask


class ParlerLogitsProcessorOracle:
    def __init__(self, eos_token_id: int, num_codebooks: int, batch_size: int):
        self.eos, self.K, self.B = eos_token_id, num_codebooks, batch_size
        self.codebook_idx = np.arange(batch_size * num_codebooks)
        self.first_unfinished = np.arange(batch_size) * num_codebooks
        self.max_codebooks = np.arange(batch_size) * num_codebooks + num_codebooks - 1

    def __call__(self, input_ids: np.ndarray, scores: np.ndarray) -> np.ndarray:
        is_eos = (input_ids == self.eos).sum(1)  # over the row's entire history (:46)
        adv = (is_eos[self.first_unfinished] > 0) & (self.first_unfinished < self.max_codebooks)
        self.first_unfinished = np.where(adv, self.first_unfinished + 1, self.first_unfinished)  # :48
        mask = self.codebook_idx > np.repeat(self.first_unfinished, self.K)  # :51
        scores[mask, self.eos] = -math.inf  # in place (:52)
        return scores


def min_new_tokens(scores, cur_len, prompt_len, min_new, eos):
    if cur_len - prompt_len < min_new:
        scores = scores.copy()
        scores[:, eos] = -math.inf
    return scores


def temperature(scores, t):
    return scores / np.float32(t)


def top_k(scores, k):
    k = min(k, scores.shape[-1])
    kth = np.sort(scores, axis=-1)[:, -k][:, None]
    return np.where(scores < kth, -np.inf, scores).astype(np.float32)


def top_p(scores, p, min_keep=1):
    t = torch.from_numpy(scores)
    sl, si = torch.sort(t, descending=False)
    cum = sl.softmax(dim=-1).cumsum(dim=-1)
    rem = cum <= (1 - p)
    rem[..., -min_keep:] = False
    rem = rem.scatter(1, si, rem)
    return t.masked_fill(rem, -math.inf).numpy()


def process_scores(scores: np.ndarray, raw_ids: np.ndarray, parler: ParlerLogitsProcessorOracle, gen: dict):
    """One step's processor chain on fp32 scores [B*K, V]; raw_ids = un-masked history (Q10)."""
    s = scores.astype(np.float32).copy()
    if gen.get("min_new_tokens", 0) > 0:
        s = min_new_tokens(s, raw_ids.shape[1], 1, gen["min_new_tokens"], parler.eos)
    s = parler(raw_ids, s)
    if gen.get("do_sample", False):
        if gen.get("temperature", 1.0) != 1.0:
            s = temperature(s, gen["temperature"])
        if gen.get("top_k", 0):
            s = top_k(s, gen["top_k"])
        if gen.get("top_p", 1.0) < 1.0:
            s = top_p(s, gen["top_p"])
    return s


def softmax_rows(s: np.ndarray) -> np.ndarray:
    return torch.from_numpy(s).softmax(dim=-1).numpy()


def generate_tokens(dec, cfg, enc_hidden, enc_mask, prompt_hidden, prompt_mask, gen: dict,
                    forced_tokens: np.ndarray | None = None, pick=None, collect_logits=False):
    """Restated `generate()` up to the raw token matrix (modeling_parler_tts.py:3449-3572 + `_sample`).

    gen: max_length (total decoder length incl. BOS), do_sample, temperature, top_k, top_p, min_new_tokens.
    forced_tokens [B*K, steps]: teacher forcing (the sampled token is replaced before the append).
    pick(step, probs_or_scores) -> next tokens, for sampled runs driven by an external uniform stream.
    Returns dict(raw_ids [B*K, 1+steps], delay_mask, logits list, scores list).
    """
    B = enc_hidden.shape[0]
    K, bos, pad, eos = cfg.num_codebooks, cfg.bos_token_id, cfg.pad_token_id, cfg.eos_token_id
    L = gen["max_length"]
    ids = np.full((B * K, 1), bos, dtype=np.int64)
    ids, delay_mask = build_delay_pattern_mask(ids, bos, pad, L, K)  # :3523
    parler = ParlerLogitsProcessorOracle(eos, K, B)
    unfinished = np.ones(B * K, dtype=np.int64)
    all_logits, all_scores = [], []
    step = 0
    while True:
        model_in = apply_delay_pattern_mask(ids, delay_mask)  # :2909
        if step == 0:
            logits = dec.prefill(torch.from_numpy(model_in), enc_hidden, enc_mask, prompt_hidden, prompt_mask)
        else:
            logits = dec.step(torch.from_numpy(model_in[:, -1:]))
        nl = logits[:, -1, :].float().numpy()
        s = process_scores(nl, ids, parler, gen)
        if collect_logits:
            all_logits.append(nl.copy())
            all_scores.append(s.copy())
        if pick is not None:
            nxt = pick(step, s)
        elif gen.get("do_sample", False):
            nxt = torch.multinomial(torch.from_numpy(softmax_rows(s)), 1).squeeze(1).numpy()
        else:
            nxt = s.argmax(-1)
        if forced_tokens is not None:
            nxt = forced_tokens[:, step]
        nxt = nxt * unfinished + pad * (1 - unfinished)
        ids = np.concatenate([ids, nxt[:, None]], axis=1)
        done = (ids[:, -1] == eos) | (ids.shape[1] >= L)  # EosTokenCriteria | MaxLengthCriteria
        unfinished = unfinished & ~done
        step += 1
        if unfinished.max() == 0:
            break
    return dict(raw_ids=ids, delay_mask=delay_mask, logits=all_logits, scores=all_scores, steps=step)


def frames_from_raw(raw_ids, delay_mask, cfg, B):
    """generate() tail (:3586-3597) -> audio codes [B, K, T']."""
    from .delay_pattern import undelay
    return undelay(raw_ids, cfg.bos_token_id, cfg.pad_token_id, cfg.num_codebooks, B, delay_mask)


def valid_frame_mask(codes_bkt: np.ndarray, codebook_size: int) -> np.ndarray:
    """Per-sample frame validity (:3630-3631): a frame survives iff no codebook id >= codebook_size."""
    return (codes_bkt >= codebook_size).sum(axis=1) == 0
