"""numpy restatement of the delay-pattern mask (test infrastructure only).

  build_delay_pattern_mask   parler_tts/modeling_parler_tts.py:214-276
  apply_delay_pattern_mask   :205-211
  de-delay after generate    :3586-3597
Integer work: results must be bit-exact with the reference (fixtures in tests/golden/delay_*.npz
were produced by executing the reference functions).
"""
from __future__ import annotations
import numpy as np


def build_delay_pattern_mask(input_ids: np.ndarray, bos_token_id: int, pad_token_id: int, max_length: int,
                             num_codebooks: int):
    ids = np.asarray(input_ids, dtype=np.int64).reshape(-1, num_codebooks, input_ids.shape[-1])
    bsz, K, seq_len = ids.shape
    shifted = -np.ones((bsz, K, max_length), dtype=np.int64)
    if max_length < 2 * K - 1:  # :242-243
        return ids.reshape(bsz * K, -1), shifted.reshape(bsz * K, -1)
    for k in range(K):  # :246-248
        hi = min(seq_len + k, max_length)
        shifted[:, k, k:hi] = ids[:, k, : hi - k]
    col = np.arange(max_length)[None, :]
    row = np.arange(K)[:, None]
    eos_pat = (col - row) >= (max_length - K + 1)  # triu(diagonal=max_length-K+1)  (:252-254)
    bos_pat = col <= row  # tril                     (:256)
    mask = ~(bos_pat | eos_pat)
    out = mask[None] * shifted + bos_pat[None] * bos_token_id + eos_pat[None] * pad_token_id  # :261
    first = out[:, 0, :]
    starts = np.nonzero(first == -1)[1]
    first_start = int(starts.min()) if len(starts) > 0 else seq_len
    pattern_mask = out.reshape(bsz * K, -1)
    return out[..., :first_start].reshape(bsz * K, -1), pattern_mask


def apply_delay_pattern_mask(input_ids: np.ndarray, mask: np.ndarray):
    m = mask[..., : input_ids.shape[-1]]
    return np.where(m == -1, input_ids, m)


def undelay(output_ids: np.ndarray, bos_token_id: int, pad_token_id: int, num_codebooks: int, batch_size: int,
            delay_mask: np.ndarray):
    """generate() tail (:3586-3597): apply stashed mask, rebuild at actual length, keep free cells -> [B,K,T']."""
    out = apply_delay_pattern_mask(output_ids, delay_mask)
    _, mask = build_delay_pattern_mask(output_ids[:, :1], bos_token_id, pad_token_id, out.shape[1], num_codebooks)
    keep = (mask != bos_token_id) & (mask != pad_token_id)
    return out[keep].reshape(batch_size, num_codebooks, -1)
