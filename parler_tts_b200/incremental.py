"""Incremental (streaming) codec decode: O(T) instead of the reference streamer's O(T^2).

The reference's `ParlerTTSStreamer` re-decodes the ENTIRE token history every `play_steps` steps and keeps the new tail
(parler_tts/streamer.py:66-122).  The DAC decoder is a stack of (dilated / transposed)
convolutions with a finite receptive field, so the samples of frame f depend only on the code frames in
[f - R, f + R].  `IncrementalDecoder` decodes a window of `new frames + 2R frames of context` per call (overlap-save) and
emits exactly the samples a decode of the whole sequence would produce for the frames that have all of their right
context -- any batch size, no carried kernel state, the unchanged `DACModel.decode` kernels underneath.

`dac_dependency_radius` derives R from the decoder's layer list (upsampling ratios; kernel 7 / dilations 1,3,9 residual
units; transposed convolutions with kernel 2s, stride s, padding ceil(s/2)) by exact interval arithmetic, so it follows
the config instead of being a constant (R = 10 frames for the 44.1 kHz DAC).
"""
from __future__ import annotations
import math
from typing import Callable, Optional, Sequence

import torch


def dac_dependency_radius(upsampling_ratios: Sequence[int], res_dilations: Sequence[int] = (1, 3, 9), kernel: int = 7) -> int:
    """Largest |f' - f| such that the output samples of code frame f depend on code frame f' (dac decoder layer stack:
    conv k7 -> per block [convT(k=2s, stride s, pad ceil(s/2)), 3 x (conv k7 dilated, conv k1)] -> conv k7)."""
    hop = 1
    for s in upsampling_ratios:
        hop *= int(s)
    radius = 0
    for f in (64, 65):  # two adjacent mid-sequence frames (the pattern is periodic in the frame index)
        lo, hi = f * hop, (f + 1) * hop - 1            # output samples of frame f
        half = (kernel - 1) // 2
        lo, hi = lo - half, hi + half                    # final conv k7
        for s in reversed([int(v) for v in upsampling_ratios]):
            for d in reversed(tuple(res_dilations)):     # residual units: conv k1 (no spread), conv k7 dilated
                lo, hi = lo - half * d, hi + half * d
            k, p = 2 * s, math.ceil(s / 2)               # transposed conv: y[n] += x[i] w[j], n = i*s + j - p
            lo, hi = -((-(lo + p - (k - 1))) // s), (hi + p) // s   # ceil / floor
        lo, hi = lo - half, hi + half                    # first conv k7
        radius = max(radius, f - lo, hi - f)
    return int(radius)


class IncrementalDecoder:
    """Feed code frames as they are generated, get the finalized audio back.

    decode_fn: codes [B, K, T] (int64) -> audio [B, hop*T]; e.g. `lambda c: dac.decode(c[None]).audio_values[:, 0]`.
    push(codes_new [B, K, t]) -> audio [B, hop * n] for the n frames that just became final (n may be 0);
    finish() -> the remaining tail.  Concatenating all returned chunks equals decode_fn(all codes) (up to the
    floating-point reassociation of decode_fn itself; identical windows give identical samples).
    """

    def __init__(self, decode_fn: Callable[[torch.Tensor], torch.Tensor], hop_length: int, radius: int, min_new_frames: int = 1):
        self.decode_fn, self.hop, self.R = decode_fn, int(hop_length), int(radius)
        self.min_new = max(1, int(min_new_frames))
        self.codes: Optional[torch.Tensor] = None   # only the last R + pending frames are kept
        self.first = 0                               # absolute index of self.codes[..., 0]
        self.done = 0                                # frames whose samples have been emitted
        self.total = 0                               # frames received

    def _emit(self, upto: int) -> Optional[torch.Tensor]:
        if upto <= self.done:
            return None
        w0 = max(0, self.done - self.R)              # left context (or the true start of the sequence)
        window = self.codes[..., w0 - self.first:]
        audio = self.decode_fn(window)
        out = audio[..., (self.done - w0) * self.hop:(upto - w0) * self.hop]
        self.done = upto
        keep_from = max(0, self.done - self.R)       # drop frames no later window can need
        if keep_from > self.first:
            self.codes = self.codes[..., keep_from - self.first:]
            self.first = keep_from
        return out

    def push(self, codes_new: torch.Tensor) -> Optional[torch.Tensor]:
        if codes_new.dim() != 3:
            raise ValueError(f"codes must be [B, K, t], got {tuple(codes_new.shape)}")
        self.codes = codes_new if self.codes is None else torch.cat([self.codes, codes_new], dim=-1)
        self.total += codes_new.shape[-1]
        final = self.total - self.R                  # frames with all of their right context present
        if final - self.done < self.min_new:
            return None
        return self._emit(final)

    def finish(self) -> Optional[torch.Tensor]:
        if self.codes is None:
            return None
        return self._emit(self.total)                # the sequence end is a true edge: zero padding is the real context
