"""ParlerTTSStreamer: same queue/iterator contract as the reference (parler_tts/streamer.py:11-146).

put(value) receives the initial [B*K, 1] ids and then one [B*K] column per step (on the CPU, as
`_sample` hands them over); every `play_steps` columns the accumulated tokens are de-delayed and decoded
with the CUDA DAC path, and the new tail (minus `stride` samples of overlap) is queued for the consumer
thread.  Batch size 1 only, like the reference (:110-112).

`incremental=True` (not in the reference) keeps the same put/iterate contract but decodes only the
NEW frames plus a fixed window of context every `play_steps` (parler_tts_b200/incremental.py): O(T) codec work instead
of O(T^2), and every emitted sample is final -- no `stride` of provisional tail is held back or re-decoded.
"""
from __future__ import annotations
import math
from queue import Queue
from typing import Optional

import numpy as np
import torch

from .incremental import IncrementalDecoder, dac_dependency_radius
from .modeling import apply_delay_pattern_mask, build_delay_pattern_mask


class ParlerTTSStreamer:
    def __init__(self, model, device: Optional[str] = None, play_steps: Optional[int] = 10, stride: Optional[int] = None,
                 timeout: Optional[float] = None, incremental: bool = False):
        self.decoder = model.decoder
        self.audio_encoder = model.audio_encoder
        self.generation_config = model.generation_config
        self.device = device if device is not None else model.device
        self.use_audio_scales = model.use_audio_scales
        self.use_4dim_audio_codes = model.use_4dim_audio_codes
        self.audio_kwargs = {"audio_scales": [None]} if self.use_audio_scales else {}
        self.play_steps = play_steps
        if stride is not None:
            self.stride = stride
        else:
            hop_length = math.floor(self.audio_encoder.config.sampling_rate / self.audio_encoder.config.frame_rate)
            self.stride = hop_length * (play_steps - self.decoder.num_codebooks) // 6
        self.token_cache = None
        self.to_yield = 0
        self.audio_queue: Queue = Queue()
        self.stop_signal = None
        self.timeout = timeout
        self._inc = None
        self._pushed = 0
        if incremental:
            cfg = self.audio_encoder.config
            hop = 1
            for r in cfg.decoder_rates:
                hop *= int(r)
            self._inc = IncrementalDecoder(
                lambda c: self.audio_encoder.decode(audio_codes=c[None, ...], **self.audio_kwargs).audio_values[:, 0],
                hop, dac_dependency_radius(cfg.decoder_rates))

    def _valid_frames(self, input_ids: torch.Tensor) -> torch.Tensor:
        """De-delay the raw token history -> complete code frames [1, K, n] (columns holding a special token are dropped)."""
        gc = self.generation_config
        ids = input_ids.to(self.device)
        _, mask = build_delay_pattern_mask(ids[:, :1], gc.bos_token_id, gc.decoder_start_token_id, ids.shape[-1],
                                           self.decoder.num_codebooks)
        ids = apply_delay_pattern_mask(ids, mask)
        keep = (mask != gc.bos_token_id) & (mask != gc.pad_token_id)
        ids = ids[keep].reshape(1, self.decoder.num_codebooks, -1)[None, ...]
        cs = self.audio_encoder.config.codebook_size
        sample = ids[:, 0]
        ok = (sample >= cs).sum(dim=(0, 1)) == 0
        return sample[:, :, ok]

    def apply_delay_pattern_mask(self, input_ids: torch.Tensor) -> np.ndarray:
        sample = self._valid_frames(input_ids)
        if sample.shape[-1] == 0:
            return np.zeros(0, dtype=np.float32)
        audio = self.audio_encoder.decode(audio_codes=sample[None, ...], **self.audio_kwargs).audio_values
        return audio[0, 0].float().cpu().numpy()

    def _push_incremental(self, final: bool):
        # below 2K-1 columns the delay-pattern mask degenerates (build_delay_pattern_mask returns the ids unmasked,
        # modeling_parler_tts.py:241-243): no complete frame can be told apart from the BOS triangle yet
        enough = self.token_cache is not None and self.token_cache.shape[-1] >= 2 * self.decoder.num_codebooks - 1
        frames = self._valid_frames(self.token_cache) if enough else None
        chunks = []
        if frames is not None and frames.shape[-1] > self._pushed:
            out = self._inc.push(frames[..., self._pushed:])
            self._pushed = frames.shape[-1]
            if out is not None:
                chunks.append(out)
        if final:
            out = self._inc.finish()
            if out is not None:
                chunks.append(out)
        audio = torch.cat(chunks, dim=-1)[0].float().cpu().numpy() if chunks else np.zeros(0, dtype=np.float32)
        self.on_finalized_audio(audio, stream_end=final)

    def put(self, value: torch.Tensor):
        batch_size = value.shape[0] // self.decoder.num_codebooks
        if batch_size > 1:
            raise ValueError("ParlerTTSStreamer only supports batch size 1")
        if self.token_cache is None:
            self.token_cache = value if value.dim() == 2 else value[:, None]
        else:
            self.token_cache = torch.concatenate([self.token_cache, value[:, None]], dim=-1)
        if self._inc is not None:
            if self.token_cache.shape[-1] % self.play_steps == 0:
                self._push_incremental(final=False)
            return
        if self.token_cache.shape[-1] % self.play_steps == 0:
            audio_values = self.apply_delay_pattern_mask(self.token_cache)
            self.on_finalized_audio(audio_values[self.to_yield: -self.stride])
            self.to_yield += len(audio_values) - self.to_yield - self.stride

    def end(self):
        if self._inc is not None:
            self._push_incremental(final=True)
            return
        if self.token_cache is not None:
            audio_values = self.apply_delay_pattern_mask(self.token_cache)
        else:
            audio_values = np.zeros(self.to_yield)
        self.on_finalized_audio(audio_values[self.to_yield:], stream_end=True)

    def on_finalized_audio(self, audio: np.ndarray, stream_end: bool = False):
        self.audio_queue.put(audio, timeout=self.timeout)
        if stream_end:
            self.audio_queue.put(self.stop_signal, timeout=self.timeout)

    def __iter__(self):
        return self

    def __next__(self):
        value = self.audio_queue.get(timeout=self.timeout)
        if not isinstance(value, np.ndarray) and value == self.stop_signal:
            raise StopIteration()
        return value
