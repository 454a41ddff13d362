"""CPU restatement of the DAC decode path (test infrastructure only).

Reference call sites: DACModel.decode -> quantizer.from_codes(...)[0] then model.decode(z)
(parler_tts/dac_wrapper/modeling_dac.py:138-139).  The arithmetic lives in descript-audio-codec
(unpinned in setup.py:24, NOT installed here) -- PARITY UNPINNED against it.  It is pinned instead
against transformers 5.5.0's DacModel, which restates the same network:
  from_codes       transformers/models/dac/modeling_dac.py:345-369
  decoder          :405-440   block :234-262   residual unit :173-207   snake :85-99
Weights arrive already weight-norm-folded (w = g * v / ||v||, modeling_dac.py:148-157 in the reference wrapper).
"""
from __future__ import annotations
import math
import torch
import torch.nn.functional as F

from .config import Cfg


def snake(x, alpha):
    return x + (alpha + 1e-9).reciprocal() * torch.sin(alpha * x).pow(2)


class OracleDAC:
    def __init__(self, cfg: Cfg, weights: dict[str, torch.Tensor], dtype=torch.float32):
        self.cfg, self.dtype = cfg, dtype
        self.w = {k: v.to(dtype) for k, v in weights.items()}

    def from_codes(self, codes_bkt: torch.Tensor) -> torch.Tensor:
        z = 0.0
        for i in range(codes_bkt.shape[1]):
            e = F.embedding(codes_bkt[:, i, :], self.w[f"quantizer.quantizers.{i}.codebook.weight"]).transpose(1, 2)
            z = z + F.conv1d(e, self.w[f"quantizer.quantizers.{i}.out_proj.weight"], self.w[f"quantizer.quantizers.{i}.out_proj.bias"])
        return z

    def _res(self, x, p, dil):
        y = F.conv1d(snake(x, self.w[p + "snake1.alpha"]), self.w[p + "conv1.weight"], self.w[p + "conv1.bias"],
                     dilation=dil, padding=3 * dil)
        y = F.conv1d(snake(y, self.w[p + "snake2.alpha"]), self.w[p + "conv2.weight"], self.w[p + "conv2.bias"])
        return x + y

    def decoder(self, z: torch.Tensor, return_intermediates=False) -> torch.Tensor:
        w = self.w
        inter = {}
        x = F.conv1d(z, w["decoder.conv1.weight"], w["decoder.conv1.bias"], padding=3)
        inter["conv1"] = x
        for bi, s in enumerate(self.cfg.upsampling_ratios):
            p = f"decoder.block.{bi}."
            x = snake(x, w[p + "snake1.alpha"])
            x = F.conv_transpose1d(x, w[p + "conv_t1.weight"], w[p + "conv_t1.bias"], stride=s, padding=math.ceil(s / 2))
            inter[f"block{bi}.convt"] = x
            for ri, dil in ((1, 1), (2, 3), (3, 9)):
                x = self._res(x, p + f"res_unit{ri}.", dil)
            inter[f"block{bi}"] = x
        x = snake(x, w["decoder.snake1.alpha"])
        x = F.conv1d(x, w["decoder.conv2.weight"], w["decoder.conv2.bias"], padding=3)
        x = torch.tanh(x)
        return (x, inter) if return_intermediates else x

    def decode(self, audio_codes_1bkt: torch.Tensor) -> torch.Tensor:
        """DACModel.decode: audio_codes [1, B, K, T] -> audio_values [B, 1, 512*T]."""
        if len(audio_codes_1bkt) != 1:
            raise ValueError(f"Expected one frame, got {len(audio_codes_1bkt)}")
        return self.decoder(self.from_codes(audio_codes_1bkt.squeeze(0)))
