"""DACModel: the reference's codec wrapper surface, backed by the sm_100a DAC decode kernels.

Mirrors parler_tts/dac_wrapper/modeling_dac.py:14-164 for the decode path:
  DACModel.decode(audio_codes, audio_scales, padding_mask=None, return_dict=None)   (:106-142)
`encode` (voice-prompt path) is out of this path's scope and raises.
Weights: either folded tensors under transformers-DacModel style keys (decoder.conv1.weight, ...) or
descript-audio-codec checkpoint keys with weight-norm parameters (weight_g / weight_v, or
parametrizations.weight.original0/1), folded here as w = g * v / ||v|| (reference :148-157).
"""
from __future__ import annotations
import ctypes as C
import math
import re
from dataclasses import dataclass

import torch

from . import _lib
from .configuration import DACConfig


@dataclass
class DACDecoderOutput:
    """Stands in for transformers' EncodecDecoderOutput (audio_values [B, 1, samples])."""
    audio_values: torch.Tensor = None

    def __getitem__(self, i):
        return (self.audio_values,)[i]


def _dac_tensor_list(cfg: DACConfig) -> list[str]:
    """(id -> key) table in the order csrc/dac.h::make_dac_layout enumerates tensors."""
    names = []
    for i in range(cfg.num_codebooks):
        q = f"quantizer.quantizers.{i}."
        names += [q + "codebook.weight", q + "out_proj.weight", q + "out_proj.bias"]
    names += ["decoder.conv1.weight", "decoder.conv1.bias"]
    for bi in range(len(cfg.decoder_rates)):
        p = f"decoder.block.{bi}."
        names += [p + "snake1.alpha", p + "conv_t1.weight", p + "conv_t1.bias"]
        for r in (1, 2, 3):
            u = p + f"res_unit{r}."
            names += [u + "snake1.alpha", u + "conv1.weight", u + "conv1.bias", u + "snake2.alpha", u + "conv2.weight", u + "conv2.bias"]
    names += ["decoder.snake1.alpha", "decoder.conv2.weight", "decoder.conv2.bias"]
    return names


def _fold_weight_norm(sd: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
    """g * v / ||v|| over all dims but 0 (torch weight_norm default dim=0), in fp32."""
    out = {}
    pairs = {}
    for k, v in sd.items():
        m = re.match(r"(.*)\.(weight_g|weight_v|parametrizations\.weight\.original0|parametrizations\.weight\.original1)$", k)
        if m:
            kind = "g" if (m.group(2).endswith("_g") or m.group(2).endswith("original0")) else "v"
            pairs.setdefault(m.group(1), {})[kind] = v
        else:
            out[k] = v
    for base, gv in pairs.items():
        g, v = gv["g"].float(), gv["v"].float()
        norm = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
        out[base + ".weight"] = g * v / norm
    return out


def _from_descript_keys(sd: dict[str, torch.Tensor], n_blocks: int) -> dict[str, torch.Tensor]:
    """descript-audio-codec module paths -> the transformers-DacModel style keys used internally.

    descript layout: decoder.model = [WNConv1d, DecoderBlock x n, Snake1d, WNConv1d, Tanh];
    DecoderBlock.block = [Snake1d, WNConvTranspose1d, ResidualUnit x 3]; ResidualUnit.block =
    [Snake1d, WNConv1d(k7), Snake1d, WNConv1d(k1)]."""
    out = {}
    for k, v in sd.items():
        if k.startswith("quantizer."):
            out[k] = v
            continue
        m = re.match(r"decoder\.model\.(\d+)\.(.*)$", k)
        if not m:
            continue
        i, rest = int(m.group(1)), m.group(2)
        if i == 0:
            out["decoder.conv1." + rest] = v
        elif 1 <= i <= n_blocks:
            bi = i - 1
            mm = re.match(r"block\.(\d+)\.(.*)$", rest)
            j, r2 = int(mm.group(1)), mm.group(2)
            p = f"decoder.block.{bi}."
            if j == 0:
                out[p + "snake1." + r2] = v
            elif j == 1:
                out[p + "conv_t1." + r2] = v
            else:
                m3 = re.match(r"block\.(\d+)\.(.*)$", r2)
                u, r3 = int(m3.group(1)), m3.group(2)
                name = {0: "snake1.", 1: "conv1.", 2: "snake2.", 3: "conv2."}[u]
                out[p + f"res_unit{j - 1}." + name + r3] = v
        elif i == n_blocks + 1:
            out["decoder.snake1." + rest] = v
        elif i == n_blocks + 2:
            out["decoder.conv2." + rest] = v
    return out


class DACModel:
    config_class = DACConfig
    main_input_name = "input_values"

    def __init__(self, config: DACConfig, device="cuda", dtype=torch.float32):
        self.config = config
        self.device = torch.device(device)
        self.dtype = dtype
        self._c = _lib.DacConfigC()
        self._c.n_codebooks = config.num_codebooks
        self._c.codebook_size = config.codebook_size
        self._c.codebook_dim = config.codebook_dim
        self._c.latent_dim = config.latent_dim
        self._c.decoder_dim = config.decoder_dim
        self._c.n_blocks = len(config.decoder_rates)
        for i, s in enumerate(config.decoder_rates):
            self._c.strides[i] = int(s)
        self._c.dtype = _lib.dtype_code(dtype)
        self.hop_length = math.prod(config.decoder_rates)
        # the packed weight blob exists from construction on (zeros until load_state_dict / a broadcast fills it), so that
        # every rank of a sharded run owns a buffer of the right size for the init broadcast (dist.broadcast_model_weights)
        nbytes = C.c_int64()
        _lib.check(_lib.lib().ptts_dac_blob_bytes(C.byref(self._c), C.byref(nbytes)))
        self.blob = torch.zeros(nbytes.value, dtype=torch.uint8, device=self.device)
        self.loaded = False
        self._ws = None

    # -- weights -----------------------------------------------------------------------------------
    def load_state_dict(self, sd: dict[str, torch.Tensor], strict: bool = True):
        sd = {(k[len("model."):] if k.startswith("model.") else k): v for k, v in sd.items()}
        sd = _fold_weight_norm(sd)
        if any(k.startswith("decoder.model.") for k in sd):
            sd = _from_descript_keys(sd, len(self.config.decoder_rates))
        names = _dac_tensor_list(self.config)
        missing = [n for n in names if n not in sd]
        if missing and strict:
            raise ValueError(f"DACModel.load_state_dict: missing keys {missing[:5]}{'...' if len(missing) > 5 else ''}")
        lib = _lib.lib()
        n = C.c_int32()
        _lib.check(lib.ptts_dac_num_tensors(C.byref(self._c), C.byref(n)))
        assert n.value == len(names), (n.value, len(names))
        self.blob.zero_()
        for i, name in enumerate(names):
            if name not in sd:
                continue
            t = sd[name].to(device=self.device)
            if t.dtype not in (torch.float32, torch.bfloat16):
                t = t.float()
            t = t.contiguous()
            _lib.check(lib.ptts_dac_pack(C.byref(self._c), _lib.ptr(self.blob), i, _lib.ptr(t), _lib.dtype_code(t.dtype),
                                         t.numel(), _lib.stream_ptr()))
        torch.cuda.current_stream().synchronize()  # staging tensors above go out of scope
        self.loaded = True
        return self

    def to(self, *args, **kwargs):
        return self

    def eval(self):
        return self

    # -- reference surface -------------------------------------------------------------------------
    def encode(self, *args, **kwargs):
        raise NotImplementedError("DACModel.encode (voice-prompt path) is outside the B200 generate() hot path")

    @torch.no_grad()
    def decode(self, audio_codes, audio_scales=None, padding_mask=None, return_dict=None):
        """audio_codes [1, B, K, T] int64 (CUDA) -> DACDecoderOutput(audio_values [B, 1, hop*T])."""
        if not self.loaded:
            raise RuntimeError("DACModel has no weights loaded")
        if len(audio_codes) != 1:
            raise ValueError(f"Expected one frame, got {len(audio_codes)}")
        codes = audio_codes.squeeze(0)
        if codes.dim() != 3 or codes.shape[1] != self.config.num_codebooks:
            raise ValueError(f"audio_codes must be [1, B, {self.config.num_codebooks}, T], got {tuple(audio_codes.shape)}")
        codes = codes.to(device=self.device, dtype=torch.int64).contiguous()
        B, K, T = codes.shape
        if T == 0 or B == 0:
            raise ValueError("audio_codes is empty")
        if bool(((codes < 0) | (codes >= self.config.codebook_size)).any()):
            raise IndexError("audio code out of range for the codebook (the reference's embedding lookup raises too)")
        lib = _lib.lib()
        need = C.c_int64()
        _lib.check(lib.ptts_dac_workspace_bytes(C.byref(self._c), B, T, C.byref(need)))
        if self._ws is None or self._ws.numel() < need.value:
            self._ws = torch.empty(need.value, dtype=torch.uint8, device=self.device)
        audio = torch.empty(B, 1, T * self.hop_length, dtype=self.dtype, device=self.device)
        _lib.check(lib.ptts_dac_decode(C.byref(self._c), _lib.ptr(self.blob), _lib.ptr(self._ws), self._ws.numel(),
                                       _lib.ptr(codes), B, T, _lib.ptr(audio), _lib.stream_ptr()))
        if return_dict is False:
            return (audio,)
        return DACDecoderOutput(audio)

    def forward(self, tensor):
        raise ValueError("`DACModel.forward` not implemented yet")
