"""Generate the golden fixtures in this directory by EXECUTING REFERENCE CODE.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box);
the outputs (*.npz) are committed so tests never read the reference at run time.

What is executed:
  * parler_tts.modeling_parler_tts.build_delay_pattern_mask / apply_delay_pattern_mask  (verbatim)
  * parler_tts.logits_processors.ParlerTTSLogitsProcessor                               (verbatim, shim: isin)
  * parler_tts.ParlerTTSForCausalLM.forward(use_cache=False) with sdpa attention        (verbatim, 3 import shims)
  * transformers.models.dac.DacModel.decode  -- stand-in for descript-audio-codec (not installed)
Import shims: stub `dac.model.DAC`; `cache_utils.SlidingWindowCache`;
`pytorch_utils.isin_mps_friendly = torch.isin`.

Usage:  python tests/golden/make_golden.py
"""
from __future__ import annotations
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")


def import_reference():
    import transformers.cache_utils as cu
    import transformers.pytorch_utils as pu
    if not hasattr(cu, "SlidingWindowCache"):
        cu.SlidingWindowCache = cu.StaticCache
    if not hasattr(pu, "isin_mps_friendly"):
        pu.isin_mps_friendly = torch.isin
    dac = types.ModuleType("dac")
    dacm = types.ModuleType("dac.model")

    class DAC(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
    dacm.DAC = DAC
    dac.model = dacm
    sys.modules["dac"], sys.modules["dac.model"] = dac, dacm
    sys.path.insert(0, "/root/reference")
    import parler_tts
    return parler_tts


def gen_delay(pt):
    from parler_tts.modeling_parler_tts import build_delay_pattern_mask, apply_delay_pattern_mask
    out = {}
    cases = [(2, 4, 1, 8), (1, 9, 1, 257), (3, 9, 1, 20), (2, 4, 3, 12), (1, 9, 1, 16), (2, 9, 1, 17), (1, 4, 1, 6)]
    g = torch.Generator().manual_seed(5)
    for ci, (B, K, seq, L) in enumerate(cases):
        ids = torch.randint(0, 60, (B * K, seq), generator=g)
        ids[:, 0] = 65
        d, m = build_delay_pattern_mask(ids, 65, 64, L, K)
        full = torch.randint(0, 60, (B * K, L), generator=g)
        app = apply_delay_pattern_mask(full, m)
        part = apply_delay_pattern_mask(full[:, : max(1, L // 2)], m)
        out[f"c{ci}_meta"] = np.array([B, K, seq, L])
        out[f"c{ci}_ids"] = ids.numpy()
        out[f"c{ci}_delayed"] = d.numpy()
        out[f"c{ci}_mask"] = m.numpy()
        out[f"c{ci}_full"] = full.numpy()
        out[f"c{ci}_applied"] = app.numpy()
        out[f"c{ci}_applied_half"] = part.numpy()
    out["n"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "delay_pattern.npz"), **out)


def gen_logits_processor(pt):
    from parler_tts.logits_processors import ParlerTTSLogitsProcessor
    out = {}
    g = torch.Generator().manual_seed(11)
    B, K, V, eos, steps = 3, 4, 96, 64, 14
    proc = ParlerTTSLogitsProcessor(eos, K, B, "cpu")
    ids = torch.full((B * K, 1), 65, dtype=torch.long)
    hist_scores_in, hist_scores_out, hist_first = [], [], []
    for s in range(steps):
        scores = torch.randn(B * K, V, generator=g)
        hist_scores_in.append(scores.numpy().copy())
        o = proc(ids, scores.clone())
        hist_scores_out.append(o.numpy().copy())
        hist_first.append(proc.first_codebooks_unfinished.numpy().copy())
        nxt = torch.randint(0, 60, (B * K,), generator=g)
        # scripted EOS events: codebook 0 of sample 0 at step 2, then cascading; sample 2 cb0 at step 5
        if s == 2:
            nxt[0] = eos
        if s >= 3:
            nxt[0] = eos  # finished row keeps emitting pad == eos (Q11)
        if s == 4:
            nxt[1] = eos
        if s == 5:
            nxt[8] = eos
        if s == 6:
            nxt[2] = eos
            nxt[5] = eos  # a non-first codebook of sample 1 (cannot advance sample 1)
        ids = torch.cat([ids, nxt[:, None]], dim=1)
    out.update(meta=np.array([B, K, V, eos, steps]), ids=ids.numpy(), scores_in=np.stack(hist_scores_in),
               scores_out=np.stack(hist_scores_out), first=np.stack(hist_first))
    np.savez_compressed(os.path.join(HERE, "logits_processor.npz"), **out)


def gen_decoder(pt):
    from parler_tts import ParlerTTSDecoderConfig, ParlerTTSForCausalLM
    from oracle.config import tiny_cfg
    from oracle.weights import make_decoder_weights
    out = {}
    for name, kw in (("abs", dict(rope_embeddings=False)), ("rope", dict(rope_embeddings=True)),
                     ("gqa", dict(rope_embeddings=True, num_attention_heads=4, num_key_value_heads=2,
                                  num_cross_attention_key_value_heads=1, hidden_size=256))):
        cfg = tiny_cfg(**kw)
        w = make_decoder_weights(cfg, seed=3)
        rc = ParlerTTSDecoderConfig(
            vocab_size=cfg.vocab_size, max_position_embeddings=cfg.max_position_embeddings,
            num_hidden_layers=cfg.num_hidden_layers, ffn_dim=cfg.ffn_dim, num_attention_heads=cfg.num_attention_heads,
            num_key_value_heads=cfg.num_key_value_heads,
            num_cross_attention_key_value_heads=cfg.num_cross_attention_key_value_heads,
            hidden_size=cfg.hidden_size, num_codebooks=cfg.num_codebooks, pad_token_id=cfg.pad_token_id,
            eos_token_id=cfg.eos_token_id, bos_token_id=cfg.bos_token_id, dropout=0.0,
            rope_embeddings=cfg.rope_embeddings, activation_function=cfg.activation_function)
        rc._attn_implementation = "sdpa"
        m = ParlerTTSForCausalLM(rc).eval()
        sd = {k[len("decoder."):]: v for k, v in w.items() if k.startswith("decoder.")}
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert all("embed_positions" in k or "rotary" in k for k in missing), missing
        g = torch.Generator().manual_seed(7)
        B, K, T, P, S = 2, cfg.num_codebooks, 6, 5, 7
        ids = torch.randint(0, cfg.codebook_size, (B * K, T), generator=g)
        ids[:, 0] = cfg.bos_token_id
        enc = torch.randn(B, S, cfg.hidden_size, generator=g)
        enc_mask = torch.ones(B, S, dtype=torch.long)
        enc_mask[1, :3] = 0  # left padding on sample 1
        enc = enc * enc_mask[..., None]
        prompt = torch.randn(B, P, cfg.hidden_size, generator=g)
        pmask = torch.ones(B, P, dtype=torch.long)
        pmask[0, :2] = 0
        with torch.no_grad():
            lo = m(input_ids=ids, encoder_hidden_states=enc, encoder_attention_mask=enc_mask,
                   prompt_hidden_states=prompt, prompt_attention_mask=pmask, use_cache=False).logits
            lo_nomask = m(input_ids=ids, encoder_hidden_states=enc, prompt_hidden_states=prompt, use_cache=False).logits
        out[f"{name}_ids"] = ids.numpy()
        out[f"{name}_enc"] = enc.numpy()
        out[f"{name}_enc_mask"] = enc_mask.numpy()
        out[f"{name}_prompt"] = prompt.numpy()
        out[f"{name}_pmask"] = pmask.numpy()
        out[f"{name}_logits"] = lo.numpy()  # [B*K, P+T, V]
        out[f"{name}_logits_nomask"] = lo_nomask.numpy()
    np.savez_compressed(os.path.join(HERE, "decoder_forward.npz"), **out)


def gen_dac():
    from transformers.models.dac import DacConfig, DacModel
    from oracle.config import tiny_dac_cfg
    from oracle.weights import make_dac_weights
    cfg = tiny_dac_cfg()
    hc = DacConfig(encoder_hidden_size=8, downsampling_ratios=[2, 4, 8, 8], decoder_hidden_size=cfg.decoder_hidden_size,
                   n_codebooks=cfg.n_codebooks, codebook_size=cfg.codebook_size, codebook_dim=cfg.codebook_dim,
                   hidden_size=cfg.hidden_size, upsampling_ratios=cfg.upsampling_ratios, sampling_rate=44100)
    m = DacModel(hc).eval()
    w = make_dac_weights(cfg, seed=2)
    sd = m.state_dict()
    loaded = 0
    for k, v in w.items():
        assert k in sd, k
        assert sd[k].shape == v.shape, (k, sd[k].shape, v.shape)
        sd[k] = v
        loaded += 1
    m.load_state_dict(sd)
    dec_keys = [k for k in m.state_dict() if k.startswith("decoder.") or (k.startswith("quantizer.") and ("codebook" in k or "out_proj" in k))]
    assert set(dec_keys) == set(w.keys()), set(dec_keys) ^ set(w.keys())
    g = torch.Generator().manual_seed(4)
    codes = torch.randint(0, cfg.codebook_size, (2, cfg.n_codebooks, 11), generator=g)
    with torch.no_grad():
        z = m.quantizer.from_codes(codes)[0]
        audio = m.decode(audio_codes=codes).audio_values
    np.savez_compressed(os.path.join(HERE, "dac_decode.npz"), codes=codes.numpy(), z=z.numpy(), audio=audio.numpy())


def gen_warpers():
    """The warper / processor classes GenerationMixin._sample applies around the Parler processor, executed from the installed
    transformers (5.5.0: same arithmetic as the 4.46.1 the reference pins, logits_process.py:225-233, 297-299, 521-533,
    581-586).  Pins oracle.sampling.{min_new_tokens, temperature, top_k, top_p}."""
    from transformers.generation.logits_process import (MinNewTokensLengthLogitsProcessor, TemperatureLogitsWarper,
                                                        TopKLogitsWarper, TopPLogitsWarper)
    g = torch.Generator().manual_seed(7)
    R, V, eos = 6, 96, 64
    out = {"meta": np.array([R, V, eos])}
    scores = torch.randn(R, V, generator=g) * 3.0
    scores[1, 10:14] = scores[1, 9]          # ties around the k-th value
    scores[2] = scores[2].round()            # many duplicates
    scores[3, :50] = -float("inf")           # already-masked entries (what the Parler processor leaves behind)
    out["scores"] = scores.numpy()
    ids = torch.zeros(R, 5, dtype=torch.long)
    for n, (cur, mn) in enumerate([(1, 3), (3, 3), (4, 3), (2, 10)]):   # cur_len - prompt_len(1) < min_new -> EOS masked
        proc = MinNewTokensLengthLogitsProcessor(prompt_length_to_skip=1, min_new_tokens=mn, eos_token_id=eos)
        out[f"minnew{n}_args"] = np.array([cur, mn])
        out[f"minnew{n}"] = proc(ids[:, :cur], scores.clone()).numpy()
    for n, t in enumerate([0.7, 1.3]):
        out[f"temp{n}_arg"] = np.array([t], dtype=np.float64)
        out[f"temp{n}"] = TemperatureLogitsWarper(t)(ids, scores.clone()).numpy()
    for n, k in enumerate([1, 5, 50, 200]):
        out[f"topk{n}_arg"] = np.array([k])
        out[f"topk{n}"] = TopKLogitsWarper(top_k=k)(ids, scores.clone()).numpy()
    for n, pp in enumerate([0.1, 0.5, 0.9, 0.999]):
        out[f"topp{n}_arg"] = np.array([pp], dtype=np.float64)
        out[f"topp{n}"] = TopPLogitsWarper(top_p=pp)(ids, scores.clone()).numpy()
    # the chain as _sample applies it: temperature -> top-k -> top-p
    chain = TopPLogitsWarper(top_p=0.8)(ids, TopKLogitsWarper(top_k=20)(ids, TemperatureLogitsWarper(0.9)(ids, scores.clone())))
    out["chain"] = chain.numpy()
    np.savez_compressed(os.path.join(HERE, "warpers.npz"), **out)


def gen_sample_loop(pt):
    """GenerationMixin._sample of the installed transformers (5.5.0; token selection, pad-after-EOS and stopping logic are
    the same statements as in the 4.46.1 the reference calls at modeling_parler_tts.py:3564) driven by SCRIPTED logits through
    a stub model, with the processor list in the order `_get_logits_processor` builds it for the reference
    ([MinNewTokens, ParlerTTSLogitsProcessor (reference class, executed), Temperature, TopK, TopP]) and the stopping criteria
    [MaxLength, EosToken].  Pins oracle.sampling.generate_tokens' loop glue (greedy and sampled)."""
    import contextlib
    from types import SimpleNamespace
    from transformers import GenerationConfig
    from transformers.generation.utils import GenerationMixin
    from transformers.generation.logits_process import (LogitsProcessorList, MinNewTokensLengthLogitsProcessor,
                                                        TemperatureLogitsWarper, TopKLogitsWarper, TopPLogitsWarper)
    from transformers.generation.stopping_criteria import StoppingCriteriaList, MaxLengthCriteria, EosTokenCriteria
    from parler_tts.logits_processors import ParlerTTSLogitsProcessor

    B, K, V, eos, pad, bos = 2, 4, 96, 64, 64, 65
    L = 16                                     # max_length incl. the BOS column
    g = torch.Generator().manual_seed(23)
    script = torch.randn(L, B * K, V, generator=g) * 2.0
    script[:, :, eos] -= 4.0                   # EOS unlikely unless planted
    script[:, :, bos] = -30.0
    # planted EOS preferences: sample 0 finishes early through the cascade, sample 1 runs into max_length
    for step, row in [(3, 0), (5, 1), (6, 2), (8, 3), (9, 4)]:
        script[step, row, eos] = 25.0

    class Stub(GenerationMixin):
        config = SimpleNamespace(is_encoder_decoder=False)

        def __init__(self):
            self.calls = 0

        def _logits(self):
            out = SimpleNamespace(logits=script[self.calls][:, None, :].clone())
            self.calls += 1
            return out

        def _valid_auto_compile_criteria(self, *a, **k):
            return False

        def _prefill(self, input_ids, generation_config, model_kwargs, **k):
            return self._logits()

        def prepare_inputs_for_generation(self, input_ids, **k):
            return {}

        def __call__(self, **k):
            return self._logits()

        def _optimize_model_for_decode(self):
            return contextlib.nullcontext()

        def _update_model_kwargs_for_generation(self, outputs, model_kwargs, **k):
            return model_kwargs

    out = {"meta": np.array([B, K, V, eos, pad, bos, L]), "script": script.numpy()}
    cases = [dict(do_sample=False, min_new=0), dict(do_sample=False, min_new=6),
             dict(do_sample=True, min_new=2, temperature=0.8, top_k=12, top_p=0.9, seed=5)]
    for n, c in enumerate(cases):
        procs = [MinNewTokensLengthLogitsProcessor(1, c["min_new"], eos)] if c["min_new"] > 0 else []
        procs.append(ParlerTTSLogitsProcessor(eos, K, B, "cpu"))
        if c["do_sample"]:
            procs += [TemperatureLogitsWarper(c["temperature"]), TopKLogitsWarper(top_k=c["top_k"]), TopPLogitsWarper(top_p=c["top_p"])]
        gc = GenerationConfig(do_sample=c["do_sample"], pad_token_id=pad, eos_token_id=eos, bos_token_id=bos, max_length=L)
        gc._pad_token_tensor = torch.tensor(pad)
        if not hasattr(gc, "is_assistant"):
            gc.is_assistant = False
        stub = Stub()
        ids = torch.full((B * K, 1), bos, dtype=torch.long)
        if c["do_sample"]:
            torch.manual_seed(c["seed"])
        seq = stub._sample(ids, LogitsProcessorList(procs), StoppingCriteriaList([MaxLengthCriteria(L), EosTokenCriteria(eos)]),
                           gc, synced_gpus=False, streamer=None, use_cache=True)
        out[f"case{n}_cfg"] = np.array([int(c["do_sample"]), c["min_new"], c.get("top_k", 0), c.get("seed", 0)])
        out[f"case{n}_fcfg"] = np.array([c.get("temperature", 1.0), c.get("top_p", 1.0)])
        out[f"case{n}_seq"] = seq.numpy()
        out[f"case{n}_calls"] = np.array([stub.calls])
    np.savez_compressed(os.path.join(HERE, "sample_loop.npz"), **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    gen_warpers()
    pt = import_reference()
    gen_delay(pt)
    gen_logits_processor(pt)
    gen_sample_loop(pt)
    gen_decoder(pt)
    gen_dac()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
