"""The oracle (CPU restatement) against fixtures produced by executing the reference (tests/golden/make_golden.py).

Bar: bit-exact for integer work (delay pattern, processor masks); fp32 forward within 2e-5 abs of the
reference's own no-cache forward (same ops, different evaluation order: cached vs full-prefix).
"""
import os

import numpy as np
import torch

from oracle.config import tiny_cfg, tiny_dac_cfg
from oracle.weights import make_decoder_weights, make_dac_weights
from oracle.decoder import OracleDecoder
from oracle.dac import OracleDAC
from oracle.delay_pattern import build_delay_pattern_mask, apply_delay_pattern_mask, undelay
from oracle.sampling import ParlerLogitsProcessorOracle, generate_tokens, frames_from_raw


def test_delay_pattern_bit_exact(golden_dir):
    z = np.load(os.path.join(golden_dir, "delay_pattern.npz"))
    for ci in range(int(z["n"])):
        B, K, seq, L = z[f"c{ci}_meta"]
        d, m = build_delay_pattern_mask(z[f"c{ci}_ids"], 65, 64, int(L), int(K))
        assert np.array_equal(d, z[f"c{ci}_delayed"]), ci
        assert np.array_equal(m, z[f"c{ci}_mask"]), ci
        assert np.array_equal(apply_delay_pattern_mask(z[f"c{ci}_full"], m), z[f"c{ci}_applied"])
        half = z[f"c{ci}_full"][:, : max(1, int(L) // 2)]
        assert np.array_equal(apply_delay_pattern_mask(half, m), z[f"c{ci}_applied_half"])


def test_delay_pattern_docstring_example():
    # modeling_parler_tts.py:219-224: K=4, L=8
    ids = np.full((4, 1), 9, dtype=np.int64)
    _, m = build_delay_pattern_mask(ids, 9, 7, 8, 4)
    B_, P_ = 9, 7
    assert m.tolist() == [[B_, -1, -1, -1, -1, P_, P_, P_], [B_, B_, -1, -1, -1, -1, P_, P_],
                          [B_, B_, B_, -1, -1, -1, -1, P_], [B_, B_, B_, B_, -1, -1, -1, -1]]


def test_undelay_q13():
    # K=9, 256 steps -> 248 frames; frame[b,k,t] == raw[b*K+k, t+k+1]
    B, K, L = 2, 9, 257
    raw = np.random.default_rng(0).integers(0, 1024, (B * K, L))
    raw[:, 0] = 1025
    _, mask = build_delay_pattern_mask(raw[:, :1], 1025, 1024, L, K)
    fr = undelay(raw, 1025, 1024, K, B, mask)
    assert fr.shape == (B, K, L - K)
    for b in range(B):
        for k in range(K):
            assert np.array_equal(fr[b, k], raw[b * K + k, k + 1: k + 1 + L - K])


def test_logits_processor_bit_exact(golden_dir):
    z = np.load(os.path.join(golden_dir, "logits_processor.npz"))
    B, K, V, eos, steps = z["meta"]
    proc = ParlerLogitsProcessorOracle(int(eos), int(K), int(B))
    for s in range(int(steps)):
        out = proc(z["ids"][:, : s + 1], z["scores_in"][s].copy())
        assert np.array_equal(out, z["scores_out"][s]), s
        assert np.array_equal(proc.first_unfinished, z["first"][s]), s


def _cfg_for(name):
    if name == "abs":
        return tiny_cfg(rope_embeddings=False)
    if name == "rope":
        return tiny_cfg(rope_embeddings=True)
    return tiny_cfg(rope_embeddings=True, num_attention_heads=4, num_key_value_heads=2,
                    num_cross_attention_key_value_heads=1, hidden_size=256)


def test_cached_decoder_matches_reference_forward(golden_dir):
    """Teacher-forced cached loop == the reference's own no-cache forward on every prefix position."""
    z = np.load(os.path.join(golden_dir, "decoder_forward.npz"))
    for name in ("abs", "rope", "gqa"):
        cfg = _cfg_for(name)
        dec = OracleDecoder(cfg, make_decoder_weights(cfg, seed=3), torch.float32)
        ids = torch.from_numpy(z[f"{name}_ids"])
        enc = torch.from_numpy(z[f"{name}_enc"])
        prompt = torch.from_numpy(z[f"{name}_prompt"])
        P = prompt.shape[1]
        for masked in (True, False):
            em = torch.from_numpy(z[f"{name}_enc_mask"]) if masked else None
            pm = torch.from_numpy(z[f"{name}_pmask"]) if masked else None
            ref = z[f"{name}_logits" if masked else f"{name}_logits_nomask"]
            lo = dec.prefill(ids[:, :1], enc, em, prompt, pm)
            got = [lo[:, -1]]
            for t in range(1, ids.shape[1]):
                got.append(dec.step(ids[:, t:t + 1])[:, -1])
            got = torch.stack(got, 1).numpy()
            err = np.abs(got - ref[:, P:]).max()
            assert err < 2e-5, (name, masked, err)
            # multi-token prefill (q > 1) must agree too
            lo2 = dec.prefill(ids[:, :3], enc, em, prompt, pm).numpy()
            valid = np.ones(lo2.shape[1], bool)
            err2 = np.abs(lo2[:, P:] - ref[:, P:P + 3]).max()
            assert err2 < 2e-5, (name, masked, err2)


def test_dac_matches_hf(golden_dir):
    z = np.load(os.path.join(golden_dir, "dac_decode.npz"))
    cfg = tiny_dac_cfg()
    dac = OracleDAC(cfg, make_dac_weights(cfg, seed=2))
    codes = torch.from_numpy(z["codes"])
    zz = dac.from_codes(codes)
    assert np.abs(zz.numpy() - z["z"]).max() < 1e-5
    audio = dac.decode(codes[None])
    assert audio.shape == (2, 1, 11 * 512)
    assert np.abs(audio.numpy()[:, 0] - z["audio"].reshape(2, -1)).max() < 1e-5
    assert float(np.sqrt((z["audio"] ** 2).mean())) > 1e-3  # fixture is not degenerate


def test_generate_loop_smoke():
    cfg = tiny_cfg()
    dec = OracleDecoder(cfg, make_decoder_weights(cfg, seed=1, head_std=0.5), torch.float32)
    g = torch.Generator().manual_seed(0)
    B, S, P = 2, 5, 3
    enc = torch.randn(B, S, cfg.hidden_size, generator=g)
    prompt = torch.randn(B, P, cfg.hidden_size, generator=g)
    out = generate_tokens(dec, cfg, enc, None, prompt, None, dict(max_length=14, do_sample=False))
    raw = out["raw_ids"]
    assert raw.shape[0] == B * cfg.num_codebooks and raw.shape[1] <= 14
    fr = frames_from_raw(raw, out["delay_mask"], cfg, B)
    assert fr.shape == (B, cfg.num_codebooks, raw.shape[1] - cfg.num_codebooks)


def test_warpers_against_transformers_classes(golden_dir):
    """oracle.sampling's restatement of the HF processors / warpers around the Parler processor, bit-exact against the
    outputs of the installed transformers classes (fixture: make_golden.gen_warpers)."""
    from oracle.sampling import min_new_tokens, temperature, top_k, top_p
    z = np.load(os.path.join(golden_dir, "warpers.npz"))
    R, V, eos = (int(v) for v in z["meta"])
    scores = z["scores"]
    n = 0
    while f"minnew{n}" in z:
        cur, mn = (int(v) for v in z[f"minnew{n}_args"])
        assert np.array_equal(min_new_tokens(scores.copy(), cur, 1, mn, eos), z[f"minnew{n}"]), ("min_new", n)
        n += 1
    assert n == 4
    for n in range(2):
        assert np.array_equal(temperature(scores.copy(), float(z[f"temp{n}_arg"][0])), z[f"temp{n}"]), ("temperature", n)
    for n in range(4):
        assert np.array_equal(top_k(scores.copy(), int(z[f"topk{n}_arg"][0])), z[f"topk{n}"]), ("top_k", n)
    for n in range(4):
        assert np.array_equal(top_p(scores.copy(), float(z[f"topp{n}_arg"][0])), z[f"topp{n}"]), ("top_p", n)
    chain = top_p(top_k(temperature(scores.copy(), 0.9), 20), 0.8)
    assert np.array_equal(chain, z["chain"])


def test_sample_loop_against_transformers_sample(golden_dir):
    """oracle.sampling.generate_tokens' loop glue (processor order, pad-after-EOS, EOS / max-length stopping, one multinomial per
    step) against GenerationMixin._sample of the installed transformers run on scripted logits with the reference's own
    ParlerTTSLogitsProcessor (fixture: make_golden.gen_sample_loop) -- greedy, greedy + min_new_tokens, and sampled."""
    from types import SimpleNamespace
    z = np.load(os.path.join(golden_dir, "sample_loop.npz"))
    B, K, V, eos, pad, bos, L = (int(v) for v in z["meta"])
    script = torch.from_numpy(z["script"])

    class ScriptedDecoder:
        def __init__(self):
            self.calls = 0

        def prefill(self, ids, *a):
            return self.step(ids)

        def step(self, ids):
            out = script[self.calls][:, None, :].clone()
            self.calls += 1
            return out

    cfg = SimpleNamespace(num_codebooks=K, bos_token_id=bos, pad_token_id=pad, eos_token_id=eos)
    enc = torch.zeros(B, 1, 8)
    n = 0
    while f"case{n}_seq" in z:
        do_sample, min_new, top_k, seed = (int(v) for v in z[f"case{n}_cfg"])
        temp, top_p = (float(v) for v in z[f"case{n}_fcfg"])
        gen = dict(max_length=L, do_sample=bool(do_sample), min_new_tokens=min_new, temperature=temp, top_k=top_k, top_p=top_p)
        dec = ScriptedDecoder()
        if do_sample:
            torch.manual_seed(seed)
        out = generate_tokens(dec, cfg, enc, None, None, None, gen)
        assert np.array_equal(out["raw_ids"], z[f"case{n}_seq"]), n
        assert dec.calls == int(z[f"case{n}_calls"][0]), n
        n += 1
    assert n == 3
