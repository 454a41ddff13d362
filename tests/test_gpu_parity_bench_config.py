"""GPU parity at the BENCHMARKED configuration (-m gpu): BASELINE.json configs[1] (Parler-TTS-Mini, bf16, batch 32, S=64, P=32).

VERDICT r01 "Next round" item 1: the dtype/shape that bench.py times needs its own oracle evidence, not only fp32 / tiny shapes:
  (a) free-running greedy decode of the fused step kernel against OracleDecoder(bf16) over >= 128 steps, tokens asserted equal
      until the oracle's own top-2 margin drops below the measured logit noise (per utterance: a different token changes the
      history of all its codebooks), logit error bound tightened to what is measured and written to gpurun_out/;
  (b) the tcgen05 DAC path against OracleDAC(bf16) at the 44.1 kHz shape, and sample-for-sample against the SIMT path at 32 x 248;
  (c) a fully-masked cross-attention row (quirk Q8).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle.config import tiny_cfg, tiny_dac_cfg, mini_cfg, dac_cfg
from oracle.weights import make_decoder_weights, make_dac_weights
from oracle.decoder import OracleDecoder
from oracle.dac import OracleDAC
from oracle.sampling import generate_tokens
from tests.helpers import build_product_model, synth_inputs, rms, product_dac_config

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _note(name, payload):
    os.makedirs("gpurun_out", exist_ok=True)
    path = os.path.join("gpurun_out", "parity_r02.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    d[name] = payload
    json.dump(d, open(path, "w"), indent=1)


@pytest.mark.timeout(900)
def test_bench_config_bf16_greedy_vs_oracle_128_steps():
    """Mini bf16, B=32, S=64, P=32 (left-padded masks), 128 greedy steps through the fused step kernel against OracleDecoder(bf16).

    With random weights the top-2 gap of 1088 logits is often inside bf16 noise (13 % of the rows per step have a gap below
    1.3 % of the logit scale), so a free-running comparison leaves every utterance's history within a few steps (reported, not
    asserted).  The assertion that survives is the strict one per step: run the kernel on the ORACLE's history (the kernel's own
    greedy choice is compared first, then the oracle's token is appended), require the logits within the measured noise bound,
    and allow the kernel's token to differ from the oracle's only where the oracle's own top-2 margin is below twice that bound.
    All 32 utterances are checked at all 128 steps (cache length 33 -> 161: several K/V ring chunks per warp)."""
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    steps = 128
    cfg = mini_cfg(max_position_embeddings=512)
    w = make_decoder_weights(cfg, seed=91, head_std=0.2)
    dcfg = tiny_dac_cfg(n_codebooks=9, codebook_size=1024)
    model = build_product_model(cfg, dcfg, w, make_dac_weights(dcfg, seed=1), dtype=torch.bfloat16)
    B, S, P = 32, 64, 32
    K = cfg.num_codebooks
    enc, enc_mask, prompt, prompt_mask = synth_inputs(cfg, B, S, P, seed=21, masks=True)
    enc, prompt = enc.bfloat16().float(), prompt.bfloat16().float()
    L = steps + 1
    ref = generate_tokens(OracleDecoder(cfg, w, torch.bfloat16), cfg, enc, enc_mask, prompt, prompt_mask,
                          dict(max_length=L, do_sample=False), collect_logits=True)
    n_steps = min(steps, ref["raw_ids"].shape[1] - 1)

    # (1) free-running, for the record: first step at which an utterance's tokens leave the oracle's
    sess = model.decoder.engine.session(B, P, S, P + L)
    sess.begin(L, do_sample=False)
    sess.prefill(prompt.to(DEV), prompt_mask, enc.to(DEV), enc_mask)
    assert sess.fused >= 1, "the benchmarked configuration must run the fused step kernel"
    kind = sess.fused
    sess.sample()
    sess.decode_steps(n_steps - 1)
    torch.cuda.synchronize()
    free = sess.raw_ids[:, : n_steps + 1].cpu().numpy()
    neq = (free != ref["raw_ids"][:, : n_steps + 1]).reshape(B, K, -1).any(1)
    first_div = [int(np.argmax(r)) - 1 if r.any() else -1 for r in neq]

    # (2) on the oracle's history: logits and greedy choice at every step
    sess.begin(L, do_sample=False)
    sess.prefill(prompt.to(DEV), prompt_mask, enc.to(DEV), enc_mask)
    max_rel_err, flips, worst_flip_margin, rows_total = 0.0, 0, 0.0, 0
    for t in range(n_steps):
        if t > 0:
            sess.decode_forward()
        a = sess.logits.float().cpu().numpy()
        b, sc = ref["logits"][t], ref["scores"][t]
        scale = float(np.abs(b).max())
        max_rel_err = max(max_rel_err, float(np.abs(a - b).max()) / scale)
        a_proc = np.where(np.isneginf(sc), -np.inf, a)          # the processors only mask (EOS gating / min_new_tokens)
        srt = np.sort(sc, axis=-1)
        margin = (srt[:, -1] - srt[:, -2]) / scale
        want = sc.argmax(-1)
        bad = a_proc.argmax(-1) != want
        flips += int(bad.sum())
        rows_total += bad.size
        if bad.any():
            worst_flip_margin = max(worst_flip_margin, float(margin[bad].max()))
        sess.sample(forced=torch.from_numpy(ref["raw_ids"][:, t + 1].copy()))
    torch.cuda.synchronize()
    assert np.array_equal(sess.raw_ids[:, : n_steps + 1].cpu().numpy(), ref["raw_ids"][:, : n_steps + 1])
    _note("bench_config_bf16_vs_oracle", dict(steps=n_steps, B=B, S=S, P=P, fused_kind=kind, max_rel_logit_err=max_rel_err,
                                              greedy_flips=flips, rows_checked=rows_total, worst_margin_at_a_flip=worst_flip_margin,
                                              free_running_first_divergence_step=first_div))
    print(f"\n[parity] bf16 bench config (fused kind {kind}): max |logit err| / max|logit| = {max_rel_err:.4f} over {n_steps} steps; "
          f"{flips} of {rows_total} greedy choices differ from the oracle's, all at oracle top-2 margins <= {worst_flip_margin:.4f} of the "
          f"logit scale; free-running histories leave the oracle's at steps {sorted(set(first_div))}")
    # measured on B200 (gpurun_out/parity_r02.json): 0.0195 of the largest logit
    assert max_rel_err < 0.03, max_rel_err
    # a token may only differ where the oracle's own decision was inside the noise: margin below twice the logit error
    assert worst_flip_margin < 2 * max(max_rel_err, 0.01), (worst_flip_margin, max_rel_err)
    assert flips < 0.08 * rows_total, (flips, rows_total)


@pytest.mark.timeout(600)
def test_dac_tensor_core_vs_bf16_oracle_real_shape(monkeypatch):
    """tcgen05 implicit-GEMM DAC at the 44.1 kHz shape against OracleDAC run in bf16 (torch CPU, same rounding points) and fp32:
    the CUDA path's error against the fp32 truth must not exceed 1.5x the error torch's own bf16 run makes."""
    from parler_tts_b200 import DACModel
    dcfg = dac_cfg()
    w = make_dac_weights(dcfg, seed=3)
    g = torch.Generator().manual_seed(9)
    codes = torch.randint(0, 1024, (2, 9, 12), generator=g)
    ref32 = OracleDAC(dcfg, w).decode(codes[None]).numpy()
    ref16 = OracleDAC(dcfg, w, torch.bfloat16).decode(codes[None]).float().numpy()
    monkeypatch.setenv("PTTS_DAC_TC", "1")
    m = DACModel(product_dac_config(dcfg), DEV, torch.bfloat16).load_state_dict(w)
    got = m.decode(codes[None].to(DEV), [None]).audio_values.float().cpu().numpy()
    e_cpu, e_gpu, e_cross = rms(ref16 - ref32), rms(got - ref32), rms(got - ref16)
    _note("dac_tc_vs_bf16_oracle", dict(rms_ref=rms(ref32), err_cpu_bf16_vs_fp32=e_cpu, err_gpu_tc_vs_fp32=e_gpu, err_gpu_tc_vs_cpu_bf16=e_cross))
    print(f"\n[parity] DAC bf16: rms(ref) {rms(ref32):.4f}; torch-bf16 vs fp32 {e_cpu:.5f}; tcgen05 vs fp32 {e_gpu:.5f}; tcgen05 vs torch-bf16 {e_cross:.5f}")
    assert e_gpu <= 1.5 * e_cpu + 1e-4, (e_gpu, e_cpu)


@pytest.mark.timeout(600)
def test_dac_tensor_core_equals_simt_bench_shape(monkeypatch):
    """The bench's DAC workload (32 utterances x 248 frames = 4.06 M samples): tcgen05 path vs the SIMT bf16 path, sample for sample."""
    from parler_tts_b200 import DACModel
    dcfg = dac_cfg()
    w = make_dac_weights(dcfg, seed=4)
    g = torch.Generator().manual_seed(10)
    codes = torch.randint(0, 1024, (32, 9, 248), generator=g)
    outs = {}
    for tc in ("1", "0"):
        monkeypatch.setenv("PTTS_DAC_TC", tc)
        m = DACModel(product_dac_config(dcfg), DEV, torch.bfloat16).load_state_dict(w)
        outs[tc] = m.decode(codes[None].to(DEV), [None]).audio_values.float().cpu().numpy()
        del m
    a, b = outs["1"], outs["0"]
    assert a.shape == b.shape == (32, 1, 248 * 512)
    d = np.abs(a - b)
    _note("dac_tc_vs_simt_32x248", dict(rms_simt=rms(b), rms_diff=rms(a - b), max_diff=float(d.max())))
    print(f"\n[parity] DAC 32x248: rms(simt) {rms(b):.4f}, rms(tc - simt) {rms(a - b):.5f}, max |diff| {d.max():.4f}")
    # both paths round every one of the ~30 layer outputs to bf16 and differ in accumulation order (fp32 FMA chain vs TMEM
    # accumulate): they sit as far from each other as either sits from the fp32 result (3 % of the signal rms measured,
    # the same as torch's own bf16 run against fp32 in the test above)
    assert rms(a - b) < 0.05 * rms(b) + 1e-4, (rms(a - b), rms(b))


def test_fully_masked_description_row_q8():
    """Quirk Q8: an utterance whose description mask is all zeros.  The reference multiplies the encoder states by the mask
    (:3092-3093) and masks them in cross-attention with finfo.min (:1693), which makes the softmax uniform over ALL keys of a
    fully-masked row; K/V of zeroed states are zero, so the cross-attention output is exactly 0 either way.  fp32, tokens exact."""
    cfg = tiny_cfg()
    w = make_decoder_weights(cfg, seed=33, head_std=0.5)
    dcfg = tiny_dac_cfg()
    model = build_product_model(cfg, dcfg, w, make_dac_weights(dcfg, seed=1), dtype=torch.float32)
    B, S, P, L = 3, 7, 4, 14
    enc, enc_mask, prompt, prompt_mask = synth_inputs(cfg, B, S, P, seed=5, masks=True)
    enc_mask[1] = 0                       # utterance 1: nothing to attend to
    enc = enc * enc_mask[..., None]
    ref = generate_tokens(OracleDecoder(cfg, w, torch.float32), cfg, enc, enc_mask, prompt, prompt_mask,
                          dict(max_length=L, do_sample=False), collect_logits=True)
    sess = model.decoder.engine.session(B, P, S, P + L)
    sess.begin(L, do_sample=False)
    sess.prefill(prompt.to(DEV), prompt_mask, enc.to(DEV), enc_mask)
    l0 = sess.logits.cpu().numpy()
    assert np.isfinite(l0).all()
    assert np.abs(l0 - ref["logits"][0]).max() < 5e-4
    sess.sample()
    sess.decode_steps(L - 2)
    torch.cuda.synchronize()
    n = ref["raw_ids"].shape[1]
    assert np.array_equal(sess.raw_ids[:, :n].cpu().numpy(), ref["raw_ids"])


def test_sampling_is_shard_invariant():
    """SURVEY 8(e) / ADVICE r01: Philox substreams are keyed by the GLOBAL row (ptts_gen_params.row_base), so an utterance draws
    the same tokens whether it is generated in a batch of 4 on one GPU or as the second half of a 2 x 2 shard."""
    cfg = tiny_cfg()
    w = make_decoder_weights(cfg, seed=44, head_std=0.5)
    dcfg = tiny_dac_cfg()
    model = build_product_model(cfg, dcfg, w, make_dac_weights(dcfg, seed=1), dtype=torch.float32)
    B, S, P, L = 4, 6, 3, 18
    K = cfg.num_codebooks
    enc, enc_mask, prompt, prompt_mask = synth_inputs(cfg, B, S, P, seed=8, masks=True)
    gen = dict(do_sample=True, top_k=10, temperature=0.8, seed=123)

    def run(lo, hi):
        sess = model.decoder.engine.session(hi - lo, P, S, P + L)
        sess.begin(L, row_base=lo * K, **gen)
        sess.prefill(prompt[lo:hi].to(DEV), prompt_mask[lo:hi], enc[lo:hi].to(DEV), enc_mask[lo:hi])
        sess.sample()
        sess.decode_steps(L - 2)
        torch.cuda.synchronize()
        return sess.raw_ids[:, : int(sess.state[0].item())].cpu().numpy().copy()

    whole = run(0, 4)
    lo, hi = run(0, 2), run(2, 4)
    n = min(whole.shape[1], lo.shape[1], hi.shape[1])
    assert n >= 6
    # (lengths can differ after an EOS cascade: a shard stops when ITS rows are finished; compare the common prefix)
    assert np.array_equal(whole[: 2 * K, :n], lo[:, :n])
    assert np.array_equal(whole[2 * K:, :n], hi[:, :n])
    assert not np.array_equal(lo[:, 1:n], hi[:, 1:n])  # different utterances really draw different streams


def test_large_layer_shape_fused_step_vs_oracle():
    """BASELINE configs[3] layer shape (Parler-TTS-Large: H 1536, 24 heads, F 6144; helpers/model_init_scripts/init_large_model.py:25-43),
    bf16, 32 rows, 2 layers: the fused step kernel (step.cu -- the cluster kernel's shape range ends at H 1024) teacher-forced on the
    oracle's greedy history, logits within the bf16 bound measured at the Mini shape."""
    cfg = mini_cfg(num_hidden_layers=2, max_position_embeddings=64, hidden_size=1536, num_attention_heads=24, ffn_dim=6144)
    w = make_decoder_weights(cfg, seed=6, head_std=0.2)
    dcfg = tiny_dac_cfg(n_codebooks=cfg.num_codebooks, codebook_size=1024)
    model = build_product_model(cfg, dcfg, w, make_dac_weights(dcfg, seed=1), dtype=torch.bfloat16)
    B, S, P, steps = 32, 12, 6, 6
    enc, enc_mask, prompt, prompt_mask = synth_inputs(cfg, B, S, P, seed=5)
    enc, prompt = enc.bfloat16().float(), prompt.bfloat16().float()
    L = steps + 1
    ref = generate_tokens(OracleDecoder(cfg, w, torch.bfloat16), cfg, enc, enc_mask, prompt, prompt_mask,
                          dict(max_length=L, do_sample=False), collect_logits=True)
    sess = model.decoder.engine.session(B, P, S, P + L)
    sess.begin(L, do_sample=False)
    sess.prefill(prompt.to(DEV), prompt_mask, enc.to(DEV), enc_mask)
    assert sess.fused >= 1, "a fused step kernel must cover the Large layer shape at 32 rows"
    worst = 0.0
    for t in range(steps):
        if t > 0:
            sess.decode_forward()
        a, b = sess.logits.float().cpu().numpy(), ref["logits"][t]
        worst = max(worst, float(np.abs(a - b).max()) / float(np.abs(b).max()))
        sess.sample(forced=torch.from_numpy(ref["raw_ids"][:, t + 1]).to(DEV))   # stay on the oracle's history
    _note("large_shape", {"fused_kind": int(sess.fused), "max_rel_logit_err": worst, "layers": 2, "rows": B})
    assert worst < 0.03, f"Large-shape bf16 logits differ from the oracle by {worst:.4f} of the logit scale"


def test_generate_batch_above_one_tile_runs_as_shards(monkeypatch):
    """generate() runs a batch larger than the fused kernels' 32-row tile as consecutive shards; the result
    must be the one the whole batch gives (tokens, pad tail of early finishers, waveform).  fp32 makes both sides bit-exact: the
    tile limit is forced to 2 and 3 rows on a batch of 5 (ragged last shard) and compared with the unsharded run."""
    cfg = tiny_cfg()
    w = make_decoder_weights(cfg, seed=45, head_std=0.5)
    dcfg = tiny_dac_cfg()
    model = build_product_model(cfg, dcfg, w, make_dac_weights(dcfg, seed=2), dtype=torch.float32)
    B, S, P = 5, 6, 3
    enc, enc_mask, prompt, prompt_mask = synth_inputs(cfg, B, S, P, seed=9, masks=True)
    kw = dict(encoder_outputs=(enc.to(DEV),), attention_mask=enc_mask, prompt_hidden_states=prompt.to(DEV), prompt_attention_mask=prompt_mask,
              do_sample=True, top_k=12, temperature=0.9, seed=77, max_new_tokens=24, return_codes=True)
    assert model._fused_batch_limit() is None          # fp32: one session for the whole batch
    audio0, out0 = model.generate(**kw)
    for limit in (2, 3):
        monkeypatch.setattr(model, "_fused_batch_limit", lambda limit=limit: limit)
        audio1, out1 = model.generate(**kw)
        assert torch.equal(out0.raw_ids, out1.raw_ids), f"shards of {limit}: token matrix differs"
        assert torch.equal(out0.audio_codes, out1.audio_codes)
        assert out0.audios_length == out1.audios_length
        assert torch.equal(audio0, audio1)
    _note("batch_shards", {"batch": B, "limits": [2, 3], "generated_columns": int(out0.raw_ids.shape[1])})


def test_causal_lm_forward_step_operator():
    """ParlerTTSForCausalLM.forward as an HF-style loop drives it (reference :1865-1974 + prepare_inputs_for_generation :2909):
    BOS column + conditioning on the first call, then one delay-masked column per call over the returned cache; logits against
    the oracle's cached loop, fp32."""
    from oracle import delay_pattern as odp
    cfg = tiny_cfg()
    w = make_decoder_weights(cfg, seed=51, head_std=0.5)
    dcfg = tiny_dac_cfg()
    model = build_product_model(cfg, dcfg, w, make_dac_weights(dcfg, seed=1), dtype=torch.float32)
    B, S, P, L = 2, 6, 3, 12
    K = cfg.num_codebooks
    enc, enc_mask, prompt, prompt_mask = synth_inputs(cfg, B, S, P, seed=6, masks=True)
    ref = generate_tokens(OracleDecoder(cfg, w, torch.float32), cfg, enc, enc_mask, prompt, prompt_mask,
                          dict(max_length=L, do_sample=False), collect_logits=True)
    lm = model.decoder
    ids = np.full((B * K, 1), cfg.bos_token_id, dtype=np.int64)
    out = lm.forward(input_ids=torch.from_numpy(ids).to(DEV), encoder_hidden_states=enc.to(DEV), encoder_attention_mask=enc_mask.to(DEV),
                     prompt_hidden_states=prompt.to(DEV), prompt_attention_mask=prompt_mask.to(DEV), max_cache_len=P + L + 4)
    assert out.logits.shape == (B * K, 1, cfg.vocab_size)
    assert np.abs(out.logits[:, 0].cpu().numpy() - ref["logits"][0]).max() < 5e-4
    cache = out.past_key_values
    raw = ref["raw_ids"]
    for t in range(1, raw.shape[1] - 1):
        masked = odp.apply_delay_pattern_mask(raw[:, : t + 1], ref["delay_mask"])[:, -1:]   # what prepare_inputs_for_generation feeds
        out = lm(input_ids=torch.from_numpy(masked.copy()).to(DEV), past_key_values=cache)
        assert np.abs(out.logits[:, 0].cpu().numpy() - ref["logits"][t]).max() < 5e-4, t
    assert cache.get_seq_length() == P + raw.shape[1] - 1
    with pytest.raises(ValueError):
        lm.forward(input_ids=torch.zeros(B * K, 2, dtype=torch.long, device=DEV), past_key_values=cache)


def test_generate_with_custom_logits_processor_and_stopping_criteria():
    """generate(logits_processor=[...], stopping_criteria=[...]) (reference merges user lists at :3540-3552): the host-driven loop
    over the same device operators.  A processor that bans one token id and a criterion that stops at 9 columns; greedy, fp32:
    tokens bit-exact against the oracle loop with the same ban."""
    cfg = tiny_cfg()
    w = make_decoder_weights(cfg, seed=52, head_std=0.5)
    dcfg = tiny_dac_cfg()
    model = build_product_model(cfg, dcfg, w, make_dac_weights(dcfg, seed=1), dtype=torch.float32)
    B, S, P, L = 2, 6, 3, 16
    enc, enc_mask, prompt, prompt_mask = synth_inputs(cfg, B, S, P, seed=7, masks=True)
    banned = 5

    def ban(input_ids, scores):
        scores[:, banned] = -float("inf")
        return scores

    seen = []

    def stop_at_9(input_ids, scores):
        seen.append(input_ids.shape[1])
        return input_ids.shape[1] >= 9

    ref = generate_tokens(OracleDecoder(cfg, w, torch.float32), cfg, enc, enc_mask, prompt, prompt_mask, dict(max_length=L, do_sample=False),
                          pick=lambda step, s: np.where(np.arange(s.shape[1])[None, :] == banned, -np.inf, s).argmax(-1))
    _, out = model.generate(encoder_outputs=(enc.to(DEV),), attention_mask=enc_mask.to(DEV), prompt_hidden_states=prompt.to(DEV),
                            prompt_attention_mask=prompt_mask.to(DEV), do_sample=False, max_length=L, logits_processor=[ban],
                            stopping_criteria=[stop_at_9], return_codes=True)
    got = out.raw_ids.cpu().numpy()
    n = min(got.shape[1], ref["raw_ids"].shape[1])
    assert got.shape[1] <= 9 and n >= 2 and seen
    # raw_ids returned by generate() have the delay mask applied (like the reference's output_ids); compare the free cells
    from oracle import delay_pattern as odp
    want = odp.apply_delay_pattern_mask(ref["raw_ids"][:, :n], ref["delay_mask"])
    full_mask = odp.build_delay_pattern_mask(np.full((got.shape[0], 1), cfg.bos_token_id, dtype=np.int64), cfg.bos_token_id, cfg.pad_token_id, L, cfg.num_codebooks)[1]
    free = full_mask[:, :n] == -1
    assert np.array_equal(got[:, :n][free], want[free])
    assert not (got[:, 1:n][free[:, 1:]] == banned).any()
    with pytest.raises(ValueError):
        model.generate(encoder_outputs=(enc.to(DEV),), do_sample=False, max_length=6, repetition_penalty=1.3)
    with pytest.raises(ValueError):
        model.generate(encoder_outputs=(enc.to(DEV),), do_sample=False, max_length=6, not_a_real_kwarg=1)


def test_streamer_incremental_batch_vs_oracle_dac():
    """ParlerTTSStreamer(incremental=True) with a BATCH of utterances (the reference streamer is batch-1 only): the [B, n] chunks
    queued while generate() runs concatenate to the ORACLE codec's decode of the generated codes (fp32, < 1e-3 RMS -- measured
    ~1e-6), i.e. windowed decoding loses nothing, and audio flows before generation ends."""
    from parler_tts_b200 import ParlerTTSStreamer
    cfg, dcfg = tiny_cfg(), tiny_dac_cfg()
    w = make_decoder_weights(cfg, seed=62, head_std=0.5)
    dw = make_dac_weights(dcfg, seed=2)
    model = build_product_model(cfg, dcfg, w, dw, dtype=torch.float32)
    B = 3
    enc, enc_mask, prompt, prompt_mask = synth_inputs(cfg, B, 6, 3, seed=10, masks=True)
    st = ParlerTTSStreamer(model, device=DEV, play_steps=7, incremental=True)
    audio, out = model.generate(encoder_outputs=(enc.to(DEV),), attention_mask=enc_mask.to(DEV), prompt_hidden_states=prompt.to(DEV),
                                prompt_attention_mask=prompt_mask.to(DEV), do_sample=False, max_length=70, streamer=st,
                                return_codes=True, _suppress_special=True)
    chunks = [c for c in st]
    assert all(c.ndim == 2 and c.shape[0] == B for c in chunks)
    total = np.concatenate(chunks, axis=1)
    ref = OracleDAC(dcfg, dw).decode(out.audio_codes.cpu()[None]).numpy()[:, 0]
    assert total.shape == ref.shape, (total.shape, ref.shape)
    assert rms(total - ref) < 1e-3, rms(total - ref)
    assert sum(c.shape[1] > 0 for c in chunks) >= 3
    assert np.abs(total - audio.float().cpu().numpy()).max() < 1e-4   # and equals what generate() itself returned


def test_text_encoder_cuda_graph_matches_eager():
    """the description path (T5 encoder -> enc_to_dec_proj -> mask multiply, reference :3048-3097) replayed from one
    CUDA graph per input shape equals the eager chain; a second call with new ids replays the same graph."""
    from transformers import T5Config, T5EncoderModel
    tc = T5Config(vocab_size=128, d_model=64, d_kv=16, d_ff=128, num_layers=2, num_heads=4, dropout_rate=0.0)
    torch.manual_seed(0)
    enc = T5EncoderModel(tc).to(DEV).eval()
    cfg, dcfg = tiny_cfg(), tiny_dac_cfg()
    model = build_product_model(cfg, dcfg, make_decoder_weights(cfg, seed=63, head_std=0.5), make_dac_weights(dcfg, seed=2), dtype=torch.float32)
    model.text_encoder = enc
    H = cfg.hidden_size
    model.enc_to_dec_proj = (torch.randn(H, 64, device=DEV) * 0.1, torch.randn(H, device=DEV) * 0.1)
    g = torch.Generator().manual_seed(3)
    for trial in range(2):
        ids = torch.randint(0, 128, (2, 9), generator=g)
        mask = torch.ones(2, 9, dtype=torch.long)
        mask[0, :3] = 0
        got = model._encode_text(ids, mask)
        want = model._encode_text_eager(ids.to(DEV), mask.to(DEV))
        hf = torch.nn.functional.linear(enc(input_ids=ids.to(DEV), attention_mask=mask.to(DEV)).last_hidden_state, *model.enc_to_dec_proj) * mask.to(DEV)[..., None]
        assert (want - hf).abs().max().item() < 1e-5   # the device-built 4-D mask equals the library's own conversion of the 2-D mask
        assert got.shape == (2, 9, H)
        assert (got - want).abs().max().item() < 1e-5
        assert (got[0, :3] == 0).all()
    assert model._enc_graph_ok and len(model._enc_graphs) == 1
    # and generate() takes the description through it
    audio = model.generate(input_ids=ids.to(DEV), attention_mask=mask.to(DEV), do_sample=False, max_length=12)
    assert torch.isfinite(audio.float()).all()
