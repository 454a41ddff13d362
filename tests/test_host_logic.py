"""CPU-side tests: the C-ABI library loads and exports every symbol include/ptts_b200.h declares, host-side
argument validation through the ABI (no compute calls), weight-name handling, and the multi-process
(gloo, world_size 2) batch-shard + weight-broadcast plumbing."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from parler_tts_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "ptts_b200.h")).read()
    declared = set(re.findall(r"\b(ptts_[a-z0-9_]+)\s*\(", hdr))
    lib = _lib.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.ptts_version() >= 100


def test_abi_validation_errors_map_to_valueerror():
    from parler_tts_b200 import _lib, ParlerTTSDecoderConfig
    from parler_tts_b200.modeling import _decoder_config_c
    lib = _lib.lib()
    good = _decoder_config_c(ParlerTTSDecoderConfig(vocab_size=1088, num_codebooks=9, max_position_embeddings=4096,
                                                   pad_token_id=1024, eos_token_id=1024, bos_token_id=1025), torch.bfloat16)
    n = C.c_int64()
    _lib.check(lib.ptts_decoder_blob_bytes(C.byref(good), C.byref(n)))
    # SURVEY: Mini decoder ~0.85 GB in bf16 (incl. prefill-only K/V projections + tables); the blob holds the layer matrices three
    # times, each in the order its consumer streams it: mma fragments (step.cu / gemm.cu), row-major (tcgen05 prefill GEMM) and
    # one slice per (phase, cluster, rank) (cluster step kernel) -- 2.37 GB of the 180
    assert 2.3e9 < n.value < 2.45e9
    bad = _decoder_config_c(ParlerTTSDecoderConfig(vocab_size=1088, num_codebooks=9), torch.bfloat16)
    bad.vocab_size = 1001
    with pytest.raises(ValueError, match="vocab_size"):
        _lib.check(lib.ptts_decoder_blob_bytes(C.byref(bad), C.byref(n)))
    with pytest.raises(ValueError):
        _lib.check(lib.ptts_workspace_bytes(C.byref(good), 0, 4, 8, 16, C.byref(n)))
    _lib.check(lib.ptts_workspace_bytes(C.byref(good), 32, 32, 64, 32 + 257, C.byref(n)))
    kv = 98304 * 32 * (32 + 257)  # kv_tok x B x Tmax  
    assert n.value > kv
    with pytest.raises(ValueError, match="head_dim"):
        _decoder_config_c(ParlerTTSDecoderConfig(hidden_size=1024, num_attention_heads=8), torch.bfloat16)
    with pytest.raises(ValueError, match="dtype"):
        _lib.dtype_code(torch.float16)


def test_cpu_tensors_are_rejected_not_silently_computed():
    from parler_tts_b200 import apply_delay_pattern_mask
    with pytest.raises(ValueError, match="CUDA"):
        apply_delay_pattern_mask(torch.zeros(4, 3, dtype=torch.long), torch.zeros(4, 8, dtype=torch.long))


def test_dac_weight_norm_fold_and_descript_keys():
    from parler_tts_b200.dac_wrapper import _fold_weight_norm, _from_descript_keys, _dac_tensor_list
    from parler_tts_b200 import DACConfig
    g = torch.Generator().manual_seed(0)
    v = torch.randn(6, 4, 7, generator=g)
    gg = torch.rand(6, 1, 1, generator=g) + 0.5
    conv = torch.nn.utils.parametrizations.weight_norm(torch.nn.Conv1d(4, 6, 7))
    with torch.no_grad():
        conv.parametrizations.weight.original0.copy_(gg)
        conv.parametrizations.weight.original1.copy_(v)
    folded = _fold_weight_norm({"decoder.model.0.weight_g": gg, "decoder.model.0.weight_v": v, "decoder.model.0.bias": torch.zeros(6)})
    assert torch.allclose(folded["decoder.model.0.weight"], conv.weight, atol=1e-6)
    cfg = DACConfig(num_codebooks=2, decoder_rates=(2, 2))
    sd = {"decoder.model.0.weight": 0, "decoder.model.0.bias": 0, "decoder.model.3.alpha": 0, "decoder.model.4.weight": 0, "decoder.model.4.bias": 0}
    for b in (1, 2):
        sd[f"decoder.model.{b}.block.0.alpha"] = 0
        sd[f"decoder.model.{b}.block.1.weight"] = 0
        sd[f"decoder.model.{b}.block.1.bias"] = 0
        for r in (2, 3, 4):
            for u, nm in ((0, "alpha"), (1, "weight"), (1, "bias"), (2, "alpha"), (3, "weight"), (3, "bias")):
                sd[f"decoder.model.{b}.block.{r}.block.{u}.{nm}"] = 0
    for i in range(2):
        for nm in ("codebook.weight", "out_proj.weight", "out_proj.bias"):
            sd[f"quantizer.quantizers.{i}.{nm}"] = 0
    mapped = _from_descript_keys(sd, 2)
    assert set(mapped) == set(_dac_tensor_list(cfg))


def test_shard_range_partitions_batch():
    from parler_tts_b200.dist import shard_range
    for n in (1, 7, 32, 256):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _dist_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from parler_tts_b200.dist import broadcast_blob, shard_batch, gather_ragged_audio
    blob = torch.arange(1000, dtype=torch.uint8) if rank == 0 else torch.zeros(1000, dtype=torch.uint8)
    broadcast_blob(blob)
    ok_blob = bool((blob == torch.arange(1000, dtype=torch.uint8)).all())
    batch = {"encoder_outputs": torch.arange(10).float()[:, None].repeat(1, 3), "flag": 7}
    mine = shard_batch(batch, rank, world)
    audio = mine["encoder_outputs"][:, :2].clone()
    lengths = [1 + (int(v) % 2) for v in mine["encoder_outputs"][:, 0]]
    full = gather_ragged_audio(audio, lengths)
    # broadcast_model_weights: rank 0 has loaded a checkpoint, rank 1 only constructed the model -- both must issue the SAME
    # collectives (a rank that skipped `is not None` branches hung NCCL) and rank 1 must learn what was loaded
    from types import SimpleNamespace
    from parler_tts_b200.dist import broadcast_model_weights, shard_row_base
    def fake(loaded):
        fill = (lambda n, v: torch.full((n,), v, dtype=torch.uint8))
        return SimpleNamespace(decoder=SimpleNamespace(engine=SimpleNamespace(blob=fill(64, 3 if loaded else 0))),
                               audio_encoder=SimpleNamespace(blob=fill(32, 5 if loaded else 0), loaded=loaded),
                               embed_prompts_weight=torch.full((4, 8), 2.0 if loaded else 0.0),
                               enc_to_dec_proj=(torch.full((8, 6), 1.5 if loaded else 0.0), torch.full((8,), 0.5 if loaded else 0.0)),
                               _side_loaded=loaded)
    m = fake(rank == 0)
    broadcast_model_weights(m)
    ok_model = (bool((m.decoder.engine.blob == 3).all()) and bool((m.audio_encoder.blob == 5).all()) and bool((m.embed_prompts_weight == 2).all())
                and bool((m.enc_to_dec_proj[0] == 1.5).all()) and bool((m.enc_to_dec_proj[1] == 0.5).all()) and m.audio_encoder.loaded and m._side_loaded)
    assert shard_row_base(10, rank, world, 9) == (0 if rank == 0 else 45)
    q.put((rank, ok_blob and ok_model, mine["encoder_outputs"][:, 0].tolist(), mine["flag"], [a.tolist() for a in full]))
    dist.destroy_process_group()


def test_world_size_2_gloo_shard_and_broadcast():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    ps = [ctx.Process(target=_dist_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert res[0][1] and res[1][1]                       # blob arrived intact on rank 1
    assert res[0][2] == [0, 1, 2, 3, 4] and res[1][2] == [5, 6, 7, 8, 9] and res[0][3] == 7
    assert res[0][4] == res[1][4] and len(res[0][4]) == 10  # ragged outputs gathered in global order on every rank
    assert [len(a) for a in res[0][4]] == [1 + (i % 2) for i in range(10)]


def test_generation_config_update_splits_model_kwargs():
    from parler_tts_b200 import GenerationConfig
    gc = GenerationConfig()
    rest = gc.update(do_sample=False, max_new_tokens=12, prompt_input_ids="x")
    assert gc.do_sample is False and gc.max_new_tokens == 12 and rest == {"prompt_input_ids": "x"}


def test_layernorm_fold_algebra():
    """The identity ptts_decoder_finalize (gemm.cu fold_layernorm_kernel) and ln_stats.cuh rely on, in float64 numpy:
    LN(x) W^T == rstd * (x W'^T - mean * c1) + c2 with W' = gamma*W, c1 = rowsum(W'), c2 = W beta, and the row statistics
    taken as mean = S1/K, var = S2/K - mean^2 from S1 = x.1 and S2 = diag(x x^T) (what the tensor-core pass accumulates)."""
    rng = np.random.default_rng(3)
    K, N, M, eps = 64, 24, 5, 1e-5
    x = rng.normal(size=(M, K)) * 2.0 + rng.normal(size=(M, 1)) * 3.0   # rows with a non-zero mean
    gamma, beta, W = rng.normal(size=K) + 1.0, rng.normal(size=K), rng.normal(size=(N, K))
    mu, var = x.mean(1, keepdims=True), x.var(1, keepdims=True)
    ref = (((x - mu) / np.sqrt(var + eps)) * gamma + beta) @ W.T
    Wp, c2 = W * gamma, W @ beta
    c1 = Wp.sum(1)
    S1, S2 = x @ np.ones(K), np.diag(x @ x.T)
    mean = S1 / K
    rstd = 1.0 / np.sqrt(np.maximum(S2 / K - mean * mean, 0.0) + eps)
    got = rstd[:, None] * (x @ Wp.T - mean[:, None] * c1) + c2
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9)


def test_incremental_codec_decode_equals_full_decode():
    """parler_tts_b200.incremental (host logic) driven by the CPU oracle DAC as decode_fn: the concatenated
    chunks equal one decode of the whole sequence; a radius one frame too small does not (so the derived radius is tight
    enough to be meaningful)."""
    from oracle.config import tiny_dac_cfg
    from oracle.dac import OracleDAC
    from oracle.weights import make_dac_weights
    from parler_tts_b200.incremental import IncrementalDecoder, dac_dependency_radius
    cfg = tiny_dac_cfg()
    dac = OracleDAC(cfg, make_dac_weights(cfg, seed=4))
    hop = int(np.prod(cfg.upsampling_ratios))
    R = dac_dependency_radius(cfg.upsampling_ratios)
    assert R == dac_dependency_radius([8, 8, 4, 2]) == 10        # the 44.1 kHz DAC stack
    g = torch.Generator().manual_seed(0)
    B, K, T = 2, cfg.n_codebooks, 61
    codes = torch.randint(0, cfg.codebook_size, (B, K, T), generator=g)
    decode_fn = lambda c: dac.decode(c[None])[:, 0]
    full = decode_fn(codes)

    def run(radius, chunks):
        inc = IncrementalDecoder(decode_fn, hop, radius)
        outs, t = [], 0
        for n in chunks:
            o = inc.push(codes[..., t:t + n]); t += n
            if o is not None:
                outs.append(o)
        o = inc.finish()
        if o is not None:
            outs.append(o)
        assert t == T
        return torch.cat(outs, dim=-1), inc

    for chunks in ([1] * T, [7, 3, 20, 1, 1, 12, 17], [T], [30, 31]):
        got, inc = run(R, chunks)
        assert got.shape == full.shape
        assert torch.allclose(got, full, atol=5e-5, rtol=0), (chunks, float((got - full).abs().max()))  # fp reassociation of torch's conv only
        assert inc.codes.shape[-1] <= R + max(chunks) + R      # bounded state: O(T) total work
    short, _ = run(R - 4, [5] * 12 + [1])
    assert float((short - full).abs().max()) > 1e-3


def test_streamer_incremental_mode_emits_the_full_decode(monkeypatch):
    """ParlerTTSStreamer(incremental=True): host logic only -- the CUDA delay-pattern ops and the DAC are replaced by the CPU
    oracle, a scripted token stream stands in for generate().  The queued chunks concatenate to exactly one decode of all
    frames, and no chunk is ever revised (the reference streamer re-decodes the whole history every `play_steps`)."""
    from types import SimpleNamespace
    import parler_tts_b200.streamer as S
    from oracle.config import tiny_dac_cfg
    from oracle.dac import OracleDAC
    from oracle.weights import make_dac_weights
    from oracle.delay_pattern import build_delay_pattern_mask as obuild, apply_delay_pattern_mask as oapply
    monkeypatch.setattr(S, "build_delay_pattern_mask",
                        lambda ids, bos, pad, L, K: tuple(torch.from_numpy(np.asarray(a)) for a in obuild(ids.numpy(), bos, pad, L, K)))
    monkeypatch.setattr(S, "apply_delay_pattern_mask", lambda ids, m: torch.from_numpy(oapply(ids.numpy(), m.numpy())))
    cfg = tiny_dac_cfg()
    dac = OracleDAC(cfg, make_dac_weights(cfg, seed=6))
    K, cs, bos, eos = cfg.n_codebooks, cfg.codebook_size, 65, 64
    enc = SimpleNamespace(config=SimpleNamespace(codebook_size=cs, decoder_rates=list(cfg.upsampling_ratios), sampling_rate=44100, frame_rate=86),
                          decode=lambda audio_codes, **kw: SimpleNamespace(audio_values=dac.decode(audio_codes)))
    model = SimpleNamespace(decoder=SimpleNamespace(num_codebooks=K), audio_encoder=enc, device="cpu", use_audio_scales=False,
                            use_4dim_audio_codes=False,
                            generation_config=SimpleNamespace(bos_token_id=bos, pad_token_id=eos, decoder_start_token_id=bos))
    F = 47
    codes = torch.randint(0, cs, (1, K, F), generator=torch.Generator().manual_seed(2))
    steps = F + K
    raw = torch.full((K, 1 + steps), eos, dtype=torch.long)
    raw[:, 0] = bos
    for k in range(K):
        for t in range(1, 1 + steps):
            f = t - 1 - k
            raw[k, t] = codes[0, k, f] if 0 <= f < F else (7 if f < 0 else eos)   # junk where the delay mask will put BOS
    st = S.ParlerTTSStreamer(model, device="cpu", play_steps=5, incremental=True)
    st.put(raw[:, :1])
    for t in range(1, 1 + steps):
        st.put(raw[:, t])
    st.end()
    chunks = list(st)
    got = np.concatenate(chunks)
    full = dac.decode(codes[None])[0, 0].numpy()
    assert got.shape == full.shape
    assert np.abs(got - full).max() < 5e-5
    assert sum(len(c) > 0 for c in chunks) >= 4          # audio flowed while tokens were still arriving
