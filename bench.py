#!/usr/bin/env python
"""bench.py -- audio codec tokens/s (all codebooks) of the generate() hot path on N x B200.

Workload (BASELINE.json configs[1]): Parler-TTS-Mini shape, bf16, batch 32 per GPU, 256 decode steps,
top-k 50 sampling, synthetic inputs (S=64 description states, P=32 prompt prefix, left-padded masks),
random-init weights (no network -> no checkpoints).  One bench "step" = one full pass of the hot path over
one batch: begin + prefill + 255 graph-replayed decode steps + sampling = 32 x 9 x 256 tokens.

  value      : tokens/s with inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        : the same through the public API with HOST (pinned) inputs and the token matrix read back
  roofline   : decode step vs the HBM roofline: algorithmic bytes per step (SURVEY.md 8d formula, T taken
               per step) / measured step time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline: the oracle port (CPU restatement of the reference loop) timed on this box's host cores on a
               bounded sample of the same workload
`--impl reference` times that CPU port alone (the reference itself cannot be imported on the GPU box).
"""
from __future__ import annotations
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MINI = dict(vocab_size=1088, max_position_embeddings=4096, num_hidden_layers=24, ffn_dim=4096, num_attention_heads=16,
            hidden_size=1024, num_codebooks=9, pad_token_id=1024, eos_token_id=1024, bos_token_id=1025)
B_PER_GPU, DECODE_STEPS, S_LEN, P_LEN = 32, 256, 64, 32
W_STEP_PARAMS = 362_498_048          # params streamed per decode step (Mini)
KV_TOK_BYTES = 98_304                # bytes per cached token per sequence (Mini, bf16)


def algorithmic_bytes_per_step(B, K, V, T, S):
    """2*W_step + B*kv_tok*(T + S) + B*kv_tok + 4*B*K*V."""
    return 2 * W_STEP_PARAMS + B * KV_TOK_BYTES * (T + S) + B * KV_TOK_BYTES + 4 * B * K * V


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, idx=0):
        self.idx, self.rows, self.p = idx, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def synthetic_state_dict(cfg, device, seed=0):
    """Reference init (normal(0, 0.02), modeling_parler_tts.py:1093-1102) generated on the GPU, bf16."""
    g = torch.Generator(device=device).manual_seed(seed)
    H, F, V, K = cfg["hidden_size"], cfg["ffn_dim"], cfg["vocab_size"], cfg["num_codebooks"]

    def n(*s):
        return (torch.randn(*s, generator=g, device=device) * 0.02).to(torch.bfloat16)

    sd = {}
    p = "decoder.model.decoder."
    for k in range(K):
        sd[f"{p}embed_tokens.{k}.weight"] = n(V + 1, H)
        sd[f"decoder.lm_heads.{k}.weight"] = n(V, H)
    for i in range(cfg["num_hidden_layers"]):
        q = f"{p}layers.{i}."
        for a in ("self_attn", "encoder_attn"):
            for m in ("q", "k", "v", "out"):
                sd[f"{q}{a}.{m}_proj.weight"] = n(H, H)
        sd[q + "fc1.weight"], sd[q + "fc2.weight"] = n(F, H), n(H, F)
        for ln in ("self_attn_layer_norm", "encoder_attn_layer_norm", "final_layer_norm"):
            sd[q + ln + ".weight"] = torch.ones(H, device=device)
            sd[q + ln + ".bias"] = torch.zeros(H, device=device)
    sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"] = torch.ones(H, device=device), torch.zeros(H, device=device)
    sd["embed_prompts.weight"] = n(32128, H)
    return sd


def synth_dac_weights(cfg, device):
    import math
    g = torch.Generator(device=device).manual_seed(0)
    sd = {}
    C = cfg.decoder_dim
    def rn(*s, scale=1.0): return torch.randn(*s, generator=g, device=device) * scale
    for i in range(cfg.num_codebooks):
        sd[f"quantizer.quantizers.{i}.codebook.weight"] = rn(cfg.codebook_size, cfg.codebook_dim)
        sd[f"quantizer.quantizers.{i}.out_proj.weight"] = rn(cfg.latent_dim, cfg.codebook_dim, 1, scale=1 / math.sqrt(8 * 9))
        sd[f"quantizer.quantizers.{i}.out_proj.bias"] = rn(cfg.latent_dim, scale=0.02)
    sd["decoder.conv1.weight"] = rn(C, cfg.latent_dim, 7, scale=1 / math.sqrt(7 * cfg.latent_dim)); sd["decoder.conv1.bias"] = rn(C, scale=0.02)
    for bi, s in enumerate(cfg.decoder_rates):
        cin, cout = C >> bi, C >> (bi + 1)
        p = f"decoder.block.{bi}."
        sd[p + "snake1.alpha"] = torch.ones(1, cin, 1, device=device)
        sd[p + "conv_t1.weight"] = rn(cin, cout, 2 * s, scale=1 / math.sqrt(2 * cin)); sd[p + "conv_t1.bias"] = rn(cout, scale=0.02)
        for r in (1, 2, 3):
            u = p + f"res_unit{r}."
            sd[u + "snake1.alpha"] = torch.ones(1, cout, 1, device=device)
            sd[u + "conv1.weight"] = rn(cout, cout, 7, scale=0.5 / math.sqrt(7 * cout)); sd[u + "conv1.bias"] = rn(cout, scale=0.02)
            sd[u + "snake2.alpha"] = torch.ones(1, cout, 1, device=device)
            sd[u + "conv2.weight"] = rn(cout, cout, 1, scale=0.5 / math.sqrt(cout)); sd[u + "conv2.bias"] = rn(cout, scale=0.02)
    cl = C >> len(cfg.decoder_rates)
    sd["decoder.snake1.alpha"] = torch.ones(1, cl, 1, device=device)
    sd["decoder.conv2.weight"] = rn(1, cl, 7, scale=1 / math.sqrt(7 * cl)); sd["decoder.conv2.bias"] = rn(1, scale=0.02)
    return sd


def synthetic_inputs(B, H, seed, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    enc_mask = torch.ones(B, S_LEN, dtype=torch.long)
    for b, ln in enumerate(torch.randint(32, S_LEN + 1, (B,), generator=g).tolist()):
        enc_mask[b, : S_LEN - ln] = 0
    enc = (torch.randn(B, S_LEN, H, generator=g) * enc_mask[..., None]).to(torch.bfloat16)
    pmask = torch.ones(B, P_LEN, dtype=torch.long)
    for b, ln in enumerate(torch.randint(16, P_LEN + 1, (B,), generator=g).tolist()):
        pmask[b, : P_LEN - ln] = 0
    prompt = (torch.randn(B, P_LEN, H, generator=g) * 0.02).to(torch.bfloat16)
    ts = [enc, enc_mask, prompt, pmask]
    if pin:
        ts = [t.pin_memory() for t in ts]
    return [t.to(device) for t in ts] if device != "cpu" else ts


def run_reference(args, rank):
    """CPU arm: the oracle port of the reference generate() loop on the host cores (fp32, like configs[0])."""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses all physical host cores regardless of the launcher
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    from oracle.config import mini_cfg
    from oracle.weights import make_decoder_weights
    from oracle.decoder import OracleDecoder
    from oracle.sampling import generate_tokens
    torch.manual_seed(0)
    cfg = mini_cfg()
    Bc, n_dec = 32, 6   # bounded sample: B=32, 6 decode steps (1 prefill + 5 cached) per bench step
    dec = OracleDecoder(cfg, make_decoder_weights(cfg, seed=0), torch.float32)
    enc, enc_mask, prompt, pmask = synthetic_inputs(Bc, cfg.hidden_size, 1)
    gen = dict(max_length=n_dec + 1, do_sample=True, top_k=50, min_new_tokens=n_dec)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        generate_tokens(dec, cfg, enc.float(), enc_mask, prompt.float(), pmask, gen)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    tot = sum(times)
    toks = Bc * cfg.num_codebooks * n_dec * len(times)
    v = toks / tot
    sample = f"Mini fp32 B={Bc}, {n_dec} decode steps (prefill + {n_dec - 1} cached) per step, top-k 50"
    line = {"impl": "reference", "metric": "audio codec tokens/sec (all codebooks)", "value": v, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"Parler-TTS-Mini bf16 batch={B_PER_GPU}/GPU {args.decode_steps} decode steps top-k=50 (BASELINE configs[1])",
                       "global_batch": args.gpus * B_PER_GPU, "prompt_len": P_LEN, "desc_len": S_LEN, "parallelism": f"batch-shard x{args.gpus}",
                       "sample": sample + " -- the reference's CPU code path computes in fp32; same shapes, prompt and description lengths"},
            "cpu_baseline": {"value": v, "unit": "tokens/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample,
                             "os_cpu_count": os.cpu_count()},
            "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# DRAM bytes of one fused decode-step launch from the committed ncu --set full capture (profiles/r01_step_ncu_full.md)
NCU_STEP_DRAM_BYTES = 1171574000 + 10182656


def cpu_baseline_quick():
    """~10-30 s of CPU work: the oracle port on a bounded sample of the same workload (rank 0, N=1)."""
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    from oracle.config import mini_cfg
    from oracle.weights import make_decoder_weights
    from oracle.decoder import OracleDecoder
    from oracle.sampling import generate_tokens
    cfg = mini_cfg()
    Bc, n_dec = 32, 32   # ~0.36 s per decode step on 64 threads -> ~12 s
    dec = OracleDecoder(cfg, make_decoder_weights(cfg, seed=0), torch.float32)
    enc, enc_mask, prompt, pmask = synthetic_inputs(Bc, cfg.hidden_size, 1)
    gen = dict(max_length=n_dec + 1, do_sample=True, top_k=50, min_new_tokens=n_dec)
    generate_tokens(dec, cfg, enc.float(), enc_mask, prompt.float(), pmask, dict(gen, max_length=3, min_new_tokens=2))  # warm-up
    t0 = time.perf_counter()
    generate_tokens(dec, cfg, enc.float(), enc_mask, prompt.float(), pmask, gen)
    dt = time.perf_counter() - t0
    return {"value": Bc * cfg.num_codebooks * n_dec / dt, "unit": "tokens/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle port, Mini fp32 B={Bc}, {n_dec} decode steps (1 prefill + {n_dec - 1} cached), top-k 50, {dt:.1f} s",
            "os_cpu_count": os.cpu_count()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dac", action="store_true")
    ap.add_argument("--decode-steps", type=int, default=DECODE_STEPS)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback in the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from parler_tts_b200 import DACConfig, ParlerTTSConfig, ParlerTTSDecoderConfig, ParlerTTSForConditionalGeneration
    from parler_tts_b200.dist import broadcast_model_weights

    dcfg = ParlerTTSDecoderConfig(**MINI)
    cfg = ParlerTTSConfig(vocab_size=32128, text_encoder={}, audio_encoder=DACConfig(), decoder=dcfg)
    model = ParlerTTSForConditionalGeneration(cfg, device=dev, dtype=torch.bfloat16)
    if rank == 0:
        model.load_state_dict(synthetic_state_dict(MINI, dev))
    else:
        model.embed_prompts_weight = torch.empty(32128, MINI["hidden_size"], dtype=torch.bfloat16, device=dev)
    if world > 1:
        broadcast_model_weights(model)   # the one collective on this path: packed weights from rank 0 (NVLink)
        torch.cuda.synchronize()
    B, K, V, H = B_PER_GPU, MINI["num_codebooks"], MINI["vocab_size"], MINI["hidden_size"]
    n_dec = args.decode_steps
    L = n_dec + 1
    eng = model.decoder.engine
    sess = eng.session(B, P_LEN, S_LEN, P_LEN + L)
    host = synthetic_inputs(B, H, seed=1 + rank, pin=True)
    enc_d, emask_d, prompt_d, pmask_d = [t.to(dev) for t in host]
    gen = dict(do_sample=True, top_k=50, temperature=1.0, top_p=1.0, min_new_tokens=n_dec, suppress_special=True, codebook_size=1024)

    def one_pass(seed, from_host):
        if from_host:
            e, em, p, pm = [t.to(dev, non_blocking=True) for t in host]
        else:
            e, em, p, pm = enc_d, emask_d, prompt_d, pmask_d
        sess.begin(L, seed=seed, **gen)
        sess.prefill(p, pm, e, em)
        sess.sample()
        sess.decode_steps(n_dec - 1)
        if from_host:
            return sess.raw_ids[:, :L].to("cpu", non_blocking=False)
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(from_host):
        for i in range(args.warmup):
            one_pass(100 + i, from_host)
        barrier()
        l0 = sess.launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(args.steps):
            one_pass(200 + i, from_host)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, sess.launches - l0

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ms, launches = timed(False)
    clk = clocks.stop() if rank == 0 else None
    ms_e2e, _ = timed(True)
    st = sess.state.cpu().tolist()
    assert st[0] == L, f"generation stopped early at length {st[0]} (expected {L})"

    # decode-only timing for the roofline: the 255 graph-replayed steps of one pass, T taken per step
    barrier()
    sess.begin(L, seed=7, **gen)
    sess.prefill(prompt_d, pmask_d, enc_d, emask_d)
    sess.sample()
    sess.decode_steps(2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    sess.decode_steps(n_dec - 3)
    e1.record()
    torch.cuda.synchronize()
    dec_ms = e0.elapsed_time(e1)
    n_timed = n_dec - 3
    # step s (1-based count of tokens already appended) attends over T = P + s + 1 keys
    byts = sum(algorithmic_bytes_per_step(B, K, V, P_LEN + s + 1, S_LEN) for s in range(3, 3 + n_timed))
    hbm_peak, peak_src = peaks()
    achieved = byts / (dec_ms * 1e-3) / 1e9

    # DAC decode of the generated frames (reported separately: the metric is the token loop)
    dac_info = None
    if rank == 0 and not args.no_dac:
        model.audio_encoder.load_state_dict(synth_dac_weights(cfg.audio_encoder, dev))
        frames = L - K
        codes = torch.randint(0, 1024, (1, B, K, frames), device=dev)
        for _ in range(2):
            model.audio_encoder.decode(codes, [None] * B)
        torch.cuda.synchronize()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(3):
            wav = model.audio_encoder.decode(codes, [None] * B).audio_values
        d1.record()
        torch.cuda.synchronize()
        dms = d0.elapsed_time(d1) / 3
        flops = 1.608e9 * B * frames
        dac_info = {"ms": dms, "frames": frames, "batch": B, "tflops": flops / dms / 1e9, "audio_seconds": B * frames * 512 / 44100,
                    "rtf": (B * frames * 512 / 44100) / (dms / 1e3), "kernel": "conv_tc_kernel (tcgen05 implicit GEMM, bf16 in / f32 TMEM accumulate)",
                    "flop_per_frame": 1.608e9}
    if rank == 0:
        tokens = world * B * K * n_dec * args.steps
        value = tokens / (ms * 1e-3)
        e2e_v = tokens / (ms_e2e * 1e-3)
        h2d = sum(t.numel() * t.element_size() for t in host)
        d2h = B * K * L * 8
        line = {
            "metric": "audio codec tokens/sec (all codebooks)", "value": value, "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"Parler-TTS-Mini bf16 batch={B}/GPU {n_dec} decode steps top-k=50 (BASELINE configs[1])",
                       "global_batch": world * B, "prompt_len": P_LEN, "desc_len": S_LEN, "parallelism": f"batch-shard x{world}",
                       "l2": "inputs larger than L2 (0.725 GB weights + KV streamed per step)",
                       "timed_region": "generate_begin + prefill + sampling + graph-replayed decode steps"},
            "clocks": clk,
            "e2e": {"value": e2e_v, "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "GenSession begin/prefill/sample/decode_steps (the calls generate() makes) with pinned host inputs, "
                           "token matrix read back to host"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": NCU_STEP_DRAM_BYTES, "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one decode_step_kernel launch "
                         "(T=73 cached keys; algorithmic 1.16e9 there), ncu --set full capture summarised in profiles/r01_step_ncu_full.md",
                         "peak_source": peak_src, "kernel": "decode_step_kernel (one persistent cooperative kernel per token: embed + 24 x 8 phases + heads + sample)",
                         "ms_per_decode_step": dec_ms / n_timed, "algorithmic_bytes_per_step_avg": byts / n_timed},
        }
        if dac_info is not None:
            line["dac_decode"] = dac_info
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_quick()
            except Exception as ex:  # pragma: no cover
                line["cpu_baseline"] = {"value": None, "error": repr(ex)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
