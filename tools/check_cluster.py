"""Cluster step kernel (step2.cu) against the legacy fused kernel (step.cu) on the Mini layer shape: teacher-forced logits,
free-running greedy tokens and decode-step time.  Usage: python tools/check_cluster.py [layers] [batch] [steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from parler_tts_b200 import DACConfig, ParlerTTSConfig, ParlerTTSDecoderConfig, ParlerTTSForConditionalGeneration

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 24
dev = torch.device("cuda", 0)
mcfg = dict(bench.MINI, num_hidden_layers=layers)
cfg = ParlerTTSConfig(vocab_size=32128, text_encoder={}, audio_encoder=DACConfig(), decoder=ParlerTTSDecoderConfig(**mcfg))
model = ParlerTTSForConditionalGeneration(cfg, device=dev, dtype=torch.bfloat16)
sd = bench.synthetic_state_dict(mcfg, dev)
for k in list(sd):   # livelier logits than the 0.02 init so that argmax comparisons mean something
    if "lm_heads" in k:
        sd[k] = sd[k] * 10
model.load_state_dict(sd)
eng = model.decoder.engine
enc, em, pr, pm = bench.synthetic_inputs(32, 1024, 1, device=dev)
enc, em, pr, pm = enc[:B], em[:B], pr[:B], pm[:B]
L = steps + 2


def run(mode, forced=None):
    os.environ["PTTS_STEP"] = mode
    eng._sessions = {}
    sess = eng.session(B, bench.P_LEN, bench.S_LEN, bench.P_LEN + L)
    sess.begin(L, do_sample=False)
    sess.prefill(pr, pm, enc, em)
    kind = sess.fused
    logits, toks = [], []
    sess.sample(forced=None if forced is None else forced[:, 0])
    toks.append(sess.raw_ids[:, 1].clone())
    for t in range(1, steps):
        sess.decode_forward()
        torch.cuda.synchronize()
        logits.append(sess.logits.float().clone())
        sess.sample(forced=None if forced is None else forced[:, t])
        toks.append(sess.raw_ids[:, t + 1].clone())
    torch.cuda.synchronize()
    return kind, torch.stack(logits), torch.stack(toks, 1)


kind_l, log_l, tok_l = run("legacy")
print(f"legacy: fused kind {kind_l}", flush=True)
kind_c, log_c, tok_c = run("cluster", forced=tok_l)
print(f"cluster: fused kind {kind_c}", flush=True)
scale = log_l.abs().amax().item()
err = (log_l - log_c).abs().amax(dim=(1, 2)) / scale
print("teacher-forced max |logit diff| / max|logit| per step:", " ".join(f"{e:.4f}" for e in err.tolist()))
print("finite:", bool(torch.isfinite(log_c).all()), " argmax agreement:", float((log_l.argmax(-1) == log_c.argmax(-1)).float().mean()))
kind_c2, _, tok_c2 = run("cluster")
agree = (tok_c2 == tok_l).all(dim=1).float().mean().item()
print(f"free-running greedy: {agree * 100:.1f} % of the {tok_l.shape[0]} rows token-identical over {steps} steps")


def timeit(mode, n=60):
    os.environ["PTTS_STEP"] = mode
    eng._sessions = {}
    LL = n + 8
    sess = eng.session(B, bench.P_LEN, bench.S_LEN, bench.P_LEN + LL)
    sess.begin(LL, do_sample=True, top_k=50, seed=1, min_new_tokens=LL - 1, suppress_special=True, codebook_size=1024)
    sess.prefill(pr, pm, enc, em)
    sess.sample()
    sess.decode_steps(4)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sess.decode_steps(n)
    e1.record()
    torch.cuda.synchronize()
    assert int(sess.state[0].item()) == n + 6, sess.state.tolist()
    return e0.elapsed_time(e1) / n * 1e3


for mode in ("legacy", "cluster"):
    print(f"{mode}: {timeit(mode):.1f} us per decode step ({layers} layers, B={B})", flush=True)
ok = bool(torch.isfinite(log_c).all()) and err.max().item() < 0.03 and kind_c == 2
print("CLUSTER CHECK", "OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
