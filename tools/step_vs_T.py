"""Decode-step time of the bench workload as a function of the cached length: event-timed windows of 32 steps.
Usage: python tools/step_vs_T.py [decode steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from parler_tts_b200 import DACConfig, ParlerTTSConfig, ParlerTTSDecoderConfig, ParlerTTSForConditionalGeneration

n_dec = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
cfg = ParlerTTSConfig(vocab_size=32128, text_encoder={}, audio_encoder=DACConfig(), decoder=ParlerTTSDecoderConfig(**bench.MINI))
model = ParlerTTSForConditionalGeneration(cfg, device=dev, dtype=torch.bfloat16)
model.load_state_dict(bench.synthetic_state_dict(bench.MINI, dev))
B, L = 32, n_dec + 1
sess = model.decoder.engine.session(B, bench.P_LEN, bench.S_LEN, bench.P_LEN + L)
enc, em, pr, pm = bench.synthetic_inputs(B, 1024, 1, device=dev)
gen = dict(do_sample=True, top_k=50, min_new_tokens=n_dec, suppress_special=True, codebook_size=1024)
for rep in range(2):
    sess.begin(L, seed=1 + rep, **gen)
    sess.prefill(pr, pm, enc, em)
    sess.sample()
    done, W, out = 0, 32, []
    torch.cuda.synchronize()
    while done + W <= n_dec - 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sess.decode_steps(W)
        e1.record()
        out.append((bench.P_LEN + done + 2, e0, e1))
        done += W
    torch.cuda.synchronize()
    print(f"fused kind {sess.fused}; us per step by window (first T of the window): " + "  ".join(f"T{t}:{e0.elapsed_time(e1) / W * 1e3:.0f}" for t, e0, e1 in out), flush=True)
