"""Per-phase timing of the fused decode-step kernel (CTA 0 clock64 stamps) for the bench workload."""
import ctypes as C
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from parler_tts_b200 import DACConfig, ParlerTTSConfig, ParlerTTSDecoderConfig, ParlerTTSForConditionalGeneration, _lib

steps_before = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda", 0)
cfg = ParlerTTSConfig(vocab_size=32128, text_encoder={}, audio_encoder=DACConfig(), decoder=ParlerTTSDecoderConfig(**bench.MINI))
model = ParlerTTSForConditionalGeneration(cfg, device=dev, dtype=torch.bfloat16)
model.load_state_dict(bench.synthetic_state_dict(bench.MINI, dev))
B, L = 32, 257
sess = model.decoder.engine.session(B, bench.P_LEN, bench.S_LEN, bench.P_LEN + L)
enc, em, pr, pm = bench.synthetic_inputs(B, 1024, 1, device=dev)
gen = dict(do_sample=True, top_k=50, min_new_tokens=256, suppress_special=True, codebook_size=1024)
sess.begin(L, seed=1, **gen)
sess.prefill(pr, pm, enc, em)
sess.sample()
sess.decode_steps(steps_before)
nph = 8 * 24 + 2
buf = torch.zeros((nph + 1) * 8, dtype=torch.int64, device=dev)
_lib.check(_lib.lib().ptts_session_set_profile(sess.h, _lib.ptr(buf)))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sess.decode_steps(1)
e1.record()
torch.cuda.synchronize()
t = buf.cpu().view(nph + 1, 8).numpy()
_lib.check(_lib.lib().ptts_session_set_profile(sess.h, None))
ghz = 1.965
names = {0: "ln+qkv", 1: "self-attn", 2: "o-proj", 3: "ln+q_cross", 4: "cross-attn", 5: "o_cross", 6: "ln+fc1", 7: "fc2"}
import numpy as np
rows = []
tot = (t[8 * 24 + 1, 6] - t[0, 0]) / ghz / 1e3
print(f"step (event) {e0.elapsed_time(e1)*1e3:.1f} us ; CTA0 clock span {tot:.1f} us")
print("phase: work = start->done on CTA0, barrier = done->released; gemm split: wait_tile / LN / MMA / epilogue (us)")
agg = {}
for ph in range(1, 8 * 24 + 1):
    sub = (ph - 1) & 7
    r = t[ph]
    work = (r[6] - r[0]) / ghz / 1e3
    barr = (r[7] - r[6]) / ghz / 1e3
    parts = [(r[1] - r[0]), (r[2] - r[1]), (r[3] - r[2]), (r[4] - r[3]), ((r[5] - r[0]) if r[5] > 0 else 0)] if sub not in (1, 4) else [0, 0, 0, 0, 0]
    agg.setdefault(sub, []).append([work, barr] + [x / ghz / 1e3 for x in parts])
out = {}
for sub, v in agg.items():
    m = np.array(v).mean(0)
    out[names[sub]] = [round(float(x), 2) for x in m]
    print(f"{names[sub]:12s} work {m[0]:6.2f}  barrier {m[1]:6.2f} | tile {m[2]:6.2f}  ln {m[3]:6.2f} (issued at {m[6]:5.2f})  mma {m[4]:6.2f}  epi {m[5]:6.2f}")
r = t[0]
print(f"embed work {(r[6]-r[0])/ghz/1e3:.2f} barrier {(r[7]-r[6])/ghz/1e3:.2f}")
r = t[8 * 24 + 1]
print(f"heads work {(r[6]-r[0])/ghz/1e3:.2f} | tile {(r[1]-r[0])/ghz/1e3:.2f} ln {(r[2]-r[1])/ghz/1e3:.2f} mma {(r[3]-r[2])/ghz/1e3:.2f} epi {(r[4]-r[3])/ghz/1e3:.2f}")
r = t[nph]
print(f"prologue (entry -> embed start) {(t[0,0]-r[3])/ghz/1e3:.2f} | post-heads barrier {(r[0]-t[nph-1,6])/ghz/1e3:.2f} | sampling {(r[1]-r[0])/ghz/1e3:.2f} | last barrier {(r[2]-r[1])/ghz/1e3:.2f} | kernel span {(r[2]-r[3])/ghz/1e3:.1f} us")
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"step_us": e0.elapsed_time(e1) * 1e3, "phases_us": out, "T": bench.P_LEN + steps_before + 2}, open("gpurun_out/step_phases.json", "w"), indent=1)
