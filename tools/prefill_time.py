import os, sys, torch
sys.path.insert(0, "/root/repo")
import bench
from parler_tts_b200 import DACConfig, ParlerTTSConfig, ParlerTTSDecoderConfig, ParlerTTSForConditionalGeneration
dev = torch.device("cuda", 0)
cfg = ParlerTTSConfig(vocab_size=32128, text_encoder={}, audio_encoder=DACConfig(), decoder=ParlerTTSDecoderConfig(**bench.MINI))
model = ParlerTTSForConditionalGeneration(cfg, device=dev, dtype=torch.bfloat16)
model.load_state_dict(bench.synthetic_state_dict(bench.MINI, dev))
B, L = 32, 257
sess = model.decoder.engine.session(B, bench.P_LEN, bench.S_LEN, bench.P_LEN + L)
enc, em, pr, pm = bench.synthetic_inputs(B, 1024, 1, device=dev)
gen = dict(do_sample=True, top_k=50, min_new_tokens=256, suppress_special=True, codebook_size=1024)
for rep in range(3):
    sess.begin(L, seed=1, **gen)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sess.prefill(pr, pm, enc, em); e1.record(); torch.cuda.synchronize()
    print(f"prefill {e0.elapsed_time(e1):.2f} ms", flush=True)
sess.sample(); sess.decode_steps(4); torch.cuda.synchronize()
print("logits finite", bool(torch.isfinite(sess.logits).all()))
