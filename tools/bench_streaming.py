"""Streaming generate() + incremental codec decode: time to first audio and real-time factor (SURVEY 8f rank 1 / BASELINE configs[4] style).
Mini shape, bf16, synthetic weights / inputs; ParlerTTSStreamer(incremental=True) consumes one token column per step on a thread.
Usage: python tools/bench_streaming.py [batch] [decode steps] [play_steps]"""
import json
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from parler_tts_b200 import DACConfig, ParlerTTSConfig, ParlerTTSDecoderConfig, ParlerTTSForConditionalGeneration, ParlerTTSStreamer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 256
play = int(sys.argv[3]) if len(sys.argv) > 3 else 20
dev = torch.device("cuda", 0)
cfg = ParlerTTSConfig(vocab_size=32128, text_encoder={}, audio_encoder=DACConfig(), decoder=ParlerTTSDecoderConfig(**bench.MINI))
model = ParlerTTSForConditionalGeneration(cfg, device=dev, dtype=torch.bfloat16)
model.load_state_dict(bench.synthetic_state_dict(bench.MINI, dev))
model.audio_encoder.load_state_dict(bench.synth_dac_weights(cfg.audio_encoder, dev))
enc, em, pr, pm = [t[:B] for t in bench.synthetic_inputs(32, 1024, 1, device=dev)]


def run():
    st = ParlerTTSStreamer(model, device=dev, play_steps=play, incremental=True)
    kw = dict(encoder_outputs=(enc,), attention_mask=em, prompt_hidden_states=pr, prompt_attention_mask=pm, do_sample=True, top_k=50,
              min_new_tokens=steps, max_new_tokens=steps, seed=3, _suppress_special=True, streamer=st)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = threading.Thread(target=lambda: model.generate(**kw))
    th.start()
    first, n = None, 0
    for chunk in st:
        m = chunk.shape[-1]
        if m > 0 and first is None:
            first = time.perf_counter() - t0
        n += m
    th.join()
    total = time.perf_counter() - t0
    return first, total, n


run()  # warm-up (session creation, tensor maps)
first, total, n = run()
audio_s = B * n / 44100
out = {"batch": B, "decode_steps": steps, "play_steps": play, "time_to_first_audio_ms": 1e3 * first, "wall_s": total, "samples_per_utterance": n,
       "rtf_all_utterances": audio_s / total, "rtf_per_utterance": (n / 44100) / total,
       "tokens_per_s": B * 9 * steps / total,
       "note": "streamer contract = one host-visible token column per step (decode_steps(1) + a .cpu() per token), incremental windows of new frames + 2 x 10 context frames"}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/streaming_b{B}.json", "w"), indent=1)
