"""DAC decode timing (codes -> waveform) for the 44.1 kHz DAC shape, bf16 tcgen05 path vs SIMT path."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parler_tts_b200 import DACConfig, DACModel

from bench import synth_dac_weights

def main():
    B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 248
    dev = torch.device("cuda", 0)
    cfg = DACConfig()
    sd = synth_dac_weights(cfg, dev)
    codes = torch.randint(0, 1024, (1, B, 9, T), device=dev)
    out = {}
    for tc in (("1",) if os.environ.get("PTTS_DAC_BENCH_TC_ONLY") == "1" else ("1", "0")):
        os.environ["PTTS_DAC_TC"] = tc
        m = DACModel(cfg, dev, torch.bfloat16).load_state_dict(sd)
        for _ in range(2): m.decode(codes, [None] * B)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3 if tc == "1" else 1
        e0.record()
        for _ in range(n): a = m.decode(codes, [None] * B).audio_values
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        flops = 1.608e9 * B * T
        out["tcgen05" if tc == "1" else "simt"] = {"ms": ms, "tflops": flops / ms / 1e9, "rtf": (B * T * 512 / 44100) / (ms / 1e3), "audio_rms": float(a.float().pow(2).mean().sqrt())}
    out["config"] = {"B": B, "frames": T, "flop_per_frame": 1.608e9}
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/dac_bench.json", "w"), indent=1)

main()
