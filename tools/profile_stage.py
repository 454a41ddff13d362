"""Debug: where the time before the activation-tile TMA goes (build with -DPTTS_PROF_STAGE)."""
import os, sys
sys.argv = [sys.argv[0], "100"]
exec(open(os.path.join(os.path.dirname(__file__), "profile_step.py")).read().split("ghz = 1.965")[0])
ghz = 1.965
import numpy as np
names = {0: "ln+qkv", 2: "o-proj", 3: "ln+q_cross", 5: "o_cross", 6: "ln+fc1", 7: "fc2"}
agg = {}
for ph in range(1, 8 * 24 + 1):
    sub = (ph - 1) & 7
    if sub in (1, 4): continue
    r = t[ph]
    agg.setdefault(sub, []).append([(r[i] - r[0]) / ghz / 1e3 for i in (2, 3, 5, 1)])
for sub, v in agg.items():
    m = np.array(v).mean(0)
    print(f"{names[sub]:12s} after syncthreads {m[0]:5.2f}  after fence.proxy.async {m[1]:5.2f}  copies issued {m[2]:5.2f}  tile landed {m[3]:5.2f}")
