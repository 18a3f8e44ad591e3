"""GPU: fused-step time against the number of envs around whole multiples of the resident CTA count
(148 SMs x 5 CTAs x 128 envs = 94 720): how much of the 131 072-env step is wave quantisation."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_narde_b200 import VecNardeEnv

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
out = {}
for n in (47360, 94720, 113664, 131072, 142080, 189440, 262144, 284160):
    env = VecNardeEnv(n, seed=0x5EED, max_actions=64)
    env.reset()
    for _ in range(300):
        env.step()
    ts = []
    for _ in range(40):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.step(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    ms = sum(ts) / len(ts)
    out[n] = {"ctas": (n + 127) // 128, "waves": (n + 127) // 128 / 740.0, "ms": ms, "p50": ts[len(ts) // 2], "ns_per_env": ms * 1e6 / n}
    print(n, out[n], flush=True)
    del env
print(json.dumps(out))
