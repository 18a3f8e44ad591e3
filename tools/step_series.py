"""GPU debug: per-step time series of the fused step: ms, deferred envs, legal actions, finished episodes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_narde_b200 import VecNardeEnv

E = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
T = int(sys.argv[2]) if len(sys.argv) > 2 else 900
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
env = VecNardeEnv(E, seed=0x5EED, max_actions=64, graph=False)
env.reset()
rows = []
prev = env.stats.clone()
for t in range(T):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); env.step(); b.record()
    torch.cuda.synchronize()
    st = env.stats.clone()
    d = (st - prev).tolist(); prev = st
    rows.append((t, a.elapsed_time(b), int((env._ws_adv if env._ws_adv is not None else env._workspaces[0])[3].item()), d[5] / E, d[0], int((env.counts > 64).sum().item())))
for r in rows:
    if r[0] % 10 == 0 or r[1] > 0.25:
        print("step %4d  ms %.4f  deferred %5d  meanA %6.2f  finished %6d  overflow %6d" % r)
import statistics
for lo in range(0, T, 100):
    w = [r[1] for r in rows[lo:lo + 100]]
    print("steps %d-%d: mean %.4f min %.4f max %.4f  deferred mean %.1f" % (lo, lo + 99, statistics.mean(w), min(w), max(w), statistics.mean(r[2] for r in rows[lo:lo+100])))
