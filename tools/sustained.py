"""GPU debug: back-to-back fused steps (no L2 flush, one device interval) vs flushed per-step timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_narde_b200 import VecNardeEnv

E = 131072
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
env = VecNardeEnv(E, seed=0x5EED, max_actions=64)
env.reset()
for _ in range(300):
    env.step()
torch.cuda.synchronize()

def interval(fn, n):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def flushed(fn, n):
    ev = []
    for _ in range(n):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    return sum(x.elapsed_time(y) for x, y in ev) / n

for rep in range(2):
    print("flushed per-step      %.4f ms" % flushed(lambda: env.step(), 100))
    print("back-to-back (graph)  %.4f ms" % interval(lambda: env.step(), 100))
    print("back-to-back x1000    %.4f ms" % interval(lambda: env.step(), 1000))
env2 = VecNardeEnv(E, seed=0x5EED, max_actions=64, graph=False)
env2.reset()
for _ in range(300):
    env2.step()
print("no graph flushed      %.4f ms" % flushed(lambda: env2.step(), 100))
print("no graph back-to-back %.4f ms" % interval(lambda: env2.step(), 300))
env3 = VecNardeEnv(E, seed=0x5EED, max_actions=64, write_actions=False)
env3.reset()
for _ in range(300):
    env3.step()
print("no lists flushed      %.4f ms" % flushed(lambda: env3.step(), 100))
print("no lists back-to-back %.4f ms" % interval(lambda: env3.step(), 300))
