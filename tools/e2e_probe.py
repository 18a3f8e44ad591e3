"""GPU debug: where does the end-to-end step time go (policy input mode x copies); CPU enqueue cost vs GPU time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_narde_b200 import VecNardeEnv

E = 131072
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timed(fn, n=60, headstart=False, do_flush=True):
    ev = []
    torch.cuda.synchronize()
    if headstart:
        for _ in range(300):
            flush.fill_(0)
    t0 = time.perf_counter()
    for _ in range(n):
        if do_flush:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    cpu = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    t = sorted(x.elapsed_time(y) for x, y in ev)
    return "gpu mean %.4f p50 %.4f min %.4f max %.4f | cpu enqueue/iter %.4f ms, wall/iter %.4f ms" % (sum(t) / n, t[n // 2], t[0], t[-1], cpu, wall)

env = VecNardeEnv(E, seed=0x5EED, max_actions=64)
env.reset()
for _ in range(300):
    env.step()
torch.cuda.synchronize()
h = torch.randint(-(1 << 31), (1 << 31) - 1, (E,), dtype=torch.int64).to(torch.int32).pin_memory()
env.action_in.copy_(h)
hr = torch.zeros(E).pin_memory(); hd = torch.zeros(E, dtype=torch.uint8).pin_memory()
def f0():
    env.step(env.action_in, fraction=True)
def f1():
    env.action_in.copy_(h, non_blocking=True)
    env.step(env.action_in, fraction=True)
def f2():
    f1()
    hr.copy_(env.reward, non_blocking=True); hd.copy_(env.done, non_blocking=True)
for _ in range(5):
    f2(); env.step()
for hs in (False, True):
    print("headstart", hs)
    print(" random (device Philox)      ", timed(lambda: env.step(), headstart=hs))
    print(" fraction, no copies         ", timed(f0, headstart=hs))
    print(" fraction + H2D              ", timed(f1, headstart=hs))
    print(" fraction + H2D + D2H        ", timed(f2, headstart=hs))
    print(" random, no flush            ", timed(lambda: env.step(), headstart=hs, do_flush=False))
    print(" fraction+H2D+D2H, no flush  ", timed(f2, headstart=hs, do_flush=False))
