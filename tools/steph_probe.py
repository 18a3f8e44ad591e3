"""GPU debug: step_host with the action DMA vs in-kernel bulk fetch from host memory, fresh policy data."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_narde_b200 import VecNardeEnv
E, POOL = 131072, 16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
pool = torch.randint(-(1 << 31), (1 << 31) - 1, (POOL, E), dtype=torch.int64).to(torch.int32).pin_memory()
def timed(fn, n=100):
    ev = []
    for _ in range(n):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in ev)
    return "mean %.4f p50 %.4f min %.4f" % (sum(t) / n, t[n // 2], t[0])
for name, kw in (("device random step", None), ("step_host dma_in=True", {"dma_in": True}), ("step_host dma_in=False", {"dma_in": False}),
                 ("step_host packed result", {"packed": True}), ("step_host obs=packed", {"obs": "packed"}),
                 ("step_host packed result + obs", {"packed": True, "obs": "packed"})):
    env = VecNardeEnv(E, seed=0x5EED, max_actions=64)
    env.reset()
    for _ in range(300):
        env.step()
    k = [0]
    def f():
        k[0] += 1
        if kw is None:
            env.step()
        else:
            env.step_host(fraction=True, actions=pool[k[0] % POOL], **kw)
    for _ in range(POOL + 2):
        f()
    torch.cuda.synchronize()
    print("%-26s" % name, timed(f), flush=True)
