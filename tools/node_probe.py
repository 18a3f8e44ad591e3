"""GPU: what one more small node in the step's CUDA graph costs (extra counter-advance kernels in front of the
fused step; the game is not affected by the Philox step index skipping values)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_narde_b200 import VecNardeEnv, _cabi

n = 131072
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for extra in (0, 2, 4, 0, 2, 4):
    env = VecNardeEnv(n, seed=0x5EED, max_actions=64)
    env.reset()
    for _ in range(300):
        env.step()
    env._step_dev.fill_(env.step_count)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(extra + 1):
            _cabi.advance_counter(env._step_dev)
        env._launch_full(None, None, _cabi.AUTORESET)
    ts = []
    for _ in range(60):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print("extra nodes", extra, "mean %.4f p50 %.4f min %.4f" % (sum(ts) / len(ts), ts[len(ts) // 2], ts[0]), flush=True)
    del env, g
