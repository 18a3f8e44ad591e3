"""GPU: the cta_group::2 scorer vs the default scorer: equality and timing."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from gym_narde_b200 import VecNardeEnv
from gym_narde_b200.mlp import AfterstateMLP

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 350208
torch.manual_seed(0)
fn = nn.Sequential(nn.Linear(198, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU()).cuda()
head = nn.Linear(256, 576).cuda()
mlp = AfterstateMLP.from_module(fn, head)
env = VecNardeEnv(rows, seed=5, write_actions=False)
env.reset()
for _ in range(40):
    env.step()
lo, hi = env.lo, env.hi
for n in (256, 129, 1, 5000, rows):
    a = mlp.score_states(lo[:n].contiguous(), hi[:n].contiguous())
    b = mlp.score_states_2sm(lo[:n].contiguous(), hi[:n].contiguous())
    torch.cuda.synchronize()
    print("rows", n, "equal", bool(torch.equal(a, b)), "max abs diff", float((a - b).abs().max()), flush=True)
sc = torch.empty(rows, device="cuda")
flop = rows * 2 * (198 * 256 + 256 * 256 + 256 * 576)
def timeit(f, n=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); f(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / n
for name, f in (("score_states", lambda: mlp.score_states(lo, hi, out=sc)), ("score_states_2sm", lambda: mlp.score_states_2sm(lo, hi, out=sc))):
    ms = timeit(f)
    print(name, "ms %.4f TFLOP/s %.1f" % (ms, flop / ms / 1e9), flush=True)
