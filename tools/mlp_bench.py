"""GPU: throughput of the tcgen05 afterstate MLP (config 5) vs torch fp32/bf16 on the same box."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from gym_narde_b200 import VecNardeEnv
from gym_narde_b200.mlp import AfterstateMLP

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 700416
torch.manual_seed(0)
fn = nn.Sequential(nn.Linear(198, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU()).cuda()
head = nn.Linear(256, 576).cuda()
move2_head = nn.Linear(256 + 576, 576).cuda()
mlp = AfterstateMLP.from_module(fn, head, move2_head)
m1 = torch.randint(0, 576, (rows,), device="cuda", dtype=torch.int32)
env = VecNardeEnv(rows, seed=5, write_actions=False)
env.reset()
for _ in range(40):
    env.step()
x = env.observe().clone()
lo, hi = env.lo, env.hi
q = torch.empty(rows, 576, device="cuda")
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
    return min(a.elapsed_time(b) for a, b in ev), sum(a.elapsed_time(b) for a, b in ev) / n

res = {"rows": rows}
for name, f in (("forward_x_q", lambda: mlp.forward(x, out=q)), ("score_x", lambda: mlp.score(x, out=sc)),
                ("forward_states_q", lambda: mlp.forward_states(lo, hi, out=q)),
                ("forward_move2_x_q", lambda: mlp.forward(x, m1, out=q)),
                ("forward_move2_states_q", lambda: mlp.forward_states(lo, hi, m1, out=q)),
                ("score_states", lambda: mlp.score_states(lo, hi, out=sc))):
    best, mean = timeit(f)
    res[name] = {"ms": mean, "best_ms": best, "tflops": flop / mean / 1e9, "rows_per_s": rows / mean * 1e3}
with torch.no_grad():
    tb, tm = timeit(lambda: head(fn(x)))
    fb = nn.Sequential(fn, head).to(torch.bfloat16)
    xb = x.to(torch.bfloat16)
    bb, bm = timeit(lambda: fb(xb))
    oh = torch.zeros(rows, 576, device="cuda", dtype=torch.bfloat16)
    m2b = move2_head.to(torch.bfloat16)
    idx = m1.long().unsqueeze(1)

    def torch_move2():
        oh.zero_()
        oh.scatter_(1, idx, 1)
        return m2b(torch.cat((fb[0](xb), oh), dim=1))
    _, m2m = timeit(torch_move2)
res["torch_bf16_move2"] = {"ms": m2m, "note": "one-hot + cat + Linear(832,576), as DecomposedDQN.forward(x, move1) does"}
res["torch_fp32"] = {"ms": tm, "tflops": flop / tm / 1e9}
res["torch_bf16"] = {"ms": bm, "tflops": flop / bm / 1e9}
print(json.dumps(res))
