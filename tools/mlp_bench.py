"""GPU: throughput of the tcgen05 afterstate MLP (config 5) vs torch fp32/bf16 on the same box."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from gym_narde_b200.mlp import AfterstateMLP

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 700416
torch.manual_seed(0)
fn = nn.Sequential(nn.Linear(198, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU()).cuda()
head = nn.Linear(256, 576).cuda()
mlp = AfterstateMLP.from_module(fn, head)
x = (torch.rand(rows, 198, device="cuda") < 0.1).float()
q = torch.empty(rows, 576, device="cuda")
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

best, mean = timeit(lambda: mlp.forward(x, out=q))
with torch.no_grad():
    tb, tm = timeit(lambda: head(fn(x)))
    fb = nn.Sequential(fn, head).to(torch.bfloat16)
    xb = x.to(torch.bfloat16)
    bb, bm = timeit(lambda: fb(xb))
print(json.dumps({"rows": rows, "ours_ms": mean, "ours_best_ms": best, "ours_tflops": flop / mean / 1e9,
                  "rows_per_s": rows / mean * 1e3, "torch_fp32_ms": tm, "torch_fp32_tflops": flop / tm / 1e9,
                  "torch_bf16_ms": bm, "torch_bf16_tflops": flop / bm / 1e9}))
