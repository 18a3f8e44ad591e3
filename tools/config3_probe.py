"""GPU probe: BASELINE config 3 (narde_enumerate_fast over 1M synthetic positions) timed with CUDA events; under
`ncu --metrics gpu__time_duration.sum` the launch list splits the call into k_step_full_v2 / k_step_deferred."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_narde_b200 import _cabi
from gym_narde_b200.workloads import config3_positions

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
lo, hi, dice, strata = config3_positions(dev, n=n, seed=1234)
cap = 64
actions = torch.zeros((n, cap), dtype=torch.int64, device=dev)
counts = torch.zeros(n, dtype=torch.int32, device=dev)
ovf = torch.zeros(n, dtype=torch.uint8, device=dev)
ws = torch.zeros(_cabi.workspace_ints(n), dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
run = lambda: _cabi.enumerate_actions_fast(lo, hi, dice, actions, counts, ovf, ws)
for _ in range(3):
    run()
ts = []
for _ in range(reps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
print("config3 n=%d: mean %.4f ms p50 %.4f min %.4f  -> %.3e positions/s; deferred %d, mean legal %.2f" % (
    n, sum(ts) / len(ts), ts[len(ts) // 2], ts[0], n / (sum(ts) / len(ts) * 1e-3), int(ws[0].item()), float(counts.float().mean())))
