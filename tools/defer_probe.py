"""GPU, timing experiment only: the fused step with the block rule switched off by the library's debug flag
(no env is handed to the exact kernel, results are WRONG where the rule matters) against the real step:
an upper bound of what the exact kernel's tail and the exact in-item arithmetic cost."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NARDE_B200_DEBUG_HOOKS"] = "1"   # libnarde_b200_debug.so: build it first with `python -m gym_narde_b200.build --debug-hooks`
import torch
from gym_narde_b200 import VecNardeEnv, _cabi

lib = _cabi.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
for flags in [int(x) for x in (sys.argv[2].split(',') if len(sys.argv) > 2 else '0,1,0,1'.split(','))]:
    env = VecNardeEnv(n, seed=0x5EED, max_actions=64)
    env.reset()
    for _ in range(300):
        env.step()
    lib.narde_debug_set_flags(flags)
    ts = []
    for _ in range(60):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.step(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    lib.narde_debug_set_flags(0)
    ts.sort()
    print("flags", flags, "mean %.4f p50 %.4f min %.4f" % (sum(ts) / len(ts), ts[len(ts) // 2], ts[0]), flush=True)
    del env
