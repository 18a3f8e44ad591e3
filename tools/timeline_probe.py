"""Debug tool (GPU): wall-clock timeline of one fused step from %globaltimer marks (debug build): when the main
kernel's CTAs start and end, when the exact kernel's CTAs start / finish their envs / see the primary complete / exit,
against the CUDA-event time of the whole step (graph replay, L2 flushed in front, as bench.py times it)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NARDE_B200_DEBUG_HOOKS"] = "1"
import numpy as np
import torch
from gym_narde_b200 import VecNardeEnv, _cabi

E = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
env = VecNardeEnv(E, seed=0x5EED, max_actions=64)
env.reset()
for _ in range(300):
    env.step()
TILE = 64 if os.environ.get('NARDE_TILE') == '64' else 128   # envs per main CTA (the library reads the same variable)
nb = min((E + TILE - 1) // TILE, 2048)
buf = torch.zeros((3072 + 2048, 16), dtype=torch.int64, device="cuda")
lib = _cabi.load()
lib.narde_debug_set_clock_buffer.argtypes = [C.c_void_p]
lib.narde_debug_set_flags(flags)
if len(sys.argv) > 3:
    lib.narde_debug_set_stagger.argtypes = [C.c_uint]
    lib.narde_debug_set_stagger(int(sys.argv[3]))          # ns between the start of an SM's CTA slots (first wave)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
rows = []
for it in range(12):
    buf.zero_()
    assert lib.narde_debug_set_clock_buffer(C.c_void_p(buf.data_ptr())) == 0
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); env.step(); b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    x = buf.cpu().numpy().astype(np.float64)
    m = x[:nb]
    d = x[3072:3072 + 1184]
    t0 = m[:, 10].min()
    main_end = m[:, 11].max()
    first_end = m[:, 11].min()
    rows.append((ms * 1e3, (m[:, 10].max() - t0) / 1e3, (first_end - t0) / 1e3, (main_end - t0) / 1e3,
                 (d[:, 0].min() - t0) / 1e3, (d[:, 0].max() - t0) / 1e3, (d[:, 2].max() - t0) / 1e3,
                 (d[:, 3].min() - t0) / 1e3, (d[:, 4].max() - t0) / 1e3))
lib.narde_debug_set_clock_buffer(None)
lib.narde_debug_set_flags(0)
print("flags", flags, "envs", E, "(us; t = 0 at the first main CTA's start)")
print("%9s %9s %9s %9s | %9s %9s %9s %9s %9s" % ("event", "lastStart", "firstEnd", "mainEnd", "defStart", "defStartL", "defSolved", "defWaitOk", "defExit"))
for r in rows[2:]:
    print("%9.1f %9.1f %9.1f %9.1f | %9.1f %9.1f %9.1f %9.1f %9.1f" % r)
med = np.median(np.array(rows[2:]), axis=0)
print("median")
print("%9.1f %9.1f %9.1f %9.1f | %9.1f %9.1f %9.1f %9.1f %9.1f" % tuple(med))
