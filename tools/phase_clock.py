"""Debug tool (GPU): per-CTA phase timing of k_step_full_v2 from clock64 marks."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NARDE_B200_DEBUG_HOOKS"] = "1"   # libnarde_b200_debug.so: build it first with `python -m gym_narde_b200.build --debug-hooks`
import numpy as np
import torch
from gym_narde_b200 import VecNardeEnv, _cabi

E = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
env = VecNardeEnv(E, seed=0x5EED, max_actions=64)
env.reset()
for _ in range(300):
    env.step()
TILE = 64 if os.environ.get('NARDE_TILE') == '64' else 128   # envs per main CTA (the library reads the same variable)
nb = min((E + TILE - 1) // TILE, 2048)
buf = torch.zeros((max(nb, 2048) + 1024, 16), dtype=torch.int64, device="cuda")
lib = _cabi.load()
lib.narde_debug_set_clock_buffer.argtypes = [C.c_void_p]
assert lib.narde_debug_set_clock_buffer(C.c_void_p(buf.data_ptr())) == 0
names = ["load", "scan+bases", "rows+L1", "scan+L2bases", "count", "scan+envbases", "emit", "finish", "obs"]
if len(sys.argv) > 2:
    lib0 = _cabi.load(); lib0.narde_debug_set_flags(int(sys.argv[2]))
acc = []
for _ in range(5):
    buf.zero_()
    env.step()
    torch.cuda.synchronize()
    acc.append(buf.cpu().numpy().copy())
lib.narde_debug_set_clock_buffer(None)
full = np.stack(acc).astype(np.float64)
a = full[:, :nb]          # [steps, blocks, 16]
d = np.diff(a[:, :, :10], axis=2)              # phase durations in cycles
tot = a[:, :, 9] - a[:, :, 0]
print("blocks", nb, "block total cycles: mean %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f" % (
    tot.mean(), np.percentile(tot, 50), np.percentile(tot, 90), np.percentile(tot, 99), tot.max()))
for k, nm in enumerate(names):
    x = d[:, :, k]
    print("%-14s mean %8.0f  p50 %8.0f  p99 %8.0f  max %9.0f  share %.1f%%" % (
        nm, x.mean(), np.percentile(x, 50), np.percentile(x, 99), x.max(), 100 * x.sum() / tot.sum()))
span = a[:, :, 9].max(axis=1) - a[:, :, 0].min(axis=1)
print("kernel span cycles per step (max end - min start, per-SM clocks differ slightly):", span)

dd = full[:, 2048:2048 + 1024]
dn = ["init", "items", "item counts", "scan", "materialise", "test", "rank", "(end of solve)", "complete", "obs"]
for st in range(dd.shape[0]):
    x = dd[st]
    x = x[(x[:, 10] > x[:, 0]) & (x[:, 0] > 0)]
    if st > 0:                      # rows not rewritten since the previous step are stale
        prev = dd[st - 1]
    if len(x) == 0:
        continue
    dur = np.diff(x[:, :11], axis=1)
    tot = x[:, 10] - x[:, 0]
    w = np.argmax(tot)
    meta = x[:, 11].astype(np.int64)
    cnt, ncand = meta & 0xFFFFFFFF, meta >> 32
    print("step", st, "deferred envs (first 1024)", len(x), "| total mean %.0f p50 %.0f p90 %.0f max %.0f | candidates mean %.0f p90 %.0f max %d" % (
        tot.mean(), np.percentile(tot, 50), np.percentile(tot, 90), tot.max(), ncand.mean(), np.percentile(ncand, 90), ncand.max()))
    print("   mean per phase:", " ".join("%s=%.0f" % (n, v) for n, v in zip(dn, dur.mean(axis=0))))
    print("   slowest env   :", " ".join("%s=%.0f" % (n, v) for n, v in zip(dn, dur[w])), "count %d candidates %d" % (cnt[w], ncand[w]))
