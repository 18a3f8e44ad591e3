"""GPU soak (not part of the pytest suites: minutes of single-core Python): the fused step against the oracle,
lock-step, over ~2 M env turns on both CTA tiles of the kernel, with auto-reset and the Philox stream."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity as P
import support

be = support.CudaBackend()
tot = 0
for n, T, seed, cap in ((17000, 60, 0xA11CE, 64), (4096, 250, 0xB0B, 64), (2048, 200, 0xC0DE, 4)):
    t0 = time.time()
    eps = P.check_step_full_lockstep(be, n, T, seed, env_base=123456, cap=cap)
    tot += n * T
    print("envs %6d x %3d steps  seed %#x cap %2d : %8d env turns, %6d finished episodes, all outputs bit-equal to the oracle  (%.0f s)"
          % (n, T, seed, cap, n * T, eps, time.time() - t0), flush=True)
print("total env turns checked:", tot)
