"""Build-container only: time the REAL Python reference (read-only /root/reference) on this container's CPU,
next to the oracle port on the same core, so that the "port" CPU baseline of bench.py can be related to the
reference's own speed.  The reference cannot travel to the GPU box; the result is committed as a fixture
(profiles/reference_python_timing_container.json).  Test infrastructure: imports oracle/.
"""
import json, os, platform, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import ref_loader, oracle as O

narde, narde_env = ref_loader.load()
np.random.seed(0)
env = narde_env.NardeEnv()
env.reset(seed=0)
# (1) NardeEnv.step with uniformly random action codes, as examples/play_random_agent.py intends
n, t0, episodes = 0, time.perf_counter(), 0
while time.perf_counter() - t0 < 8.0:
    a = (int(np.random.randint(0, 576)), int(np.random.randint(0, 576)))
    obs, r, done, trunc, info = env.step(a)
    n += 1
    if done or n % 1000 == 0:
        env.reset()
        episodes += 1
ref_step = n / (time.perf_counter() - t0)
# (2) valid-random codes: get_valid_moves on the env's own dice is not possible (step rolls inside), so the
#     evaluate_model.py pattern: enumerate with a fresh roll, then step
env.reset(seed=1)
n, t0 = 0, time.perf_counter()
while time.perf_counter() - t0 < 8.0:
    dice = [int(np.random.randint(1, 7)), int(np.random.randint(1, 7))]
    vm = env.game.get_valid_moves(dice, env.current_player)
    def code(m):
        return m[0] * 24 + (0 if m[1] == 'off' else m[1])
    a = (code(vm[np.random.randint(len(vm))]), code(vm[np.random.randint(len(vm))])) if vm else (0, 0)
    obs, r, done, trunc, info = env.step(a)
    n += 1
    if done or n % 1000 == 0:
        env.reset()
ref_valid = n / (time.perf_counter() - t0)
# (3) the oracle port on ONE core of the same machine: full-rules self-play (what bench.py's CPU arm runs)
O.build()
t0 = time.perf_counter()
turns, acts, eps = O.selfplay(0x5EED, 0, 64, 3000)
port = turns / (time.perf_counter() - t0)
out = {"machine": platform.processor() or platform.machine(), "cpu_count": os.cpu_count(), "python": platform.python_version(),
       "reference_NardeEnv_step_random_codes_per_s_per_core": ref_step,
       "reference_get_valid_moves_plus_step_per_s_per_core": ref_valid,
       "oracle_port_full_rules_turns_per_s_per_core": port,
       "note": "reference = /root/reference gym_narde/envs (gymnasium stubbed), single process; oracle port = oracle/narde_oracle.c "
               "o_selfplay (full legal-turn enumeration per turn, more work per step than the reference's single half-move lists)"}
print(json.dumps(out, indent=1))
