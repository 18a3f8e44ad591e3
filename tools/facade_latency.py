"""GPU: latency of the single-env facade (the reference's own call pattern, one env, one call at a time) -- a kernel
launch + one stream synchronisation per call on pinned (mapped) host buffers -- beside the unmodified Python reference
from baseline/_ref on the same box (when it was vendored there).  This path is launch-latency bound by construction;
the batched VecNardeEnv is the throughput path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def drive(make_env, steps=3000):
    env = make_env()
    np.random.seed(0)
    env.reset()
    t_moves = t_step = 0.0
    n = 0
    for _ in range(steps):
        dice = [int(np.random.randint(1, 7)), int(np.random.randint(1, 7))]
        t0 = time.perf_counter()
        moves = env.unwrapped.game.get_valid_moves(dice, env.unwrapped.current_player)
        t1 = time.perf_counter()
        if moves:
            m1 = moves[np.random.randint(len(moves))]
            code = m1[0] * 24 + (0 if m1[1] == 'off' else m1[1])
            action = (code, 0)
        else:
            action = (0, 0)
        t2 = time.perf_counter()
        out = env.step(action)
        t3 = time.perf_counter()
        t_moves += t1 - t0
        t_step += t3 - t2
        n += 1
        if out[2] or (len(out) > 4 and out[3]):
            env.reset()
    return 1e6 * t_moves / n, 1e6 * t_step / n


if __name__ == "__main__":
    import gym_narde_b200
    a = drive(lambda: gym_narde_b200.make("narde-v0"))
    print("gym_narde_b200 facade : get_valid_moves %.1f us/call, step %.1f us/call" % a)
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    if os.path.isdir(os.path.join(ref, "gym_narde")):
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(ref)), "tests"))
        import gymnasium_stub  # (the image has no gymnasium; the reference only needs Env / spaces / register)
        gymnasium_stub.install()
        sys.path.insert(0, ref)
        for k in [k for k in sys.modules if k.startswith("gym_narde") and not k.startswith("gym_narde_b200")]:
            del sys.modules[k]
        from gym_narde.envs.narde_env import NardeEnv as RefEnv
        b = drive(lambda: RefEnv())
        print("Python reference      : get_valid_moves %.1f us/call, step %.1f us/call" % b)
    else:
        print("Python reference      : baseline/_ref not present")
