"""GPU debug: fused-step time with outputs switched off (which stores cost what)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_narde_b200 import VecNardeEnv, _cabi

E = 131072
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timed(fn, n=80):
    ev = []
    for _ in range(n):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in ev)
    return "mean %.4f p50 %.4f min %.4f" % (sum(t) / n, t[n // 2], t[0])

for name, kw in (("all outputs", {}), ("no action lists", {"write_actions": False})):
    env = VecNardeEnv(E, seed=0x5EED, max_actions=64, graph=False, **kw)
    env.reset()
    for _ in range(300):
        env.step()
    print("%-28s" % name, timed(lambda: env.step()))
    if name == "all outputs":
        # no Box(198): call the C ABI directly without obs
        def no_obs():
            env.step_count += 1
            env._step_dev.fill_(env.step_count)
            _cabi.step_full(env.lo, env.hi, env.env_base, env.seed, 0, actions=env.actions, counts=env.counts,
                            dice_out=env.dice, chosen=env.chosen, obs198=None, reward=env.reward, done=env.done,
                            stats=env.stats, flags=2, max_episode_steps=1000, truncated=env.trunc,
                            workspace=env._workspaces[0], step_dev=env._step_dev)
        print("%-28s" % "no Box(198)", timed(no_obs))
        def no_obs_no_act():
            env.step_count += 1
            env._step_dev.fill_(env.step_count)
            _cabi.step_full(env.lo, env.hi, env.env_base, env.seed, 0, actions=None, counts=env.counts,
                            dice_out=env.dice, chosen=env.chosen, obs198=None, reward=env.reward, done=env.done,
                            stats=env.stats, flags=2, max_episode_steps=1000, truncated=env.trunc,
                            workspace=env._workspaces[0], step_dev=env._step_dev)
        print("%-28s" % "no Box(198), no lists", timed(no_obs_no_act))
