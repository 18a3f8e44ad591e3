"""GPU debug: cost of the action-input path and of host-resident inputs / outputs of the fused step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_narde_b200 import VecNardeEnv, _cabi

E = 131072
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, n=100):
    ev = []
    for _ in range(n):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in ev)
    return "mean %.4f p50 %.4f min %.4f" % (sum(t) / n, t[n // 2], t[0])

def fresh():
    env = VecNardeEnv(E, seed=0x5EED, max_actions=64, graph=False)
    env.reset()
    for _ in range(300):
        env.step()
    torch.cuda.synchronize()
    return env

frac = torch.randint(-(1 << 31), (1 << 31) - 1, (E,), dtype=torch.int64).to(torch.int32)
h_act = frac.clone().pin_memory(); d_act = frac.cuda()
h_rew = torch.zeros(E).pin_memory(); h_done = torch.zeros(E, dtype=torch.uint8).pin_memory(); h_tr = torch.zeros(E, dtype=torch.uint8).pin_memory()

def make(env, act, rew, done, tr):
    def f():
        env.step_count += 1
        env._step_dev.fill_(env.step_count)
        _cabi.step_full(env.lo, env.hi, env.env_base, env.seed, 0, action_idx=act, actions=env.actions, counts=env.counts,
                        dice_out=env.dice, chosen=env.chosen, obs198=env.obs, reward=rew, done=done, stats=env.stats,
                        flags=2 | (32 if act is not None else 0), max_episode_steps=1000, truncated=tr,
                        workspace=env._workspaces[0], step_dev=env._step_dev)
    return f

for name, (a, host_out) in (("random (Philox), device results", (None, False)), ("device actions, device results", (d_act, False)),
                            ("host actions,   device results", (h_act, False)), ("device actions, host results  ", (d_act, True)),
                            ("host actions,   host results  ", (h_act, True)), ("random (Philox), host results  ", (None, True))):
    env = fresh()
    f = make(env, a, h_rew if host_out else env.reward, h_done if host_out else env.done, h_tr if host_out else env.trunc)
    for _ in range(5):
        f()
    print("%-34s" % name, timed(f), flush=True)

# fresh uniform fractions every step (pre-generated), device-resident: is the slowdown above the constant policy?
env = fresh()
pool = [torch.randint(-(1 << 31), (1 << 31) - 1, (E,), dtype=torch.int64, device="cuda").to(torch.int32) for _ in range(110)]
it = iter(pool)
def g():
    a = next(it)
    env.step_count += 1
    env._step_dev.fill_(env.step_count)
    _cabi.step_full(env.lo, env.hi, env.env_base, env.seed, 0, action_idx=a, actions=env.actions, counts=env.counts,
                    dice_out=env.dice, chosen=env.chosen, obs198=env.obs, reward=env.reward, done=env.done, stats=env.stats,
                    flags=2 | 32, max_episode_steps=1000, truncated=env.trunc, workspace=env._workspaces[0], step_dev=env._step_dev)
for _ in range(5):
    g()
print("%-34s" % "fresh device fractions each step", timed(g, 100), flush=True)
