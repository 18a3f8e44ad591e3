"""GPU debug: which half of the zero-copy host step costs what -- the action words read from pinned host memory, or
the results written into it (direct C-ABI calls, graph-free, L2 flushed in front of every timed turn)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_narde_b200 import VecNardeEnv, _cabi
E, POOL = 131072, 16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
pool_h = torch.randint(-(1 << 31), (1 << 31) - 1, (POOL, E), dtype=torch.int64).to(torch.int32).pin_memory()
pool_d = pool_h.cuda()
rew_h, done_h, tr_h = torch.zeros(E).pin_memory(), torch.zeros(E, dtype=torch.uint8).pin_memory(), torch.zeros(E, dtype=torch.uint8).pin_memory()
for name, host_in, host_out in (("device in / device out", 0, 0), ("HOST in / device out", 1, 0), ("device in / HOST out", 0, 1), ("HOST in / HOST out", 1, 1)):
    env = VecNardeEnv(E, seed=0x5EED, max_actions=64, graph=False)
    env.reset()
    for _ in range(300):
        env.step()
    k = [0]
    def f():
        k[0] += 1
        env.step_count += 1
        env._step_dev.fill_(env.step_count)
        src = (pool_h if host_in else pool_d)[k[0] % POOL]
        _cabi.step_full(env.lo, env.hi, env.env_base, env.seed, 0, action_idx=src, actions=env.actions, counts=env.counts,
                        dice_out=env.dice, chosen=env.chosen, obs198=env.obs, reward=rew_h if host_out else env.reward,
                        done=done_h if host_out else env.done, stats=env.stats, flags=2 | 32, max_episode_steps=1000,
                        truncated=tr_h if host_out else env.trunc, workspace=env._workspaces[0], step_dev=env._step_dev)
    for _ in range(5):
        f()
    ev = []
    for _ in range(80):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k[0] += 1
        env.step_count += 1
        env._step_dev.fill_(env.step_count)
        a.record()
        src = (pool_h if host_in else pool_d)[k[0] % POOL]
        _cabi.step_full(env.lo, env.hi, env.env_base, env.seed, 0, action_idx=src, actions=env.actions, counts=env.counts,
                        dice_out=env.dice, chosen=env.chosen, obs198=env.obs, reward=rew_h if host_out else env.reward,
                        done=done_h if host_out else env.done, stats=env.stats, flags=2 | 32, max_episode_steps=1000,
                        truncated=tr_h if host_out else env.trunc, workspace=env._workspaces[0], step_dev=env._step_dev)
        b.record(); ev.append((a, b))
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in ev)
    print("%-26s mean %.4f p50 %.4f min %.4f" % (name, sum(t) / len(t), t[len(t) // 2], t[0]), flush=True)
