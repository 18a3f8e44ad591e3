"""GPU (-m gpu): oracle parity AT THE BENCHMARKED CONFIGURATION.

bench.py times `VecNardeEnv.step()` on 131 072 envs per GPU (BASELINE config 4's shard): CUDA-graph replay,
`k_step_full_v2<128, true>` + its programmatic dependent `k_step_deferred`, NARDE_DEVICE_ADVANCE, action
capacity 64, env_base = rank * 131072.  Here exactly that call plays 131 072 envs x 300 steps (39 M env turns,
env_base != 0) and EVERY env turn is compared with the C oracle's bulk trace (oracle/narde_oracle.c
o_selfplay_trace, played on all host cores): packed state after the turn, legal-action count, checksum of the
stored action list, chosen action, dice, reward, terminated / truncated bits, the accumulated statistics, and the
Box(198) rows (re-derived from the compared states by an independent torch encoder that is itself checked against
the oracle's o_obs198).  The same for the host-facing `step_host` and for config 3's synthetic positions.
"""
import os

import numpy as np
import pytest

from gym_narde_b200 import state as S
from oracle import oracle as O

pytestmark = pytest.mark.gpu

N_FULL = 131072


def _torch_obs198(torch, lo, hi):
    """README.md:44-102 as plain torch ops on the packed planes (test-side, shares nothing with the kernels)."""
    b = torch.cat([lo.view(torch.int8), hi.view(torch.int8)[:, :8]], 1).to(torch.int32)       # [n, 24]
    off15 = torch.from_numpy(np.array([np.float32(k / 15.0) for k in range(16)], np.float32)).to(lo.device)
    out = []
    for colour, sign in ((0, 1), (1, -1)):
        c = (b * sign).clamp(min=0)
        f = torch.stack([(c >= 1).float(), (c >= 2).float(), (c >= 3).float(), (c - 3).clamp(min=0).float() * 0.5], 2)
        out.append(f.reshape(-1, 96))
        out.append(torch.zeros(b.shape[0], 1, device=lo.device))
        out.append(off15[hi[:, 8 + colour].long()][:, None])
    white = (hi.view(torch.int8)[:, 10] == 1).float()[:, None]
    out += [white, 1.0 - white]
    return torch.cat(out, 1)


def _gpu_list_hash(torch, actions, counts, w):
    cap = actions.shape[1]
    keep = torch.arange(cap, device=actions.device)[None, :] < counts.clamp(max=cap)[:, None]
    return (torch.where(keep, actions, torch.zeros_like(actions)) * w[None, :]).sum(1)


class _Recorder:
    """Per-step copies of everything a step produced, device-resident: [T, N, ...]."""

    def __init__(self, torch, env, T):
        n, dev = env.num_envs, env.device
        self.torch, self.env, self.t = torch, env, 0
        self.lo = torch.zeros((T, n, 16), dtype=torch.uint8, device=dev)
        self.hi = torch.zeros((T, n, 16), dtype=torch.uint8, device=dev)
        self.chosen = torch.zeros((T, n), dtype=torch.int64, device=dev)
        self.count = torch.zeros((T, n), dtype=torch.int32, device=dev)
        self.dice = torch.zeros((T, n, 2), dtype=torch.uint8, device=dev)
        self.done = torch.zeros((T, n), dtype=torch.uint8, device=dev)
        self.reward = torch.zeros((T, n), dtype=torch.float32, device=dev)
        self.hash = torch.zeros((T, n), dtype=torch.int64, device=dev)
        self.w = torch.from_numpy(O.list_weights(env.max_actions)).to(dev)
        self.obs_ok = True

    def record(self, reward, done, trunc):
        torch, env, t = self.torch, self.env, self.t
        self.lo[t], self.hi[t] = env.lo, env.hi
        self.chosen[t], self.count[t], self.dice[t] = env.chosen, env.counts, env.dice
        self.done[t] = done.to(env.device) | (trunc.to(env.device) << 1)
        self.reward[t] = reward.to(env.device)
        self.hash[t] = _gpu_list_hash(torch, env.actions, env.counts, self.w)
        self.obs_ok = self.obs_ok and bool(torch.equal(env.obs, _torch_obs198(torch, env.lo, env.hi)))
        self.t += 1

    def compare(self, seed, env_base, chunk=8192, init=None, words=None, word_mode=1, step0=0):
        """Oracle trace chunk by chunk (host threads), compared on the device.  Returns env turns checked."""
        torch, env, T = self.torch, self.env, self.t
        n = env.num_envs
        stats = np.zeros(8, np.int64)
        turns = 0
        for b in range(0, n, chunk):
            e = min(b + chunk, n)
            tr = O.selfplay_trace(seed, env_base + b, e - b, T, step0=step0,
                                  init=None if init is None else (init[0][b:e], init[1][b:e]),
                                  words=None if words is None else words[:, b:e], word_mode=word_mode,
                                  cap=env.max_actions, reward_mode=1 if env.reward_mode == "mover12" else 0,
                                  autoreset=env.autoreset, max_episode_steps=env.max_episode_steps)
            for key in ("count", "dice", "hash", "chosen", "reward", "done", "lo", "hi"):
                got = getattr(self, key)[:T, b:e]
                want = torch.from_numpy(tr[key]).to(env.device)
                if not torch.equal(got, want):
                    bad = (got != want).reshape(T, e - b, -1).any(2).nonzero()[0].tolist()
                    raise AssertionError("%s differs from the oracle first at step %d, env %d (global id %d): got %s want %s" % (
                        key, step0 + bad[0] + 1, b + bad[1], env_base + b + bad[1],
                        got[bad[0], bad[1]].tolist(), want[bad[0], bad[1]].tolist()))
            stats[:6] += tr["stats"][:6]
            stats[6] = max(stats[6], tr["stats"][6])
            stats[7] += tr["stats"][7]
            turns += tr["turns"]
        return turns, stats


def test_benchmark_config_131072_envs_300_graph_replayed_steps_vs_oracle():
    import torch
    from gym_narde_b200 import VecNardeEnv
    n, T, seed, base = N_FULL, 300, 0x5EED, 3 * N_FULL
    env = VecNardeEnv(n, seed=seed, max_actions=64, env_base=base)             # bench.py's constructor call (rank 3)
    env.reset()
    rec = _Recorder(torch, env, T)
    for t in range(T):
        obs, rew, term, trunc, info = env.step()
        rec.record(rew, env.done, env.trunc)
    # the timed path: one graph of two nodes (main kernel + programmatic dependent) that advances its own counter
    assert "random" in env._graphs and env._ws_adv is not None and env.use_graph
    assert int(env._step_dev.item()) == T and env._ws_adv[:3].tolist() == [0, 0, 0]
    turns, stats = rec.compare(seed, base)
    assert turns == n * T
    assert rec.obs_ok, "a Box(198) batch differed from the rows re-derived from the state planes"
    got = env.stats.cpu().numpy()
    assert (got[:8] == stats).all() and got[8] == 0, (got, stats)      # (slot 8: clamped caller indices -- none here)
    assert stats[0] > 2 * n and stats[7] > 0        # every env finished games on the way; the capacity overflowed somewhere
    print("benchmark-config parity: %d env turns bit-equal to the oracle; stats %s" % (turns, stats.tolist()))


def test_benchmark_config_step_host_vs_oracle():
    """The e2e path of bench.py: step_host(fraction=True) -- action words fetched from pinned host memory inside the
    kernel, reward / done / truncated written into pinned host memory -- 131 072 envs x 100 steps from a steady-state
    position (200 random steps first), every turn against the oracle trace driven by the same words."""
    import torch
    from gym_narde_b200 import VecNardeEnv
    n, T, seed, base, t0 = N_FULL, 100, 0xFACE, N_FULL, 200
    env = VecNardeEnv(n, seed=seed, max_actions=64, env_base=base)
    env.reset()
    for _ in range(t0):
        env.step()
    init = (env.lo.cpu().numpy(), env.hi.cpu().numpy())
    rng = np.random.default_rng(5)
    words = rng.integers(0, 1 << 32, (T, n), dtype=np.uint64).astype(np.uint32)
    pool = torch.from_numpy(words[:8].view(np.int32).copy()).pin_memory()      # a ring of 8 pinned rows, refilled on the way
    rec = _Recorder(torch, env, T)
    for t in range(T):
        row = pool[t % 8]
        row.copy_(torch.from_numpy(words[t].view(np.int32)))
        io = env.step_host(fraction=True, actions=row)
        torch.cuda.synchronize()
        rec.record(io["reward"], io["done"], io["truncated"])
    turns, stats = rec.compare(seed, base, init=init, words=words, word_mode=1, step0=t0)
    assert turns == n * T and rec.obs_ok
    print("step_host parity: %d env turns bit-equal to the oracle" % turns)


def test_truncation_and_small_tile_graph_replay_vs_oracle():
    """TimeLimit truncation + auto-reset through the graph-replayed step on the 32-env tile (n <= 16384) and a
    ragged size on the 128-env tile."""
    import torch
    from gym_narde_b200 import VecNardeEnv
    for n, T, mes in ((4096, 260, 60), (20011, 150, 45)):
        env = VecNardeEnv(n, seed=77, max_actions=16, env_base=9, max_episode_steps=mes)
        env.reset()
        rec = _Recorder(torch, env, T)
        for t in range(T):
            env.step()
            rec.record(env.reward, env.done, env.trunc)
        turns, stats = rec.compare(77, 9, chunk=4096)
        assert turns == n * T and rec.obs_ok
        assert (env.stats.cpu().numpy()[:8] == stats).all()
        assert int(rec.done.bitwise_and(2).ne(0).sum().item()) > n   # truncations happened


def test_torch_obs198_encoder_equals_oracle():
    """The test-side encoder used above, pinned against o_obs198 (README.md:44-102)."""
    import torch
    import parity as P
    lo, hi = P.pack_corpus(P.selfplay_corpus(20, 7))
    u = S.unpack_states(lo, hi)
    got = _torch_obs198(torch, torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()).cpu().numpy()
    for i in range(lo.shape[0]):
        ref = O.obs198(u["board"][i], int(u["off_w"][i]), int(u["off_b"][i]), int(u["turn"][i]))
        assert (ref == got[i]).all(), i


def test_step_host_packed_observation_is_lossless():
    """step_host(obs="packed"): the state planes the kernel mirrors into pinned host memory equal the device planes,
    and expanding them on the host gives the device's Box(198) batch bit for bit (deferred envs included)."""
    import torch
    import gym_narde_b200
    from gym_narde_b200 import VecNardeEnv
    for n in (N_FULL, 5000):
        env = VecNardeEnv(n, seed=21, max_actions=64, env_base=17)
        env.reset()
        for _ in range(60):
            env.step()
        rows = torch.randint(-(1 << 31), (1 << 31) - 1, (4, n), dtype=torch.int64).to(torch.int32).pin_memory()
        for t in range(12):
            io = env.step_host(fraction=True, actions=rows[t % 4], obs="packed")
            torch.cuda.synchronize()
            assert torch.equal(io["lo"], env.lo.cpu()) and torch.equal(io["hi"], env.hi.cpu()), (n, t)
            host_rows = gym_narde_b200.expand_obs198(io["lo"].numpy(), io["hi"].numpy())
            assert (host_rows == env.obs.cpu().numpy()).all(), (n, t)
            dev_rows = gym_narde_b200.expand_obs198(env.lo, env.hi)
            assert torch.equal(dev_rows, env.obs)
        twin = VecNardeEnv(n, seed=21, max_actions=64, env_base=17)      # the form bench.py's e2e_with_obs times: + one result byte
        twin.load_state_dict(env.state_dict())
        for t in range(6):
            io = env.step_host(fraction=True, actions=rows[t % 4], obs="packed", packed=True)
            twin.step(rows[t % 4].cuda(), fraction=True)
            torch.cuda.synchronize()
            assert torch.equal(io["lo"], twin.lo.cpu()) and torch.equal(io["hi"], twin.hi.cpu()), (n, t)
            want = twin.done.cpu() | (twin.trunc.cpu() << 1) | (twin.reward.cpu().to(torch.uint8) << 2)
            assert torch.equal(io["result"], want), (n, t)
        for t in range(8):                                               # obs="compact": one 20-byte record per env
            io = env.step_host(fraction=True, actions=rows[t % 4], obs="compact")
            twin.step(rows[t % 4].cuda(), fraction=True)
            torch.cuda.synchronize()
            lo_h, hi_h, res = S.unpack_compact(io["rec"].numpy())
            assert (lo_h == twin.lo.cpu().numpy()).all() and (hi_h == twin.hi.cpu().numpy()).all(), (n, t)
            want = (twin.done.cpu() | (twin.trunc.cpu() << 1) | (twin.reward.cpu().to(torch.uint8) << 2)).numpy()
            assert (res == want).all(), (n, t)
            assert torch.equal(env.obs, twin.obs)


def test_config3_synthetic_positions_vs_oracle():
    """BASELINE config 3's three strata (self-play / forced doubles / bear-off), 262 144 positions through
    narde_enumerate_fast (main kernel + exact kernel for the order-dependent doubles), every position's count and
    stored list against the oracle's o_enumerate_batch."""
    import torch
    from gym_narde_b200 import _cabi
    from gym_narde_b200.workloads import config3_positions
    n, cap = 1 << 18, 64
    lo, hi, dice, strata = config3_positions("cuda", n=n, seed=1234)
    actions = torch.zeros((n, cap), dtype=torch.int64, device="cuda")
    counts = torch.zeros(n, dtype=torch.int32, device="cuda")
    ovf = torch.zeros(n, dtype=torch.uint8, device="cuda")
    ws = torch.zeros(_cabi.workspace_ints(n), dtype=torch.int32, device="cuda")
    _cabi.enumerate_actions_fast(lo, hi, dice, actions, counts, ovf, ws)
    w = torch.from_numpy(O.list_weights(cap)).cuda()
    got_hash = _gpu_list_hash(torch, actions, counts, w).cpu().numpy()
    want_counts, want_hash = O.enumerate_batch(lo.cpu().numpy(), hi.cpu().numpy(), dice.cpu().numpy(), cap)
    assert (counts.cpu().numpy() == want_counts).all()
    assert (got_hash == want_hash).all()
    assert (ovf.cpu().numpy() == (want_counts > cap)).all()
    assert int(ws[0].item()) > n // 200                     # the exact kernel had its ~2 % share
    for name, (b, e) in strata.items():
        assert want_counts[b:e].max() > 20, name
