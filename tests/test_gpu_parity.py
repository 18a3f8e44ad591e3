"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle and the golden fixtures."""
import numpy as np
import pytest

import parity as P
from gym_narde_b200 import state as S

pytestmark = pytest.mark.gpu


def test_gpu_golden_valid_moves(cuda_backend):
    assert P.check_golden_valid_moves(cuda_backend) > 1000


def test_gpu_golden_step_traces(cuda_backend):
    assert P.check_golden_step_traces(cuda_backend) > 2000


def test_gpu_tier_n_kat(cuda_backend):
    assert P.check_tier_n_kat(cuda_backend) > 300


def test_gpu_half_moves_vs_oracle(cuda_backend):
    lo, hi = P.pack_corpus(P.selfplay_corpus(40, 101))
    n = lo.shape[0]
    rng = np.random.RandomState(0)
    dice4 = np.zeros((n, 4), np.uint8)
    k = rng.randint(0, 3, size=n)
    d = rng.randint(1, 7, size=(n, 4))
    dice4[:, 0] = d[:, 0]
    dice4[k >= 1, 1] = d[k >= 1, 1]
    dice4[k == 2] = d[k == 2, :1]
    P.check_half_moves_vs_oracle(cuda_backend, lo, hi, dice4)
    b, off, ft = P.synthetic_boards(6000, 102)
    lo, hi, _, _ = P.pack_mover_boards(b, off, ft, 103)
    dice4 = np.zeros((6000, 4), np.uint8)
    dice4[:, :2] = P.random_dice(6000, 104)
    P.check_half_moves_vs_oracle(cuda_backend, lo, hi, dice4)


def test_gpu_enumerate_selfplay(cuda_backend):
    lo, hi = P.pack_corpus(P.selfplay_corpus(60, 105))
    assert P.check_enumerate_vs_oracle(cuda_backend, lo, hi, P.random_dice(lo.shape[0], 106, 0.3)) > 0


def test_gpu_enumerate_synthetic_block_rule_heavy(cuda_backend):
    b, off, ft = P.synthetic_boards(12000, 107)
    lo, hi, _, _ = P.pack_mover_boards(b, off, ft, 108)
    assert P.check_enumerate_vs_oracle(cuda_backend, lo, hi, P.random_dice(12000, 109, 0.5)) > 0


def test_gpu_enumerate_overflow(cuda_backend):
    b, off, ft = P.synthetic_boards(1000, 110)
    lo, hi, _, _ = P.pack_mover_boards(b, off, ft, 111)
    P.check_enumerate_vs_oracle(cuda_backend, lo, hi, P.random_dice(1000, 112, 0.7), cap=8)


def test_gpu_enumerate_fast_path(cuda_backend):
    """narde_enumerate_fast (what VecNardeEnv.get_valid_actions calls) against the oracle; states untouched."""
    cuda_backend.enumerate_fast = True
    try:
        lo, hi = P.pack_corpus(P.selfplay_corpus(60, 105))
        assert P.check_enumerate_vs_oracle(cuda_backend, lo, hi, P.random_dice(lo.shape[0], 106, 0.3)) > 0
        b, off, ft = P.synthetic_boards(12000, 107)
        lo, hi, _, _ = P.pack_mover_boards(b, off, ft, 108)
        assert P.check_enumerate_vs_oracle(cuda_backend, lo, hi, P.random_dice(12000, 109, 0.5)) > 0
        P.check_enumerate_vs_oracle(cuda_backend, lo[:1000].copy(), hi[:1000].copy(), P.random_dice(1000, 112, 0.7), cap=8)
    finally:
        cuda_backend.enumerate_fast = False


def test_gpu_step_ref_lockstep(cuda_backend):
    assert P.check_step_ref_lockstep(cuda_backend, 200, 500, 42) > 0


def test_gpu_step_full_lockstep_config2_4096(cuda_backend):
    """BASELINE config 2: 4096 lock-step envs, identical (Philox) dice, every env checked against the
    oracle at every step: legal-action lists, chosen action, next state, reward, done, Box(198)."""
    assert P.check_step_full_lockstep(cuda_backend, 4096, 110, 0x5EED, env_base=0, cap=64) > 1000


def test_gpu_step_full_mover_reward_no_autoreset(cuda_backend):
    P.check_step_full_lockstep(cuda_backend, 300, 220, 99, cap=4, flags=1)


def test_gpu_step_full_caller_actions_index_and_fraction(cuda_backend):
    P.check_step_full_lockstep(cuda_backend, 512, 150, 31, action_mode="index")
    P.check_step_full_lockstep(cuda_backend, 512, 150, 32, action_mode="fraction")


def test_gpu_obs198(cuda_backend):
    lo, hi = P.pack_corpus(P.selfplay_corpus(30, 113))
    P.check_obs198(cuda_backend, lo, hi)
    # ragged sizes around the CTA size (128) exercise the 8- and 16-byte store paths
    for n in (1, 2, 127, 129, 255):
        P.check_obs198(cuda_backend, lo[:n].copy(), hi[:n].copy())


def test_gpu_empty_batch(cuda_backend):
    """n == 0 is a no-op, not an error (reference: an empty list of envs)."""
    import torch
    from gym_narde_b200 import _cabi
    lo = torch.zeros((0, 16), dtype=torch.uint8, device="cuda")
    hi = torch.zeros((0, 16), dtype=torch.uint8, device="cuda")
    _cabi.obs198(lo, hi, torch.zeros((0, 198), device="cuda"))
    _cabi.reset(lo, hi, 0, 0, 0)
    _cabi.step_full(lo, hi, 0, 0, 1)
    torch.cuda.synchronize()


def _hash_state(lo, hi):
    import torch
    return int((lo.to(torch.int64).sum() * 1000003 + (hi.to(torch.int64) * torch.arange(1, 17, device=hi.device)).sum()).item())


def test_gpu_full_size_properties_131072():
    """Config 4 shard size (131072 envs/GPU): size-independent properties over 300 fused steps."""
    import torch
    from gym_narde_b200 import VecNardeEnv
    n, T = 131072, 300
    env = VecNardeEnv(n, seed=1234, max_actions=32)
    env.reset()
    sum_counts = 0
    for t in range(T):
        obs, rew, term, trunc, info = env.step()
        if t % 50 == 49:
            u_lo, u_hi = env.lo.cpu().numpy(), env.hi.cpu().numpy()
            u = S.unpack_states(u_lo, u_hi)
            white = np.maximum(u["board"], 0).sum(1) + u["off_w"]
            black = np.maximum(-u["board"], 0).sum(1) + u["off_b"]
            assert (white == 15).all() and (black == 15).all()           # checker conservation
            assert ((u["turn"] == 1) | (u["turn"] == -1)).all()
            o = obs.cpu().numpy()
            assert o.min() >= 0 and o[:, 196:].sum(1).min() == 1 and o[:, 196:].sum(1).max() == 1
            # Box(198) rows re-derived from the state planes
            chk = o[:, 0:96:4].sum(1) + o[:, 98:194:4].sum(1)
            assert (chk == (u["board"] != 0).sum(1)).all()
    st = env.episode_stats()
    assert st["episodes"] > n and st["white_wins"] + st["black_wins"] == st["episodes"]
    assert 60 < st["episode_steps"] / st["episodes"] < 140               # ~95 turns per game
    assert abs(st["white_wins"] / st["episodes"] - 0.5) < 0.05
    h1 = _hash_state(env.lo, env.hi)
    # determinism: same seed -> identical trajectory
    env2 = VecNardeEnv(n, seed=1234, max_actions=32)
    env2.reset()
    for t in range(T):
        env2.step()
    assert _hash_state(env2.lo, env2.hi) == h1 and env2.episode_stats() == st
    # sharding invariance: two half-size shards with global env ids == one run
    a = VecNardeEnv(n // 2, seed=1234, max_actions=32, env_base=0)
    b = VecNardeEnv(n // 2, seed=1234, max_actions=32, env_base=n // 2)
    a.reset()
    b.reset()
    for t in range(T):
        a.step()
        b.step()
    assert torch.equal(torch.cat([a.lo, b.lo]), env.lo) and torch.equal(torch.cat([a.hi, b.hi]), env.hi)


def test_gpu_vec_env_api_and_policy_loop():
    """roll -> get_valid_actions -> step(action_idx, dice) equals the fused random step when the
    same indices are chosen; reference-rules vec env returns int32[24] observations."""
    import torch
    from gym_narde_b200 import VecNardeEnv
    n = 2048
    fused = VecNardeEnv(n, seed=7, max_actions=512)
    loop = VecNardeEnv(n, seed=7, max_actions=512)
    fused.reset()
    loop.reset()
    for t in range(60):
        fused.step()
        dice = loop.roll().clone()
        acts, counts, ovf = loop.get_valid_actions(dice)
        assert not ovf.any()
        assert torch.equal(dice, fused.dice) and torch.equal(counts, fused.counts)
        # pick the same action the fused kernel picked, by searching the enumerated list
        match = (acts == fused.chosen[:, None]) & (torch.arange(512, device="cuda")[None, :] < counts[:, None])
        idx = match.float().argmax(1).to(torch.int32)
        assert (match.any(1) | (counts == 0)).all()
        loop.step(idx, dice=dice)
        assert torch.equal(loop.lo, fused.lo) and torch.equal(loop.hi, fused.hi)
        assert torch.equal(loop.obs, fused.obs) and torch.equal(loop.reward, fused.reward)
    ref = VecNardeEnv(256, seed=3, rules="reference")
    obs, _ = ref.reset()
    assert obs.shape == (256, 24) and obs.dtype == torch.int32
    codes = torch.randint(0, 576, (256, 2), dtype=torch.int32, device="cuda")
    obs, rew, term, trunc, info = ref.step(codes)
    assert obs.abs().sum(1).eq(30).all() and rew.dtype == torch.int32


def test_gpu_block_kernel_equals_per_thread_kernel(cuda_backend):
    """k_step_full_v2 (CTA-cooperative) and k_step_full (thread per env) agree bit for bit."""
    from test_core_hostsim import _v1_v2_equal
    b, off, ft = P.synthetic_boards(20001, 31)
    lo, hi, _, _ = P.pack_mover_boards(b, off, ft, 32)
    dice = P.random_dice(20001, 33, 0.4)
    o = _v1_v2_equal(cuda_backend, lo, hi, seed=5, step=9, dice_in=dice, cap=48, flags=0)
    assert o["counts"].max() > 48
    idx = np.random.RandomState(1).randint(-2, 80, size=20001).astype(np.int32)
    _v1_v2_equal(cuda_backend, lo, hi, seed=5, step=9, dice_in=dice, action_idx=idx, cap=16, flags=1)
    _v1_v2_equal(cuda_backend, lo, hi, seed=5, step=9, cap=0, flags=2, want_actions=False)
    for n in (1, 127, 128, 129, 300):
        _v1_v2_equal(cuda_backend, lo[:n].copy(), hi[:n].copy(), seed=n, step=3, cap=64, flags=2)
    lo, hi = P.pack_corpus(P.selfplay_corpus(60, 41))
    for step in range(1, 6):
        _v1_v2_equal(cuda_backend, lo, hi, seed=77, step=step, cap=32, flags=2)


def test_gpu_edge_sizes_and_ragged_batches(cuda_backend):
    """Ragged batch sizes around both CTA tiles (32 / 128 envs) and the small-batch switch (16384):
    the fused step must agree with the thread-per-env kernel bit for bit; cap = 0 (count only) and cap = 1."""
    from test_core_hostsim import _v1_v2_equal
    b, off, ft = P.synthetic_boards(16500, 77)
    lo, hi, _, _ = P.pack_mover_boards(b, off, ft, 78)
    for n in (1, 31, 32, 33, 127, 129, 16383, 16384, 16385):
        _v1_v2_equal(cuda_backend, lo[:n].copy(), hi[:n].copy(), seed=n, step=5, cap=24, flags=2)
    _v1_v2_equal(cuda_backend, lo[:700].copy(), hi[:700].copy(), seed=3, step=2, cap=0, flags=0, want_actions=False)
    _v1_v2_equal(cuda_backend, lo[:700].copy(), hi[:700].copy(), seed=3, step=2, cap=1, flags=1)


def test_device_advance_mode_equals_host_stepped_calls():
    """NARDE_DEVICE_ADVANCE through the C ABI: the step index is *step_dev + 1, the kernels store it back and leave the
    workspace header (list length, arrival counters) zero -- six consecutive calls with no memset / counter kernel between them
    give exactly the states and outputs of six calls with the step number passed from the host; both CTA tiles."""
    import torch
    from gym_narde_b200 import _cabi
    dev = torch.device("cuda")
    for n in (3000, 40000):
        def alloc():
            lo = torch.zeros((n, 16), dtype=torch.uint8, device=dev)
            hi = torch.zeros((n, 16), dtype=torch.uint8, device=dev)
            _cabi.reset(lo, hi, 7, 0xD1CE, 0)
            return dict(lo=lo, hi=hi, actions=torch.zeros((n, 16), dtype=torch.int64, device=dev),
                        counts=torch.zeros(n, dtype=torch.int32, device=dev), dice=torch.zeros((n, 2), dtype=torch.uint8, device=dev),
                        chosen=torch.zeros(n, dtype=torch.int64, device=dev), obs=torch.zeros((n, 198), device=dev),
                        rew=torch.zeros(n, device=dev), done=torch.zeros(n, dtype=torch.uint8, device=dev),
                        stats=torch.zeros(9, dtype=torch.int64, device=dev))

        def call(b, step, flags, ws, step_dev):
            _cabi.step_full(b["lo"], b["hi"], 7, 0xD1CE, step, actions=b["actions"], counts=b["counts"], dice_out=b["dice"],
                            chosen=b["chosen"], obs198=b["obs"], reward=b["rew"], done=b["done"], stats=b["stats"],
                            flags=flags, max_episode_steps=0, workspace=ws, step_dev=step_dev)

        a, b = alloc(), alloc()
        ws_a = torch.zeros(n + 8, dtype=torch.int32, device=dev)
        ws_b = torch.zeros(n + 8, dtype=torch.int32, device=dev)
        for k in range(40):                                   # into the middle game, where turns get deferred
            call(a, 1 + k, _cabi.AUTORESET, ws_a, None)
            call(b, 1 + k, _cabi.AUTORESET, ws_a, None)
        ctr = torch.full((1,), 40, dtype=torch.int64, device=dev)
        deferred = 0
        for k in range(6):
            call(a, 41 + k, _cabi.AUTORESET, ws_a, None)
            call(b, 0, _cabi.AUTORESET | _cabi.DEVICE_ADVANCE, ws_b, ctr)
            deferred += int(ws_a[0].item())
            assert int(ctr.item()) == 41 + k
            assert ws_b[:3].tolist() == [0, 0, 0]                                 # the counters are left clean
            assert int(ws_b[3].item()) == int(ws_a[0].item())                     # ... the count kept for diagnostics
            for key in a:
                assert torch.equal(a[key], b[key]), (n, k, key)
        assert deferred > 0          # the exact kernel had work on the way
    # the mode needs its workspace and counter, and is not an enumerate-only call
    lib = _cabi.load()
    with pytest.raises(_cabi.NardeCudaError):
        call(b, 0, _cabi.DEVICE_ADVANCE, ws_b, None)
    with pytest.raises(_cabi.NardeCudaError):
        call(b, 0, _cabi.DEVICE_ADVANCE, ws_a[:n + 2], ctr)     # shorter than NARDE_WORKSPACE_INTS(n)


@pytest.mark.gpu
def test_chunked_multi_stream_step_equals_unchunked():
    """VecNardeEnv(chunks=k): the step launched as k sub-batches on k streams (each with its own deferred-turn list,
    early programmatic trigger and exact kernel) plays exactly the turns of the single launch."""
    import torch
    from gym_narde_b200 import VecNardeEnv
    n = 70000
    ref = VecNardeEnv(n, seed=41, max_actions=32, graph=False)
    ref.reset()
    others = [VecNardeEnv(n, seed=41, max_actions=32, chunks=k, graph=g) for k, g in ((2, False), (3, True))]
    for e in others:
        e.reset()
    for t in range(90):
        ref.step()
        for e in others:
            e.step()
            assert torch.equal(ref.lo, e.lo) and torch.equal(ref.hi, e.hi), t
            assert torch.equal(ref.obs, e.obs) and torch.equal(ref.reward, e.reward) and torch.equal(ref.done, e.done), t
            assert torch.equal(ref.counts, e.counts) and torch.equal(ref.chosen, e.chosen), t
    for e in others:
        assert ref.episode_stats() == e.episode_stats()
