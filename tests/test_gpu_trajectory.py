"""GPU (-m gpu): trajectory ring, exact checkpoint/resume, trainer action-code bridge."""
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_ring_records_and_exact_resume():
    import torch
    from gym_narde_b200 import VecNardeEnv
    from gym_narde_b200.trajectory import TrajectoryRing
    n = 5000
    env = VecNardeEnv(n, seed=77, max_actions=32)
    env.reset()
    ring = TrajectoryRing(env, capacity=64)
    for t in range(50):
        ring.step()
        v = ring.view(t)
        assert torch.equal(v["lo"], env.lo) and torch.equal(v["hi"], env.hi)
        assert torch.equal(v["action"], env.chosen) and torch.equal(v["reward"], env.reward)
        assert torch.equal(v["dice"], env.dice)
        assert torch.equal(v["done"], env.done | (env.trunc << 1))
    buf = io.BytesIO()
    torch.save(ring.state_dict(), buf)
    for t in range(40):                       # wraps the ring (capacity 64)
        ring.step()
    lo_a, hi_a, rec_a, stats_a = env.lo.clone(), env.hi.clone(), ring.records.clone(), env.stats.clone()
    with pytest.raises(IndexError):
        ring.view(10)
    # resume from the checkpoint in a fresh env: bit-identical continuation (Philox keyed on (seed, env, step))
    env2 = VecNardeEnv(n, seed=1, max_actions=32)
    env2.reset()
    ring2 = TrajectoryRing(env2, capacity=64)
    buf.seek(0)
    ring2.load_state_dict(torch.load(buf))
    for t in range(40):
        ring2.step()
    assert torch.equal(env2.lo, lo_a) and torch.equal(env2.hi, hi_a)
    assert torch.equal(ring2.records, rec_a) and torch.equal(env2.stats, stats_a)


def test_action_codes_match_reference_convention():
    import torch
    from gym_narde_b200 import VecNardeEnv
    from gym_narde_b200 import state as S
    from gym_narde_b200.trajectory import action_codes
    env = VecNardeEnv(2000, seed=5, max_actions=48)
    env.reset()
    for _ in range(45):
        env.step()
    acts, counts, _ = env.get_valid_actions()
    codes = action_codes(env).cpu().numpy()
    a, c = acts.cpu().numpy().view(np.uint64), counts.cpu().numpy()
    seen_off = seen_double = 0
    for i in range(0, 2000, 7):
        for k in range(48):
            if k >= c[i]:
                assert (codes[i, k] == 0).all()
                continue
            moves = S.decode_action(int(a[i, k]))
            code = lambda m: m[0] * 24 + (0 if m[1] == 'off' else m[1])      # train_deepq_pytorch.py:432-437
            want = [code(moves[0]) if moves else 0, code(moves[1]) if len(moves) > 1 else 0, len(moves)]
            assert codes[i, k].tolist() == want, (i, k, moves)
            seen_off += any(m[1] == 'off' for m in moves)
            seen_double += len(moves) > 2
    assert seen_double > 0
