"""GPU (-m gpu): the batched afterstate-greedy actor (enumerate -> afterstates -> tcgen05 score -> argmax -> step).

Integer pieces are bit-exact: every afterstate row equals the state the (oracle-validated) fused step
produces when that action is played; the greedy index equals torch's segment arg-max of the same scores."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _net():
    import torch
    import torch.nn as nn
    torch.manual_seed(0)
    fn = nn.Sequential(nn.Linear(198, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU()).cuda()
    head = nn.Linear(256, 576).cuda()
    return fn, head


def test_afterstates_equal_fused_step_results():
    import torch
    from gym_narde_b200 import VecNardeEnv, AfterstateMLP, AfterstateActor
    fn, head = _net()
    mlp = AfterstateMLP.from_module(fn, head)
    n, cap = 3000, 96
    env = VecNardeEnv(n, seed=21, max_actions=cap, autoreset=False, max_episode_steps=0)
    actor = AfterstateActor(env, mlp)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    checked = 0
    for t in range(130):
        dice = env.roll().clone()
        acts, counts, ovf = env.get_valid_actions(dice)
        acts, counts = acts.clone(), counts.clone()
        off = actor.afterstates(acts, counts).clone()
        c = counts.clamp(max=cap).long()
        assert int(actor.rows_dev.item()) == int(c.sum().item())
        idx = (torch.rand(n, device="cuda", generator=g) * c.clamp(min=1)).long().clamp(max=cap - 1)
        done_before = env.done.bool().clone()
        env.step(idx.to(torch.int32), dice=dice)
        live = (c > 0) & ~done_before
        rows = (off + idx)[live]
        assert torch.equal(actor.as_lo[rows], env.lo[live]) and torch.equal(actor.as_hi[rows], env.hi[live]), t
        checked += int(live.sum().item())
    assert checked > 200000


def test_actor_choice_is_segment_argmax_and_step_plays_it():
    import torch
    from gym_narde_b200 import VecNardeEnv, AfterstateMLP, AfterstateActor
    fn, head = _net()
    mlp = AfterstateMLP.from_module(fn, head)
    n, cap = 2048, 64
    for mode in ("max", "white_value"):
        env = VecNardeEnv(n, seed=9, max_actions=cap)
        actor = AfterstateActor(env, mlp, mode=mode)
        env.reset()
        for t in range(60):
            turn = env.hi[:, 10].view(torch.int8).clone()          # +1 WHITE / -1 BLACK to move
            choice, dice = actor.choose()
            counts = env.counts.clone()
            c = counts.clamp(max=cap).long()
            K = int(actor.rows_dev.item())
            # scores of the packed afterstate rows == row max of forward_states on the same rows
            q = mlp.forward_states(actor.as_lo[:K].contiguous(), actor.as_hi[:K].contiguous())
            assert torch.equal(actor.scores[:K], q.max(1).values)
            # dense [n, cap] view of the ragged scores, -inf padded
            dense = torch.full((n, cap), float("-inf"), device="cuda")
            col = torch.arange(cap, device="cuda")[None, :].expand(n, cap)
            mask = col < c[:, None]
            dense[mask] = actor.scores[:K]
            if mode == "white_value":
                dense = torch.where(mask, dense * turn.float()[:, None], dense)
            want = dense.argmax(1)
            live = c > 0
            assert torch.equal(choice.long()[live], want[live]), (mode, t)
            chosen_act = env.actions[torch.arange(n, device="cuda"), choice.long()].clone()
            env.step(choice, dice=dice)
            assert torch.equal(env.chosen[live], chosen_act[live])


def test_actor_graph_replay_equals_eager_turns():
    """AfterstateActor.step_graph (one CUDA-graph replay per greedy turn, step number read from the device
    counter) plays exactly the turns of the eager step()."""
    import torch
    from gym_narde_b200 import VecNardeEnv, AfterstateMLP, AfterstateActor
    fn, head = _net()
    mlp = AfterstateMLP.from_module(fn, head)
    n, cap = 4096, 64
    envs = [VecNardeEnv(n, seed=13, max_actions=cap) for _ in range(2)]
    actors = [AfterstateActor(e, mlp) for e in envs]
    for e in envs:
        e.reset()
        for _ in range(30):
            e.step()
    for t in range(50):
        actors[0].step()
        actors[1].step_graph()
        assert torch.equal(envs[0].lo, envs[1].lo) and torch.equal(envs[0].hi, envs[1].hi), t
        assert torch.equal(envs[0].obs, envs[1].obs) and torch.equal(envs[0].reward, envs[1].reward)
        assert torch.equal(envs[0].chosen, envs[1].chosen)
    assert envs[0].episode_stats() == envs[1].episode_stats()
