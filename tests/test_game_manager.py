"""The interactive turn manager (gym_narde_b200/narde_game_manager.py).  CPU: its host logic
(tree of partial turns, max-dice / higher-die / head budget composition) with the two rules kernels supplied by the
test-only host build of the same device headers, against the oracle.  GPU: the same checks through the C ABI, the
reference's own manager tests (tests/test_narde_game_manager.py:77-129, 237-260) and a full interactive game."""
import random

import numpy as np
import pytest

import parity as P
from gym_narde_b200 import state as S


class _Game:
    """Host-side fields of Narde (narde.py:21-29) for the CPU tests, where the CUDA facade cannot be constructed."""

    def __init__(self):
        self.board = np.zeros(24, dtype=np.int32)
        self.board[23], self.board[11] = 15, -15
        self.borne_off_white = self.borne_off_black = 0
        self.first_turn_white = self.first_turn_black = True


class _Env:
    def __init__(self):
        self.game = _Game()


def _corpora():
    lo, hi = P.pack_corpus(P.selfplay_corpus(6, 11))
    dice = P.random_dice(lo.shape[0], 12, 0.35)
    b, off, ft = P.synthetic_boards(250, 13)
    lo2, hi2, _, _ = P.pack_mover_boards(b, off, ft, 14)
    return (lo, hi, dice), (lo2, hi2, P.random_dice(250, 15, 0.5))


def _play_one_game(mgr, rng, max_turns=400):
    """Random interactive game through roll_dice / make_move only; returns (turns, half-moves)."""
    color, turns, halves = "white", 0, 0
    while turns < max_turns:
        dice, by_piece = mgr.roll_dice(color)
        turns += 1
        while by_piece:
            src = rng.choice(sorted(by_piece))
            mv = rng.choice(by_piece[src])
            assert mv[1] == "off" or mv[1] in mgr.get_valid_moves_for_position(src, color)
            res = mgr.make_move(mv[0], mv[1], color)
            assert "error" not in res, res
            halves += 1
            by_piece = res.get("valid_moves_by_piece", {}) if res.get("needs_next_move") else {}
            assert res.get("turn_complete") or by_piece
        over, winner = mgr.is_game_over()
        if over:
            assert winner == color
            return turns, halves
        color = "black" if color == "white" else "white"
    raise AssertionError("game did not finish")


def _manager_kats(make):
    # second head move in the same turn is rejected (reference tests/test_narde_game_manager.py:77-95)
    mgr = make()
    mgr.game.first_turn_white = False
    dice, by = mgr.roll_dice("white", dice=(6, 5))
    assert dice == [6, 5] and by == {23: [(23, 17), (23, 18)]}
    assert "error" not in mgr.make_move(23, 17, "white")
    bad = mgr.make_move(23, 18, "white")
    assert "head position" in bad["error"].lower()
    assert mgr.get_game_state()["dice_remaining"] == [5]
    res = mgr.make_move(17, 12, "white")
    assert res.get("turn_complete") and res["board"][12] == 1 and res["board"][23] == 14
    # opening 6-6: two head moves, the third is rejected (:97-129); Black's 11 blocks 17->11, so the turn ends
    mgr = make()
    dice, by = mgr.roll_dice("white", dice=(6, 6))
    assert mgr.is_first_turn_special_doubles and mgr.max_head_moves["white"] == 2
    r1 = mgr.make_move(23, 17, "white")
    assert r1.get("needs_next_move") and r1["total_moves"] == 4
    r2 = mgr.make_move(23, 17, "white")
    assert "error" not in r2
    if not r2.get("turn_complete"):
        assert "head position" in mgr.make_move(23, 17, "white")["error"].lower()
    else:
        assert r2["board"][23] == 13 and r2["board"][17] == 2
    # errors leave the position alone; undo goes back to the roll
    mgr = make()
    mgr.game.first_turn_white = False
    mgr.game.board[23], mgr.game.board[20], mgr.game.board[18] = 13, 1, 1
    mgr.roll_dice("white", dice=(3, 5))
    before = mgr.game.board.copy()
    assert "piece" in mgr.make_move(10, 5, "white")["error"]
    assert "die" in mgr.make_move(20, 16, "white")["error"]
    assert (mgr.game.board == before).all()
    assert "error" not in mgr.make_move(20, 15, "white")
    undone = mgr.undo_moves()
    assert undone["board"] == before.tolist() and undone["dice"] == [5, 3]
    assert sorted(mgr.get_valid_moves_for_position(20, "white")) == [15, 17]
    # Black moves in its own frame (head 23 = absolute 11); the board in the response is absolute
    mgr = make()
    dice, by = mgr.roll_dice("black", dice=(2, 1))
    res = mgr.make_move(23, 21, "black")
    assert "error" not in res and res["board"][9] == -1 and res["board"][11] == -14
    # game over <=> 15 borne off (:237-260), bear-off reported as -1
    mgr = make()
    g = mgr.game
    g.board[:] = 0
    g.board[2], g.board[14] = 1, -15
    g.borne_off_white, g.first_turn_white, g.first_turn_black = 14, False, False
    assert mgr.is_game_over() == (False, None)
    mgr.roll_dice("white", dice=(6, 4))
    assert mgr.get_valid_moves_for_position(2, "white") == [-1]
    assert mgr.make_move(2, -1, "white").get("turn_complete")
    assert mgr.is_game_over() == (True, "white") and mgr.get_game_state()["winner"] == "white"


def _rule_kats(make):
    # no playable die: the roll offers nothing and the position stays as it is (a passed turn)
    mgr = make()
    g = mgr.game
    g.board[:] = 0
    g.board[23], g.board[17], g.board[18] = 15, -7, -8
    g.first_turn_white = g.first_turn_black = False
    dice, by = mgr.roll_dice("white", dice=(6, 5))
    assert by == {} and mgr.get_valid_moves_for_position(23, "white") == []
    assert "error" in mgr.make_move(23, 17, "white") and g.board[23] == 15
    # either die can be played but not both: the higher one must be (narde.py:6 rule 4)
    mgr = make()
    g = mgr.game
    g.board[:] = 0
    g.board[23], g.board[10], g.board[17], g.board[18] = 1, 1, -7, -8
    g.borne_off_white, g.first_turn_white, g.first_turn_black = 13, False, False
    dice, by = mgr.roll_dice("white", dice=(5, 6))
    assert dice == [6, 5] and by == {10: [(10, 4)]}
    assert "error" in mgr.make_move(10, 5, "white")
    assert mgr.make_move(10, 4, "white").get("turn_complete") and g.board[4] == 1
    # both dice when possible: a first half-move that would strand the second die is not offered
    mgr = make()
    g = mgr.game
    g.board[:] = 0
    g.board[23], g.board[10], g.board[17], g.board[18], g.board[4] = 1, 1, -7, -7, -1
    g.borne_off_white, g.first_turn_white, g.first_turn_black = 13, False, False
    dice, by = mgr.roll_dice("white", dice=(6, 5))      # 10->4 is blocked; 10->5 then nothing for the 6
    assert by == {10: [(10, 5)]}


def test_manager_rule_kats_cpu(hostsim):
    from gym_narde_b200.narde_game_manager import NardeGameManager
    _rule_kats(lambda: NardeGameManager(_Env(), _ops=hostsim))


def test_turn_tree_vs_oracle_cpu(hostsim):
    for lo, hi, dice in _corpora():
        assert P.check_turn_tree_vs_oracle(hostsim, lo, hi, dice) > 0


def test_manager_kats_and_game_cpu(hostsim):
    from gym_narde_b200.narde_game_manager import NardeGameManager
    _manager_kats(lambda: NardeGameManager(_Env(), _ops=hostsim))
    turns, halves = _play_one_game(NardeGameManager(_Env(), _ops=hostsim), random.Random(3))
    assert turns > 20 and halves > turns


class _EnvObs(_Env):
    def _get_obs(self):               # what the reference env hands the model (narde_env.py:24-25)
        return np.asarray(self.game.board, dtype=np.int32)


def _ai_turns(make):
    """execute_ai_moves (my_game/narde_game_manager.py:890-1099): response keys, the greedy Q[from*24+to] choice
    among the offered half-moves, the no-move branch and the game-over branch."""
    import torch

    class Net(torch.nn.Module):       # a fixed scorer: prefers low sources, then short moves -> the choice is predictable
        def forward(self, x):
            q = torch.zeros(1, 576)
            for f in range(24):
                for t in range(24):
                    q[0, f * 24 + t] = -f + 0.01 * t
            return q

    net = Net()
    mgr = make()
    res = mgr.execute_ai_moves(net, "cpu", dice=(6, 5))
    assert set(res) >= {"board", "current_player", "dice", "valid_moves_by_piece", "ai_moves", "ai_moved_from_head",
                        "ai_dice", "borne_off"}
    assert res["ai_dice"] == [6, 5] and res["current_player"] == "white" and res["ai_moved_from_head"] is True
    # opening: only the head has checkers; one head move per turn, then the moved checker carries on
    assert res["ai_moves"] == [{"from": 23, "to": 17}, {"from": 17, "to": 12}] or \
        res["ai_moves"] == [{"from": 23, "to": 18}, {"from": 18, "to": 12}]
    assert res["board"][11] == -14 and res["board"][0] == -1          # black checker: mover point 12 = absolute point 0
    assert mgr.current_player == "white" and res["valid_moves_by_piece"]
    # among several offered half-moves the arg-max of Q[from*24 + to] is played: source 3 beats source 10
    mgr = make()
    g = mgr.game
    g.board[:] = 0
    g.board[12 + 10], g.board[12 + 3], g.board[5] = -1, -1, 15       # black's points 10 and 3 (mover frame), white far away
    g.borne_off_black, g.first_turn_white, g.first_turn_black = 13, False, False
    res = mgr.execute_ai_moves(net, "cpu", dice=(2, 1))
    assert res["ai_moves"][0]["from"] == 3
    # no playable die: ai_had_no_moves and White is on roll
    mgr = make()
    g = mgr.game
    g.board[:] = 0
    g.board[11], g.board[5], g.board[6] = -15, 7, 8                   # black head (mover 23) blocked at 17 and 18
    g.first_turn_white = g.first_turn_black = False
    res = mgr.execute_ai_moves(net, "cpu", dice=(6, 5))
    assert res.get("ai_had_no_moves") is True and res["current_player"] == "white" and "ai_moves" not in res
    # last checker borne off: game over, winner black
    mgr = make()
    g = mgr.game
    g.board[:] = 0
    g.board[12 + 2], g.board[20] = -1, 15
    g.borne_off_black, g.first_turn_white, g.first_turn_black = 14, False, False
    res = mgr.execute_ai_moves(net, "cpu", dice=(6, 4))
    assert res["game_over"] is True and res["winner"] == "black" and res["ai_moves"] == [{"from": 2, "to": -1}]


def test_execute_ai_moves_cpu(hostsim):
    from gym_narde_b200.narde_game_manager import NardeGameManager
    _ai_turns(lambda: NardeGameManager(_EnvObs(), _ops=hostsim))


def test_manager_needs_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from gym_narde_b200 import _cabi
    from gym_narde_b200.narde_game_manager import NardeGameManager
    with pytest.raises(_cabi.NardeCudaError):
        NardeGameManager(_Env())


@pytest.mark.gpu
def test_turn_tree_vs_oracle_gpu():
    from gym_narde_b200.narde_game_manager import _CudaOps
    ops = _CudaOps()
    for lo, hi, dice in _corpora():
        assert P.check_turn_tree_vs_oracle(ops, lo, hi, dice) > 0


@pytest.mark.gpu
def test_manager_on_the_facade_gpu():
    import gym_narde_b200
    from gym_narde_b200.narde_game_manager import NardeGameManager

    def make():
        env = gym_narde_b200.make("narde-v0")
        env.reset(seed=1)
        return NardeGameManager(env)

    _manager_kats(make)
    _rule_kats(make)
    _ai_turns(make)
    # the AI turn driven by the tcgen05 scorer (AfterstateMLP.forward_states on the packed node)
    import torch
    import torch.nn as nn
    from gym_narde_b200 import AfterstateMLP
    torch.manual_seed(4)
    fn = nn.Sequential(nn.Linear(198, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU()).cuda()
    mlp = AfterstateMLP.from_module(fn, nn.Linear(256, 576).cuda())
    mgr = make()
    res = mgr.execute_ai_moves(mlp, dice=(4, 2))
    assert len(res["ai_moves"]) == 2 and res["current_player"] == "white"
    mgr = make()
    turns, halves = _play_one_game(mgr, random.Random(5))
    assert turns > 20 and halves > turns
    # the offered first half-moves are exactly the first half-moves of some ordering of get_valid_actions(roll)
    mgr = make()
    mgr.game.first_turn_white = False
    mgr.game.board[23], mgr.game.board[20], mgr.game.board[18] = 13, 1, 1
    _, by = mgr.roll_dice("white", dice=(3, 5))
    acts = mgr.game.get_valid_actions((3, 5), 1)
    assert len(mgr._tree.ends) == len(acts)
    assert {a[0] for a in acts} <= {m for ms in by.values() for m in ms}
