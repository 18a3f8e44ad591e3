"""Backend-agnostic parity checks: a backend (support.HostSim on the CPU, support.CudaBackend on
the GPU through the C ABI) is compared with the oracle and with the committed golden fixtures."""
from __future__ import annotations

import json
import os
import random

import numpy as np

from gym_narde_b200 import state as S
from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def moves_from_buffer(moves, count):
    return [(int(moves[k, 0]), -1 if moves[k, 1] == 255 else int(moves[k, 1])) for k in range(count)]


def act_moves(a):
    return [(f, -1 if t == 'off' else t) for f, t in S.decode_action(a)]


# ------------------------------------------------------------------------------------------
# position corpora
# ------------------------------------------------------------------------------------------
def selfplay_corpus(n_games, seed, uniform=True):
    """States along oracle full-rules random self-play games: list of state tuples."""
    rng = random.Random(seed)
    out = []
    for _ in range(n_games):
        oe = O.OracleEnv()
        oe.reset([rng.randint(1, 6) for _ in range(40)])
        for _ in range(400):
            out.append(oe.state_tuple())
            d = (rng.randint(1, 6), rng.randint(1, 6))
            st = oe.state_tuple()
            pl = st[5]
            b = np.array(st[0])
            mover = b if pl == 1 else S.rotate_board(b)
            _, n = O.turn_enumerate(mover, st[1] if pl == 1 else st[2], d[0], d[1], st[3] if pl == 1 else st[4], cap=1)
            idx = rng.randrange(n) if n else 0
            _, _, done, _ = oe.full_step(d, idx, want_obs=False)
            if done:
                break
    return out


def pack_corpus(corpus):
    boards = np.array([c[0] for c in corpus])
    lo, hi = S.pack_states(boards, np.array([c[1] for c in corpus]), np.array([c[2] for c in corpus]),
                           np.array([c[5] for c in corpus]), np.array([c[3] for c in corpus]).astype(bool),
                           np.array([c[4] for c in corpus]).astype(bool))
    return lo, hi


def synthetic_boards(n, seed):
    """Random (possibly unreachable) mover-frame boards stressing block / bear-off / head rules.
    Returns mover-frame boards, mover off counts, first_turn flags."""
    rng = random.Random(seed)
    boards, offs, fts = [], [], []
    for _ in range(n):
        kind = rng.random()
        b = [0] * 24
        nown, nopp = rng.randint(1, 15), rng.randint(1, 15)
        if kind < 0.35:
            base, width = rng.randint(0, 14), rng.randint(4, 10)
            po = [min(23, base + rng.randrange(width)) for _ in range(nown)]
            hi_only = rng.random() < 0.6
            pp = [rng.randint(12, 23) if hi_only else rng.randint(0, 23) for _ in range(nopp)]
        elif kind < 0.6:
            po = [rng.randint(0, 5) if rng.random() < 0.9 else rng.randint(6, 9) for _ in range(nown)]
            pp = [rng.randint(12, 17) for _ in range(nopp)]
        elif kind < 0.8:
            po = [23 if rng.random() < 0.6 else rng.randint(10, 22) for _ in range(nown)]
            pp = [rng.randint(0, 23) for _ in range(nopp)]
        else:
            po = [rng.randint(0, 23) for _ in range(nown)]
            pp = [rng.randint(0, 23) for _ in range(nopp)]
        for p in po:
            b[p] += 1
        for p in pp:
            if b[p] <= 0:
                b[p] -= 1
        boards.append(b)
        offs.append(15 - sum(x for x in b if x > 0))
        fts.append(rng.random() < 0.2)
    return np.array(boards), np.array(offs), np.array(fts)


def pack_mover_boards(boards, offs, fts, seed):
    """Place mover-frame boards into the absolute frame with a random colour to move."""
    rng = np.random.RandomState(seed)
    turn = np.where(rng.rand(len(boards)) < 0.5, 1, -1)
    absb = np.where(turn[:, None] == 1, boards, S.rotate_board(boards))
    offw = np.where(turn == 1, offs, 0)
    offb = np.where(turn == 1, 0, offs)
    lo, hi = S.pack_states(absb, offw, offb, turn, fts, fts)
    return lo, hi, turn, absb


def random_dice(n, seed, p_double=0.3):
    rng = np.random.RandomState(seed)
    d = rng.randint(1, 7, size=(n, 2)).astype(np.uint8)
    dbl = rng.rand(n) < p_double
    d[dbl, 1] = d[dbl, 0]
    return d


# ------------------------------------------------------------------------------------------
# checks
# ------------------------------------------------------------------------------------------
def check_golden_valid_moves(backend):
    """Narde.get_valid_moves golden lists (generated from the real reference)."""
    cases = load_golden("ref_valid_moves.json")
    boards = np.array([c["board"] for c in cases])
    lo, hi = S.pack_states(boards, [c["off_w"] for c in cases], [c["off_b"] for c in cases],
                           [c["player"] for c in cases], [c["first_w"] for c in cases], [c["first_b"] for c in cases])
    dice4 = np.zeros((len(cases), 4), np.uint8)
    for i, c in enumerate(cases):
        dice4[i, :len(c["roll"])] = c["roll"]
    moves, counts = backend.half_moves(lo, hi, dice4)
    for i, c in enumerate(cases):
        assert moves_from_buffer(moves[i], counts[i]) == [tuple(m) for m in c["moves"]], (i, c)
    return len(cases)


def check_golden_step_traces(backend):
    """NardeEnv.reset/step traces of the real reference, replayed in lock-step (all traces as one batch)."""
    traces = load_golden("ref_step_traces.json")
    n = len(traces)
    start = np.zeros((n, 24), np.int64)
    start[:, 23], start[:, 11] = 15, -15
    lo, hi = S.pack_states(start, 0, 0, [t["player0"] for t in traces], True, True)
    obs0 = backend.obs24(lo, hi)
    for i, t in enumerate(traces):
        assert obs0[i].tolist() == t["reset_obs"]
    T = max(len(t["steps"]) for t in traces)
    nsteps = 0
    for k in range(T):
        dice = np.ones((n, 2), np.uint8)
        codes = np.zeros((n, 2), np.int32)
        live = []
        for i, t in enumerate(traces):
            if k < len(t["steps"]):
                dice[i] = t["steps"][k]["dice"]
                codes[i] = t["steps"][k]["action"]
                live.append(i)
        before = (lo.copy(), hi.copy())
        obs, rew, done = backend.step_ref(lo, hi, dice, codes)
        u = S.unpack_states(lo, hi)
        for i in range(n):
            if i not in live:  # finished traces: the env is terminated and must be left untouched
                assert (lo[i] == before[0][i]).all() and (hi[i] == before[1][i]).all()
                continue
            s = traces[i]["steps"][k]
            assert obs[i].tolist() == s["obs"], (i, k)
            assert int(rew[i]) == s["reward"] and bool(done[i] & 1) == s["done"], (i, k)
            assert u["board"][i].tolist() == s["board"], (i, k)
            assert (int(u["off_w"][i]), int(u["off_b"][i]), bool(u["first_w"][i]), bool(u["first_b"][i]),
                    int(u["turn"][i])) == (s["off_w"], s["off_b"], s["first_w"], s["first_b"], s["player"]), (i, k)
            nsteps += 1
    return nsteps


def check_tier_n_kat(backend):
    """Afterstate sets composed from the reference's own primitives (tests/golden/tier_n_kat.json)."""
    cases = load_golden("tier_n_kat.json")
    boards = np.array([c["board_mover"] for c in cases])
    lo, hi = S.pack_states(boards, [c["off"] for c in cases], 0, 1, [c["first_turn"] for c in cases], False)
    dice = np.array([c["dice"] for c in cases], np.uint8)
    cap = 256
    acts, counts, ovf = backend.enumerate(lo, hi, dice, cap)
    for i, c in enumerate(cases):
        assert counts[i] == len(c["afterstates"]) and not ovf[i], (i, counts[i], len(c["afterstates"]))
        got = set()
        for a in acts[i, :counts[i]]:
            b = list(c["board_mover"])
            for f, t in act_moves(a):
                b[f] -= 1
                if t >= 0:
                    b[t] += 1
            got.add(tuple(b))
        assert got == set(tuple(a) for a in c["afterstates"]), i
    return len(cases)


def check_half_moves_vs_oracle(backend, lo, hi, dice4):
    u = S.unpack_states(lo, hi)
    moves, counts = backend.half_moves(lo, hi, dice4)
    og = O.OracleNarde()
    for i in range(lo.shape[0]):
        for k in range(24):
            og.g.board[k] = int(u["board"][i, k])
        og.g.first_turn_white, og.g.first_turn_black = int(u["first_w"][i]), int(u["first_b"][i])
        ref = og.get_valid_moves([int(x) for x in dice4[i] if x], int(u["turn"][i]))
        ref = [(f, -1 if t == 'off' else t) for f, t in ref]
        assert ref == moves_from_buffer(moves[i], counts[i]), (i, dice4[i], ref)


def check_enumerate_vs_oracle(backend, lo, hi, dice, cap=4096):
    """Exact equality of the canonical action lists (order, representative sequences, counts)."""
    u = S.unpack_states(lo, hi)
    acts, counts, ovf = backend.enumerate(lo, hi, dice, cap)
    total = 0
    for i in range(lo.shape[0]):
        pl = int(u["turn"][i])
        b = u["board"][i] if pl == 1 else S.rotate_board(u["board"][i])
        ft = u["first_w"][i] if pl == 1 else u["first_b"][i]
        moff = u["off_w"][i] if pl == 1 else u["off_b"][i]
        ref, nref = O.turn_enumerate(b, int(moff), int(dice[i, 0]), int(dice[i, 1]), bool(ft), cap=cap)
        assert nref == counts[i], (i, nref, counts[i], dice[i], b.tolist())
        assert bool(ovf[i]) == (nref > cap)
        k = min(nref, cap)
        got = [act_moves(a) for a in acts[i, :k]]
        assert got == [list(map(tuple, r["moves"])) for r in ref[:k]], (i, dice[i], b.tolist())
        total += nref
    return total


def check_step_ref_lockstep(backend, n, T, seed):
    """backend.step_ref vs the oracle's NardeEnv.step, identical dice and codes."""
    rng = random.Random(seed)
    lo, hi = backend.reset(n, env_base=1000, seed=seed, step=0)
    u = S.unpack_states(lo, hi)
    envs = [O.OracleEnv() for _ in range(n)]
    for i, e in enumerate(envs):
        pl = O.opening_player(seed, 1000 + i, 0)
        assert pl == u["turn"][i]
        e.reset([6, 1] if pl == 1 else [1, 6])
    finished = [False] * n
    for t in range(T):
        dice = np.array([[rng.randint(1, 6), rng.randint(1, 6)] for _ in range(n)], np.uint8)
        codes = np.zeros((n, 2), np.int32)
        for i, e in enumerate(envs):
            if i % 2 == 0:
                codes[i] = (rng.randrange(576), rng.randrange(576))
            else:
                og = O.OracleNarde()
                og.g = e.e.game
                v = og.get_valid_moves([int(dice[i, 0]), int(dice[i, 1])], e.e.current_player)
                if v:
                    m1, m2 = rng.choice(v), rng.choice(v)
                    codes[i] = (m1[0] * 24 + (0 if m1[1] == 'off' else m1[1]), m2[0] * 24 + (0 if m2[1] == 'off' else m2[1]))
        obs, rew, done = backend.step_ref(lo, hi, dice, codes, max_episode_steps=1000)
        u = S.unpack_states(lo, hi)
        for i, e in enumerate(envs):
            if finished[i]:
                assert done[i] & 1
                continue
            oobs, orew, odone = e.step(dice[i], codes[i])
            assert (oobs == obs[i]).all() and orew == rew[i] and odone == bool(done[i] & 1), (t, i)
            assert e.state_tuple() == (tuple(u["board"][i]), u["off_w"][i], u["off_b"][i], int(u["first_w"][i]),
                                       int(u["first_b"][i]), u["turn"][i]), (t, i)
            finished[i] = odone
    return sum(finished)


def check_step_full_lockstep(backend, n, T, seed, env_base=77, cap=32, flags=2, action_mode=None):
    """Fused full-rules step (Philox dice + Philox action + auto-reset) vs the oracle.
    action_mode: None = Philox-uniform choice inside the step; "index" = caller's int32 indices
    (negative / too large values are clamped); "fraction" = caller's u32 fractions (flag 32)."""
    rng = np.random.default_rng(seed ^ 0xAC71)
    lo, hi = backend.reset(n, env_base=env_base, seed=seed, step=0)
    envs = [O.OracleEnv() for _ in range(n)]
    for i, e in enumerate(envs):
        pl = O.opening_player(seed, env_base + i, 0)
        e.reset([6, 1] if pl == 1 else [1, 6])
    episodes = 0
    reward_mode = 1 if (flags & 1) else 0
    idle = [False] * n  # without auto-reset a finished env idles: done stays set, nothing else changes
    for t in range(T):
        given = None
        if action_mode == "index":
            given = rng.integers(-3, 40, n).astype(np.int32)
        elif action_mode == "fraction":
            given = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32).view(np.int32)
        out = backend.step_full(lo, hi, env_base=env_base, seed=seed, step=t + 1, cap=cap, action_idx=given,
                                flags=flags | (32 if action_mode == "fraction" else 0))
        u = S.unpack_states(lo, hi)
        exp_stats = np.zeros(9, np.int64)
        for i, e in enumerate(envs):
            if idle[i]:
                assert out["done"][i] & 1 and out["counts"][i] == 0 and out["reward"][i] == 0
                assert e.state_tuple()[:3] == (tuple(u["board"][i]), u["off_w"][i], u["off_b"][i])
                continue
            d1, d2, w = O.turn_dice(seed, env_base + i, t + 1)
            assert (d1, d2) == tuple(out["dice"][i])
            st = e.state_tuple()
            pl = st[5]
            b = np.array(st[0])
            mv = b if pl == 1 else S.rotate_board(b)
            acts, nact = O.turn_enumerate(mv, st[1] if pl == 1 else st[2], d1, d2, st[3] if pl == 1 else st[4])
            assert nact == out["counts"][i], (t, i, nact, out["counts"][i])
            idx = (w * nact) >> 32 if nact else 0
            if nact and action_mode == "index":
                idx = min(max(int(given[i]), 0), nact - 1)
            elif nact and action_mode == "fraction":
                idx = (int(given[i].view(np.uint32)) * nact) >> 32
            obs, rew, done, _ = e.full_step((d1, d2), idx, reward_mode)
            assert rew == out["reward"][i] and done == bool(out["done"][i] & 1), (t, i)
            exp_stats[5] += nact
            exp_stats[6] = max(exp_stats[6], nact)
            exp_stats[7] += nact > cap
            if nact and action_mode == "index":
                exp_stats[8] += int(given[i]) < 0 or int(given[i]) >= nact      # NARDE_STAT_CLAMPED_ACTIONS
            if nact:
                assert act_moves(out["chosen"][i]) == list(map(tuple, acts[idx]["moves"]))
                k = min(nact, cap)
                assert [act_moves(a) for a in out["actions"][i, :k]] == [list(map(tuple, a["moves"])) for a in acts[:k]]
            if done:
                episodes += 1
                exp_stats[0] += 1
                exp_stats[1 if pl == 1 else 2] += 1
                loser_off = e.e.game.borne_off_black if pl == 1 else e.e.game.borne_off_white
                exp_stats[3] += loser_off == 0
                exp_stats[4] += int(u["steps"][i]) if not (flags & 2) else 0
                if flags & 2:
                    pl2 = O.opening_player(seed, env_base + i, t + 1)
                    e.reset([6, 1] if pl2 == 1 else [1, 6])
                    obs = O.obs198(e.board, 0, 0, e.e.current_player)
                else:
                    idle[i] = True
            assert (obs == out["obs198"][i]).all(), (t, i)
            assert e.state_tuple() == (tuple(u["board"][i]), u["off_w"][i], u["off_b"][i], int(u["first_w"][i]),
                                       int(u["first_b"][i]), u["turn"][i]), (t, i)
        for k in (0, 1, 2, 3, 5, 6, 7, 8):
            assert exp_stats[k] == out["stats"][k], (t, k, exp_stats, out["stats"])
    return episodes


def check_obs198(backend, lo, hi):
    u = S.unpack_states(lo, hi)
    got = backend.obs198(lo, hi)
    for i in range(lo.shape[0]):
        ref = O.obs198(u["board"][i], int(u["off_w"][i]), int(u["off_b"][i]), int(u["turn"][i]))
        assert (ref == got[i]).all(), i


def check_turn_tree_vs_oracle(ops, lo, hi, dice):
    """The turn manager's tree of partial turns (gym_narde_b200/narde_game_manager.py:TurnTree, built from batched
    half-move / apply calls through `ops`) against the oracle's full-turn enumeration: the legal end-of-turn
    positions are the same SET, every canonical sequence of the oracle can be played half-move by half-move,
    and every half-move offered at the root starts some legal turn."""
    from gym_narde_b200.narde_game_manager import TurnTree

    u = S.unpack_states(lo, hi)
    total = 0
    for i in range(lo.shape[0]):
        pl = int(u["turn"][i])
        b = u["board"][i] if pl == 1 else S.rotate_board(u["board"][i])
        ft = bool(u["first_w"][i] if pl == 1 else u["first_b"][i])
        moff = int(u["off_w"][i] if pl == 1 else u["off_b"][i])
        d1, d2 = int(dice[i, 0]), int(dice[i, 1])
        ref, nref = O.turn_enumerate(b, moff, d1, d2, ft, cap=8192)
        tree = TurnTree(lo[i].tobytes(), hi[i].tobytes(), (d1, d2), ft, ops=ops)
        ends = set()
        for nd in tree.ends:
            v = S.unpack_states(np.frombuffer(nd.lo, np.uint8), np.frombuffer(nd.hi, np.uint8))
            eb = v["board"][0] if pl == 1 else S.rotate_board(v["board"][0])
            ends.add((tuple(int(x) for x in eb), int(v["off_w"][0] if pl == 1 else v["off_b"][0])))
        if nref == 0:
            assert tree.max_depth == 0 and tree.moves_from([tree.root]) == [], (i, dice[i], b.tolist())
            continue
        want = {(tuple(int(x) for x in r["after"]), int(r["after_off"])) for r in ref}
        assert ends == want, (i, dice[i], b.tolist(), len(ends), len(want))
        firsts = set()
        for r in ref:                                   # each canonical sequence is playable step by step
            nds = [tree.root]
            for f, t in r["moves"]:
                mv = (int(f), "off" if t in (-1, 255, "off") else int(t))
                assert mv in tree.moves_from(nds), (i, dice[i], b.tolist(), r["moves"], mv)
                nds = tree.children(nds, *mv)
            assert any(nd in tree.ends for nd in nds) and tree.moves_from(nds) == []
            firsts.add((int(r["moves"][0][0]), "off" if r["moves"][0][1] in (-1, 255, "off") else int(r["moves"][0][1])))
        assert firsts <= set(tree.moves_from([tree.root]))
        assert tree.launches <= 2 * (4 if d1 == d2 else 2)
        total += nref
    return total


def list_hash_np(actions, counts, cap):
    """sum_k act[k] * weight(k) mod 2^64 over the first min(count, cap) stored actions (o_selfplay_trace's checksum)."""
    w = O.list_weights(cap).view(np.uint64)
    k = np.minimum(counts, cap)
    mask = np.arange(cap)[None, :] < k[:, None]
    with np.errstate(over="ignore"):
        return (np.where(mask, actions.view(np.uint64), 0) * w[None, :]).sum(1, dtype=np.uint64).view(np.int64)


def check_trace_vs_backend(backend, n, T, seed, env_base=5, cap=16, flags=2, action_mode=None, max_episode_steps=0,
                           threads=4):
    """The oracle's bulk trace (o_selfplay_trace, what the full-size GPU parity test and bench.py compare against)
    versus step-by-step calls of the backend: every record field at every step, and the accumulated stats."""
    rng = np.random.default_rng(seed ^ 0x7A5E)
    words = None
    if action_mode == "index":
        words = rng.integers(-3, 40, (T, n)).astype(np.int32)
    elif action_mode == "fraction":
        words = rng.integers(0, 1 << 32, (T, n), dtype=np.uint64).astype(np.uint32).view(np.int32)
    tr = O.selfplay_trace(seed, env_base, n, T, step0=0, words=words, word_mode=0 if action_mode == "index" else 1,
                          cap=cap, reward_mode=flags & 1, autoreset=bool(flags & 2),
                          max_episode_steps=max_episode_steps, threads=threads)
    lo, hi = backend.reset(n, env_base=env_base, seed=seed, step=0)
    stats = np.zeros(8, np.int64)
    for t in range(T):
        out = backend.step_full(lo, hi, env_base=env_base, seed=seed, step=t + 1, cap=cap,
                                action_idx=None if words is None else words[t],
                                flags=flags | (32 if action_mode == "fraction" else 0), max_episode_steps=max_episode_steps)
        assert (lo == tr["lo"][t]).all() and (hi == tr["hi"][t]).all(), t
        assert (out["counts"] == tr["count"][t]).all() and (out["dice"] == tr["dice"][t]).all(), t
        assert (out["chosen"].view(np.int64) == tr["chosen"][t]).all(), t
        assert (out["reward"] == tr["reward"][t]).all(), t
        assert ((out["done"] & 1) | (out["trunc"] << 1) == tr["done"][t]).all(), t
        assert (list_hash_np(out["actions"], out["counts"], cap) == tr["hash"][t]).all(), t
        st = out["stats"]
        stats[:6] += st[:6]
        stats[6] = max(stats[6], st[6])
        stats[7] += st[7]
    assert (stats == tr["stats"]).all(), (stats, tr["stats"])
    return tr["turns"], int(tr["stats"][0])
