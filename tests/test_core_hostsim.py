"""CPU: the CUDA path's per-environment core (narde_core.cuh / narde_env.cuh compiled by g++ into the
test-only host harness) against the oracle and the golden fixtures.  Same checks as the -m gpu
parity tests, so the rules arithmetic is proven before it ever reaches the GPU."""
import numpy as np

import parity as P


def test_golden_valid_moves(hostsim):
    assert P.check_golden_valid_moves(hostsim) > 1000


def test_golden_step_traces(hostsim):
    assert P.check_golden_step_traces(hostsim) > 2000


def test_tier_n_kat(hostsim):
    assert P.check_tier_n_kat(hostsim) > 300


def test_half_moves_selfplay_and_synthetic(hostsim):
    lo, hi = P.pack_corpus(P.selfplay_corpus(25, 1))
    n = lo.shape[0]
    rng = np.random.RandomState(0)
    dice4 = np.zeros((n, 4), np.uint8)
    k = rng.randint(0, 3, size=n)
    d = rng.randint(1, 7, size=(n, 4))
    dice4[:, 0] = d[:, 0]
    dice4[k >= 1, 1] = d[k >= 1, 1]
    four = k == 2
    dice4[four] = d[four, :1]
    P.check_half_moves_vs_oracle(hostsim, lo, hi, dice4)
    b, off, ft = P.synthetic_boards(3000, 2)
    lo, hi, _, _ = P.pack_mover_boards(b, off, ft, 3)
    dice4 = np.zeros((3000, 4), np.uint8)
    dice4[:, :2] = P.random_dice(3000, 4)
    P.check_half_moves_vs_oracle(hostsim, lo, hi, dice4)


def test_enumerate_selfplay(hostsim):
    lo, hi = P.pack_corpus(P.selfplay_corpus(25, 5))
    assert P.check_enumerate_vs_oracle(hostsim, lo, hi, P.random_dice(lo.shape[0], 6, 0.3)) > 0


def test_enumerate_synthetic_block_rule_heavy(hostsim):
    b, off, ft = P.synthetic_boards(5000, 7)
    lo, hi, _, _ = P.pack_mover_boards(b, off, ft, 8)
    assert P.check_enumerate_vs_oracle(hostsim, lo, hi, P.random_dice(5000, 9, 0.5)) > 0


def test_enumerate_overflow_reports_true_count(hostsim):
    b, off, ft = P.synthetic_boards(500, 17)
    lo, hi, _, _ = P.pack_mover_boards(b, off, ft, 18)
    P.check_enumerate_vs_oracle(hostsim, lo, hi, P.random_dice(500, 19, 0.7), cap=8)


def test_step_ref_lockstep(hostsim):
    assert P.check_step_ref_lockstep(hostsim, 64, 500, 42) > 0


def test_step_full_lockstep_autoreset(hostsim):
    assert P.check_step_full_lockstep(hostsim, 96, 260, 0x5EED) > 50


def test_step_full_lockstep_mover_reward_no_autoreset(hostsim):
    P.check_step_full_lockstep(hostsim, 48, 220, 99, cap=4, flags=1)


def test_obs198(hostsim):
    lo, hi = P.pack_corpus(P.selfplay_corpus(10, 21))
    P.check_obs198(hostsim, lo, hi)
