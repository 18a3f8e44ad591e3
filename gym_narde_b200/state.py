"""Host-side packing of the HBM state record (include/narde_b200.h "State record").

Two SoA planes of 16-byte lanes per environment:
  lo[i] : int8 points 0..15 (absolute / White frame, +white -black)
  hi[i] : int8 points 16..23 | off_white | off_black | turn (+1/-1) | flags | u16 steps | u16 rsvd
which replaces the reference's Narde fields (gym_narde/envs/narde.py:21-29) and
NardeEnv.current_player (gym_narde/envs/narde_env.py:14).
"""
from __future__ import annotations

import numpy as np

FLAG_FIRST_W = 1
FLAG_FIRST_B = 2
FLAG_DONE = 4

OFF = 255          # half-move destination byte meaning 'off'
EMPTY_SLOT = 0xFFFF


def pack_states(boards, off_w=0, off_b=0, turn=1, first_w=False, first_b=False, done=False, steps=0):
    """boards: [n,24] ints in [-15,15].  Scalars broadcast.  Returns (lo, hi) uint8 [n,16]."""
    boards = np.asarray(boards, dtype=np.int64).reshape(-1, 24)
    n = boards.shape[0]
    lo = np.zeros((n, 16), dtype=np.uint8)
    hi = np.zeros((n, 16), dtype=np.uint8)
    b8 = boards.astype(np.int8).view(np.uint8)
    lo[:, :] = b8[:, :16]
    hi[:, :8] = b8[:, 16:]
    hi[:, 8] = np.broadcast_to(np.asarray(off_w, dtype=np.int64), (n,)).astype(np.uint8)
    hi[:, 9] = np.broadcast_to(np.asarray(off_b, dtype=np.int64), (n,)).astype(np.uint8)
    hi[:, 10] = np.broadcast_to(np.asarray(turn, dtype=np.int64), (n,)).astype(np.int8).view(np.uint8)
    flags = (np.broadcast_to(np.asarray(first_w, dtype=bool), (n,)).astype(np.uint8) * FLAG_FIRST_W
             | np.broadcast_to(np.asarray(first_b, dtype=bool), (n,)).astype(np.uint8) * FLAG_FIRST_B
             | np.broadcast_to(np.asarray(done, dtype=bool), (n,)).astype(np.uint8) * FLAG_DONE)
    hi[:, 11] = flags
    st = np.broadcast_to(np.asarray(steps, dtype=np.int64), (n,)).astype(np.uint16)
    hi[:, 12] = (st & 0xFF).astype(np.uint8)
    hi[:, 13] = (st >> 8).astype(np.uint8)
    return lo, hi


def unpack_states(lo, hi):
    """Inverse of pack_states.  lo, hi: uint8 [n,16] (numpy).  Returns a dict of arrays."""
    lo = np.asarray(lo, dtype=np.uint8).reshape(-1, 16)
    hi = np.asarray(hi, dtype=np.uint8).reshape(-1, 16)
    board = np.concatenate([lo, hi[:, :8]], axis=1).view(np.int8).astype(np.int32)
    flags = hi[:, 11]
    return {
        "board": board,
        "off_w": hi[:, 8].astype(np.int32),
        "off_b": hi[:, 9].astype(np.int32),
        "turn": hi[:, 10].view(np.int8).astype(np.int32),
        "first_w": (flags & FLAG_FIRST_W) != 0,
        "first_b": (flags & FLAG_FIRST_B) != 0,
        "done": (flags & FLAG_DONE) != 0,
        "steps": hi[:, 12].astype(np.int32) | (hi[:, 13].astype(np.int32) << 8),
    }


def decode_action(a):
    """u64 turn action -> list of (from, to) with to == 'off' for bear-off (mover frame)."""
    a = int(a)
    out = []
    for k in range(4):
        h = (a >> (16 * k)) & 0xFFFF
        if h == EMPTY_SLOT:
            break
        frm, to = h & 0xFF, h >> 8
        out.append((frm, 'off' if to == OFF else to))
    return out


def encode_action(moves):
    """list of (from, to|'off') (<= 4) -> u64 turn action."""
    a = 0xFFFFFFFFFFFFFFFF
    for k, (frm, to) in enumerate(moves):
        t = OFF if to == 'off' else int(to)
        a &= ~(0xFFFF << (16 * k))
        a |= ((int(frm) & 0xFF) | (t << 8)) << (16 * k)
    return a


def rotate_board(board):
    """Mover-frame view for Black (gym_narde/envs/narde.py:16-17)."""
    board = np.asarray(board)
    return np.concatenate((-board[..., 12:], -board[..., :12]), axis=-1)
