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


COMPACT_RECORD_BYTES = 20   # NARDE_COMPACT_RECORD_BYTES (include/narde_b200.h)


def unpack_compact(rec):
    """Decode the compact host records VecNardeEnv.step_host(obs="compact") receives ([n,20] uint8; layout in
    include/narde_b200.h) -> (lo, hi, result): the two [n,16] uint8 state planes exactly as the device holds them
    (feed them to expand_obs198 / unpack_states) and the turn's result byte (bit 0 terminated, bit 1 truncated,
    bits 2-3 the reward 0/1/2).  Numpy on the host: a decoder of the wire format, nothing in the product calls it."""
    rec = np.ascontiguousarray(np.asarray(rec, dtype=np.uint8).reshape(-1, COMPACT_RECORD_BYTES))
    n = rec.shape[0]
    x = rec[:, :16].copy().view("<u8")                     # [n,2]
    x0, x1 = x[:, 0], x[:, 1]
    m60 = np.uint64((1 << 60) - 1)
    a = x0 & m60
    b = ((x0 >> np.uint64(60)) | (x1 << np.uint64(4))) & m60
    pts = np.empty((n, 24), np.int16)
    for p in range(12):
        pts[:, p] = ((a >> np.uint64(5 * p)) & np.uint64(31)).astype(np.int16)
        pts[:, p + 12] = ((b >> np.uint64(5 * p)) & np.uint64(31)).astype(np.int16)
    pts = ((pts ^ 16) - 16).astype(np.int8)                # 5-bit two's complement
    w4 = rec[:, 16:20].copy().view("<u4")[:, 0]
    lo = np.zeros((n, 16), np.uint8)
    hi = np.zeros((n, 16), np.uint8)
    lo[:, :] = pts[:, :16].view(np.uint8)
    hi[:, :8] = pts[:, 16:].view(np.uint8)
    hi[:, 8] = ((x1 >> np.uint64(56)) & np.uint64(15)).astype(np.uint8)
    hi[:, 9] = ((x1 >> np.uint64(60)) & np.uint64(15)).astype(np.uint8)
    hi[:, 10] = np.where(w4 & 1, 1, -1).astype(np.int8).view(np.uint8)
    hi[:, 11] = ((w4 >> 1) & 7).astype(np.uint8)
    steps = (w4 >> 8) & 0xFFFF
    hi[:, 12] = (steps & 0xFF).astype(np.uint8)
    hi[:, 13] = (steps >> 8).astype(np.uint8)
    return lo, hi, ((w4 >> 4) & 15).astype(np.uint8)


_OBS_LUT = None


def expand_obs198(lo, hi, out=None):
    """Decode packed state records into Box(198) rows (README.md:44-102): the observation a host-side consumer of
    VecNardeEnv.step_host(obs="packed") reads, bit-equal to what narde_obs198 / the fused step write on the device.

    CUDA tensors [n,16] uint8 -> the narde_obs198 kernel (returns a float32 CUDA tensor [n,198]);
    numpy arrays / host tensors -> a table lookup on the host (returns float32 numpy [n,198]): this is a decoder of
    the wire format, not an alternative to the kernels (nothing in the product calls it)."""
    try:
        import torch
        is_t = isinstance(lo, torch.Tensor)
    except ImportError:
        is_t = False
    if is_t and lo.is_cuda:
        from . import _cabi
        if out is None:
            out = torch.empty((lo.shape[0], 198), dtype=torch.float32, device=lo.device)
        _cabi.obs198(lo, hi, out)
        return out
    global _OBS_LUT
    if _OBS_LUT is None:
        # per signed point count v in [-15, 15]: WHITE's four features then BLACK's ([n>=1, n>=2, n>=3, (n-3)/2])
        lut = np.zeros((256, 8), np.float32)
        for v in range(-15, 16):
            for c, n in ((0, max(v, 0)), (1, max(-v, 0))):
                lut[v & 0xFF, 4 * c:4 * c + 4] = (n >= 1, n >= 2, n >= 3, (n - 3) / 2.0 if n > 3 else 0.0)
        _OBS_LUT = (lut, np.array([np.float32(k / 15.0) for k in range(256)], np.float32))
    lut, off15 = _OBS_LUT
    lo = np.asarray(lo, dtype=np.uint8).reshape(-1, 16)
    hi = np.asarray(hi, dtype=np.uint8).reshape(-1, 16)
    n = lo.shape[0]
    if out is None:
        out = np.zeros((n, 198), np.float32)
    f = lut[np.concatenate([lo, hi[:, :8]], axis=1)]            # [n, 24, 8]
    out[:, 0:96] = f[:, :, :4].reshape(n, 96)
    out[:, 96] = 0.0                                             # bar: no hitting in Narde (narde.py:71)
    out[:, 97] = off15[hi[:, 8]]
    out[:, 98:194] = f[:, :, 4:].reshape(n, 96)
    out[:, 194] = 0.0
    out[:, 195] = off15[hi[:, 9]]
    white = hi[:, 10].view(np.int8) == 1
    out[:, 196] = white
    out[:, 197] = ~white
    return out
