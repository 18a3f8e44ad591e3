"""Afterstate-scoring MLP (BASELINE config 5): host side of narde_mlp_forward.

Architecture = the reference's DecomposedDQN.forward(x) with state_size 198
(train_deepq_pytorch.py:184-236): Linear(198,256)-ReLU-Linear(256,256)-ReLU-Linear(256,576).
The CUDA kernel (csrc/narde_mlp.cu) runs bf16 operands / fp32 accumulation on tcgen05 tensor cores;
weights are re-packed once on the host into the shared-memory operand layout it streams.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi

IN, HID, OUT, K1, KC = 198, 256, 576, 208, 32


def _pack_block(w_blk, two_sm=False):
    """w_blk: [nb, K] float32 (K a multiple of 16) -> bf16 stages of 32 K (the last may be 16) in the K-major
    interleave layout: offset(n, k) = (k//8)*(nb*16) + (n//8)*128 + (n%8)*16 + (k%8)*2 within a stage.
    two_sm: every stage is stored as [rows 0..nb/2) | rows nb/2..nb)], each half in that layout on its own
    (the B operand halves of a cta_group::2 MMA, one per CTA of the pair)."""
    import torch

    nb, k = w_blk.shape
    stages = []
    for k0 in range(0, k, KC):
        klen = min(KC, k - k0)
        for rows in ((slice(0, nb // 2), slice(nb // 2, nb)) if two_sm else (slice(0, nb),)):
            s = w_blk[rows, k0:k0 + klen].to(torch.bfloat16)             # [n, klen]
            n = s.shape[0]
            s = s.reshape(n // 8, 8, klen // 8, 8).permute(2, 0, 1, 3)   # [kchunk, rowgroup, row, k]
            stages.append(s.contiguous().view(torch.int16).reshape(-1))
    return torch.cat(stages)


def pack_weights(w1, b1, w2, b2, w3, b3, two_sm=False):
    """torch Linear weights ([out, in]) and biases -> (wpack int16 tensor, bias float32 [1088])."""
    import torch

    assert tuple(w1.shape) == (HID, IN) and tuple(w2.shape) == (HID, HID) and tuple(w3.shape) == (OUT, HID)
    w1p = torch.zeros((HID, K1), dtype=torch.float32, device=w1.device)
    w1p[:, :IN] = w1.float()
    parts = [_pack_block(w1p, two_sm), _pack_block(w2.float(), two_sm)]
    for n0 in range(0, OUT, 256):
        parts.append(_pack_block(w3.float()[n0:min(n0 + 256, OUT)], two_sm))
    wpack = torch.cat(parts).contiguous()
    bias = torch.cat([b1.float(), b2.float(), b3.float()]).contiguous()
    return wpack, bias


class AfterstateMLP:
    """DecomposedDQN.forward(x) with state_size 198 on the tcgen05 tensor cores.

    forward(x)            x float32 [K,198]            -> q float32 [K,576]   (move1 Q-values)
    score(x)              x float32 [K,198]            -> max_a q[:, a]  float32 [K]
    forward_states(lo,hi) packed states [K,16] uint8 x2 -> q   (Box(198) encoded inside the kernel)
    score_states(lo,hi)   packed states                -> max_a q[:, a]       (the afterstate score)"""

    def __init__(self, w1, b1, w2, b2, w3, b3):
        torch = _cabi.require_cuda()
        lib = _cabi.load()
        self.torch, self.lib = torch, lib
        self.wpack, self.bias = pack_weights(w1.cuda(), b1.cuda(), w2.cuda(), b2.cuda(), w3.cuda(), b3.cuda())
        self.wpack2, _ = pack_weights(w1.cuda(), b1.cuda(), w2.cuda(), b2.cuda(), w3.cuda(), b3.cuda(), two_sm=True)

    @classmethod
    def from_module(cls, feature_network, move1_head):
        """feature_network = nn.Sequential(Linear, ReLU, Linear, ReLU), move1_head = Linear (reference names)."""
        l1, l2 = feature_network[0], feature_network[2]
        return cls(l1.weight.data, l1.bias.data, l2.weight.data, l2.bias.data, move1_head.weight.data, move1_head.bias.data)

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def _check_x(self, x):
        t = self.torch
        if not (x.is_cuda and x.dtype == t.float32 and x.is_contiguous() and x.dim() == 2 and x.shape[1] == IN):
            raise _cabi.NardeCudaError("x must be a contiguous CUDA float32 [K,198] tensor")

    def _check_states(self, lo, hi):
        t = self.torch
        for a in (lo, hi):
            if not (a.is_cuda and a.dtype == t.uint8 and a.is_contiguous() and a.dim() == 2 and a.shape[1] == 16):
                raise _cabi.NardeCudaError("states must be contiguous CUDA uint8 [K,16] planes")
        if lo.shape[0] != hi.shape[0]:
            raise _cabi.NardeCudaError("lo/hi planes differ in length")

    def _run(self, fn, args, what):
        rc = fn(*args, self._stream())
        if rc != 0:
            raise _cabi.NardeCudaError("%s failed: %d" % (what, rc))

    def forward(self, x, out=None):
        t = self.torch
        self._check_x(x)
        k = x.shape[0]
        if out is None:
            out = t.empty((k, OUT), dtype=t.float32, device=x.device)
        self._run(self.lib.narde_mlp_forward, (C.c_void_p(x.data_ptr()), k, C.c_void_p(self.wpack.data_ptr()),
                                               C.c_void_p(self.bias.data_ptr()), C.c_void_p(out.data_ptr())), "narde_mlp_forward")
        return out

    def score(self, x, out=None):
        t = self.torch
        self._check_x(x)
        k = x.shape[0]
        if out is None:
            out = t.empty(k, dtype=t.float32, device=x.device)
        self._run(self.lib.narde_mlp_score, (C.c_void_p(x.data_ptr()), k, C.c_void_p(self.wpack.data_ptr()),
                                             C.c_void_p(self.bias.data_ptr()), C.c_void_p(out.data_ptr())), "narde_mlp_score")
        return out

    def forward_states(self, lo, hi, out=None):
        t = self.torch
        self._check_states(lo, hi)
        k = lo.shape[0]
        if out is None:
            out = t.empty((k, OUT), dtype=t.float32, device=lo.device)
        self._run(self.lib.narde_mlp_forward_states,
                  (C.c_void_p(lo.data_ptr()), C.c_void_p(hi.data_ptr()), k, C.c_void_p(self.wpack.data_ptr()),
                   C.c_void_p(self.bias.data_ptr()), C.c_void_p(out.data_ptr())), "narde_mlp_forward_states")
        return out

    def score_states(self, lo, hi, out=None, rows_dev=None):
        """rows_dev: optional int64 CUDA tensor [1]; only min(rows_dev, len(lo)) rows are scored (no host sync)."""
        t = self.torch
        self._check_states(lo, hi)
        k = lo.shape[0]
        if out is None:
            out = t.empty(k, dtype=t.float32, device=lo.device)
        self._run(self.lib.narde_mlp_score_states,
                  (C.c_void_p(lo.data_ptr()), C.c_void_p(hi.data_ptr()), k,
                   C.c_void_p(rows_dev.data_ptr() if rows_dev is not None else None), C.c_void_p(self.wpack.data_ptr()),
                   C.c_void_p(self.bias.data_ptr()), C.c_void_p(out.data_ptr())), "narde_mlp_score_states")
        return out

    def score_states_2sm(self, lo, hi, out=None, rows_dev=None):
        """score_states through the cta_group::2 kernel (clusters of two CTAs, M = 256 per tcgen05.mma)."""
        t = self.torch
        self._check_states(lo, hi)
        k = lo.shape[0]
        if out is None:
            out = t.empty(k, dtype=t.float32, device=lo.device)
        fn = self.lib.narde_debug_mlp_score_states_2sm
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        fn.restype = C.c_int
        self._run(fn, (C.c_void_p(lo.data_ptr()), C.c_void_p(hi.data_ptr()), k,
                       C.c_void_p(rows_dev.data_ptr() if rows_dev is not None else None), C.c_void_p(self.wpack2.data_ptr()),
                       C.c_void_p(self.bias.data_ptr()), C.c_void_p(out.data_ptr())), "narde_debug_mlp_score_states_2sm")
        return out

    __call__ = forward
