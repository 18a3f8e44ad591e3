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

IN, HID, OUT, KPAD, KC = 198, 256, 576, 256, 64


def _pack_block(w_blk):
    """w_blk: [nb, 256] float32 (K padded) -> bytes of 4 K-chunk stages in K-major interleave layout:
    offset(n, k) = (k//8)*(nb*16) + (n//8)*128 + (n%8)*16 + (k%8)*2 within a stage of nb x 64."""
    import torch

    nb = w_blk.shape[0]
    stages = []
    for c in range(KPAD // KC):
        s = w_blk[:, c * KC:(c + 1) * KC].to(torch.bfloat16)          # [nb, 64]
        s = s.reshape(nb // 8, 8, KC // 8, 8).permute(2, 0, 1, 3)      # [kchunk, rowgroup, row, k]
        stages.append(s.contiguous().view(torch.int16).reshape(-1))
    return torch.cat(stages)


def pack_weights(w1, b1, w2, b2, w3, b3):
    """torch Linear weights ([out, in]) and biases -> (wpack int16 tensor, bias float32 [1088])."""
    import torch

    assert tuple(w1.shape) == (HID, IN) and tuple(w2.shape) == (HID, HID) and tuple(w3.shape) == (OUT, HID)
    w1p = torch.zeros((HID, KPAD), dtype=torch.float32, device=w1.device)
    w1p[:, :IN] = w1.float()
    parts = [_pack_block(w1p), _pack_block(w2.float())]
    for n0 in range(0, OUT, 256):
        parts.append(_pack_block(w3.float()[n0:min(n0 + 256, OUT)]))
    wpack = torch.cat(parts).contiguous()
    bias = torch.cat([b1.float(), b2.float(), b3.float()]).contiguous()
    return wpack, bias


class AfterstateMLP:
    """q = forward(x): x float32 [K,198] on the GPU -> float32 [K,576] (move1 Q-values)."""

    def __init__(self, w1, b1, w2, b2, w3, b3):
        torch = _cabi.require_cuda()
        lib = _cabi.load()
        lib.narde_mlp_forward.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.narde_mlp_forward.restype = C.c_int
        self.torch, self.lib = torch, lib
        self.wpack, self.bias = pack_weights(w1.cuda(), b1.cuda(), w2.cuda(), b2.cuda(), w3.cuda(), b3.cuda())

    @classmethod
    def from_module(cls, feature_network, move1_head):
        """feature_network = nn.Sequential(Linear, ReLU, Linear, ReLU), move1_head = Linear (reference names)."""
        l1, l2 = feature_network[0], feature_network[2]
        return cls(l1.weight.data, l1.bias.data, l2.weight.data, l2.bias.data, move1_head.weight.data, move1_head.bias.data)

    def forward(self, x, out=None):
        t = self.torch
        if not (x.is_cuda and x.dtype == t.float32 and x.is_contiguous() and x.shape[1] == IN):
            raise _cabi.NardeCudaError("x must be a contiguous CUDA float32 [K,198] tensor")
        k = x.shape[0]
        if out is None:
            out = t.empty((k, OUT), dtype=t.float32, device=x.device)
        rc = self.lib.narde_mlp_forward(C.c_void_p(x.data_ptr()), k, C.c_void_p(self.wpack.data_ptr()),
                                        C.c_void_p(self.bias.data_ptr()), C.c_void_p(out.data_ptr()),
                                        C.c_void_p(t.cuda.current_stream().cuda_stream))
        if rc != 0:
            raise _cabi.NardeCudaError("narde_mlp_forward failed: %d" % rc)
        return out

    __call__ = forward
