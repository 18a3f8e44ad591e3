"""Synthetic workloads of BASELINE.json's configs (generated on the device with the product's own kernels)."""
from __future__ import annotations


def config3_positions(device, n=1 << 20, seed=1234):
    """BASELINE config 3: `n` synthetic positions in three strata --
    A 40 % self-play states sampled at a uniformly random ply in [0, 90] of full-rules random games;
    B 30 % stratum-A states with the dice forced to doubles;
    C 30 % bear-off races: 15-k mover checkers over points 0..5, 15-k' opponent checkers over the mover-frame points
      12..17, first_turn False, uniform dice.
    Returns (lo, hi, dice, strata) with strata = {name: (begin, end)}."""
    import torch
    from .vec_env import VecNardeEnv

    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    nA, nB = int(0.4 * n), int(0.3 * n)
    nC = n - nA - nB
    env = VecNardeEnv(nA, seed=seed, max_actions=1, device=dev, write_actions=False)
    env.reset()
    target = torch.randint(0, 91, (nA,), device=dev, generator=g)
    lo_a, hi_a = env.lo.clone(), env.hi.clone()
    for t in range(1, 91):
        env.step()
        m = (target == t)[:, None]
        lo_a = torch.where(m, env.lo, lo_a)
        hi_a = torch.where(m, env.hi, hi_a)
    dice_a = torch.randint(1, 7, (nA, 2), device=dev, generator=g).to(torch.uint8)
    pick = torch.randint(0, nA, (nB,), device=dev, generator=g)
    d = torch.randint(1, 7, (nB, 1), device=dev, generator=g).to(torch.uint8)
    lo_b, hi_b, dice_b = lo_a[pick], hi_a[pick], d.expand(nB, 2).contiguous()
    # stratum C, absolute frame with WHITE (= mover) to move
    k_m = torch.randint(0, 15, (nC,), device=dev, generator=g)
    k_o = torch.randint(0, 15, (nC,), device=dev, generator=g)
    board = torch.zeros((nC, 24), dtype=torch.int32, device=dev)
    slots = torch.arange(15, device=dev)[None, :]
    pm = torch.randint(0, 6, (nC, 15), device=dev, generator=g)
    po = torch.randint(12, 18, (nC, 15), device=dev, generator=g)
    board.scatter_add_(1, pm, (slots < (15 - k_m)[:, None]).to(torch.int32))
    board.scatter_add_(1, po, -(slots < (15 - k_o)[:, None]).to(torch.int32))
    planes = torch.zeros((nC, 32), dtype=torch.uint8, device=dev)
    planes[:, :24] = board.to(torch.int8).view(torch.uint8)
    planes[:, 24] = k_m.to(torch.uint8)
    planes[:, 25] = k_o.to(torch.uint8)
    planes[:, 26] = 1                                              # WHITE to move; flags 0 (first_turn False)
    lo_c, hi_c = planes[:, :16].contiguous(), planes[:, 16:].contiguous()
    dice_c = torch.randint(1, 7, (nC, 2), device=dev, generator=g).to(torch.uint8)
    lo = torch.cat([lo_a, lo_b, lo_c]).contiguous()
    hi = torch.cat([hi_a, hi_b, hi_c]).contiguous()
    dice = torch.cat([dice_a, dice_b, dice_c]).contiguous()
    return lo, hi, dice, {"A_selfplay": (0, nA), "B_doubles": (nA, nA + nB), "C_bearoff": (nA + nB, n)}
