"""Short GPU programs for `ncu --set full` captures (kept short: ncu replays every captured launch ~40 times).

    python tools/ncu_target.py step  [envs] [burn_in]   steady-state fused steps (k_step_full_v2<128,true> + k_step_deferred)
    python tools/ncu_target.py actor [envs] [burn_in]   greedy turns (enumerate, afterstates, k_mlp, argmax, step)
    python tools/ncu_target.py c3                       BASELINE config 3 enumeration over 1 M positions

With `step`, burn-in steps launch 2 kernels each: `-k regex:k_step -s <2*burn_in> -c 4` captures two steady-state turns."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_narde_b200 import VecNardeEnv

what = sys.argv[1] if len(sys.argv) > 1 else "step"
E = int(sys.argv[2]) if len(sys.argv) > 2 else (131072 if what == "step" else 65536)   # what: step | actor | actor_graph | c3
B = int(sys.argv[3]) if len(sys.argv) > 3 else 300
torch.cuda.set_device(0)
if what == "c3":
    from gym_narde_b200 import _cabi
    from gym_narde_b200.workloads import config3_positions
    dev = torch.device("cuda:0")
    n = 1 << 20
    lo, hi, dice, _ = config3_positions(dev, n=n, seed=1234)
    actions = torch.zeros((n, 64), dtype=torch.int64, device=dev)
    counts = torch.zeros(n, dtype=torch.int32, device=dev)
    ovf = torch.zeros(n, dtype=torch.uint8, device=dev)
    ws = torch.zeros(_cabi.workspace_ints(n), dtype=torch.int32, device=dev)
    for _ in range(3):
        _cabi.enumerate_actions_fast(lo, hi, dice, actions, counts, ovf, ws)
    torch.cuda.synchronize()
    print("c3 done; deferred", int(ws[0].item()))
    sys.exit(0)
if what == "mlp":      # the tcgen05 scorer alone: 350 208 packed states of a self-play batch, three launches
    import torch.nn as nn
    from gym_narde_b200 import AfterstateMLP
    env = VecNardeEnv(43776, seed=3)
    env.reset()
    for _ in range(40):
        env.step()
    lo, hi = env.lo.repeat(8, 1).contiguous(), env.hi.repeat(8, 1).contiguous()
    torch.manual_seed(0)
    fn = nn.Sequential(nn.Linear(198, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU()).cuda()
    mlp = AfterstateMLP.from_module(fn, nn.Linear(256, 576).cuda())
    out = torch.zeros(lo.shape[0], dtype=torch.float32, device="cuda")
    for _ in range(3):
        mlp.score_states(lo, hi, out=out)
    torch.cuda.synchronize()
    print("mlp done", lo.shape[0], float(out.mean()))
    sys.exit(0)
env = VecNardeEnv(E, seed=0x5EED, max_actions=64, graph=False)
env.reset()
for _ in range(B):
    env.step()
torch.cuda.synchronize()
if what == "step":
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        flush.fill_(1)
        env.step()
    torch.cuda.synchronize()
    print("step done", env.episode_stats())
else:
    import torch.nn as nn
    from gym_narde_b200 import AfterstateMLP, AfterstateActor
    torch.manual_seed(0)
    fn = nn.Sequential(nn.Linear(198, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU()).cuda()
    head = nn.Linear(256, 576).cuda()
    actor = AfterstateActor(env, AfterstateMLP.from_module(fn, head))
    for _ in range(4):
        actor.step_graph() if what == "actor_graph" else actor.step()
    torch.cuda.synchronize()
    print("actor done", env.episode_stats())
