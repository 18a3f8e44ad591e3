"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck): every kernel of the path once or more."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from gym_narde_b200 import VecNardeEnv, AfterstateMLP, AfterstateActor
from gym_narde_b200.trajectory import TrajectoryRing, action_codes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for N in (n, 20000):                      # both CTA tiles of the fused step (32-env and 128-env)
    env = VecNardeEnv(N, seed=3, max_actions=32, graph=False)
    env.reset()
    ring = TrajectoryRing(env, 4)
    for _ in range(steps if N == n else 6):
        ring.step()
    a, c, o = env.get_valid_actions()
    action_codes(env)
    idx = torch.zeros(N, dtype=torch.int32, device="cuda")
    env.step(idx, dice=env.roll())
torch.manual_seed(0)
fn = nn.Sequential(nn.Linear(198, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU()).cuda()
head = nn.Linear(256, 576).cuda()
mlp = AfterstateMLP.from_module(fn, head)
env = VecNardeEnv(n, seed=4, max_actions=32, graph=False)
env.reset()
actor = AfterstateActor(env, mlp)
for _ in range(3):
    actor.step()
x = env.observe()[:300].contiguous()
q = mlp(x); s = mlp.score(x); qs = mlp.forward_states(env.lo[:300].contiguous(), env.hi[:300].contiguous())
ref = VecNardeEnv(512, seed=5, rules="reference")
ref.reset()
ref.step(torch.randint(0, 576, (512, 2), dtype=torch.int32, device="cuda"))
ref.get_valid_moves(torch.tensor([[3, 5, 0, 0]] * 512, dtype=torch.uint8, device="cuda"))
torch.cuda.synchronize()
print("sanitize workload ok", float(q.abs().max()), env.episode_stats())
