"""GPU debug: HostPipeline cost breakdown."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_narde_b200 import VecNardeEnv, _cabi
from gym_narde_b200 import vec_env as V

E, D = 131072, 8
def tm(fn, reps=20):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / D

def fresh():
    env = VecNardeEnv(E, seed=0x5EED, max_actions=64)
    env.reset()
    for _ in range(300):
        env.step()
    torch.cuda.synchronize()
    return env

env = fresh()
pipe = env.host_pipeline(depth=D, fraction=True)
pipe.actions.copy_(torch.randint(-(1 << 31), (1 << 31) - 1, (D, E), dtype=torch.int64).to(torch.int32))
print("pipeline (H2D + step + D2H)      %.4f ms/turn" % tm(pipe.run))
# variants by monkeypatching the copies away
class NoCopy(V.HostPipeline):
    mode = "none"
    def _capture(self):
        env, t, Dd = self.env, self.env.torch, self.depth
        flags = 2 | 32
        env._step_dev.fill_(env.step_count)
        t.cuda.synchronize(env.device)
        g = t.cuda.CUDAGraph()
        with t.cuda.graph(g):
            main = t.cuda.current_stream(env.device)
            self._s_in.wait_stream(main); self._s_out.wait_stream(main)
            ev_in = [t.cuda.Event() for _ in range(Dd)]; ev_c = [t.cuda.Event() for _ in range(Dd)]; ev_out = [t.cuda.Event() for _ in range(Dd)]
            for k in range(Dd):
                b = k & 1
                if self.mode in ("h2d", "both"):
                    with t.cuda.stream(self._s_in):
                        if k >= 2: self._s_in.wait_event(ev_c[k - 2])
                        self._d_act[b].copy_(self.actions[k], non_blocking=True); ev_in[k].record(self._s_in)
                    main.wait_event(ev_in[k])
                if self.mode in ("d2h", "both") and k >= 2: main.wait_event(ev_out[k - 2])
                _cabi.advance_counter(env._step_dev)
                env._launch_full(self._d_act[b], None, flags)
                if self.mode in ("d2h", "both", "stage"):
                    self._d_rew[b].copy_(env.reward, non_blocking=True)
                    t.bitwise_or(env.done, env.trunc << 1, out=self._d_done[b])
                ev_c[k].record(main)
                if self.mode in ("d2h", "both"):
                    with t.cuda.stream(self._s_out):
                        self._s_out.wait_event(ev_c[k])
                        self.reward[k].copy_(self._d_rew[b], non_blocking=True); self.done[k].copy_(self._d_done[b], non_blocking=True)
                        ev_out[k].record(self._s_out)
            main.wait_stream(self._s_in); main.wait_stream(self._s_out)
        return g
for mode in ("none", "stage", "h2d", "d2h", "both"):
    env = fresh()
    p = NoCopy(env, D, True); p.mode = mode
    p.actions.copy_(pipe.actions)
    print("mode %-6s                      %.4f ms/turn" % (mode, tm(p.run)))
env = fresh()
print("plain graph step                 %.4f ms/turn" % (tm(lambda: [env.step() for _ in range(D)])))
