#!/usr/bin/env python
"""bench.py -- legal-move-enumerated Narde env steps/s on N B200s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port on host cores)

Workload (config.workload): BASELINE config 4's per-GPU shard -- 131072 lock-step environments per
GPU (1M environments at 8 GPUs), full-rules random self-play: per env turn = Philox dice ->
full legal-turn enumeration (written to HBM) -> uniform action -> apply -> termination/reward ->
auto-reset -> Box(198) observation.  One "step" = one fused kernel launch over all envs of the GPU.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "legal-move-enumerated env steps/sec"
UNIT = "env_steps/s"
SEED = 0x5EED


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


_POLLER = r"""
import sys, time
import pynvml as n
n.nvmlInit()
pci = sys.argv[1]
h = n.nvmlDeviceGetHandleByPciBusId(pci.encode()) if pci != "-" else n.nvmlDeviceGetHandleByIndex(int(sys.argv[2]))
print("max", n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM), flush=True)
while True:
    try:
        print(time.time(), n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM), int(n.nvmlDeviceGetCurrentClocksEventReasons(h)),
              n.nvmlDeviceGetPowerUsage(h) / 1000.0, flush=True)
    except Exception as e:
        print("err", e, flush=True)
    time.sleep(0.001)
"""


class ClockSampler:
    """SM clock + clock-event (throttle) reasons sampled DURING the timed region.

    A helper PROCESS polls NVML every ~1 ms for the whole run (a thread would be starved by the launch
    loop holding the GIL; `nvidia-smi -lms` is too coarse for a timed region of tens of milliseconds);
    mark()/stop() keep the samples whose timestamps fall inside the timed region."""
    BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
            ("sw_power_cap", 0x4), ("hw_power_brake_slowdown", 0x80))

    def __init__(self, gpu_index, pci_bus_id=None):
        self.rows = []
        self.sm_max = None
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _POLLER, pci_bus_id or "-", str(gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = line.split()
            try:
                if f[0] == "max":
                    self.sm_max = float(f[1])
                elif f[0] != "err":
                    self.rows.append((float(f[0]), float(f[1]), int(f[2]), float(f[3])))
            except (ValueError, IndexError):
                pass

    def wait_ready(self, timeout=20.0):
        t = time.time()
        while self.proc is not None and not self.rows and time.time() - t < timeout and self.proc.poll() is None:
            time.sleep(0.01)

    def start(self):
        self.t0 = time.time()

    def stop(self):
        self.t1 = time.time()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        time.sleep(0.01)
        self.proc.terminate()
        rows = [r for r in self.rows if self.t0 <= r[0] <= self.t1]
        bits = 0
        for r in rows:
            bits |= r[2]
        return {"sm_mhz": statistics.median(r[1] for r in rows) if rows else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(k for k, b in self.BITS if bits & b), "samples": len(rows),
                "power_w_max": max(r[3] for r in rows) if rows else None,
                "source": "NVML polled every ~1 ms by a helper process; samples inside the timed region only"}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port (C restatement of the reference's algorithm) on the host cores
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, env_base, n_envs, n_steps = args
    from oracle import oracle as O
    t0 = time.perf_counter()
    n, a, e = O.selfplay(seed, env_base, n_envs, n_steps)
    return n, a, e, time.perf_counter() - t0


def cpu_selfplay(cores, envs_per_core, steps):
    """Full-rules random self-play on `cores` processes; returns (env_steps/s, total steps, mean A)."""
    import multiprocessing as mp
    from oracle import oracle as O
    O.build()
    jobs = [(SEED, c * envs_per_core, envs_per_core, steps) for c in range(cores)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    acts = sum(r[1] for r in res)
    return total / wall, total, acts / max(total, 1), wall


def run_reference_arm(args):
    """--impl reference: the reference algorithm's CPU implementation on all host cores.  The
    reference is pure Python and cannot travel to the GPU box (no /root/reference there), so this
    arm times the oracle port (oracle/narde_oracle.c), as the tier contract prescribes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    envs_per_core, sub_steps = 64, 100      # one "step" = 64*100 env turns per core (bounded sample)
    for _ in range(max(args.warmup, 1)):
        cpu_selfplay(cores, envs_per_core, 10)
    vals, tot, A = [], 0, 0.0
    t_all = 0.0
    for _ in range(args.steps):
        v, n, a, wall = cpu_selfplay(cores, envs_per_core, sub_steps)
        vals.append(v)
        tot += n
        A = a
        t_all += wall
    value = tot / t_all
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "full-rules random self-play (same dice/action stream as the CUDA arm)",
                   "envs_per_core": envs_per_core, "turns_per_env_per_step": sub_steps, "mean_legal_actions": A},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d procs x %d envs x %d turns per step, %d steps" % (cores, envs_per_core, sub_steps, args.steps)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------
def run_cuda_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0 and world > 1:
        print("warning: --gpus %d but WORLD_SIZE %d" % (args.gpus, world), file=sys.stderr)

    from gym_narde_b200 import VecNardeEnv, _cabi

    E, K, W = args.envs_per_gpu, args.steps, args.warmup
    env = VecNardeEnv(E, seed=SEED, max_actions=args.cap, env_base=rank * E, device=dev, chunks=args.chunks,
                      graph=not args.no_graph)
    env.reset()
    for _ in range(args.burn_in):           # de-correlate game phases: steady-state self-play mix
        env.step()
    torch.cuda.synchronize()

    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(step_fn, n_steps):
        """Per-step CUDA events on the launching stream; L2 flushed (untimed) between steps."""
        evs = []
        for _ in range(n_steps):
            if flush is not None:
                flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    # ---- device-resident arm: `value` ----
    for _ in range(W):
        env.step()
    stats0 = env.stats.clone()
    props = torch.cuda.get_device_properties(dev)
    try:
        pci = "%08X:%02X:%02X.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
    except Exception:
        pci = None
    sampler = ClockSampler(local_rank, pci)
    sampler.wait_ready()
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    ms = timed_loop(lambda: env.step(), K)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t_wall0)
    clocks = sampler.stop()
    dstats = (env.stats - stats0).cpu().tolist()
    total_ms = sum(ms)
    A = dstats[5] / float(E * K)

    # ---- end-to-end arm: host action indices in (pinned) -> step -> reward/done out (pinned) ----
    # uniformly random u32 fractions (fraction=True), FRESH every step from a pinned pool the "host policy" filled
    # ahead: the same i.i.d. uniform self-play policy as the device arm (a constant fraction per env is a different,
    # more expensive game: 0.147 instead of 0.110 ms/step with device-resident inputs)
    POOL = 16
    h_pool = torch.randint(-(1 << 31), (1 << 31) - 1, (POOL, E), dtype=torch.int64).to(torch.int32).pin_memory()
    h_idx = h_pool[0]
    e2e_t = [0]
    d_idx = env.action_in                                    # persistent device input of VecNardeEnv.step
    h_rew = torch.zeros(E, dtype=torch.float32).pin_memory()
    h_done = torch.zeros(E, dtype=torch.uint8).pin_memory()

    def e2e_step_copies():
        e2e_t[0] += 1
        d_idx.copy_(h_pool[e2e_t[0] % POOL], non_blocking=True)        # H2D: this step's inputs (action choices)
        obs, rew, term, trunc, info = env.step(d_idx, fraction=True)   # public API call (graph replay)
        h_rew.copy_(rew, non_blocking=True)                  # D2H: this step's results
        h_done.copy_(env.done, non_blocking=True)

    io = env.host_io()                                       # pinned host result buffers of the host-facing step

    def e2e_step():
        # public API call with HOST buffers, zero-copy: the fused step fetches this step's action choices from pinned
        # host memory (one bulk copy per CTA) and writes reward / done / truncated into pinned host memory
        e2e_t[0] += 1
        env.step_host(fraction=True, actions=h_pool[e2e_t[0] % POOL])

    for _ in range(max(W, POOL + 1)):          # one graph per pool buffer is captured on first use: all of them now
        e2e_step()
    barrier()
    ms_e2e = timed_loop(e2e_step, K)
    barrier()
    total_e2e = sum(ms_e2e)
    for _ in range(W):
        e2e_step_copies()
    ms_e2e_copies = timed_loop(e2e_step_copies, K)
    barrier()

    # ---- end-to-end, pipelined: VecNardeEnv.host_pipeline -- windows of 8 turns as ONE CUDA graph whose per-turn
    # H2D (action choices) / D2H (reward, done bits) copies run on copy streams, double-buffered on the device, so
    # the copies of neighbouring turns overlap the kernels.  Every turn still copies its own inputs in and its own
    # results out; timed as one device interval over K turns; no L2 flush is possible inside it, but a turn writes
    # ~120 MB of outputs (Box(198) 104 MB + action lists + ...) on top of the previous turn's, more than the L2. ----
    depth = 8
    pipe = env.host_pipeline(depth=depth, fraction=True)
    pipe.actions.copy_(h_pool[:depth])
    n_rep = max(1, (K + depth - 1) // depth)

    def e2e_pipelined(reps):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            pipe.run()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1)

    e2e_pipelined(2)
    barrier()
    ms_pipe = e2e_pipelined(n_rep) * K / (n_rep * depth)      # scaled to K turns (n_rep * depth >= K turns were timed)
    barrier()

    # ---- end-to-end with the whole Box(198) batch copied to the host as well (a host-side policy) ----
    h_obs = torch.empty((E, 198), dtype=torch.float32).pin_memory()

    def e2e_obs_step():
        e2e_step_copies()
        h_obs.copy_(env.obs, non_blocking=True)

    k_obs = max(3, min(K, 20))
    for _ in range(2):
        e2e_obs_step()
    barrier()
    ms_e2e_obs = timed_loop(e2e_obs_step, k_obs)
    barrier()

    # ---- config 2 side measurement (4096 lock-step envs, same kernel) ----
    small = VecNardeEnv(4096, seed=SEED, max_actions=args.cap, env_base=0, device=dev)
    small.reset()
    for _ in range(args.burn_in):
        small.step()
    torch.cuda.synchronize()
    ms_small = timed_loop(lambda: small.step(), min(K, 50))

    # ---- Tier R side measurement: the reference CODE's exact NardeEnv.step (narde_env.py:27-103) batched ----
    ref_env = VecNardeEnv(E, seed=SEED, rules="reference", device=dev)
    ref_env.reset()
    codes = torch.randint(0, 576, (E, 2), dtype=torch.int32, device=dev)
    for _ in range(30):
        ref_env.step(codes)
    torch.cuda.synchronize()
    ms_ref = timed_loop(lambda: ref_env.step(codes), min(K, 50))
    tier_r = {"value": E * len(ms_ref) / (sum(ms_ref) * 1e-3), "unit": "reference-exact env steps/s (rank 0)", "envs": E,
              "ms_per_step": sum(ms_ref) / len(ms_ref),
              "note": "k_roll_dice + k_step_ref: NardeEnv.step semantics of the reference code (2 dice, <= 2 half-moves, its "
                      "quirks), int32[24] observation, uniformly random action codes as examples/play_random_agent.py samples them"}

    # ---- config 5 side measurement: afterstate scoring with the tcgen05 MLP, 64K envs (rank 0's GPU) ----
    cfg5 = cfg3 = None
    if not args.no_config5 and rank == 0:
        cfg5 = config5_afterstate_scoring(torch, dev, args, timed_loop)
    if not args.no_config3 and rank == 0:
        cfg3 = config3_enumeration_microbench(torch, dev, args, timed_loop, measured_peaks()[0])

    # ---- policy-weight broadcast (config 4: NCCL broadcast of the packed DecomposedDQN(198) weights) ----
    bcast_ms = None
    if world > 1:
        from gym_narde_b200 import dist as ndist
        wp = torch.zeros(557056 // 2 + 16, dtype=torch.int16, device=dev)   # packed operand stages (~0.56 MB bf16)
        bs = torch.zeros(1088, dtype=torch.float32, device=dev)
        if rank == 0:
            wp.fill_(7)
            bs.fill_(0.5)
        ndist.broadcast_policy([wp, bs], src=0)                               # warm-up / communicator setup
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ndist.broadcast_policy([wp, bs], src=0)
        b.record()
        torch.cuda.synchronize()
        assert int(wp[0].item()) == 7 and float(bs[0].item()) == 0.5
        bcast_ms = a.elapsed_time(b)

    # ---- reduce over ranks (MAX time), gather episode stats with NCCL ----
    tmax = torch.tensor([total_ms, total_e2e, ms_pipe], dtype=torch.float64, device=dev)
    st = env.stats.clone()
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(st) for _ in range(world)]
        dist.all_gather(gathered, st)                        # config 4: allgather of episode stats
        st_all = torch.stack(gathered)
    else:
        st_all = st[None]
    total_ms_max, total_e2e_max, ms_pipe_max = tmax.cpu().tolist()

    if rank == 0:
        peak, peak_src = measured_peaks()
        units = world * E * K
        value = units / (total_ms_max * 1e-3)
        e2e_value = units / (total_e2e_max * 1e-3)
        bytes_per_unit = 871 + 8 * A                         # B_step(A)
        kernel_ms = total_ms / K                             # rank-0 kernel: one launch per step
        achieved = E * bytes_per_unit / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01f_step_full_v2.json")   # ncu --set full of the same kernel/workload
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp))["launches"][0].get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8/u32 bit arithmetic (obs f32)", "data": "synthetic",
            "config": {"workload": "config4 shard: %d lock-step envs/GPU full-rules random self-play (1M envs at 8 GPUs)" % E,
                       "envs_per_gpu": E, "action_capacity": args.cap, "burn_in_steps": args.burn_in,
                       "mean_legal_actions": A, "max_legal_actions": int(st_all[:, 6].max().item()),
                       "l2": "flushed between timed steps (256 MiB fill, untimed)" if flush is not None else "not flushed",
                       "parallelism": "env-sharded x%d, no data-path collective" % world},
            "roofline": {"bound": "hbm", "kernel": "k_step_full_v2<128,true> + its programmatic dependent k_step_deferred<256> (order-dependent doubles turns, overlaps the tail), timed together as one step", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": "ncu --set full, profiles/r01f_step_full_v2.json (per launch)", "peak_source": peak_src,
                         "algorithmic_bytes_per_env_step": bytes_per_unit, "kernel_ms": kernel_ms},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * E, "d2h_bytes_per_step": 6 * E,
                    "ms_per_step": total_e2e_max / K,
                    "note": "VecNardeEnv.step_host(fraction=True, actions=pool row), zero-copy, one CUDA-graph replay per step: every CTA of the fused step bulk-copies its envs' int32 action choices (fresh u32 fractions of the legal list, pinned pool) from host memory into shared memory, and reward f32 / done u8 / truncated u8 are written by the kernel straight into pinned host memory; Box(198) stays in HBM for the device-resident policy"},
            "e2e_explicit_copies": {"value": world * E * K / (sum(ms_e2e_copies) * 1e-3), "unit": UNIT, "ms_per_step": sum(ms_e2e_copies) / K,
                                    "note": "the same turn with cudaMemcpyAsync H2D / D2H around VecNardeEnv.step (rank 0's time)"},
            "e2e_pipelined": {"value": units / (ms_pipe_max * 1e-3), "unit": UNIT, "ms_per_step": ms_pipe_max / K,
                              "h2d_bytes_per_step": 4 * E, "d2h_bytes_per_step": 6 * E,
                              "note": "VecNardeEnv.host_pipeline(depth=8, fraction=True): 8 turns per CUDA-graph replay, every turn with its "
                                      "own DMA of pinned action choices on a copy-in stream (double-buffered on the device) and its results "
                                      "written by the kernel straight into pinned host rows; one device interval, no L2 flush (a turn's "
                                      "outputs exceed the L2)"},
            "e2e_obs_to_host": {"value": world * E * k_obs / (sum(ms_e2e_obs) * 1e-3), "unit": UNIT, "steps": k_obs,
                                "d2h_bytes_per_step": 5 * E + 792 * E, "ms_per_step": sum(ms_e2e_obs) / k_obs,
                                "note": "same as e2e plus the full Box(198) float32 batch copied D2H every step (PCIe-bound; rank 0's time)"},
            # k_step_full_v2 + k_step_deferred per step (+ k_advance_counter when a chunked env replays a graph)
            "gpu_launches": K * ((3 if (env.use_graph and env._ws_adv is None) else 2) * len(env._chunks)),
            "wall_ms": wall_ms, "clocks": clocks,
            "config2_4096_envs": {"value": 4096 * len(ms_small) / (sum(ms_small) * 1e-3), "unit": UNIT,
                                  "ms_per_step": sum(ms_small) / len(ms_small)},
            "episode_stats": {k: int(v) for k, v in zip(_cabi.STAT_NAMES, st_all.sum(0).tolist())},
        }
        line["tier_r_reference_rules_step"] = tier_r
        if bcast_ms is not None:
            line["policy_broadcast_ms"] = bcast_ms
        if cfg5 is not None:
            line["config5_afterstate_scoring"] = cfg5
        if cfg3 is not None:
            line["config3_enumeration_microbench"] = cfg3
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            cpu_selfplay(cores, 64, 150)                        # pool / library warm-up
            v, n, a, wall = cpu_selfplay(cores, 64, 2000)       # rate probe (~1-2 s)
            # size the reported sample to ~12 s of CPU work from the probed rate
            v, n, a, wall = cpu_selfplay(cores, 64, max(2000, min(200000, int(12.0 * v / (cores * 64)))))
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d procs x 64 envs, %d env turns total in %.1f s (oracle/narde_oracle.c o_selfplay)" % (cores, n, wall)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def config3_enumeration_microbench(torch, dev, args, timed_loop, peak):
    """BASELINE config 3: get_valid_actions over 1M synthetic positions, seed 1234, three strata:
    A 40% self-play states sampled at a uniformly random ply in [0,90]; B 30% stratum-A states with the dice
    forced to doubles; C 30% bear-off races (15-k mover checkers over points 0..5, 15-k' opponent checkers over
    the mover-frame points 12..17, first_turn False, uniform dice)."""
    from gym_narde_b200 import VecNardeEnv, _cabi
    g = torch.Generator(device=dev).manual_seed(1234)
    n = 1 << 20
    nA, nB = int(0.4 * n), int(0.3 * n)
    nC = n - nA - nB
    env = VecNardeEnv(nA, seed=1234, max_actions=1, device=dev, write_actions=False)
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
    cap = args.cap
    actions = torch.zeros((n, cap), dtype=torch.int64, device=dev)
    counts = torch.zeros(n, dtype=torch.int32, device=dev)
    ovf = torch.zeros(n, dtype=torch.uint8, device=dev)
    ws = torch.zeros(_cabi.workspace_ints(n), dtype=torch.int32, device=dev)
    run = lambda: _cabi.enumerate_actions_fast(lo, hi, dice, actions, counts, ovf, ws)
    for _ in range(3):
        run()
    k = max(5, min(args.steps, 20))
    ms = timed_loop(run, k)
    mean_ms = sum(ms) / k
    # cross-check against the independent thread-per-env enumerator on a 10k-position subsample
    sub = torch.randperm(n, device=dev, generator=g)[:10000]
    a2 = torch.zeros((10000, cap), dtype=torch.int64, device=dev)
    c2 = torch.zeros(10000, dtype=torch.int32, device=dev)
    _cabi.enumerate_actions(lo[sub].contiguous(), hi[sub].contiguous(), dice[sub].contiguous(), a2, c2, None)
    keep = torch.arange(cap, device=dev)[None, :] < c2.clamp(max=cap)[:, None]
    same = bool(torch.equal(c2, counts[sub]) and torch.equal(a2[keep], actions[sub][keep]))
    A = float(counts.float().mean().item())
    bytes_per = 38 + 8 * float(counts.clamp(max=cap).float().mean().item())
    strata = {}
    for name, a, b in (("A_selfplay", 0, nA), ("B_doubles", nA, nA + nB), ("C_bearoff", nA + nB, n)):
        strata[name] = {"positions": b - a, "mean_legal": float(counts[a:b].float().mean().item()),
                        "max_legal": int(counts[a:b].max().item())}
    return {"positions": n, "positions_per_s": n / (mean_ms * 1e-3), "ms": mean_ms, "cap": cap, "mean_legal_actions": A,
            "max_legal_actions": int(counts.max().item()), "overflow_positions": int(ovf.sum().item()),
            "deferred_exact_positions": int(ws[0].item()), "strata": strata,
            "subsample_equals_thread_per_env_enumerator": same,
            "achieved_GBps": n * bytes_per / (mean_ms * 1e-3) / 1e9, "hbm_frac": n * bytes_per / (mean_ms * 1e-3) / 1e9 / peak,
            "algorithmic_bytes_per_position": bytes_per}


def config5_afterstate_scoring(torch, dev, args, timed_loop):
    """BASELINE config 5: DecomposedDQN(198) (train_deepq_pytorch.py:184-236, torch.manual_seed(0) random init) over
    all legal afterstates of 65536 envs: enumerate -> afterstates -> Box(198) encode + 3-layer MLP + max_a Q
    on the tcgen05 tensor cores -> greedy choice -> env step, all on the device (AfterstateActor)."""
    import torch.nn as nn
    from gym_narde_b200 import VecNardeEnv, AfterstateMLP, AfterstateActor
    torch.manual_seed(0)
    fn = nn.Sequential(nn.Linear(198, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU()).to(dev)
    head = nn.Linear(256, 576).to(dev)
    mlp = AfterstateMLP.from_module(fn, head)
    E5 = 65536
    env = VecNardeEnv(E5, seed=SEED, max_actions=args.cap, device=dev)
    actor = AfterstateActor(env, mlp)
    env.reset()
    for _ in range(60):                      # random self-play burn-in, then greedy turns
        env.step()
    for _ in range(5):
        actor.step_graph()
    torch.cuda.synchronize()
    k = max(5, min(args.steps, 30))
    rows = []

    def greedy_step():
        actor.step_graph()                   # one CUDA-graph replay per greedy turn
        rows.append(actor.rows_dev.clone())

    ms_actor = timed_loop(greedy_step, k)
    mean_rows = float(torch.stack(rows).float().mean().item())
    # the scorer alone on the last step's afterstates (same rows, L2 flushed between launches)
    ms_mlp = timed_loop(lambda: mlp.score_states(actor.as_lo, actor.as_hi, out=actor.scores, rows_dev=actor.rows_dev), k)
    last_rows = int(actor.rows_dev.item())
    flop_row = 2 * (198 * 256 + 256 * 256 + 256 * 576)
    mlp_ms = sum(ms_mlp) / len(ms_mlp)
    tfl = last_rows * flop_row / (mlp_ms * 1e-3) / 1e12
    peak_tf = 1361.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        try:
            peak_tf = float(json.load(open(pk))["bf16_tflops_sustained"])
        except Exception:
            pass
    return {"envs": E5, "greedy_env_steps_per_s": E5 * k / (sum(ms_actor) * 1e-3), "ms_per_greedy_step": sum(ms_actor) / k,
            "afterstate_rows_per_step": mean_rows, "scorer_rows": last_rows, "scorer_ms": mlp_ms,
            "scorer_rows_per_s": last_rows / (mlp_ms * 1e-3), "scorer_tflops": tfl,
            "scorer_frac_of_bf16_peak": tfl / peak_tf, "bf16_peak_tflops": peak_tf,
            "flops_per_row": flop_row, "dtype": "bf16 operands, fp32 accumulate (tcgen05)",
            "note": "k_mlp<states in, row-max out>: Box(198) encoded in-kernel from 32-byte afterstates; weights random init"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=131072)
    ap.add_argument("--cap", type=int, default=64)
    ap.add_argument("--burn-in", type=int, default=300)
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--chunks", type=int, default=None)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--no-config3", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_cuda_arm(args)


if __name__ == "__main__":
    main()
