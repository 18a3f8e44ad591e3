"""Batched afterstate-greedy actor: the device-resident replacement of
the per-move loop of DQNAgent.act (train_deepq_pytorch.py:411-600).

Per lock-step turn, entirely on the GPU and without a host synchronisation:
    roll (Philox) -> get_valid_actions (full legal-turn enumeration) -> afterstate of every legal
    action (narde_afterstates) -> Box(198) encoding + DecomposedDQN(198) forward + max_a Q on the
    tcgen05 tensor cores (narde_mlp_score_states) -> greedy choice per env (narde_segment_argmax)
    -> VecNardeEnv.step(action_idx, dice).
"""
from __future__ import annotations

import ctypes as C

from . import _cabi


class _SideBatch:
    """Buffers of the second pass over the environments whose legal list exceeds env.max_actions."""

    def __init__(self, t, dev, m, cap):
        self.m, self.cap = m, cap
        self.lo = t.zeros((m, 16), dtype=t.uint8, device=dev)
        self.hi = t.zeros((m, 16), dtype=t.uint8, device=dev)
        self.dice = t.ones((m, 2), dtype=t.uint8, device=dev)
        self.idx = t.full((m,), -1, dtype=t.int32, device=dev)
        self.ctrl = t.zeros(4, dtype=t.int32, device=dev)
        self.actions = t.zeros((m, cap), dtype=t.int64, device=dev)
        self.counts = t.zeros(m, dtype=t.int32, device=dev)
        self.counts_eff = t.zeros(m, dtype=t.int32, device=dev)
        self.ovf = t.zeros(m, dtype=t.uint8, device=dev)
        self.ws = t.zeros(_cabi.workspace_ints(m), dtype=t.int32, device=dev)
        self.scan_ws = t.zeros((m + 127) // 128 + 4, dtype=t.int64, device=dev)
        self.offsets = t.zeros(m, dtype=t.int64, device=dev)
        self.rows_dev = t.zeros(1, dtype=t.int64, device=dev)
        self.choice = t.zeros(m, dtype=t.int32, device=dev)
        self.value = t.zeros(m, dtype=t.float32, device=dev)


class AfterstateActor:
    def __init__(self, env, mlp, mode="max", overflow_slots=None, overflow_cap=3072, overflow_rows=None):
        """env: VecNardeEnv(rules="full"); mlp: AfterstateMLP; mode: "max" (the mover maximises the score)
        or "white_value" (the net scores positions for WHITE: WHITE maximises, BLACK minimises).

        Every legal action is scored, as DQNAgent.act does (train_deepq_pytorch.py:430-507): environments whose list
        is longer than env.max_actions (doubles turns, up to ~1300 actions; 0.5-1 % of the envs in self-play) get a
        second pass with capacity `overflow_cap` (default 3072 >= C(18, 4) = 3060, the number of 4-multisets of 15 sources:
        every doubles turn fits; 2018 legal turns have been seen in self-play) in a side batch of `overflow_slots` environments (default N/16, at
        least 256) sharing `overflow_rows` afterstate rows (default 192 per slot).  Environments that do not fit --
        more overflowing envs than slots, a list longer than overflow_cap, or the row pool exhausted -- are counted in
        `uncovered_envs()` and keep the choice among their first max_actions actions; overflow_slots=0 switches the
        second pass off."""
        if env.rules != "full":
            raise ValueError("AfterstateActor needs rules='full'")
        if mode not in ("max", "white_value"):
            raise ValueError("mode must be 'max' or 'white_value'")
        t = env.torch
        self.env, self.mlp, self.mode = env, mlp, 0 if mode == "max" else 1
        n, cap, dev = env.num_envs, env.max_actions, env.device
        self.cap = cap
        self.as_lo = t.zeros((n * cap, 16), dtype=t.uint8, device=dev)   # afterstate planes, ragged rows packed
        self.as_hi = t.zeros((n * cap, 16), dtype=t.uint8, device=dev)
        self.scores = t.zeros(n * cap, dtype=t.float32, device=dev)
        self.offsets = t.zeros(n, dtype=t.int64, device=dev)
        self.rows_dev = t.zeros(1, dtype=t.int64, device=dev)
        self._scan_ws = t.zeros((n + 127) // 128 + 4, dtype=t.int64, device=dev)   # NARDE_AFTERSTATE_SCRATCH_WORDS(n)
        self.choice = t.zeros(n, dtype=t.int32, device=dev)
        self.value = t.zeros(n, dtype=t.float32, device=dev)
        self.dice = env.dice                                        # the turn's dice (written by the enumeration)
        self.act_override = t.zeros(n, dtype=t.int64, device=dev)   # chosen actions beyond the stored lists (side batch)
        self.lib = _cabi.load()
        m = max(256, n // 16) if overflow_slots is None else int(overflow_slots)
        m = -(-m // 128) * 128            # whole tiles of the side batch's kernels
        self.side = None
        if m > 0:
            self.side = sb = _SideBatch(t, dev, m, int(overflow_cap))
            rows = int(overflow_rows) if overflow_rows is not None else 192 * m
            sb.rows_cap = rows
            sb.as_lo = t.zeros((rows, 16), dtype=t.uint8, device=dev)
            sb.as_hi = t.zeros((rows, 16), dtype=t.uint8, device=dev)
            sb.scores = t.zeros(rows, dtype=t.float32, device=dev)
        self._graph = None
        self._s2 = t.cuda.Stream(device=dev) if self.side is not None else None
        self._enum_ws = t.zeros(_cabi.workspace_ints(n), dtype=t.int32, device=dev)

    def _stream(self):
        return C.c_void_p(self.env.torch.cuda.current_stream().cuda_stream)

    def afterstates(self, actions, counts):
        """Fill as_lo/as_hi with the afterstates of actions[i, :min(counts[i], cap)]; returns offsets [N]."""
        env = self.env
        # one launch: the device-wide exclusive scan of min(counts, cap) (decoupled look-back), the row count and the
        # rows themselves (narde_afterstates_scan) -- no host-side prefix sum, no torch ops
        rc = self.lib.narde_afterstates_scan(C.c_void_p(env.lo.data_ptr()), C.c_void_p(env.hi.data_ptr()),
                                             C.c_void_p(actions.data_ptr()), C.c_void_p(counts.data_ptr()),
                                             env.num_envs, self.cap, C.c_void_p(self.offsets.data_ptr()),
                                             C.c_void_p(self.rows_dev.data_ptr()), C.c_void_p(self.as_lo.data_ptr()),
                                             C.c_void_p(self.as_hi.data_ptr()), None,
                                             C.c_void_p(self._scan_ws.data_ptr()), 0, None, self._stream())
        if rc != 0:
            raise _cabi.NardeCudaError("narde_afterstates_scan failed: %d" % rc)
        return self.offsets

    def _side_prepare(self, dice):
        """Second pass (see __init__): gather -> enumerate with the large capacity -> afterstate rows -> score -> arg-max.  It only
        needs the first pass's enumeration (overflow flags), so it runs on a side stream BESIDE the first pass's rows / scorer /
        arg-max (small latency-bound kernels next to the one-CTA-per-SM scorer); fork / join with stream waits, which a
        CUDA-graph capture records as a branch."""
        sb, env, P, t = self.side, self.env, C.c_void_p, self.env.torch
        if sb is None:
            return
        main = t.cuda.current_stream(env.device)
        self._s2.wait_stream(main)
        with t.cuda.stream(self._s2):
            self._side_prepare_launches(dice)

    def _side_prepare_launches(self, dice):
        sb, env, P = self.side, self.env, C.c_void_p
        st = self._stream()
        rc = self.lib.narde_gather_overflow(P(env.lo.data_ptr()), P(env.hi.data_ptr()), P(dice.data_ptr()),
                                            P(env.overflow.data_ptr()), env.num_envs, sb.m, P(sb.lo.data_ptr()),
                                            P(sb.hi.data_ptr()), P(sb.dice.data_ptr()), P(sb.idx.data_ptr()),
                                            P(sb.ctrl.data_ptr()), st)
        if rc != 0:
            raise _cabi.NardeCudaError("narde_gather_overflow failed: %d" % rc)
        _cabi.enumerate_actions_fast(sb.lo, sb.hi, sb.dice, sb.actions, sb.counts, sb.ovf, sb.ws)
        rc = self.lib.narde_afterstates_scan(P(sb.lo.data_ptr()), P(sb.hi.data_ptr()), P(sb.actions.data_ptr()),
                                             P(sb.counts.data_ptr()), sb.m, sb.cap, P(sb.offsets.data_ptr()),
                                             P(sb.rows_dev.data_ptr()), P(sb.as_lo.data_ptr()), P(sb.as_hi.data_ptr()),
                                             None, P(sb.scan_ws.data_ptr()), sb.rows_cap, P(sb.counts_eff.data_ptr()), st)
        if rc != 0:
            raise _cabi.NardeCudaError("narde_afterstates_scan failed: %d" % rc)
        # (the scorer keeps no state between launches, so this one may be queued beside the first pass's: its CTAs get
        # their SMs when those leave)
        self.mlp.score_states(sb.as_lo, sb.as_hi, out=sb.scores, rows_dev=sb.rows_dev)
        rc = self.lib.narde_segment_argmax(P(sb.scores.data_ptr()), P(sb.offsets.data_ptr()), P(sb.counts_eff.data_ptr()),
                                           P(sb.hi.data_ptr()), sb.m, sb.cap, self.mode, P(sb.choice.data_ptr()),
                                           P(sb.value.data_ptr()), st)
        if rc != 0:
            raise _cabi.NardeCudaError("narde_segment_argmax failed: %d" % rc)

    def _side_finish(self):
        """Second pass, after the join: scatter the side batch's choices into self.choice / self.value /
        self.act_override."""
        sb, env, P, t = self.side, self.env, C.c_void_p, self.env.torch
        if sb is None:
            return
        t.cuda.current_stream(env.device).wait_stream(self._s2)
        st = self._stream()
        rc = self.lib.narde_scatter_choice(P(sb.choice.data_ptr()), P(sb.value.data_ptr()), P(sb.idx.data_ptr()),
                                           P(sb.counts_eff.data_ptr()), P(sb.counts.data_ptr()), sb.m, sb.cap,
                                           P(self.choice.data_ptr()), P(self.value.data_ptr()), P(sb.ctrl.data_ptr()),
                                           P(sb.actions.data_ptr()), P(self.act_override.data_ptr()), st)
        if rc != 0:
            raise _cabi.NardeCudaError("narde_scatter_choice failed: %d" % rc)

    def uncovered_envs(self):
        """Env turns whose choice was made among the first max_actions actions only (one D2H copy)."""
        return int(self.side.ctrl[2].item()) if self.side is not None else -1

    def choose(self):
        """roll -> enumerate -> afterstates -> score -> greedy index.  Returns (choice [N] int32, dice [N,2])."""
        env = self.env
        env.roll()                        # (writes env.dice, which self.dice is)
        actions, counts, _ = env.get_valid_actions(self.dice)
        self._side_prepare(self.dice)
        self.afterstates(actions, counts)
        self.mlp.score_states(self.as_lo, self.as_hi, out=self.scores, rows_dev=self.rows_dev)
        rc = self.lib.narde_segment_argmax(C.c_void_p(self.scores.data_ptr()), C.c_void_p(self.offsets.data_ptr()),
                                           C.c_void_p(counts.data_ptr()), C.c_void_p(env.hi.data_ptr()), env.num_envs,
                                           self.cap, self.mode, C.c_void_p(self.choice.data_ptr()),
                                           C.c_void_p(self.value.data_ptr()), self._stream())
        if rc != 0:
            raise _cabi.NardeCudaError("narde_segment_argmax failed: %d" % rc)
        self._side_finish()
        return self.choice, self.dice

    def _play_chosen(self):
        """narde_step_chosen: the lists of this turn are in env.actions / env.counts and the choice is made -- apply it and
        complete the turn (reward, termination, auto-reset, statistics, Box(198)) without enumerating the position again."""
        env, P = self.env, C.c_void_p
        flags = (_cabi.REWARD_MOVER12 if env.reward_mode == "mover12" else 0) | (_cabi.AUTORESET if env.autoreset else 0)
        rc = self.lib.narde_step_chosen(P(env.lo.data_ptr()), P(env.hi.data_ptr()), env.num_envs, env.env_base, env.seed, 0,
                                        P(self.dice.data_ptr()), P(self.choice.data_ptr()), P(env.actions.data_ptr()), self.cap,
                                        P(env.counts.data_ptr()), P(self.act_override.data_ptr()), P(env.chosen.data_ptr()),
                                        P(env.obs.data_ptr()), P(env.reward.data_ptr()), P(env.done.data_ptr()),
                                        P(env.trunc.data_ptr()), P(env.stats.data_ptr()), flags, env.max_episode_steps,
                                        P(env._step_dev.data_ptr()), self._stream())
        if rc != 0:
            raise _cabi.NardeCudaError("narde_step_chosen failed: %d" % rc)

    def step(self):
        """One greedy lock-step turn for all envs; returns VecNardeEnv.step's tuple."""
        env = self.env
        self.choose()
        env.step_count += 1
        env._step_dev.fill_(env.step_count)
        self._play_chosen()
        return env.obs, env.reward, env.terminated, env.truncated, env.info

    # -- the same turn as ONE CUDA-graph replay ---------------------------------------------------
    def _enqueue_turn(self):
        """All launches of a greedy turn with the step number read from the env's device counter, so that the
        captured graph can be replayed: enumerate (Philox dice of the turn) -> afterstates -> score -> argmax ->
        fused step with the chosen indices (which re-derives the same dice from the same counter)."""
        env, t = self.env, self.env.torch
        flags = (_cabi.REWARD_MOVER12 if env.reward_mode == "mover12" else 0) | (_cabi.AUTORESET if env.autoreset else 0)
        _cabi.advance_counter(env._step_dev)
        _cabi.step_full(env.lo, env.hi, env.env_base, env.seed, 0, actions=env.actions, counts=env.counts,
                        dice_out=self.dice, done=env.overflow, flags=_cabi.ENUMERATE_ONLY, workspace=self._enum_ws,
                        step_dev=env._step_dev)
        self._side_prepare(self.dice)
        self.afterstates(env.actions, env.counts)
        self.mlp.score_states(self.as_lo, self.as_hi, out=self.scores, rows_dev=self.rows_dev)
        rc = self.lib.narde_segment_argmax(C.c_void_p(self.scores.data_ptr()), C.c_void_p(self.offsets.data_ptr()),
                                           C.c_void_p(env.counts.data_ptr()), C.c_void_p(env.hi.data_ptr()), env.num_envs,
                                           self.cap, self.mode, C.c_void_p(self.choice.data_ptr()),
                                           C.c_void_p(self.value.data_ptr()), self._stream())
        if rc != 0:
            raise _cabi.NardeCudaError("narde_segment_argmax failed: %d" % rc)
        self._side_finish()
        self._play_chosen()

    def step_graph(self):
        """step() as one CUDA-graph replay (captured on first use; needs the env built with chunks=1)."""
        env, t = self.env, self.env.torch
        if len(env._chunks) != 1:
            raise ValueError("step_graph needs an unchunked env")
        env.step_count += 1
        if self._graph is not None and self._graph_seed != env.seed:
            self._graph = None           # the seed is a frozen kernel argument of the captured turn
        if self._graph is None:
            self._enqueue_turn_warmup()
            env._step_dev.fill_(env.step_count - 1)
            t.cuda.synchronize(env.device)
            g = t.cuda.CUDAGraph()
            with t.cuda.graph(g):
                self._enqueue_turn()
            self._graph = g
            self._graph_seed = env.seed
        self._graph.replay()
        return env.obs, env.reward, env.terminated, env.truncated, env.info

    def _enqueue_turn_warmup(self):
        """First-use work that must not happen during capture (function attributes, lazy allocations): run the
        read-only part of a turn once, outside the graph."""
        env = self.env
        saved = env._step_dev.clone()
        _cabi.step_full(env.lo, env.hi, env.env_base, env.seed, 0, actions=env.actions, counts=env.counts,
                        dice_out=self.dice, done=env.overflow, flags=_cabi.ENUMERATE_ONLY, workspace=self._enum_ws,
                        step_dev=env._step_dev)
        self.afterstates(env.actions, env.counts)
        self.mlp.score_states(self.as_lo, self.as_hi, out=self.scores, rows_dev=self.rows_dev)
        env._step_dev.copy_(saved)
