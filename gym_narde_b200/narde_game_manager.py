"""Turn manager for interactive play: the reference's NardeGameManager surface
(my_game/narde_game_manager.py:13-1099 -- roll_dice :127, make_move :190, get_valid_moves_for_position :745,
undo_moves :798, is_game_over :844, get_game_state :857, set_current_player :881) over the GPU rules engine.

The reference manager re-derives half-move legality in Python on top of Narde.get_valid_moves and hands the
colour STRING to a function that compares it with 1 (narde.py:60 vs narde_game_manager.py:161), so White is
shown Black's view.  This one keeps no rules of its own.  When the dice are rolled it builds the whole tree of
partial turns with BATCHED device calls -- per ply one narde_half_moves launch over (position, die) rows
(Narde.get_valid_moves([die]), narde.py:58-92) and one narde_apply_actions(HALF_MOVES_ONLY) launch over the
candidate rows (execute_rotated_move, narde.py:36-56) -- and composes the turn-level rules on the host exactly as
the Tier-N enumeration does (per-turn head budget, both dice orders, maximal dice usage, the higher die when only
one can be played).  A half-move is offered iff it lies on a path to a legal end-of-turn position; those end
positions are, as a set, the afterstates of get_valid_actions(roll) (tests/test_game_manager.py checks that on
self-play and synthetic positions).

Positions are in the MOVER's frame, as Narde.get_valid_moves returns them (own head = 23, moving towards 0,
'off' = borne off; get_valid_moves_for_position reports 'off' as -1 like the reference); boards in the responses
are in the absolute White frame (game.board).  The manager mutates env.game like the reference does.
"""
from __future__ import annotations

import random
from collections import defaultdict

import numpy as np

from . import _cabi
from . import state as S
from .envs.narde import rotate_board

HEAD = 23
HEAD_POSITIONS = {"white": HEAD, "black": HEAD}          # mover's frame
_TURN = {"white": 1, "black": -1}


class _Node:
    __slots__ = ("lo", "hi", "remaining", "head_used", "depth", "edges", "parents", "viable")

    def __init__(self, lo, hi, remaining, head_used, depth):
        self.lo, self.hi = lo, hi
        self.remaining, self.head_used, self.depth = remaining, head_used, depth
        self.edges = {}          # (from, to, die) -> child node
        self.parents = []
        self.viable = False


class _CudaOps:
    """The two batched rules kernels the tree is built from, numpy in / numpy out through the C ABI."""

    def __init__(self):
        self.torch = _cabi.require_cuda()
        _cabi.load()
        self.dev = self.torch.device("cuda")

    def _up(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    def half_moves(self, lo, hi, dice4):
        t = self.torch
        n = lo.shape[0]
        moves = t.zeros((n, _cabi.MAX_HALF_MOVES, 2), dtype=t.uint8, device=self.dev)
        counts = t.zeros(n, dtype=t.int32, device=self.dev)
        _cabi.half_moves(self._up(lo), self._up(hi), self._up(dice4), moves, counts)
        return moves.cpu().numpy(), counts.cpu().numpy()

    def apply_actions(self, lo, hi, acts, flags=0):
        tlo, thi = self._up(lo), self._up(hi)
        _cabi.apply_actions(tlo, thi, self._up(np.asarray(acts, dtype=np.uint64).view(np.int64)), flags=flags)
        lo[...] = tlo.cpu().numpy()
        hi[...] = thi.cpu().numpy()


class TurnTree:
    """Every partial turn of one roll.  Nodes are (position, remaining dice); positions reached twice are merged.
    `ops` (test-only) swaps the object the two kernels are called through; the product always uses the CUDA one
    and raises without a GPU."""

    def __init__(self, lo, hi, dice, first_turn, ops=None):
        self.ops = ops if ops is not None else _CudaOps()
        d = sorted((int(dice[0]), int(dice[1])), reverse=True)
        self.doubles = d[0] == d[1]
        self.max_head = 2 if (first_turn and self.doubles and d[0] in (3, 4, 6)) else 1   # narde.py:100-103
        remaining = tuple(d * 2) if self.doubles else tuple(d)
        self.root = _Node(bytes(lo), bytes(hi), remaining, 0, 0)
        self.nodes = {(self.root.lo, self.root.hi[:12], remaining): self.root}
        self.launches = 0
        frontier = [self.root]
        while frontier:
            frontier = self._expand(frontier)
        self._mark_legal(d)

    @staticmethod
    def _stack(rows):
        lo = np.frombuffer(b"".join(r[0] for r in rows), dtype=np.uint8).reshape(-1, 16).copy()
        hi = np.frombuffer(b"".join(r[1] for r in rows), dtype=np.uint8).reshape(-1, 16).copy()
        return lo, hi

    def _expand(self, frontier):
        rows = [(nd, die) for nd in frontier for die in sorted(set(nd.remaining), reverse=True)]
        if not rows:
            return []
        lo, hi = self._stack([(nd.lo, nd.hi) for nd, _ in rows])
        dice4 = np.zeros((len(rows), 4), dtype=np.uint8)
        dice4[:, 0] = [die for _, die in rows]
        moves_h, counts_h = self.ops.half_moves(lo, hi, dice4)       # one launch: every (position, die) of this ply
        self.launches += 1
        cand = []                                                    # (row, from, to)
        for r, (nd, die) in enumerate(rows):
            for f, to in moves_h[r, :counts_h[r]]:
                f = int(f)
                if f == HEAD and nd.head_used >= self.max_head:      # per-turn head budget
                    continue
                cand.append((r, f, "off" if int(to) == S.OFF else int(to)))
        if not cand:
            return []
        idx = np.array([c[0] for c in cand], dtype=np.int64)
        clo_h, chi_h = lo[idx].copy(), hi[idx].copy()
        acts = np.array([S.encode_action([(f, to)]) for _, f, to in cand], dtype=np.uint64)
        self.ops.apply_actions(clo_h, chi_h, acts, flags=_cabi.HALF_MOVES_ONLY)   # one launch: every candidate
        self.launches += 1
        new = []
        for k, (r, f, to) in enumerate(cand):
            nd, die = rows[r]
            rem = list(nd.remaining)
            rem.remove(die)
            rem = tuple(rem)
            lo_b, hi_b = clo_h[k].tobytes(), chi_h[k].tobytes()
            key = (lo_b, hi_b[:12], rem)
            child = self.nodes.get(key)
            if child is None:
                child = _Node(lo_b, hi_b, rem, nd.head_used + (f == HEAD), nd.depth + 1)
                self.nodes[key] = child
                new.append(child)
            nd.edges[(f, to, die)] = child
            child.parents.append(nd)
        return new

    def _mark_legal(self, d):
        depth = max(nd.depth for nd in self.nodes.values())
        self.max_depth = depth
        ends = [nd for nd in self.nodes.values() if nd.depth == depth] if depth else [self.root]
        if depth == 1 and not self.doubles:                      # rule 4 (narde.py:6): the higher die if either plays
            high = [nd for nd in ends if nd.remaining == (d[1],)]
            ends = high or ends
        self.ends = ends
        stack = list(ends)
        while stack:
            nd = stack.pop()
            if not nd.viable:
                nd.viable = True
                stack.extend(nd.parents)

    def moves_from(self, nodes):
        """Half-moves (from, to) that keep a legal end of turn reachable, in die-descending, point-ascending order.
        `nodes`: the candidates for "where the turn stands" -- more than one only after a bear-off that either
        remaining die could have paid for (the dice are not committed until a later half-move needs one of them)."""
        seen, out = set(), []
        edges = [kv for nd in nodes for kv in nd.edges.items()]
        for (f, to, die), child in sorted(edges, key=lambda kv: (-kv[0][2], kv[0][0])):
            if child.viable and (f, to) not in seen:
                seen.add((f, to))
                out.append((f, to))
        return out

    def children(self, nodes, f, to):
        """The candidates after half-move (f, to), the one that paid with the smallest die first."""
        opts = {}
        for nd in nodes:
            for (ff, tt, die), ch in nd.edges.items():
                if ff == f and tt == to and ch.viable:
                    opts.setdefault(id(ch), (die, ch))
        return [ch for _, ch in sorted(opts.values(), key=lambda o: o[0])]

    def end_states(self):
        return {(nd.lo, nd.hi[:10]) for nd in self.ends}


class NardeGameManager:
    """Drop-in for my_game/narde_game_manager.py:NardeGameManager (same constructor, methods and response keys)."""

    def __init__(self, narde_env, _ops=None):
        self._ops = _ops if _ops is not None else _CudaOps()      # raises without CUDA: there is no CPU path
        self.env = narde_env
        self.game = narde_env.game
        self.current_player = "white"
        self.dice_state = {"original": [], "expanded": [], "remaining": [], "used": []}
        self.valid_moves = []
        self.valid_moves_by_piece = {}
        self.first_move_made = False
        self.move_options = []
        self.head_move_made = {"white": False, "black": False}
        self.head_moves_count = {"white": 0, "black": 0}
        self.saved_state = None
        self.is_first_turn_special_doubles = False
        self.max_head_moves = {"white": 1, "black": 1}
        self.turn_started = False
        self.moves_count = 0
        self._tree = None
        self._nodes = []

    # ---- helpers -------------------------------------------------------------------------
    @property
    def is_doubles_roll(self):
        o = self.dice_state["original"]
        return len(o) == 2 and o[0] == o[1]

    def _pack(self, player_color):
        g = self.game
        return S.pack_states(np.asarray(g.board, dtype=np.int64), g.borne_off_white, g.borne_off_black,
                             _TURN[player_color], bool(g.first_turn_white), bool(g.first_turn_black))

    def _adopt(self, node):
        u = S.unpack_states(np.frombuffer(node.lo, dtype=np.uint8), np.frombuffer(node.hi, dtype=np.uint8))
        g = self.game
        board = u["board"][0].astype(np.int32)
        try:
            g.board[:] = board
        except Exception:
            g.board = board
        g.borne_off_white, g.borne_off_black = int(u["off_w"][0]), int(u["off_b"][0])
        g.first_turn_white, g.first_turn_black = bool(u["first_w"][0]), bool(u["first_b"][0])

    def _mover_board(self, player_color):
        b = np.asarray(self.game.board, dtype=np.int32)
        return b if player_color == "white" else rotate_board(b)

    def _refresh(self):
        self.valid_moves = self._tree.moves_from(self._nodes) if self._tree else []
        by = defaultdict(list)
        for m in self.valid_moves:
            by[m[0]].append(m)
        self.valid_moves_by_piece = dict(by)
        self.move_options = list(self.valid_moves)
        self.dice_state["remaining"] = list(self._nodes[0].remaining) if self._nodes else []

    def _borne_off(self):
        return {"white": self.game.borne_off_white, "black": self.game.borne_off_black}

    # ---- the reference surface -----------------------------------------------------------
    def roll_dice(self, player_color, dice=None):
        """my_game/narde_game_manager.py:127-188.  `dice` injects a roll (tests, replays); otherwise random.randint
        as in the reference.  Returns (dice sorted descending, valid_moves_by_piece)."""
        if player_color not in _TURN:
            raise ValueError("player_color must be 'white' or 'black'")
        dice = [random.randint(1, 6), random.randint(1, 6)] if dice is None else [int(dice[0]), int(dice[1])]
        if any(d < 1 or d > 6 for d in dice):
            raise ValueError("dice must be in 1..6")
        dice.sort(reverse=True)
        self.current_player = player_color
        first = self.game.first_turn_white if player_color == "white" else self.game.first_turn_black
        lo, hi = self._pack(player_color)
        self._tree = TurnTree(lo[0].tobytes(), hi[0].tobytes(), dice, bool(first), self._ops)
        self._nodes = [self._tree.root]
        expanded = [dice[0]] * 4 if dice[0] == dice[1] else list(dice)
        self.dice_state = {"original": list(dice), "expanded": expanded, "remaining": list(expanded), "used": []}
        self.is_first_turn_special_doubles = self._tree.max_head == 2
        self.max_head_moves[player_color] = self._tree.max_head
        self.head_move_made[player_color] = False
        self.head_moves_count[player_color] = 0
        self.moves_count = 0
        self.first_move_made = False
        self._save_state_for_undo()          # the position the dice were rolled on (undo_moves returns to it)
        self.turn_started = False
        self._refresh()
        return dice, self.valid_moves_by_piece

    def make_move(self, from_pos, to_pos, player_color):
        """my_game/narde_game_manager.py:190-234: one half-move; {'error': ...} and nothing changes if it is not
        legal at this point of the turn."""
        if self._tree is None or player_color != self.current_player:
            return {"error": "Roll the dice first"}
        to = "off" if to_pos in ("off", -1) else int(to_pos)
        from_pos = int(from_pos)
        board = self._mover_board(player_color)
        if not (0 <= from_pos < 24) or board[from_pos] <= 0:
            return {"error": f"No {player_color} piece at position {from_pos}"}
        if to != "off" and (not (0 <= to < 24) or board[to] < 0):
            return {"error": "Cannot land on opponent's piece"}
        if from_pos == HEAD and self._nodes[0].head_used >= self._tree.max_head:
            return {"error": "Only one checker may leave the head position per turn"}
        dist = from_pos - to if to != "off" else None
        if dist is not None and not any(dist in nd.remaining for nd in self._nodes):
            return {"error": f"No die with value {dist} available"}
        nxt = self._tree.children(self._nodes, from_pos, to)
        if not nxt:
            return {"error": "Invalid move"}
        self.turn_started = True
        self._nodes = nxt
        self._adopt(nxt[0])
        self.moves_count += 1
        self.dice_state["used"] = list(self.dice_state["expanded"])
        for d in nxt[0].remaining:
            self.dice_state["used"].remove(d)
        if from_pos == HEAD:
            self.head_move_made[player_color] = True
            self.head_moves_count[player_color] += 1
        self._refresh()
        if self.valid_moves:
            self.first_move_made = True
            return {"board": np.asarray(self.game.board).tolist(), "first_move_complete": True,
                    "needs_next_move": True, "valid_moves_by_piece": self.valid_moves_by_piece,
                    "dice_remaining": self.dice_state["remaining"], "move_number": self.moves_count,
                    "total_moves": 4 if self.is_doubles_roll else 2, "borne_off": self._borne_off()}
        return self._complete_turn(player_color)

    def _complete_turn(self, player_color):
        self.first_move_made = False
        self.moves_count = 0
        self.head_move_made[player_color] = False
        self._tree, self._nodes = None, []
        self.valid_moves, self.valid_moves_by_piece, self.move_options = [], {}, []
        self.turn_started = False
        return {"board": np.asarray(self.game.board).tolist(), "turn_complete": True, "borne_off": self._borne_off()}

    def _save_state_for_undo(self):
        self.saved_state = {"board": np.asarray(self.game.board).copy(), "borne_off_white": self.game.borne_off_white,
                            "borne_off_black": self.game.borne_off_black,
                            "first_turn_white": self.game.first_turn_white,
                            "first_turn_black": self.game.first_turn_black}
        self.turn_started = True

    def get_valid_moves_for_position(self, position, player_color):
        """my_game/narde_game_manager.py:745-796: destinations of the piece at `position`; 'off' is reported as -1."""
        if self._tree is None or player_color != self.current_player:
            return []
        return [-1 if m[1] == "off" else m[1] for m in self.valid_moves_by_piece.get(int(position), [])]

    def undo_moves(self):
        """my_game/narde_game_manager.py:798-831: back to the position the dice were rolled on."""
        if not self.saved_state or self._tree is None:
            return {"error": "No moves to undo"}
        self._nodes = [self._tree.root]
        self._adopt(self._tree.root)
        self.first_move_made = False
        self.moves_count = 0
        self.head_move_made = {"white": False, "black": False}
        self.head_moves_count[self.current_player] = 0
        self.dice_state["used"] = []
        self.turn_started = False
        self._refresh()
        return {"board": np.asarray(self.game.board).tolist(), "current_player": self.current_player,
                "dice": self.dice_state["original"], "valid_moves_by_piece": self.valid_moves_by_piece,
                "borne_off": self._borne_off()}

    def is_game_over(self):
        """my_game/narde_game_manager.py:844-855."""
        if self.game.borne_off_white == 15:
            return True, "white"
        if self.game.borne_off_black == 15:
            return True, "black"
        return False, None

    def get_game_state(self):
        """my_game/narde_game_manager.py:857-879."""
        over, winner = self.is_game_over()
        return {"board": np.asarray(self.game.board).tolist(), "current_player": self.current_player,
                "dice": self.dice_state["original"], "dice_remaining": self.dice_state["remaining"],
                "valid_moves_by_piece": self.valid_moves_by_piece, "first_move_made": self.first_move_made,
                "borne_off": self._borne_off(), "game_over": over, "winner": winner}

    # ---- the AI's turn (my_game/narde_game_manager.py:890-1099) ---------------------------------------
    def _q_values(self, ai_model, device):
        """Q-values [576] of the position the AI is looking at.  A DecomposedDQN-like torch module gets the env's
        observation as the reference passes it (narde_game_manager.py:1045-1050: FloatTensor(env._get_obs())[None]);
        an AfterstateMLP (the tcgen05 kernel, 198 inputs) gets the packed state of the turn's current node and encodes
        Box(198) itself."""
        import torch
        node = self._nodes[0]
        if hasattr(ai_model, "forward_states"):
            lo = torch.frombuffer(bytearray(node.lo), dtype=torch.uint8).reshape(1, 16).cuda()
            hi = torch.frombuffer(bytearray(node.hi), dtype=torch.uint8).reshape(1, 16).cuda()
            return ai_model.forward_states(lo, hi)[0].float().cpu()
        env = getattr(self.env, "unwrapped", self.env)      # make() may have wrapped the env (TimeLimit)
        obs = torch.as_tensor(np.asarray(env._get_obs(), dtype=np.float32)).unsqueeze(0).to(device)
        with torch.no_grad():
            return ai_model.forward(obs)[0].float().cpu()

    def _select_ai_move(self, valid_moves, ai_model, device, move_index):
        """narde_game_manager.py:1027-1072: the legal half-move with the largest Q[from*24 + to] ('off' -> from*24);
        ties go to the first move in the offered order, as Python's max does."""
        q = self._q_values(ai_model, device)
        best, best_v = None, None
        for f, to in valid_moves:
            v = float(q[f * 24 + (0 if to == "off" else to)])
            if best is None or v > best_v:
                best, best_v = (f, to), v
        return best

    def execute_ai_moves(self, ai_model, device=None, dice=None):
        """my_game/narde_game_manager.py:890-939: BLACK's whole turn -- roll, then half-move by half-move the legal
        move the model scores highest, then either the game-over response or White's roll.  Same response keys.
        The half-moves offered at every point are those of the turn tree (a legal end of turn stays reachable), so the
        AI cannot strand a die the way the reference's per-move filter can.  `dice` injects the AI's roll (tests)."""
        dice_rolled, _ = self.roll_dice("black", dice)
        ai_dice = list(self.dice_state["original"])
        if not self.valid_moves:
            self._tree, self._nodes = None, []
            self.set_current_player("white")
            d, by_piece = self.roll_dice("white")
            return {"board": np.asarray(self.game.board).tolist(), "current_player": "white", "dice": d,
                    "valid_moves_by_piece": by_piece, "ai_had_no_moves": True, "borne_off": self._borne_off()}
        ai_moves, moved_from_head, k = [], False, 0
        while self._tree is not None and self.valid_moves:
            f, to = self._select_ai_move(self.valid_moves, ai_model, device, k)
            r = self.make_move(f, to, "black")
            if "error" in r:                      # cannot happen: the move came from the offered list
                raise RuntimeError("AI move rejected: %s" % r["error"])
            ai_moves.append({"from": f, "to": -1 if to == "off" else to})
            moved_from_head = moved_from_head or f == HEAD
            k += 1
        over, winner = self.is_game_over()
        if over:
            return {"board": np.asarray(self.game.board).tolist(), "game_over": True, "winner": winner,
                    "ai_moves": ai_moves, "ai_moved_from_head": moved_from_head, "ai_dice": ai_dice,
                    "borne_off": self._borne_off()}
        self.set_current_player("white")
        d, by_piece = self.roll_dice("white")
        return {"board": np.asarray(self.game.board).tolist(), "current_player": "white", "dice": d,
                "valid_moves_by_piece": by_piece, "ai_moves": ai_moves, "ai_moved_from_head": moved_from_head,
                "ai_dice": ai_dice, "borne_off": self._borne_off()}

    def set_current_player(self, player_color):
        """my_game/narde_game_manager.py:881-888."""
        if player_color not in _TURN:
            raise ValueError("player_color must be 'white' or 'black'")
        self.current_player = player_color
