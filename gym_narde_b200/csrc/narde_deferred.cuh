// narde_deferred.cuh -- exact doubles turns when the 6-prime block rule bites (CTA per env).
//
// The block kernel (narde_block.cuh) hands over the rare doubles turns (~0.2% of env steps) in
// which some board of the move tree violates the block rule (narde.py:139-184): there the set of
// playable multisets depends on the ORDER of the half-moves.  Here one CTA owns one such
// environment and does what the rule text says, level by level: the multisets of j sources that
// are reachable by j legal half-moves through legal boards are expanded to level j+1 by every
// legal half-move (narde.py:64-89).  The board depends only on the multiset, so a level is a SET
// of multisets, kept as a dense bitmap indexed by the multiset's canonical key (base-24 digits
// 23 - source, ascending): atomicOr is the de-duplication, the bitmap is already in canonical
// order, and its population count is the number of legal turns.  The deepest non-empty level
// (<= 4) is the answer (max-dice rule).  Only the first `cap` actions and the chosen one are
// materialised (rank -> multiset by a select on the bitmap), with the representative ordering =
// the first legal ordering when higher sources are tried first.
//
// Written as phases like narde_block.cuh; the host harness emulates the CTA.
#pragma once
#include "narde_block.cuh"

namespace narde {

// Multisets of j sources are sorted tuples of digits x = 23 - source (ascending), i.e. j-combinations
// with repetition of 24 values: C(24+j-1, j) = 24, 300, 2600, 17550 for j = 1..4.  Their rank in
// lexicographic order (combinatorial number system) is the bitmap index, so the level-4 set takes
// 549 words instead of 24^4 / 32, and a CTA needs ~16 KB of shared memory.
constexpr int kDefCap = 2600;                     // multisets per expandable level (<= C(26,3))
constexpr int kDefBmWords = (17550 + 31) / 32;    // 549

NHD uint32_t sm_fetch_add(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, v);
#else
  uint32_t o = *p;
  *p += v;
  return o;
#endif
}
NHD uint32_t sm_fetch_or(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
  return atomicOr(p, v);
#else
  uint32_t o = *p;
  *p |= v;
  return o;
#endif
}
NHD void sm_min(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
  atomicMin(p, v);
#else
  if (v < *p) *p = v;
#endif
}

// binomials C(v, i), i = 1..4 (v < 32)
NHD uint32_t binom(uint32_t v, int i) {
  switch (i) {
    case 1: return v;
    case 2: return v * (v - 1u) / 2u;
    case 3: return v < 2u ? 0u : v * (v - 1u) * (v - 2u) / 6u;
    default: return v < 3u ? 0u : v * (v - 1u) * (v - 2u) * (v - 3u) / 24u;
  }
}
NHD uint32_t multiset_total(int j) { return binom((uint32_t)(24 + j - 1), j); }

template <int BLK>
struct DeferredSharedT {
  State st;
  uint32_t own, opp, nlo0, nlo1, nhi;
  int32_t d, H, player, d1, d2;
  uint32_t rnd;
  uint32_t n_cur, n_next, depth, any4, which;  // which: 0 -> current level list in a[], 1 -> in b[]
  uint32_t count, idx;
  uint64_t chosen;
  uint32_t part[BLK], base[BLK], part2[33];
  uint32_t bin[4][32];                         // bin[k][v] = C(v, k + 1): the rank <-> multiset arithmetic by lookup
  uint32_t bm[kDefBmWords];
  uint16_t a[kDefCap], b[kDefCap];             // level lists (levels 1..3: codes of <= 15 bits)
};

template <int BLK>
struct DeferredStep {
  typedef DeferredSharedT<BLK> Sh;
  static constexpr int WARPS = BLK / 32;

  static NHD Pos base_pos(const Sh& sh) {
    Pos P;
    P.lo = (uint64_t)sh.nlo0 | ((uint64_t)sh.nlo1 << 32);
    P.hi = sh.nhi;
    P.own = sh.own;
    P.opp = sh.opp;
    return P;
  }
  // code of a level-j multiset: j digits of 5 bits, digit = 23 - source, ascending digits (= sources
  // descending), first digit most significant -> numeric order = canonical key order
  static NHD void unpack(uint32_t code, int j, int* src) {
    for (int i = 0; i < j; i++) src[i] = 23 - (int)((code >> (5 * (j - 1 - i))) & 31u);
  }
  static NHD uint32_t insert(uint32_t code, int j, int s) {  // level j -> j+1
    uint32_t x = (uint32_t)(23 - s), out = 0;
    bool placed = false;
    for (int i = 0; i < j; i++) {
      uint32_t dg = (code >> (5 * (j - 1 - i))) & 31u;
      if (!placed && x <= dg) {
        out = (out << 5) | x;
        placed = true;
      }
      out = (out << 5) | dg;
    }
    if (!placed) out = (out << 5) | x;
    return out;
  }
  // lexicographic rank of the sorted digit tuple among all j-multisets of 24 values: with
  // y_i = x_i + i (strictly increasing, < n = 24 + j - 1) and z_m = n - 1 - y_{j-1-m},
  // rank = C(n, j) - 1 - sum_m C(z_m, m + 1)
  static NHD uint32_t rank_of(const Sh& sh, uint32_t code, int j) {
    const uint32_t n = (uint32_t)(24 + j - 1);
    uint32_t colex = 0;
    for (int m = 0; m < j; m++) {
      int i = j - 1 - m;  // digit index, most significant first
      uint32_t y = ((code >> (5 * (j - 1 - i))) & 31u) + (uint32_t)i;
      colex += sh.bin[m][n - 1u - y];
    }
    return multiset_total(j) - 1u - colex;
  }
  static NHD uint32_t code_of_rank(const Sh& sh, uint32_t r, int j) {
    const uint32_t n = (uint32_t)(24 + j - 1);
    uint32_t c = multiset_total(j) - 1u - r, code = 0;
    for (int m = j - 1; m >= 0; m--) {  // greedy: largest z_m with C(z_m, m+1) <= c
      uint32_t v = (uint32_t)m, hi = n - 1u;  // binom(., m+1) is non-decreasing: binary search
      while (v < hi) {
        uint32_t mid = (v + hi + 1u) >> 1;
        if (sh.bin[m][mid] <= c)
          v = mid;
        else
          hi = mid - 1u;
      }
      c -= sh.bin[m][v];
      int i = j - 1 - m;                      // z_m belongs to digit i
      uint32_t x = (n - 1u - v) - (uint32_t)i;
      code |= x << (5 * (j - 1 - i));
    }
    return code;
  }
  static NHD uint32_t bm_words(int level) { return (multiset_total(level) + 31u) >> 5; }

  // ---- phase 0 (one thread): load, dice, decode ------------------------------------------
  static NHD void ph_init(int tid, Sh& sh, const State& s_in, int64_t i, const StepFullArgs& A) {
    static_assert(BLK >= 128, "the lookup table is filled by 128 threads");
    if (tid >= BLK - 128) {
      const int t = tid - (BLK - 128);
      sh.bin[t >> 5][t & 31] = binom((uint32_t)(t & 31), (t >> 5) + 1);
    }
    if (tid != 0) return;
    sh.st = s_in;
    uint32_t env = (uint32_t)(A.env_base + i);
    U4 rnd = turn_random(A.seed, env, A.step);
    int d1, d2;
    if (A.dice_in) {
      d1 = A.dice_in[2 * i];
      d2 = A.dice_in[2 * i + 1];
    } else {
      d1 = die_from_word(rnd.x);
      d2 = die_from_word(rnd.y);
    }
    sh.rnd = rnd.z;
    sh.d1 = d1;
    sh.d2 = d2;
    int player = s_in.turn();
    bool first_turn = (s_in.flags() & (player == 1 ? FLAG_FIRST_W : FLAG_FIRST_B)) != 0;
    Pos P = decode_pos(s_in, player);
    sh.player = player;
    sh.d = d1;  // doubles: d1 == d2
    sh.H = (first_turn && (d1 == 3 || d1 == 4 || d1 == 6)) ? 2 : 1;
    sh.own = P.own;
    sh.opp = P.opp;
    sh.nlo0 = (uint32_t)P.lo;
    sh.nlo1 = (uint32_t)(P.lo >> 32);
    sh.nhi = P.hi;
    sh.a[0] = 0;  // level 0: the empty multiset
    sh.n_cur = 1;
    sh.n_next = 0;
    sh.depth = 0;
    sh.any4 = 0;
    sh.which = 0;
    sh.count = 0;
    sh.idx = 0;
    sh.chosen = ACT_EMPTY;
  }
  // ---- clear the bitmap of `level` -----------------------------------------------------------
  static NHD void ph_clear(int tid, Sh& sh, int level) {
    uint32_t nw = bm_words(level);
    for (uint32_t k = (uint32_t)tid; k < nw; k += BLK) sh.bm[k] = 0;
    if (tid == 0) sh.n_next = 0;
  }
  // one legal half-move s from the level-j node (code, board P): test the after-board, mark the child
  static NHD void visit_child(Sh& sh, uint16_t* nxt, uint32_t code, int j, const Pos& P, int s, bool risky) {
    if (risky && violates_block(after_mask(P, s, s - sh.d), P.opp)) return;  // narde.py:78-89
    uint32_t child = insert(code, j, s);
    uint32_t r = rank_of(sh, child, j + 1);
    uint32_t bit = 1u << (r & 31u);
    uint32_t old = sm_fetch_or(&sh.bm[r >> 5], bit);
    if (old & bit) return;  // this multiset was already reached through another ordering
    if (j + 1 < 4) {
      nxt[sm_fetch_add(&sh.n_next, 1u)] = (uint16_t)child;  // distinct multisets: cannot exceed kDefCap
    } else {
      sh.any4 = 1u;
    }
  }
  static NHD uint32_t node_board(const Sh& sh, uint32_t code, int j, Pos* Pout, uint32_t* risky) {
    int src[4];
    unpack(code, j, src);
    Pos P = base_pos(sh);
    int heads = 0;
    for (int k = 0; k < j; k++) {  // sources descending: arrivals precede departures
      P.move(src[k], src[k] - sh.d);
      heads += src[k] == 23;
    }
    uint32_t m = cand_mask(P.own, P.opp, sh.d, heads < sh.H);  // narde.py:64-77
    *risky = violates_block(P.own, P.opp) ? m : (m & (completing_points(P.own, P.opp) << sh.d));
    *Pout = P;
    return m;
  }
  // ---- expand level `level-1` -> `level`: narrow levels one warp per node (lane = source point),
  // wide levels one thread per node ---------------------------------------------------------------
  static NHD void ph_expand(int tid, Sh& sh, int level) {
    const uint16_t* cur = sh.which ? sh.b : sh.a;
    uint16_t* nxt = sh.which ? sh.a : sh.b;
    const int j = level - 1;
    const uint32_t n_cur = sh.n_cur;
    Pos P;
    uint32_t risky;
    // narrow level: node = warp, one candidate source per lane; wide level: node = thread, all candidates
    const bool narrow = n_cur <= (uint32_t)(4 * WARPS);
    const int lane = tid & 31;
    if (narrow && lane >= 24) return;
    const uint32_t start = narrow ? (uint32_t)(tid >> 5) : (uint32_t)tid, stride = narrow ? (uint32_t)WARPS : (uint32_t)BLK;
    const uint32_t lane_mask = narrow ? (1u << lane) : 0xFFFFFFu;
#pragma unroll 1
    for (uint32_t pi = start; pi < n_cur; pi += stride) {
      uint32_t code = cur[pi];
      uint32_t m = node_board(sh, code, j, &P, &risky) & lane_mask;
#pragma unroll 1
      for (; m; m &= m - 1) {
        int s = ctz32(m);
        visit_child(sh, nxt, code, j, P, s, ((risky >> s) & 1u) != 0);
      }
    }
  }
  static NHD bool level_found(const Sh& sh, int level) { return level < 4 ? sh.n_next > 0 : sh.any4 != 0; }
  static NHD void ph_advance(int tid, Sh& sh, int level) {
    if (tid != 0) return;
    if (!level_found(sh, level)) return;
    sh.depth = (uint32_t)level;
    if (level < 4) {
      sh.which ^= 1u;
      sh.n_cur = sh.n_next;
    }
  }
  // the search stopped below level 4: the bitmap now belongs to the (empty, all-zero) next level;
  // set the bits of the deepest level from its list
  static NHD void ph_rebuild(int tid, Sh& sh) {
    const uint16_t* cur = sh.which ? sh.b : sh.a;
    int j = (int)sh.depth;
    for (uint32_t k = (uint32_t)tid; k < sh.n_cur; k += BLK) {
      uint32_t r = rank_of(sh, cur[k], j);
      sm_fetch_or(&sh.bm[r >> 5], 1u << (r & 31u));
    }
  }
  // ---- count the final level: per-thread word ranges, block exclusive scan ---------------------
  static NHD void bm_range(const Sh& sh, int tid, uint32_t* w0, uint32_t* w1) {
    uint32_t nw = bm_words((int)sh.depth);
    uint32_t per = (nw + BLK - 1) / BLK;
    uint32_t lo = (uint32_t)tid * per, hi = lo + per;
    *w0 = lo < nw ? lo : nw;
    *w1 = hi < nw ? hi : nw;
  }
  static NHD void ph_bm_count(int tid, Sh& sh) {
    uint32_t w0, w1, c = 0;
    bm_range(sh, tid, &w0, &w1);
    for (uint32_t w = w0; w < w1; w++) c += (uint32_t)popc32(sh.bm[w]);
    sh.part[tid] = c;
  }
  static NHD void ph_bm_scan1(int tid, Sh& sh) {
    if (tid < 32) {
      uint32_t r = 0;
      for (int k = 0; k < BLK / 32; k++) r += sh.part[tid * (BLK / 32) + k];
      sh.part2[tid] = r;
    }
  }
  static NHD void ph_bm_scan2(int tid, Sh& sh) {
    if (tid != 0) return;
    uint32_t r = 0;
    for (int k = 0; k < 32; k++) {
      uint32_t t = sh.part2[k];
      sh.part2[k] = r;
      r += t;
    }
    sh.part2[32] = r;
  }
  // base[tid] = rank of the first multiset in this thread's word range; also the pick (thread 0)
  static NHD void ph_bm_scan3(int tid, Sh& sh, int64_t i, const StepFullArgs& A) {
    int g = tid / (BLK / 32);
    uint32_t r = sh.part2[g];
    for (int k = g * (BLK / 32); k < tid; k++) r += sh.part[k];
    sh.base[tid] = r;
    if (tid == 0) {
      uint32_t count = sh.depth ? sh.part2[32] : 0u;
      sh.count = count;
      sh.idx = pick_action_index(A, i, sh.rnd, count);
    }
  }
  // the multiset of canonical rank r (r < count): select on the bitmap
  static NHD_NOINLINE uint32_t select_code(const Sh& sh, uint32_t r) {
    int lo = 0, hi = BLK - 1;
    while (lo < hi) {  // last thread range whose first rank is <= r
      int mid = (lo + hi + 1) >> 1;
      if (sh.base[mid] <= r)
        lo = mid;
      else
        hi = mid - 1;
    }
    uint32_t w0, w1, k = r - sh.base[lo];
    bm_range(sh, lo, &w0, &w1);
    uint32_t w = w0, word = sh.bm[w];
    for (;;) {
      uint32_t c = (uint32_t)popc32(word);
      if (k < c) break;
      k -= c;
      word = sh.bm[++w];
    }
    for (; k; k--) word &= word - 1u;
    return code_of_rank(sh, (w << 5) + (uint32_t)ctz32(word), (int)sh.depth);
  }
  // ---- materialise the first `cap` actions of the canonical list and the chosen one -------------
  // Representative ordering = the lexicographically first legal ordering of the (descending) sources,
  // higher sources tried first: highest-source-first itself in the common case, else a depth-first search.
  static NHD uint32_t emit_total(const Sh& sh, const StepFullArgs& A) {  // ranks to materialise (+1: the chosen one)
    uint32_t n = sh.count;
    uint32_t lim = A.actions ? (n < (uint32_t)A.cap ? n : (uint32_t)A.cap) : 0u;
    return lim + ((n && sh.idx >= lim) ? 1u : 0u);
  }
  static NHD uint32_t emit_rank(const Sh& sh, const StepFullArgs& A, uint32_t k) {
    uint32_t n = sh.count;
    uint32_t lim = A.actions ? (n < (uint32_t)A.cap ? n : (uint32_t)A.cap) : 0u;
    return k < lim ? k : sh.idx;
  }
  static NHD_NOINLINE bool sequence_legal(const Sh& sh, const int* order, int j) {
    Pos P = base_pos(sh);
    int heads = 0;
    for (int t = 0; t < j; t++) {
      int s = order[t];
      uint32_t m = cand_mask(P.own, P.opp, sh.d, heads < sh.H);
      if (!((m >> s) & 1u)) return false;
      P.move(s, s - sh.d);
      heads += s == 23;
      if (violates_block(P.own, P.opp)) return false;
    }
    return true;
  }
  static NHD void emit_store(Sh& sh, int64_t i, const StepFullArgs& A, uint32_t k, const int* order, int j) {
    uint64_t* slice = A.actions ? A.actions + i * (int64_t)A.cap : nullptr;
    uint32_t n = sh.count;
    uint32_t lim = slice ? (n < (uint32_t)A.cap ? n : (uint32_t)A.cap) : 0u;
    uint64_t act = ACT_EMPTY;
    for (int t = 0; t < j; t++) act = act_set(act, t, order[t], order[t] - sh.d);
    if (k < lim) slice[k] = act;
    if (emit_rank(sh, A, k) == sh.idx) sh.chosen = act;
  }
  // the lexicographically first legal ordering of a multiset that cannot be played highest-source-first:
  // depth-first over the <= 4! orderings, higher sources tried first (dbl_order_search's order and its rule
  // for equal sources), as a compile-time recursion with the sources and the ordering packed 5 bits each, so
  // that the positions of the path stay in registers instead of a local-memory stack
  template <int D>
  static NHD bool order_dfs(const Sh& sh, const Pos& P, int heads, uint32_t srcp, int j, uint32_t used, uint32_t* orderp) {
    if constexpr (D >= 4) {
      return true;
    } else {
      const uint32_t m = cand_mask(P.own, P.opp, sh.d, heads < sh.H);
      for (int c = 0; c < j; c++) {
        if ((used >> c) & 1u) continue;
        const int s = (int)((srcp >> (5 * c)) & 31u);
        bool same = false;  // an equal source already tried at this depth is the same move
        for (int e = 0; e < c; e++)
          if (!((used >> e) & 1u) && (int)((srcp >> (5 * e)) & 31u) == s) same = true;
        if (same || !((m >> s) & 1u)) continue;
        Pos nx = P;
        nx.move(s, s - sh.d);
        if (violates_block(nx.own, nx.opp)) continue;
        *orderp = (*orderp & ~(31u << (5 * D))) | ((uint32_t)s << (5 * D));
        if (D + 1 == j) return true;
        if (order_dfs<D + 1>(sh, nx, heads + (s == 23), srcp, j, used | (1u << c), orderp)) return true;
      }
      return false;
    }
  }
  static NHD_NOINLINE void first_legal_order(const Sh& sh, const int* src, int j, int* order) {
    uint32_t srcp = 0, orderp = 0;
    for (int t = 0; t < j; t++) srcp |= (uint32_t)src[t] << (5 * t);
    if (!order_dfs<0>(sh, base_pos(sh), 0, srcp, j, 0u, &orderp)) orderp = srcp;  // unreachable: the multiset was reached legally
    for (int t = 0; t < j; t++) order[t] = (int)((orderp >> (5 * t)) & 31u);
  }
  // one thread per emitted rank: rank -> multiset -> representative ordering -> store
  static NHD void ph_emit(int tid, Sh& sh, int64_t i, const StepFullArgs& A) {
    const uint32_t total = emit_total(sh, A);
    const int j = (int)sh.depth;
    for (uint32_t k = (uint32_t)tid; k < total; k += BLK) {
      uint32_t code = select_code(sh, emit_rank(sh, A, k));
      int src[4], order[4];
      unpack(code, j, src);
      if (sequence_legal(sh, src, j)) {
          emit_store(sh, i, A, k, src, j);
      } else {
          first_legal_order(sh, src, j, order);
        emit_store(sh, i, A, k, order, j);
      }
    }
  }
};

}  // namespace narde
