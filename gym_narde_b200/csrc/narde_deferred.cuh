// narde_deferred.cuh -- exact doubles turns when the 6-prime block rule bites (CTA per env).
//
// The block kernel (narde_block.cuh) hands over the rare doubles turns (~0.05% of env steps) in
// which some board of the move tree violates the block rule (narde.py:139-184): there the set of
// playable multisets depends on the ORDER of the half-moves.  Here one CTA owns one such
// environment and does what the rule text says, level by level: the multisets of j sources that
// are reachable by j legal half-moves through legal boards are expanded to level j+1 by every
// legal half-move (narde.py:64-89), de-duplicated in a shared-memory hash set (the board depends
// only on the multiset).  The deepest non-empty level (<= 4) is the answer (max-dice rule),
// sorted by the canonical key; representative orderings come from dbl_order_search.
//
// Written as phases like narde_block.cuh; the host harness emulates the CTA.
#pragma once
#include "narde_block.cuh"

namespace narde {

constexpr int kDefCap = 4096;    // multisets per level (observed maximum ~1500)
constexpr int kDefHash = 8192;   // open addressing, power of two

NHD uint32_t sm_cas(uint32_t* p, uint32_t expect, uint32_t val) {
#if defined(__CUDA_ARCH__)
  return atomicCAS(p, expect, val);
#else
  uint32_t o = *p;
  if (o == expect) *p = val;
  return o;
#endif
}
NHD uint32_t sm_fetch_add(uint32_t* p, uint32_t v) {
#if defined(__CUDA_ARCH__)
  return atomicAdd(p, v);
#else
  uint32_t o = *p;
  *p += v;
  return o;
#endif
}

struct DeferredShared {
  State st;
  uint32_t own, opp, nlo0, nlo1, nhi;
  int32_t d, H, player, d1, d2;
  uint32_t rnd;
  uint32_t n_cur, n_next, depth, overflow, which;  // which: 0 -> level list in a[], 1 -> in b[]
  uint32_t count, idx;
  uint64_t chosen;
  uint32_t part[1024], base[1024], part2[32];
  // a | b | hash are contiguous: once the search is over, b..hash is reused as a 24^4-bit bitmap
  uint32_t a[kDefCap], b[kDefCap];
  uint32_t hash[kDefHash];
};
static_assert(kDefCap + kDefHash >= (24 * 24 * 24 * 24 + 31) / 32, "bitmap must fit in b + hash");

template <int BLK>
struct DeferredStep {
  typedef DeferredShared Sh;

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

  // ---- phase 0 (one thread): load, dice, decode ------------------------------------------
  static NHD void ph_init(int tid, Sh& sh, const State& s_in, int64_t i, const StepFullArgs& A) {
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
    sh.overflow = 0;
    sh.which = 0;
    sh.count = 0;
    sh.idx = 0;
    sh.chosen = ACT_EMPTY;
  }
  static NHD void ph_clear(int tid, Sh& sh) {
    for (int k = tid; k < kDefHash; k += BLK) sh.hash[k] = 0;
    if (tid == 0) sh.n_next = 0;
  }
  // ---- expand level `level-1` -> `level` ---------------------------------------------------
  static NHD void ph_expand(int tid, Sh& sh, int level) {
    const uint32_t* cur = sh.which ? sh.b : sh.a;
    uint32_t* nxt = sh.which ? sh.a : sh.b;
    const int j = level - 1;
    const int d = sh.d, H = sh.H;
    for (uint32_t pi = (uint32_t)tid; pi < sh.n_cur; pi += BLK) {
      uint32_t code = cur[pi];
      int src[4];
      unpack(code, j, src);
      Pos P = base_pos(sh);
      int heads = 0;
      for (int k = 0; k < j; k++) {  // sources descending: arrivals precede departures
        P.move(src[k], src[k] - d);
        heads += src[k] == 23;
      }
      uint32_t m = cand_mask(P.own, P.opp, d, heads < H);  // narde.py:64-77
      if (m == 0) continue;
      uint32_t risky = violates_block(P.own, P.opp) ? m : (m & (completing_points(P.own, P.opp) << d));
      for (; m; m &= m - 1) {
        int s = ctz32(m);
        if (((risky >> s) & 1u) && violates_block(after_mask(P, s, s - d), P.opp)) continue;  // narde.py:78-89
        uint32_t child = insert(code, j, s);
        uint32_t key = child + 1u;  // 0 = empty slot
        uint32_t h = (child * 2654435761u) >> (32 - 13);
        for (;;) {
          uint32_t o = sm_cas(&sh.hash[h], 0u, key);
          if (o == 0u) {  // new multiset
            uint32_t at = sm_fetch_add(&sh.n_next, 1u);
            if (at < (uint32_t)kDefCap)
              nxt[at] = child;
            else
              sh.overflow = 1;
            break;
          }
          if (o == key) break;  // already present
          h = (h + 1) & (kDefHash - 1);
        }
      }
    }
  }
  // returns true when the level is non-empty and the search continues
  static NHD void ph_advance(int tid, Sh& sh, int level) {
    if (tid != 0) return;
    if (sh.n_next > 0) {
      sh.which ^= 1u;
      sh.n_cur = sh.n_next < (uint32_t)kDefCap ? sh.n_next : (uint32_t)kDefCap;
      sh.depth = (uint32_t)level;
    }
  }
  // ---- sort the final level ascending: set one bit per multiset in a dense bitmap (index = the
  // base-24 value of the digits, same order as the key), then compact the bitmap in order --------
  static NHD uint32_t bm_words(const Sh& sh) {
    uint32_t bits = 1;
    for (uint32_t k = 0; k < sh.depth; k++) bits *= 24u;
    return (bits + 31u) >> 5;
  }
  static NHD void ph_to_a(int tid, Sh& sh) {  // the final list must live in a[] (b..hash become the bitmap)
    if (sh.which)
      for (uint32_t k = (uint32_t)tid; k < sh.n_cur; k += BLK) sh.a[k] = sh.b[k];
  }
  static NHD void ph_bm_clear(int tid, Sh& sh) {
    uint32_t* bm = sh.b;
    uint32_t nw = bm_words(sh);
    for (uint32_t k = (uint32_t)tid; k < nw; k += BLK) bm[k] = 0;
    if (tid == 0) sh.which = 0;
  }
  static NHD void ph_bm_set(int tid, Sh& sh) {
    uint32_t* bm = sh.b;
    int j = (int)sh.depth;
    for (uint32_t k = (uint32_t)tid; k < sh.n_cur; k += BLK) {
      uint32_t code = sh.a[k], dense = 0;
      for (int i = 0; i < j; i++) dense = dense * 24u + ((code >> (5 * (j - 1 - i))) & 31u);
#if defined(__CUDA_ARCH__)
      atomicOr(&bm[dense >> 5], 1u << (dense & 31u));
#else
      bm[dense >> 5] |= 1u << (dense & 31u);
#endif
    }
  }
  static NHD void bm_range(const Sh& sh, int tid, uint32_t* w0, uint32_t* w1) {
    uint32_t nw = bm_words(sh);
    uint32_t per = (nw + BLK - 1) / BLK;
    uint32_t lo = (uint32_t)tid * per, hi = lo + per;
    *w0 = lo < nw ? lo : nw;
    *w1 = hi < nw ? hi : nw;
  }
  static NHD void ph_bm_count(int tid, Sh& sh) {
    const uint32_t* bm = sh.b;
    uint32_t w0, w1, c = 0;
    bm_range(sh, tid, &w0, &w1);
    for (uint32_t w = w0; w < w1; w++) c += (uint32_t)popc32(bm[w]);
    sh.part[tid] = c;
  }
  // block exclusive scan of part[0..BLK) in three short phases (part2 = per-32-group sums)
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
  }
  static NHD void ph_bm_scan3(int tid, Sh& sh) {
    int g = tid / (BLK / 32);
    uint32_t r = sh.part2[g];
    for (int k = g * (BLK / 32); k < tid; k++) r += sh.part[k];
    sh.base[tid] = r;
  }
  static NHD void ph_bm_emit(int tid, Sh& sh) {
    const uint32_t* bm = sh.b;
    int j = (int)sh.depth;
    uint32_t w0, w1, at = sh.base[tid];
    bm_range(sh, tid, &w0, &w1);
    for (uint32_t w = w0; w < w1; w++) {
      for (uint32_t m = bm[w]; m; m &= m - 1) {
        uint32_t dense = (w << 5) + (uint32_t)ctz32(m), code = 0;
        for (int i = 0; i < j; i++) {  // base-24 digits, least significant first
          code |= (dense % 24u) << (5 * i);
          dense /= 24u;
        }
        sh.a[at++] = code;
      }
    }
  }
  static NHD uint64_t representative(const Sh& sh, uint32_t code) {
    int src[4], order[4];
    int j = (int)sh.depth;
    unpack(code, j, src);
    Pos P = base_pos(sh);
    uint64_t act = ACT_EMPTY;
    if (dbl_order_search(P, src, j, sh.d, sh.H, order))
      for (int k = 0; k < j; k++) act = act_set(act, k, order[k], order[k] - sh.d);
    return act;
  }
  // ---- pick, write the canonical list -------------------------------------------------------
  static NHD void ph_pick(int tid, Sh& sh, int64_t i, const StepFullArgs& A) {
    if (tid != 0) return;
    uint32_t count = sh.depth ? sh.n_cur : 0u;
    sh.count = count;
    sh.idx = pick_action_index(A, i, sh.rnd, count);
  }
  // safety net (a level overflowed kDefCap; never observed): one thread walks the tree exactly
  static NHD void ph_fallback(int tid, Sh& sh, int64_t i, const StepFullArgs& A) {
    if (tid != 0) return;
    Pos P = base_pos(sh);
    bool ft = sh.H == 2;  // H == 2 only on a first turn (and then d is 3, 4 or 6)
    uint64_t* slice = A.actions ? A.actions + i * (int64_t)A.cap : nullptr;
    StoreSink sk = {slice, slice ? A.cap : 0, 1, 0};
    uint32_t count = (uint32_t)enumerate_turn(P, sh.d, sh.d, ft, sk);
    sh.count = count;
    sh.depth = 4;
    uint32_t idx = pick_action_index(A, i, sh.rnd, count);
    if (count) {
      PickSink pk = {(int)idx, 0, ACT_EMPTY};
      enumerate_turn(P, sh.d, sh.d, ft, pk);
      sh.chosen = pk.picked;
    }
    sh.idx = idx;
  }
  static NHD void ph_emit(int tid, Sh& sh, int64_t i, const StepFullArgs& A) {
    const uint32_t* cur = sh.which ? sh.b : sh.a;
    uint64_t* slice = A.actions ? A.actions + i * (int64_t)A.cap : nullptr;
    uint32_t n = sh.count;
    uint32_t lim = slice ? (n < (uint32_t)A.cap ? n : (uint32_t)A.cap) : 0u;
    for (uint32_t k = (uint32_t)tid; k < lim; k += BLK) {
      uint64_t act = representative(sh, cur[k]);
      slice[k] = act;
      if (k == sh.idx) sh.chosen = act;
    }
    if (tid == 0 && n && sh.idx >= lim) sh.chosen = representative(sh, cur[sh.idx]);
  }
};

}  // namespace narde
