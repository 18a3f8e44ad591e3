// narde_block.cuh -- the fused full-rules step, CTA-cooperative version.
//
// One CTA of BLK threads owns BLK environments.  Instead of one thread walking one environment's
// whole move tree (badly divergent: doubles vs non-doubles, 1 vs 1000 legal turns), the CTA
// flattens the enumeration into uniform work items that are dealt out evenly to its threads:
//   non-doubles : item = (env, p)  = "source p plays the higher die"      (~9 per env)
//   doubles     : item = (env, s1) = "highest source played is s1"        (~4 per env)
// Items are counted, prefix-summed in shared memory, and every thread takes a contiguous chunk of
// ceil(items / BLK).  Legal-action counts per item are prefix-summed again to give every item its
// slot in the env's canonical action list, so the list is written in order without sorting.
// Environments for which the 6-prime block rule could matter (~4%) and doubles turns that cannot
// use all four dice (~3%) take the sequential per-thread path of narde_core.cuh.
//
// The code is written as PHASES: plain functions of (tid, shared block state) that only read what
// earlier phases wrote.  The kernel runs them with __syncthreads() in between; the test-only host
// harness runs "for tid in 0..BLK" per phase, which is the same thing, so this file is verified
// against the oracle on the CPU before it reaches the GPU.
#pragma once
#include "narde_env.cuh"

namespace narde {

enum : uint8_t { K_NONE = 0, K_DONE = 1, K_ND = 2, K_D = 3 };

template <int BLK>
struct BlockShared {
  State st[BLK];
  // per-env position in the mover frame
  uint32_t own[BLK], opp[BLK], ones[BLK];
  uint32_t nlo0[BLK], nlo1[BLK], nhi[BLK];
  uint32_t Ca[BLK], S[BLK], qhome[BLK];  // non-doubles: candidates of die a, die b; "q makes all home"
  uint32_t rnd[BLK];
  int32_t count[BLK];
  uint64_t chosen[BLK];
  uint8_t a[BLK], b[BLK], kind[BLK], first[BLK], d1[BLK], d2[BLK], blk[BLK];
  // item masks / bases: ND rows and doubles first sources
  uint32_t rowmask[BLK], dmask[BLK];
  uint32_t ibase[BLK + 1], dbase[BLK + 1];
  uint32_t pres[BLK * 24];   // ND: pairs (p, q) legal in some order, one mask per row
  uint8_t icnt[BLK * 24];    // ND: de-duplicated pairs per row
  uint32_t dcnt[BLK * 16];   // doubles: 4-move leaves under each first source
  uint32_t eG[BLK], eEnd[BLK];
  // scan scratch
  uint32_t partA[BLK], partB[BLK], baseA[BLK], baseB[BLK];
  uint32_t wsA[40], wsB[40];
};

struct ItemIter {  // items (env e, bit p) over masks[e], envs ascending, bits descending
  int e, j, j1;
  uint32_t rem;
  NHD void init(const uint32_t* masks, const uint32_t* base, int nenv, int j0, int j1_) {
    j = j0;
    j1 = j1_;
    e = 0;
    rem = 0;
    if (j >= j1) return;
    int lo = 0, hi = nenv - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if ((int)base[mid] <= j0)
        lo = mid;
      else
        hi = mid - 1;
    }
    e = lo;
    rem = masks[e];
    for (int k = j0 - (int)base[e]; k > 0; k--) rem &= ~(1u << fls32(rem));
  }
  NHD bool next(const uint32_t* masks, int* eo, int* po, bool* first, bool* last) {
    if (j >= j1) return false;
    while (rem == 0u) {
      e++;
      rem = masks[e];
    }
    *first = rem == masks[e];
    int p = fls32(rem);
    rem &= ~(1u << p);
    *last = rem == 0u;
    *eo = e;
    *po = p;
    j++;
    return true;
  }
};

template <int BLK>
struct BlockStep {
  typedef BlockShared<BLK> Sh;
  static constexpr int PER = BLK / 32;  // partial sums per scan lane

  static NHD Pos pos_of(const Sh& sh, int e) {
    Pos P;
    P.lo = (uint64_t)sh.nlo0[e] | ((uint64_t)sh.nlo1[e] << 32);
    P.hi = sh.nhi[e];
    P.own = sh.own[e];
    P.opp = sh.opp[e];
    return P;
  }
  static NHD int head_budget(const Sh& sh, int e) {
    int d = sh.a[e];
    return (sh.first[e] && (d == 3 || d == 4 || d == 6)) ? 2 : 1;
  }

  // ---- phase 1: load, dice, decode, classify, item masks --------------------------------
  static NHD void ph_load(int tid, Sh& sh, bool valid, const State& s_in, int64_t i, const StepFullArgs& A) {
    sh.partA[tid] = 0;
    sh.partB[tid] = 0;
    sh.rowmask[tid] = 0;
    sh.dmask[tid] = 0;
    sh.count[tid] = 0;
    sh.chosen[tid] = ACT_EMPTY;
    sh.eG[tid] = 0;
    sh.eEnd[tid] = 0;
    sh.kind[tid] = K_NONE;
    if (!valid) return;
    sh.st[tid] = s_in;
    if (s_in.flags() & FLAG_DONE) {
      sh.kind[tid] = K_DONE;
      return;
    }
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
    sh.rnd[tid] = rnd.z;
    sh.d1[tid] = (uint8_t)d1;
    sh.d2[tid] = (uint8_t)d2;
    int player = s_in.turn();
    bool first_turn = (s_in.flags() & (player == 1 ? FLAG_FIRST_W : FLAG_FIRST_B)) != 0;
    uint32_t ones;
    Pos P = decode_pos(s_in, player, &ones);
    int a = d1 > d2 ? d1 : d2, b = d1 > d2 ? d2 : d1;
    sh.a[tid] = (uint8_t)a;
    sh.b[tid] = (uint8_t)b;
    sh.first[tid] = first_turn ? 1 : 0;
    sh.own[tid] = P.own;
    sh.opp[tid] = P.opp;
    sh.ones[tid] = ones;
    sh.nlo0[tid] = (uint32_t)P.lo;
    sh.nlo1[tid] = (uint32_t)(P.lo >> 32);
    sh.nhi[tid] = P.hi;
    // the 6-prime block rule can only matter for ~4% of positions; those take the exact
    // (nibble-board) row arithmetic, everything else the mask-only fast path
    bool blk = !block_rule_irrelevant(P, d1, d2);
    sh.blk[tid] = blk ? 1 : 0;
    if (a != b) {
      sh.kind[tid] = K_ND;
      uint32_t Ca = cand_mask(P.own, P.opp, a, true);
      uint32_t S = cand_mask(P.own, P.opp, b, true);
      if (blk) {
        Ca = block_filter(P, Ca, a);
        S = block_filter(P, S, b);
      }
      uint32_t q_home = 0, outside = P.own & ~0x3Fu;
      if (outside == 0u) {
        q_home = S;
      } else if ((outside & (outside - 1u)) == 0u) {
        int qo = ctz32(outside);
        if ((ones & outside) && qo - b < 6) q_home = S & outside;
      }
      sh.Ca[tid] = Ca;
      sh.S[tid] = S;
      sh.qhome[tid] = q_home;
      uint32_t rows = (P.own | (S >> b)) & 0xFFFFFFu;
      sh.rowmask[tid] = rows;
      sh.partA[tid] = (uint32_t)popc32(rows);
    } else {
      sh.kind[tid] = K_D;
      uint32_t m = cand_mask(P.own, P.opp, a, true);
      sh.dmask[tid] = m;
      sh.partB[tid] = (uint32_t)popc32(m);
    }
  }

  // ---- block exclusive scan of partA / partB (3 phases) ----------------------------------
  static NHD void ph_scan1(int tid, Sh& sh) {
    if (tid < 32) {
      uint32_t sa = 0, sb = 0;
      for (int k = 0; k < PER; k++) {
        sa += sh.partA[tid * PER + k];
        sb += sh.partB[tid * PER + k];
      }
      sh.wsA[tid] = sa;
      sh.wsB[tid] = sb;
    }
  }
  static NHD void ph_scan2(int tid, Sh& sh) {
    if (tid == 0) {
      uint32_t ra = 0, rb = 0;
      for (int k = 0; k < 32; k++) {
        uint32_t ta = sh.wsA[k], tb = sh.wsB[k];
        sh.wsA[k] = ra;
        sh.wsB[k] = rb;
        ra += ta;
        rb += tb;
      }
      sh.wsA[32] = ra;
      sh.wsB[32] = rb;
    }
  }
  static NHD void ph_scan3(int tid, Sh& sh) {
    int g = tid / PER;
    uint32_t ra = sh.wsA[g], rb = sh.wsB[g];
    for (int k = g * PER; k < tid; k++) {
      ra += sh.partA[k];
      rb += sh.partB[k];
    }
    sh.baseA[tid] = ra;
    sh.baseB[tid] = rb;
  }
  // after the first scan: bases of the item lists
  static NHD void ph_item_bases(int tid, Sh& sh) {
    sh.ibase[tid] = sh.baseA[tid];
    sh.dbase[tid] = sh.baseB[tid];
    if (tid == 0) {
      sh.ibase[BLK] = sh.wsA[32];
      sh.dbase[BLK] = sh.wsB[32];
    }
  }
  static NHD void chunk(int total, int tid, int* j0, int* j1) {
    int c = (total + BLK - 1) / BLK;
    int a = tid * c, b = a + c;
    *j0 = a < total ? a : total;
    *j1 = b < total ? b : total;
  }

  // ---- non-doubles row arithmetic (masks only; the block rule is known not to matter) -----
  static NHD void nd_row(const Sh& sh, int e, int p, uint32_t* m1_out, uint32_t* m2_out) {
    uint32_t own = sh.own[e], opp = sh.opp[e], ones = sh.ones[e], S = sh.S[e];
    int a = sh.a[e], b = sh.b[e];
    uint32_t bp = 1u << p;
    int ta = p - a;
    uint32_t m1 = 0;
    if (sh.Ca[e] & bp) {  // p moves a first, then q moves b on the resulting board
      uint32_t own1 = (own & ~(ones & bp)) | (ta >= 0 ? (1u << ta) : 0u);
      m1 = cand_mask(own1, opp, b, p != 23);
    }
    uint32_t m2;  // q in S moves b first, then p moves a
    if (own & bp) {
      m2 = S & ~(ones & bp);
    } else {
      m2 = p + b < 24 ? (S & (1u << (p + b))) : 0u;
    }
    if (ta >= 0) {
      if ((opp >> ta) & 1u) m2 = 0;
    } else {
      m2 &= sh.qhome[e];
    }
    if (p == 23) m2 &= ~(1u << 23);
    if (sh.blk[e]) {  // exact path: every intermediate and final board is tested (narde.py:78-89)
      Pos P = pos_of(sh, e);
      m1 = 0;
      if (sh.Ca[e] & bp) {
        Pos P1 = P;
        P1.move(p, ta);
        m1 = block_filter(P1, cand_mask(P1.own, P1.opp, b, p != 23), b);
      }
      uint32_t keep = 0;
      for (uint32_t mm = m2 & ~m1; mm; mm &= mm - 1) {
        int q = ctz32(mm);
        Pos Q = P;
        Q.move(q, q - b);
        if (!violates_block(after_mask(Q, p, ta), Q.opp)) keep |= 1u << q;
      }
      m2 = keep | (m2 & m1);
    }
    *m1_out = m1;
    *m2_out = m2;
  }
  // pairs of row p that duplicate a pair of a higher row (see enum_nondouble)
  static NHD uint32_t nd_dups(const Sh& sh, int e, int p) {
    int a = sh.a[e], b = sh.b[e];
    uint32_t dup = 0;
    int x = p + b;
    if (x < 24 && x - a >= 0 && ((sh.pres[e * 24 + x] >> (x - a)) & 1u)) dup |= 1u << x;
    if (p < b)
      for (int q = p + 1; q < b; q++)
        if ((sh.pres[e * 24 + q] >> p) & 1u) dup |= 1u << q;
    return dup;
  }

  // ---- doubles sub-tree below (s1): count 4-move leaves --------------------------------
  static NHD uint32_t dbl_count_under(const Pos& P, int d, int H, int s1) {
    Pos P1 = P;
    P1.move(s1, s1 - d);
    int h1 = s1 == 23;
    uint32_t leaves = 0;
    uint32_t m2 = cand_mask(P1.own, P1.opp, d, h1 < H) & ((2u << s1) - 1u);
    while (m2) {
      int s2 = fls32(m2);
      m2 &= ~(1u << s2);
      Pos P2 = P1;
      P2.move(s2, s2 - d);
      int h2 = h1 + (s2 == 23);
      uint32_t m3 = cand_mask(P2.own, P2.opp, d, h2 < H) & ((2u << s2) - 1u);
      while (m3) {
        int s3 = fls32(m3);
        m3 &= ~(1u << s3);
        Pos P3 = P2;
        P3.move(s3, s3 - d);
        int h3 = h2 + (s3 == 23);
        leaves += (uint32_t)popc32(cand_mask(P3.own, P3.opp, d, h3 < H) & ((2u << s3) - 1u));
      }
    }
    return leaves;
  }
  // emit the leaves below s1 at list positions off, off+1, ...; capture the idx-th
  static NHD void dbl_emit_under(const Pos& P, int d, int H, int s1, uint32_t off, uint32_t cnt, uint64_t* slice, int cap,
                                 uint32_t idx, uint64_t* chosen) {
    if (!slice && !(idx >= off && idx < off + cnt)) return;
    Pos P1 = P;
    P1.move(s1, s1 - d);
    int h1 = s1 == 23;
    uint64_t a1 = act_set(ACT_EMPTY, 0, s1, s1 - d);
    uint32_t k = off;
    uint32_t m2 = cand_mask(P1.own, P1.opp, d, h1 < H) & ((2u << s1) - 1u);
    while (m2) {
      int s2 = fls32(m2);
      m2 &= ~(1u << s2);
      Pos P2 = P1;
      P2.move(s2, s2 - d);
      int h2 = h1 + (s2 == 23);
      uint64_t a2 = act_set(a1, 1, s2, s2 - d);
      uint32_t m3 = cand_mask(P2.own, P2.opp, d, h2 < H) & ((2u << s2) - 1u);
      while (m3) {
        int s3 = fls32(m3);
        m3 &= ~(1u << s3);
        Pos P3 = P2;
        P3.move(s3, s3 - d);
        int h3 = h2 + (s3 == 23);
        uint64_t a3 = act_set(a2, 2, s3, s3 - d);
        uint32_t m4 = cand_mask(P3.own, P3.opp, d, h3 < H) & ((2u << s3) - 1u);
        while (m4) {
          int s4 = fls32(m4);
          m4 &= ~(1u << s4);
          uint64_t a4 = act_set(a3, 3, s4, s4 - d);
          if (slice && (int)k < cap) slice[k] = a4;
          if (k == idx) *chosen = a4;
          k++;
        }
      }
    }
  }

  // exact (block-rule aware) variants: the sequential walker of narde_core.cuh restricted to
  // highest source s1
  struct OffsetSink {
    uint64_t* slice;
    int cap;
    uint32_t k, idx;
    uint64_t* chosen;
    NHD void operator()(uint64_t a) {
      if (slice && (int)k < cap) slice[k] = a;
      if (k == idx) *chosen = a;
      k++;
    }
  };
  static NHD uint32_t dbl_count_exact(const Pos& P, int d, int H, int s1) {
    DblCtx cx;
    cx.d = d;
    cx.H = H;
    cx.target = 4;
    cx.blockchk = true;
    cx.maxdepth = 0;
    cx.n = 0;
    cx.first_mask = 1u << s1;
    int src[4];
    CountSink ck;
    DblLevel<0, CountSink>::run(P, P, cx, 0, 23, true, true, ACT_EMPTY, src, ck);
    return (uint32_t)cx.n;
  }
  static NHD void dbl_emit_exact(const Pos& P, int d, int H, int s1, uint32_t off, uint32_t cnt, uint64_t* slice, int cap,
                                 uint32_t idx, uint64_t* chosen) {
    if (!slice && !(idx >= off && idx < off + cnt)) return;
    DblCtx cx;
    cx.d = d;
    cx.H = H;
    cx.target = 4;
    cx.blockchk = true;
    cx.maxdepth = 0;
    cx.n = 0;
    cx.first_mask = 1u << s1;
    int src[4];
    OffsetSink sk = {slice, cap, off, idx, chosen};
    DblLevel<0, OffsetSink>::run(P, P, cx, 0, 23, true, true, ACT_EMPTY, src, sk);
  }

  // ---- phase 3a: ND rows -> pres ; doubles items -> leaf counts --------------------------
  static NHD void ph_rows(int tid, Sh& sh) {
    int j0, j1, e, p;
    bool f, l;
    chunk((int)sh.ibase[BLK], tid, &j0, &j1);
    ItemIter it;
    it.init(sh.rowmask, sh.ibase, BLK, j0, j1);
    while (it.next(sh.rowmask, &e, &p, &f, &l)) {
      uint32_t m1, m2;
      nd_row(sh, e, p, &m1, &m2);
      sh.pres[e * 24 + p] = m1 | m2;
    }
    chunk((int)sh.dbase[BLK], tid, &j0, &j1);
    it.init(sh.dmask, sh.dbase, BLK, j0, j1);
    uint32_t sum = 0;
    int j = j0;
    while (it.next(sh.dmask, &e, &p, &f, &l)) {
      uint32_t c = sh.blk[e] ? dbl_count_exact(pos_of(sh, e), sh.a[e], head_budget(sh, e), p)
                             : dbl_count_under(pos_of(sh, e), sh.a[e], head_budget(sh, e), p);
      sh.dcnt[j < BLK * 16 ? j : 0] = c;
      sum += c;
      j++;
    }
    sh.partB[tid] = sum;
  }
  // ---- phase 3b: ND de-duplicated counts -------------------------------------------------
  static NHD void ph_nd_count(int tid, Sh& sh) {
    int j0, j1, e, p;
    bool f, l;
    chunk((int)sh.ibase[BLK], tid, &j0, &j1);
    ItemIter it;
    it.init(sh.rowmask, sh.ibase, BLK, j0, j1);
    uint32_t sum = 0;
    while (it.next(sh.rowmask, &e, &p, &f, &l)) {
      uint32_t c = (uint32_t)popc32(sh.pres[e * 24 + p] & ~nd_dups(sh, e, p));
      sh.icnt[e * 24 + p] = (uint8_t)c;
      sum += c;
    }
    sh.partA[tid] = sum;
  }
  // ---- phase 5: per-env list bounds from the scanned item counts --------------------------
  static NHD void ph_offsets(int tid, Sh& sh) {
    int j0, j1, e, p;
    bool f, l;
    chunk((int)sh.ibase[BLK], tid, &j0, &j1);
    ItemIter it;
    it.init(sh.rowmask, sh.ibase, BLK, j0, j1);
    uint32_t G = sh.baseA[tid];
    while (it.next(sh.rowmask, &e, &p, &f, &l)) {
      if (f) sh.eG[e] = G;
      G += sh.icnt[e * 24 + p];
      if (l) sh.eEnd[e] = G;
    }
    chunk((int)sh.dbase[BLK], tid, &j0, &j1);
    it.init(sh.dmask, sh.dbase, BLK, j0, j1);
    G = sh.baseB[tid];
    int j = j0;
    while (it.next(sh.dmask, &e, &p, &f, &l)) {
      if (f) sh.eG[e] = G;
      G += sh.dcnt[j < BLK * 16 ? j : 0];
      if (l) sh.eEnd[e] = G;
      j++;
    }
  }
  static NHD uint32_t pick_index(const Sh& sh, int e, int64_t i, uint32_t count, const StepFullArgs& A) {
    if (count == 0) return 0;
    if (A.action_idx) {
      int idx = A.action_idx[i];
      if (idx < 0) idx = 0;
      if (idx >= (int)count) idx = (int)count - 1;
      return (uint32_t)idx;
    }
    return mulhi32(sh.rnd[e], count);
  }
  // ---- phase 6: write the action lists in canonical order, capture the chosen action ------
  static NHD void ph_emit(int tid, Sh& sh, int64_t row0, const StepFullArgs& A) {
    int j0, j1, e, p;
    bool f, l;
    chunk((int)sh.ibase[BLK], tid, &j0, &j1);
    ItemIter it;
    it.init(sh.rowmask, sh.ibase, BLK, j0, j1);
    uint32_t G = sh.baseA[tid];
    while (it.next(sh.rowmask, &e, &p, &f, &l)) {
      uint32_t cnt = sh.icnt[e * 24 + p];
      uint32_t off = G - sh.eG[e];
      G += cnt;
      if (cnt == 0) continue;
      uint32_t total = sh.eEnd[e] - sh.eG[e];
      uint32_t idx = pick_index(sh, e, row0 + e, total, A);
      uint64_t* slice = A.actions ? A.actions + (row0 + e) * (int64_t)A.cap : nullptr;
      bool want = idx >= off && idx < off + cnt;
      if (!slice && !want) continue;
      uint32_t m1, m2;
      nd_row(sh, e, p, &m1, &m2);
      uint32_t nd = sh.pres[e * 24 + p] & ~nd_dups(sh, e, p);
      int a = sh.a[e], b = sh.b[e], ta = p - a;
      uint32_t k = off;
      while (nd) {
        int q = fls32(nd);
        nd &= ~(1u << q);
        uint64_t act = ACT_EMPTY;
        if ((m1 >> q) & 1u) {
          act = act_set(act, 0, p, ta);
          act = act_set(act, 1, q, q - b);
        } else {
          act = act_set(act, 0, q, q - b);
          act = act_set(act, 1, p, ta);
        }
        if (slice && (int)k < A.cap) slice[k] = act;
        if (k == idx) sh.chosen[e] = act;
        k++;
      }
    }
    chunk((int)sh.dbase[BLK], tid, &j0, &j1);
    it.init(sh.dmask, sh.dbase, BLK, j0, j1);
    G = sh.baseB[tid];
    int j = j0;
    while (it.next(sh.dmask, &e, &p, &f, &l)) {
      uint32_t cnt = sh.dcnt[j < BLK * 16 ? j : 0];
      uint32_t off = G - sh.eG[e];
      G += cnt;
      j++;
      if (cnt == 0) continue;
      uint32_t total = sh.eEnd[e] - sh.eG[e];
      uint32_t idx = pick_index(sh, e, row0 + e, total, A);
      uint64_t* slice = A.actions ? A.actions + (row0 + e) * (int64_t)A.cap : nullptr;
      if (sh.blk[e])
        dbl_emit_exact(pos_of(sh, e), sh.a[e], head_budget(sh, e), p, off, cnt, slice, A.cap, idx, &sh.chosen[e]);
      else
        dbl_emit_under(pos_of(sh, e), sh.a[e], head_budget(sh, e), p, off, cnt, slice, A.cap, idx, &sh.chosen[e]);
    }
  }

  // ---- phase 7: per-env completion: rare sequential cases, apply, outputs ------------------
  static NHD void ph_finish(int tid, Sh& sh, bool valid, int64_t i, const StepFullArgs& A, StepFullLocal& L) {
    L.count = 0;
    L.finished = L.white_win = L.black_win = L.mars = L.ep_len = L.overflow = 0;
    if (!valid) return;
    uint8_t kind = sh.kind[tid];
    if (kind == K_DONE) {
      if (A.counts) A.counts[i] = 0;
      if (A.dice_out) A.dice_out[2 * i] = A.dice_out[2 * i + 1] = 0;
      if (A.chosen) A.chosen[i] = ACT_EMPTY;
      if (A.reward) A.reward[i] = 0.0f;
      if (A.done) A.done[i] = 1;
      if (A.truncated) A.truncated[i] = 0;
      return;
    }
    State s = sh.st[tid];
    int player = s.turn();
    int a = sh.a[tid], b = sh.b[tid];
    uint64_t* slice = A.actions ? A.actions + (int64_t)i * A.cap : nullptr;
    uint32_t count = sh.eEnd[tid] - sh.eG[tid];
    uint64_t act = sh.chosen[tid];
    bool sequential = kind == K_D && count == 0;
    if (kind == K_ND && count == 0) {
      // maximal length 1: the higher die if it can be played (narde.py:6 rule 4), else the lower
      uint32_t m = sh.Ca[tid] ? sh.Ca[tid] : sh.S[tid];
      int d = sh.Ca[tid] ? a : b;
      count = (uint32_t)popc32(m);
      uint32_t idx = pick_index(sh, tid, i, count, A);
      uint32_t k = 0;
      while (m) {
        int sp = fls32(m);
        m &= ~(1u << sp);
        uint64_t one = act_set(ACT_EMPTY, 0, sp, sp - d);
        if (slice && (int)k < A.cap) slice[k] = one;
        if (k == idx) act = one;
        k++;
      }
    } else if (sequential) {
      // a doubles turn that cannot use all four dice (~3% of turns): exact per-thread walk
      Pos P = pos_of(sh, tid);
      bool ft = sh.first[tid] != 0;
      // dice order is irrelevant to enumerate_turn (it sorts); a >= b
      if (slice) {
        StoreSink sk = {slice, A.cap, 1, 0};
        count = (uint32_t)enumerate_turn(P, a, b, ft, sk);
      } else {
        CountSink ck;
        count = (uint32_t)enumerate_turn(P, a, b, ft, ck);
      }
      act = ACT_EMPTY;
      if (count) {
        uint32_t idx = pick_index(sh, tid, i, count, A);
        if (slice && (int)idx < A.cap) {
          act = slice[idx];
        } else {
          PickSink pk = {(int)idx, 0, ACT_EMPTY};
          enumerate_turn(P, a, b, ft, pk);
          act = pk.picked;
        }
      }
    }
    L.count = (int)count;
    L.overflow = (slice && (int)count > A.cap) ? 1 : 0;
    if (count) apply_action(s, player, act);
    float rew;
    int dn;
    finish_turn(s, player, (A.flags & F_REWARD_MOVER12) ? 1 : 0, &rew, &dn);
    int bits = dn ? DONE_TERMINATED : 0;
    if (!dn && A.max_episode_steps > 0 && (int)s.steps() >= A.max_episode_steps) bits |= DONE_TRUNCATED;
    if (bits) {
      L.finished = 1;
      L.ep_len = (int)s.steps();
      if (dn) {
        if (player == 1)
          L.white_win = 1;
        else
          L.black_win = 1;
        L.mars = (player == 1 ? s.off_b() : s.off_w()) == 0 ? 1 : 0;
      }
      if (A.flags & F_AUTORESET) s = reset_env(A.seed, (uint32_t)(A.env_base + i), A.step);
    }
    sh.st[tid] = s;
    if (A.counts) A.counts[i] = (int32_t)count;
    if (A.dice_out) {
      int d1 = sh.d1[tid], d2 = sh.d2[tid];  // roll order
      A.dice_out[2 * i] = (uint8_t)d1;
      A.dice_out[2 * i + 1] = (uint8_t)d2;
    }
    if (A.chosen) A.chosen[i] = count ? act : ACT_EMPTY;
    if (A.reward) A.reward[i] = rew;
    if (A.done) A.done[i] = (bits & DONE_TERMINATED) ? 1 : 0;
    if (A.truncated) A.truncated[i] = (bits & DONE_TRUNCATED) ? 1 : 0;
  }
};

}  // namespace narde
