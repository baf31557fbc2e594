// Tile plan of the tcgen05 scan: which worker scans which (row unit, table tile range), which
// partial-result slot it ends in, and which slots a row block owns.  Closed-form, so the scan
// kernel (every role of every CTA walks its own segments), the merge kernel (every row finds
// its slots) and the host (sizing, CPU tests through mcl_plan_segments) evaluate the same few
// integer formulas and nothing has to be uploaded.
//
// Vocabulary:
//   worker    one cluster of `cs` CTAs (a CTA pair when cs = 2); `workers` <= SMs / cs
//   row unit  `cs` consecutive row blocks of 128 queries, one per CTA of a worker
//   wave      `gu` row units the chip works on at the same time (their query tiles stay in L2)
//   node      one level of a wave's plan: R row units x table tiles [a, T) on W >= R workers.
//             nfull = W div R groups of R workers -- one worker per row unit -- walk the tiles
//             [a + g*tpc, a + (g+1)*tpc) SIDE BY SIDE, so a tile is fetched from HBM once per
//             group.  The wr = W - nfull*R workers that do not fill another group take the
//             tail tiles [t0, T) of all R row units: `passes` = R div wr passes, in each of
//             which the wr workers walk the tail side by side on wr more row units; the
//             R mod wr row units still missing the tail form the next node (R' < wr workers
//             again fill whole groups ...), Euclid's algorithm on (row units, workers).
//             tpc is chosen so that a group member and a tail worker finish together.
// Without the tail workers 64 row blocks on 148 SMs leave 20 SMs idle (C3, C5).
// Every segment (one worker, one row unit, one contiguous tile range) ends in its own slot: a
// row unit collects one slot per group of every node it passes through, plus one for its tail
// pass.  A row unit's slot count therefore varies; S is the allocation stride.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MCL_HD __host__ __device__ __forceinline__
#else
#define MCL_HD inline
#endif

namespace mcl {

constexpr int kMaxNodes = 4;
constexpr int kJointWords = 8;   // joint-threshold words per query row (see rowstate.cuh)

struct Node {
  int r0, R;       // row units [r0, r0 + R) of the wave
  int a;           // first tile (the node covers [a, T))
  int w0;          // first worker
  int nfull, tpc;  // groups, tiles per group
  int t0;          // first tail tile (= T: no tail)
  int wr, passes;  // tail workers (0 = none), tail passes
  int jbase;       // slots a row unit has collected in the nodes above
  int sync0;       // first drift counter of the node within its wave
  int nwin_g, nwin_t;   // drift windows of a group / of a tail pass
};

struct Chain {
  int n;
  Node nd[kMaxNodes];
};

struct TcPlan {
  int num_rb, num_vt, num_kb;
  int cs, workers, gu, ru, waves;
  int S;       // slots reserved per row block and column half
  int win;     // drift window in tiles
  int nsync;   // drift counters per wave
  Chain full, last;   // waves 0..waves-2 / the last wave
};

struct Seg {
  int unit;      // row unit
  int vt0, vt1;  // table tiles [vt0, vt1)
  int j;         // slot of the row unit this segment ends in
  int sync;      // first drift counter of this segment's group or tail pass
  int members;   // CTAs walking these tiles side by side
  int ng2;       // joint threshold: words in use for this wave's rows (0 = off)
  int jw;        // joint threshold: word this segment's column half 0 publishes to (-1 = reads only)
};

MCL_HD const Chain& plan_chain(const TcPlan& p, int wave) {
  return (wave == p.waves - 1) ? p.last : p.full;
}

// Walks the segments of one worker in execution order (a few integer divisions per segment;
// every role of the scan kernel runs its own copy).
struct SegIter {
  int w, v, node, pass;   // worker; next wave, node of its chain, tail pass of that node
};
MCL_HD void seg_iter_init(SegIter& it, int worker) { it.w = worker; it.v = 0; it.node = 0; it.pass = 0; }
MCL_HD bool seg_iter_next(const TcPlan& p, SegIter& it, Seg& s) {
  for (;;) {
    if (it.v >= p.waves) return false;
    const Chain& ch = plan_chain(p, it.v);
    if (it.node >= ch.n) { ++it.v; it.node = 0; it.pass = 0; continue; }
    const Node& nd = ch.nd[it.node];
    const int rel = it.w - nd.w0;
    const int ng = nd.nfull * nd.R;
    const int u0 = it.v * p.gu + nd.r0;
    const int c0 = it.v * p.nsync + nd.sync0;
    // joint threshold (rowstate.cuh): the column halves of the first groups of the wave's top node
    // publish, every segment of the wave reads
    int ng2 = 2 * ch.nd[0].nfull;
    if (ng2 > kJointWords) ng2 = kJointWords;
    if (ng2 < 4) ng2 = 0;
    s.ng2 = ng2;
    if (rel < ng) {                            // member of a group: one segment ends the wave
      const int g = rel / nd.R, m = rel - g * nd.R;
      const int a = nd.a + g * nd.tpc;
      const int b = (a + nd.tpc < nd.t0) ? a + nd.tpc : nd.t0;
      const bool top = it.node == 0;
      ++it.v; it.node = 0; it.pass = 0;
      if (rel >= 0 && a < b) {
        s.unit = u0 + m; s.vt0 = a; s.vt1 = b; s.j = nd.jbase + g;
        s.jw = (top && 2 * g < ng2) ? 2 * g : -1;
        s.sync = c0 + g * nd.nwin_g; s.members = nd.R * p.cs;
        return true;
      }
      continue;
    }
    const int m = rel - ng;
    if (m >= nd.wr) { ++it.v; it.node = 0; it.pass = 0; continue; }   // idle in this wave
    if (it.pass < nd.passes) {                 // tail worker: next pass over [t0, T)
      s.unit = u0 + it.pass * nd.wr + m; s.vt0 = nd.t0; s.vt1 = p.num_vt; s.j = nd.jbase + nd.nfull;
      s.sync = c0 + nd.nfull * nd.nwin_g + it.pass * nd.nwin_t; s.members = nd.wr * p.cs;
      s.jw = -1;
      ++it.pass;
      return true;
    }
    ++it.node; it.pass = 0;                    // the row units the passes did not reach
  }
}

// Number of segments of worker w; writes the first `cap` of them.
MCL_HD int plan_segments(const TcPlan& p, int w, Seg* out, int cap) {
  SegIter it;
  seg_iter_init(it, w);
  Seg s;
  int n = 0;
  while (seg_iter_next(p, it, s)) {
    if (n < cap) out[n] = s;
    ++n;
  }
  return n;
}

// Slots written for row unit u: 0 .. n-1.
MCL_HD int plan_unit_slots(const TcPlan& p, int u) {
  const int v = u / p.gu, uw = u - v * p.gu;
  const Chain& ch = plan_chain(p, v);
  int n = 0;
  for (int i = 0; i < ch.n; ++i) {
    const Node& nd = ch.nd[i];
    n = nd.jbase + nd.nfull;
    if (nd.wr == 0) break;
    if (uw < nd.r0 + nd.passes * nd.wr) { ++n; break; }
  }
  return n;
}

// How the merge finds the slots of a row block: slot0 = rb*stride, count = uniform, or (tcgen05
// plan) two column halves per planned slot of the row block's unit.
struct SlotMap {
  int stride;
  int uniform;   // > 0: every row block owns exactly this many slots (CUDA-core engine)
  TcPlan plan;
};
MCL_HD int slotmap_count(const SlotMap& m, int rb) {
  return m.uniform > 0 ? m.uniform : 2 * plan_unit_slots(m.plan, rb / m.plan.cs);
}

}  // namespace mcl
