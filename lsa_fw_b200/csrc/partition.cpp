// Split of one multifrontal factorisation / solve over the GPUs of a node: sub-tree -> GPU mapping and the
// per-rank ("local") symbolic structures.
//
// The reference reaches several ranks through PETSc/SLEPc/MUMPS on PETSc.COMM_WORLD (Solver/utils.py:196-203:
// "fully MPI-parallelized", README.md:43); MUMPS distributes the assembly tree by proportional mapping.  Here
// (stage 1 of SURVEY 8e): the assembly tree is cut so that the sub-trees below the cut balance the GPUs under a
// work model, every sub-tree belongs to ONE GPU (no communication inside), and the fronts above the cut are
// replicated.  This file only decides and re-indexes; the exchange steps are in factor.cu / solve.cu / krylov.cu.
#include "lsa_internal.h"

#include <algorithm>
#include <numeric>
#include <queue>
#include <stdexcept>

namespace lsa {
namespace {

// seconds-like work model of one front: FP64 flops of its partial LU at ~10 TFLOP/s (complex) and ~100 sweeps
// over its factor entries at ~3 TB/s -- the two phases of one eigensolve
double front_weight(const Front& f) {
  const double k = f.k, r = f.r;
  const double flops = (2.0 / 3.0) * k * k * k + 2.0 * k * k * r + 2.0 * k * r * r;
  const double entries = k * k + 2.0 * k * r;
  return flops * 4e-13 + entries * 5.3e-10;
}

// makespan of longest-processing-time-first packing of `w` (descending) into `bins` bins
double lpt_makespan(std::vector<double> w, int bins, std::vector<int>* assign = nullptr) {
  std::vector<int> idx(w.size());
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return w[a] > w[b]; });
  std::vector<double> load(bins, 0.0);
  if (assign) assign->assign(w.size(), 0);
  for (int i : idx) {
    int best = 0;
    for (int b = 1; b < bins; ++b)
      if (load[b] < load[best]) best = b;
    load[best] += w[i];
    if (assign) (*assign)[i] = best;
  }
  return *std::max_element(load.begin(), load.end());
}

}  // namespace

void partition(const Symbolic& g, int rank, int world, Symbolic& loc, Partition& part) {
  if (world < 1 || rank < 0 || rank >= world) throw std::runtime_error("partition: rank / world out of range");
  const int ns = g.ns;
  part = Partition();
  part.rank = rank;
  part.world = world;
  part.ns_global = ns;
  part.nnz_lu_global = g.nnz_lu;
  part.flops_global = g.flops;

  // ---- sub-tree weights (children precede parents: one ascending pass)
  std::vector<double> fw(ns), sw(ns);
  for (int s = 0; s < ns; ++s) sw[s] = fw[s] = front_weight(g.fronts[s]);
  for (int s = 0; s < ns; ++s)
    if (g.fronts[s].parent >= 0) sw[g.fronts[s].parent] += sw[s];

  // ---- cut: start from the roots, open the heaviest sub-tree (its root joins the replicated top) as long as
  // that lowers the modelled time  top work + makespan of the sub-trees on `world` GPUs
  std::vector<char> is_top(ns, 0);
  std::vector<int> subs;
  for (int s = 0; s < ns; ++s)
    if (g.fronts[s].parent < 0) subs.push_back(s);
  double top_w = 0.0;
  auto model = [&](const std::vector<int>& roots, double tw) {
    std::vector<double> w;
    for (int s : roots) w.push_back(sw[s]);
    return tw + (w.empty() ? 0.0 : lpt_makespan(w, world));
  };
  if (world > 1) {
    // Open the heaviest sub-tree again and again (its root joins the replicated top) and remember the modelled
    // time after every step; keep the prefix of openings with the smallest one.  (One step alone rarely pays: the
    // sibling of the opened sub-tree still bounds the makespan until it is opened too.)  A GPU may end up without a
    // sub-tree: with a replicated top, one more level of a 3-D tree can cost every GPU more than it returns.
    std::vector<int> cur(subs), opened;
    double tw = 0.0, best = model(cur, 0.0);
    size_t best_steps = 0;
    while ((int)cur.size() <= 64 * world) {
      int hi = 0;
      for (int i = 1; i < (int)cur.size(); ++i)
        if (sw[cur[i]] > sw[cur[hi]]) hi = i;
      const int s = cur[hi];
      if (g.fronts[s].nchild == 0) break;   // the heaviest sub-tree is a single front
      cur.erase(cur.begin() + hi);
      for (int c = 0; c < g.fronts[s].nchild; ++c) cur.push_back(g.child_idx[g.fronts[s].child0 + c]);
      tw += fw[s];
      opened.push_back(s);
      const double t = model(cur, tw);
      if (t < best) {
        best = t;
        best_steps = opened.size();
      }
    }
    for (size_t q = 0; q < best_steps; ++q) {
      const int s = opened[q];
      is_top[s] = 1;
      top_w += fw[s];
      subs.erase(std::find(subs.begin(), subs.end(), s));
      for (int c = 0; c < g.fronts[s].nchild; ++c) subs.push_back(g.child_idx[g.fronts[s].child0 + c]);
    }
  }
  std::sort(subs.begin(), subs.end());
  std::vector<double> w;
  for (int s : subs) w.push_back(sw[s]);
  std::vector<int> assign;
  part.weight_max = w.empty() ? 0.0 : lpt_makespan(w, world, &assign);
  part.weight_top = top_w;
  part.weight_total = top_w + std::accumulate(w.begin(), w.end(), 0.0);

  // ---- owner of every front: descend from the sub-tree roots (parents have larger ids than children)
  part.owner.assign(ns, -2);
  for (int s = 0; s < ns; ++s)
    if (is_top[s]) part.owner[s] = -1;
  for (size_t i = 0; i < subs.size(); ++i) part.owner[subs[i]] = world > 1 ? assign[i] : 0;
  for (int s = ns - 1; s >= 0; --s)
    if (part.owner[s] == -2) {
      const int p = g.fronts[s].parent;
      if (p < 0 || part.owner[p] < 0) throw std::runtime_error("partition: front without an owner");
      part.owner[s] = part.owner[p];
    }
  for (size_t i = 0; i < subs.size(); ++i) {
    if (part.owner[subs[i]] == rank) part.weight_mine += w[i];
    if (g.fronts[subs[i]].parent >= 0) part.cut_roots.push_back(subs[i]);
  }

  // ---- local fronts: top + own sub-trees + ghosts (other ranks' sub-tree roots hanging below a top front)
  auto is_cut_root = [&](int s) { return part.owner[s] >= 0 && g.fronts[s].parent >= 0 && part.owner[g.fronts[s].parent] < 0; };
  part.g2l.assign(ns, -1);
  for (int s = 0; s < ns; ++s) {
    const int o = part.owner[s];
    if (o == -1 || o == rank || is_cut_root(s)) {
      part.g2l[s] = (int)part.l2g.size();
      part.l2g.push_back(s);
    }
  }
  const int nl = (int)part.l2g.size();
  int top_levels = 0, own_min_level = 1 << 30;
  for (int s = 0; s < ns; ++s) {
    if (part.owner[s] == -1) top_levels = std::max(top_levels, g.fronts[s].level + 1);
    if (part.owner[s] == rank && is_cut_root(s)) own_min_level = std::min(own_min_level, g.fronts[s].level);
    if (part.owner[s] == rank && g.fronts[s].parent < 0) own_min_level = std::min(own_min_level, g.fronts[s].level);
  }
  part.n_top_levels = top_levels;

  if (g.symmetric) throw std::runtime_error("the symmetric factorisation is not available in the partitioned solve");
  loc = Symbolic();
  loc.n = g.n;
  loc.n_iso = g.n_iso;
  loc.perm = g.perm;
  loc.iperm = g.iperm;
  loc.ns = nl;
  loc.fronts.assign(nl, Front());
  loc.sn_ptr.assign(nl + 1, 0);
  loc.sn_of.assign(g.n, -1);
  loc.st_ptr.assign(nl + 1, 0);
  int nlev = 0;
  for (int l = 0; l < nl; ++l) {
    const int s = part.l2g[l];
    const Front& gf = g.fronts[s];
    Front& f = loc.fronts[l];
    const bool ghost = part.owner[s] >= 0 && part.owner[s] != rank;
    f.flags = ghost ? FRONT_GHOST : (is_cut_root(s) ? FRONT_CUT : FRONT_REGULAR);
    f.k = ghost ? 0 : gf.k;
    f.r = gf.r;
    f.col0 = gf.col0;
    f.parent = gf.parent >= 0 ? part.g2l[gf.parent] : -1;
    if (gf.parent >= 0 && f.parent < 0) throw std::runtime_error("partition: parent of a local front is not local");
    // local levels: the replicated top keeps its depth, this rank's sub-trees come below ALL top levels
    f.level = part.owner[s] == -1 ? gf.level : ghost ? -1 : top_levels + (gf.level - own_min_level);
    if (!ghost) nlev = std::max(nlev, f.level + 1);
    loc.sn_ptr[l] = gf.col0;
    loc.st_ptr[l + 1] = loc.st_ptr[l] + gf.r;
    f.st0 = loc.st_ptr[l];
    if (!ghost)
      for (int c = gf.col0; c < gf.col0 + gf.k; ++c) loc.sn_of[c] = l;
  }
  loc.sn_ptr[nl] = g.n;
  loc.nlevels = nlev;
  loc.st_idx.resize(loc.st_ptr[nl]);
  loc.ea_map.resize(loc.st_ptr[nl]);
  for (int l = 0; l < nl; ++l) {
    const Front& gf = g.fronts[part.l2g[l]];
    std::copy(g.st_idx.begin() + gf.st0, g.st_idx.begin() + gf.st0 + gf.r, loc.st_idx.begin() + loc.fronts[l].st0);
    std::copy(g.ea_map.begin() + gf.st0, g.ea_map.begin() + gf.st0 + gf.r, loc.ea_map.begin() + loc.fronts[l].st0);
  }
  // children (ascending local id = ascending global id)
  {
    std::vector<int> cnt(nl, 0);
    for (int l = 0; l < nl; ++l)
      if (loc.fronts[l].parent >= 0) cnt[loc.fronts[l].parent]++;
    int off = 0;
    for (int l = 0; l < nl; ++l) {
      loc.fronts[l].child0 = off;
      loc.fronts[l].nchild = 0;
      off += cnt[l];
    }
    loc.child_idx.assign(off, -1);
    for (int l = 0; l < nl; ++l) {
      const int p = loc.fronts[l].parent;
      if (p >= 0) loc.child_idx[loc.fronts[p].child0 + loc.fronts[p].nchild++] = l;
    }
  }
  // levels (ghosts are in no level: they are never factored or swept here)
  loc.lvl_ptr.assign(nlev + 1, 0);
  for (int l = 0; l < nl; ++l)
    if (loc.fronts[l].flags != FRONT_GHOST) loc.lvl_ptr[loc.fronts[l].level + 1]++;
  for (int d = 0; d < nlev; ++d) loc.lvl_ptr[d + 1] += loc.lvl_ptr[d];
  loc.lvl_front.resize(loc.lvl_ptr[nlev]);
  {
    std::vector<int> pos(loc.lvl_ptr.begin(), loc.lvl_ptr.end() - 1);
    for (int l = 0; l < nl; ++l)
      if (loc.fronts[l].flags != FRONT_GHOST) loc.lvl_front[pos[loc.fronts[l].level]++] = l;
    for (int d = 0; d < nlev; ++d)
      std::stable_sort(loc.lvl_front.begin() + loc.lvl_ptr[d], loc.lvl_front.begin() + loc.lvl_ptr[d + 1],
                       [&](int a, int b) { return loc.fronts[a].k > loc.fronts[b].k; });
  }
  // storage: factor store of the local fronts, level pools, cut pool
  auto align = [](long long x) { return (x + 3) & ~3LL; };
  long long off = 0, cut = 0;
  loc.nnz_lu = loc.n_iso;
  for (int l = 0; l < nl; ++l) {
    Front& f = loc.fronts[l];
    if (f.flags != FRONT_GHOST) {
      const long long k = f.k, r = f.r, m = k + r;
      f.p_off = off;
      off = align(off + m * k);
      f.q_off = off;
      off = align(off + k * r);
      loc.nnz_lu += k * k + 2 * k * r;
      loc.flops += (2.0 / 3.0) * k * k * k + 2.0 * k * k * r + 2.0 * k * r * r;
      loc.max_k = std::max(loc.max_k, f.k);
      loc.max_r = std::max(loc.max_r, f.r);
      loc.max_m = std::max(loc.max_m, f.k + f.r);
    }
    if (f.flags != FRONT_REGULAR) {
      f.c_off = cut;
      cut = align(cut + (long long)f.r * f.r);
    }
  }
  loc.diag_off = off;
  off = align(off + loc.n_iso);
  loc.fac_size = off;
  part.cut_pool_size = cut;
  for (int d = 0; d < nlev; ++d) {
    long long c = 0;
    for (int q = loc.lvl_ptr[d]; q < loc.lvl_ptr[d + 1]; ++q) {
      Front& f = loc.fronts[loc.lvl_front[q]];
      if (f.flags != FRONT_REGULAR) continue;
      f.c_off = c;
      c = align(c + (long long)f.r * f.r);
    }
    loc.pool_size[d & 1] = std::max(loc.pool_size[d & 1], c);
  }
  // row ranges: replicated rows (decoupled pivots + top fronts) and this rank's sub-tree rows, merged
  auto push = [](std::vector<int>& lo, std::vector<int>& hi, int a, int b) {
    if (b <= a) return;
    if (!hi.empty() && hi.back() == a) hi.back() = b;
    else {
      lo.push_back(a);
      hi.push_back(b);
    }
  };
  push(part.top_lo, part.top_hi, 0, g.n_iso);
  for (int s = 0; s < ns; ++s) {
    const Front& gf = g.fronts[s];
    if (part.owner[s] == -1) push(part.top_lo, part.top_hi, gf.col0, gf.col0 + gf.k);
    else if (part.owner[s] == rank) push(part.own_lo, part.own_hi, gf.col0, gf.col0 + gf.k);
  }
  for (size_t i = 0; i < part.top_lo.size(); ++i) part.n_top_rows += part.top_hi[i] - part.top_lo[i];
  for (int q = 0; q < 4; ++q) loc.seconds[q] = g.seconds[q];
}

}  // namespace lsa
