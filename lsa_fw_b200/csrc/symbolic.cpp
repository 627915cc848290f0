// Host symbolic phase of the multifrontal LU: runs once per sparsity pattern.
//
// Replaces what the reference obtains implicitly from PETSc/MUMPS when `iEpsSolver.solve()`
// sets up `PC lu` (reference Solver/utils.py:261-270): fill-reducing ordering, elimination
// (assembly) tree, supernode partition and row structures.  Design choices made for the GPU
// numeric phase that follows:
//   * nested dissection on the graph of A + A^T, every dissection node (leaf domain or
//     separator) becomes ONE dense front -> a static binary assembly tree whose levels are
//     batches of similar-sized dense problems;
//   * bisection key: with coordinates, the longest axis; without, the difference of BFS
//     distances to two pseudo-peripheral vertices (a geometric-like bisector from the graph alone);
//   * decoupled unknowns (Dirichlet identity rows) are split off as 1x1 pivots;
//   * unknowns flagged `order_last` (structurally zero diagonal: pressure) go last inside
//     their front, so the restricted partial pivoting of the numeric phase meets them after the
//     velocities they are coupled to.
#include "lsa_internal.h"

#include <omp.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <stdexcept>

namespace lsa {
namespace {

double now() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

struct Graph {
  int n = 0;
  std::vector<long long> xadj;
  std::vector<int> adj;
};

// Pattern of A + A^T without the diagonal, sorted and de-duplicated.
Graph build_graph(int n, const long long* rowptr, const int* colidx) {
  Graph g;
  g.n = n;
  std::vector<long long> cnt(n + 1, 0);
  for (int i = 0; i < n; ++i)
    for (long long e = rowptr[i]; e < rowptr[i + 1]; ++e) {
      int j = colidx[e];
      if (j == i) continue;
      if (j < 0 || j >= n) throw std::runtime_error("column index out of range");
      cnt[i + 1]++;
      cnt[j + 1]++;
    }
  for (int i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
  std::vector<int> raw(cnt[n]);
  std::vector<long long> pos(cnt.begin(), cnt.end() - 1);
  for (int i = 0; i < n; ++i)
    for (long long e = rowptr[i]; e < rowptr[i + 1]; ++e) {
      int j = colidx[e];
      if (j == i) continue;
      raw[pos[i]++] = j;
      raw[pos[j]++] = i;
    }
  std::vector<long long> deg(n + 1, 0);
#pragma omp parallel for schedule(dynamic, 1024)
  for (int i = 0; i < n; ++i) {
    int* b = raw.data() + cnt[i];
    int* e = raw.data() + cnt[i + 1];
    std::sort(b, e);
    deg[i + 1] = std::unique(b, e) - b;
  }
  for (int i = 0; i < n; ++i) deg[i + 1] += deg[i];
  g.xadj = deg;
  g.adj.resize(deg[n]);
#pragma omp parallel for schedule(dynamic, 1024)
  for (int i = 0; i < n; ++i)
    std::copy(raw.data() + cnt[i], raw.data() + cnt[i] + (deg[i + 1] - deg[i]), g.adj.data() + deg[i]);
  return g;
}

struct NDResult {
  std::vector<int> order;     // vertices in elimination order
  std::vector<int> sn_sizes;  // consecutive supernode sizes covering `order`
};

class Dissector {
 public:
  Dissector(const Graph& g, const AnalyzeOptions& opt)
      : g_(g), opt_(opt), mark_(g.n, -1), da_(g.n, -1), db_(g.n, -1), side_(g.n, 0), token_(0) {}

  void run(std::vector<int>& verts, NDResult& out) {
#pragma omp parallel
#pragma omp single nowait
    rec(verts, 0, out);
  }

 private:
  const Graph& g_;
  const AnalyzeOptions& opt_;
  std::vector<int> mark_, da_, db_;
  std::vector<unsigned char> side_;
  std::atomic<int> token_;

  // BFS inside the subgraph {v : mark_[v] == tok}; dist must be -1 on entry for its vertices.
  int bfs(int start, int tok, std::vector<int>& dist, std::vector<int>& queue) {
    size_t head = queue.size();
    queue.push_back(start);
    dist[start] = 0;
    while (head < queue.size()) {
      int v = queue[head++];
      for (long long e = g_.xadj[v]; e < g_.xadj[v + 1]; ++e) {
        int u = g_.adj[e];
        if (mark_[u] == tok && dist[u] < 0) {
          dist[u] = dist[v] + 1;
          queue.push_back(u);
        }
      }
    }
    return queue.back();
  }

  void bfs_multi(const std::vector<int>& sources, int tok, std::vector<int>& dist, std::vector<int>& queue) {
    queue.clear();
    for (int v : sources) {
      dist[v] = 0;
      queue.push_back(v);
    }
    for (size_t head = 0; head < queue.size(); ++head) {
      const int v = queue[head];
      for (long long e = g_.xadj[v]; e < g_.xadj[v + 1]; ++e) {
        const int u = g_.adj[e];
        if (mark_[u] == tok && dist[u] < 0) {
          dist[u] = dist[v] + 1;
          queue.push_back(u);
        }
      }
    }
  }

  static void emit_leaf(const std::vector<int>& verts, NDResult& out) {
    out.order.insert(out.order.end(), verts.begin(), verts.end());
    out.sn_sizes.push_back((int)verts.size());
  }

  // Hopcroft-Karp maximum matching + Koenig construction on the cut graph (SL x SR).
  std::vector<int> min_vertex_cover(const std::vector<int>& SL, const std::vector<int>& SR, int tok) {
    const int nl = (int)SL.size(), nr = (int)SR.size();
    if (nl == 0 || nr == 0) return {};
    // local ids of the right side via db_ (scratch, reset afterwards)
    for (int j = 0; j < nr; ++j) db_[SR[j]] = -2 - j;
    std::vector<int> xadj(nl + 1, 0), adj;
    for (int i = 0; i < nl; ++i) {
      const int v = SL[i];
      for (long long e = g_.xadj[v]; e < g_.xadj[v + 1]; ++e) {
        const int u = g_.adj[e];
        if (mark_[u] == tok && side_[u] == 1 && db_[u] <= -2) adj.push_back(-2 - db_[u]);
      }
      xadj[i + 1] = (int)adj.size();
    }
    for (int j = 0; j < nr; ++j) db_[SR[j]] = -1;
    std::vector<int> matchL(nl, -1), matchR(nr, -1), dist(nl), queue, it(nl);
    const int INF = 0x3fffffff;
    auto bfs_layers = [&]() {
      queue.clear();
      bool found = false;
      for (int i = 0; i < nl; ++i) {
        if (matchL[i] < 0) {
          dist[i] = 0;
          queue.push_back(i);
        } else {
          dist[i] = INF;
        }
      }
      for (size_t qh = 0; qh < queue.size(); ++qh) {
        const int i = queue[qh];
        for (int e = xadj[i]; e < xadj[i + 1]; ++e) {
          const int i2 = matchR[adj[e]];
          if (i2 < 0) found = true;
          else if (dist[i2] == INF) {
            dist[i2] = dist[i] + 1;
            queue.push_back(i2);
          }
        }
      }
      return found;
    };
    // iterative DFS along the layered graph
    std::vector<int> stack;
    auto try_augment = [&](int root) {
      stack.clear();
      stack.push_back(root);
      while (!stack.empty()) {
        const int i = stack.back();
        bool advanced = false;
        while (it[i] < xadj[i + 1]) {
          const int j = adj[it[i]++];
          const int i2 = matchR[j];
          if (i2 < 0) {
            // augment along the stack
            int jj = j;
            for (int q = (int)stack.size() - 1; q >= 0; --q) {
              const int ii = stack[q];
              const int prev = matchL[ii];
              matchL[ii] = jj;
              matchR[jj] = ii;
              jj = prev;
            }
            return true;
          }
          if (dist[i2] == dist[i] + 1) {
            stack.push_back(i2);
            advanced = true;
            break;
          }
        }
        if (!advanced) {
          dist[i] = INF;
          stack.pop_back();
        }
      }
      return false;
    };
    while (bfs_layers()) {
      for (int i = 0; i < nl; ++i) it[i] = xadj[i];
      for (int i = 0; i < nl; ++i)
        if (matchL[i] < 0) try_augment(i);
    }
    // Koenig: Z = vertices reachable from unmatched left vertices by alternating paths
    std::vector<char> zl(nl, 0), zr(nr, 0);
    queue.clear();
    for (int i = 0; i < nl; ++i)
      if (matchL[i] < 0) {
        zl[i] = 1;
        queue.push_back(i);
      }
    for (size_t qh = 0; qh < queue.size(); ++qh) {
      const int i = queue[qh];
      for (int e = xadj[i]; e < xadj[i + 1]; ++e) {
        const int j = adj[e];
        if (zr[j] || matchL[i] == j) continue;
        zr[j] = 1;
        const int i2 = matchR[j];
        if (i2 >= 0 && !zl[i2]) {
          zl[i2] = 1;
          queue.push_back(i2);
        }
      }
    }
    std::vector<int> cover;
    for (int i = 0; i < nl; ++i)
      if (!zl[i]) cover.push_back(SL[i]);
    for (int j = 0; j < nr; ++j)
      if (zr[j]) cover.push_back(SR[j]);
    return cover;
  }

  void rec(std::vector<int>& verts, int depth, NDResult& out) {
    const int nv = (int)verts.size();
    if (nv == 0) return;
    if (nv <= opt_.leaf_size) {
      emit_leaf(verts, out);
      return;
    }
    const int tok = token_.fetch_add(1);
    for (int v : verts) {
      mark_[v] = tok;
      da_[v] = -1;
      db_[v] = -1;
    }
    std::vector<int> queue;
    queue.reserve(nv);
    int far0 = bfs(verts[0], tok, da_, queue);
    std::vector<int> L, R, S;
    if ((int)queue.size() < nv) {
      // disconnected: distribute the components over two bins, no separator needed
      std::vector<std::pair<int, int>> comps;  // (start, end) in queue
      comps.emplace_back(0, (int)queue.size());
      for (int v : verts)
        if (da_[v] < 0) {
          int s = (int)queue.size();
          bfs(v, tok, da_, queue);
          comps.emplace_back(s, (int)queue.size());
        }
      std::sort(comps.begin(), comps.end(),
                [](auto& a, auto& b) { return (a.second - a.first) > (b.second - b.first); });
      for (auto& c : comps) {
        auto& bin = (L.size() <= R.size()) ? L : R;
        bin.insert(bin.end(), queue.begin() + c.first, queue.begin() + c.second);
      }
    } else {
      std::vector<double> key(nv);
      bool have_key = false;
      if (opt_.dim > 0 && opt_.coords) {
        int best = 0;
        double ext = -1;
        for (int a = 0; a < opt_.dim; ++a) {
          double lo = 1e300, hi = -1e300;
          for (int v : verts) {
            double c = opt_.coords[(size_t)v * opt_.dim + a];
            lo = std::min(lo, c);
            hi = std::max(hi, c);
          }
          if (hi - lo > ext) {
            ext = hi - lo;
            best = a;
          }
        }
        if (ext > 0) {
          for (int i = 0; i < nv; ++i) key[i] = opt_.coords[(size_t)verts[i] * opt_.dim + best];
          have_key = true;
        }
      }
      if (!have_key) {
        // two pseudo-peripheral vertices a, b; key = d(a, v) - d(b, v)
        for (int v : verts) da_[v] = -1;
        queue.clear();
        int b = bfs(far0, tok, da_, queue);
        queue.clear();
        int a2 = bfs(b, tok, db_, queue);
        // one more sweep improves the pair on elongated domains
        for (int v : verts) da_[v] = -1;
        queue.clear();
        bfs(a2, tok, da_, queue);
        // da_ = d(a, .), db_ = d(b, .).  The bisector of two POINTS is slanted on meshes whose hop metric
        // is anisotropic (structured triangulations); the bisector of the two END CAPS
        //   A = {v : d(b, v) >= (1 - eps) D},  B = {v : d(a, v) >= (1 - eps) D}
        // is perpendicular to the long axis whatever the metric: multi-source BFS from the caps.
        if (opt_.cap_fraction > 0.0) {
          int D = 0;
          for (int v : verts) D = std::max(D, da_[v]);
          const int thr = (int)std::floor((1.0 - opt_.cap_fraction) * D);
          std::vector<int> capA, capB;
          for (int v : verts) {
            if (db_[v] >= thr) capA.push_back(v);
            if (da_[v] >= thr) capB.push_back(v);
          }
          if (!capA.empty() && !capB.empty()) {
            for (int v : verts) {
              da_[v] = -1;
              db_[v] = -1;
            }
            bfs_multi(capA, tok, da_, queue);
            bfs_multi(capB, tok, db_, queue);
          }
        }
        for (int i = 0; i < nv; ++i) key[i] = (double)da_[verts[i]] - (double)db_[verts[i]];
      }
      std::vector<double> tmp(key);
      std::nth_element(tmp.begin(), tmp.begin() + nv / 2, tmp.end());
      const double t = tmp[nv / 2];
      int n_lt = 0, n_le = 0;
      for (double k : key) {
        n_lt += k < t;
        n_le += k <= t;
      }
      bool use_le;
      if (n_lt == 0)
        use_le = true;
      else if (n_le == nv)
        use_le = false;
      else
        use_le = std::abs(n_le - nv / 2) < std::abs(n_lt - nv / 2);
      int nL = use_le ? n_le : n_lt;
      if (nL == 0 || nL == nv) {  // all keys equal: cannot bisect, keep as one dense front
        emit_leaf(verts, out);
        return;
      }
      for (int i = 0; i < nv; ++i) side_[verts[i]] = use_le ? (key[i] > t) : (key[i] >= t);
      std::vector<int> SL, SR;
      for (int v : verts) {
        const unsigned char sv = side_[v];
        bool cut = false;
        for (long long e = g_.xadj[v]; e < g_.xadj[v + 1] && !cut; ++e) {
          int u = g_.adj[e];
          cut = (mark_[u] == tok) && (side_[u] != sv);
        }
        if (cut) (sv ? SR : SL).push_back(v);
      }
      // minimum vertex separator for this edge cut = minimum vertex cover of the bipartite graph of
      // cut edges between the two boundary layers (Koenig's theorem via Hopcroft-Karp matching).
      // On FE graphs (every element is a clique) this is about half of either boundary layer.
      S = min_vertex_cover(SL, SR, tok);
      if ((double)S.size() > 0.45 * nv) {
        emit_leaf(verts, out);
        return;
      }
      for (int v : S) side_[v] = 2;
      for (int v : verts) {
        if (side_[v] == 0) L.push_back(v);
        else if (side_[v] == 1) R.push_back(v);
      }
    }
    std::vector<int>().swap(queue);
    NDResult outL, outR;
    const bool spawn = nv > 20000;
#pragma omp task shared(L, outL) firstprivate(depth) if (spawn)
    rec(L, depth + 1, outL);
#pragma omp task shared(R, outR) firstprivate(depth) if (spawn)
    rec(R, depth + 1, outR);
#pragma omp taskwait
    out.order.reserve(out.order.size() + nv);
    out.order.insert(out.order.end(), outL.order.begin(), outL.order.end());
    out.order.insert(out.order.end(), outR.order.begin(), outR.order.end());
    out.sn_sizes.insert(out.sn_sizes.end(), outL.sn_sizes.begin(), outL.sn_sizes.end());
    out.sn_sizes.insert(out.sn_sizes.end(), outR.sn_sizes.begin(), outR.sn_sizes.end());
    if (!S.empty()) emit_leaf(S, out);
  }
};

}  // namespace

void analyze(int n, const long long* rowptr, const int* colidx, const AnalyzeOptions& opt, Symbolic& sym) {
  if (opt.nthreads > 0) omp_set_num_threads(opt.nthreads);
  double t0 = now();
  sym = Symbolic();
  sym.n = n;
  sym.symmetric = opt.symmetric;
  Graph g = build_graph(n, rowptr, colidx);
  double t1 = now();
  sym.seconds[0] = t1 - t0;

  // ---- decoupled unknowns and nested dissection of the rest
  std::vector<int> iso, rest;
  for (int v = 0; v < n; ++v) (g.xadj[v + 1] == g.xadj[v] ? iso : rest).push_back(v);
  sym.n_iso = (int)iso.size();
  NDResult nd;
  {
    Dissector d(g, opt);
    d.run(rest, nd);
  }
  // ---- constrained placement of structurally-zero-diagonal unknowns (pressure): such an unknown
  // must not be eliminated before at least one of the regular unknowns it is coupled to, otherwise
  // its pivot column is still empty when it is reached (MUMPS would delay the pivot dynamically; a
  // static structure has to settle it here).  It is moved up to the front of its earliest regular
  // neighbour when all of them lie in ancestor separators.
  {
    std::vector<int> sn_tmp(n, -1);
    {
      int pos = 0;
      for (size_t s = 0; s < nd.sn_sizes.size(); ++s)
        for (int q = 0; q < nd.sn_sizes[s]; ++q) sn_tmp[nd.order[pos++]] = (int)s;
    }
    const int ns0 = (int)nd.sn_sizes.size();
    std::vector<int> target(n, -1);
    std::vector<int> nb;
    int moved = 0;
    if (opt.order_last) {
      for (int v : nd.order) {
        target[v] = sn_tmp[v];
        if (!opt.order_last[v]) continue;
        // front index such that at least `frac` of the regular neighbours are eliminated no later
        nb.clear();
        for (long long e = g.xadj[v]; e < g.xadj[v + 1]; ++e) {
          const int u = g.adj[e];
          if (!opt.order_last[u]) nb.push_back(sn_tmp[u]);
        }
        if (nb.empty()) continue;
        std::sort(nb.begin(), nb.end());
        size_t q = (size_t)std::ceil(opt.coupled_fraction * (double)nb.size());
        if (q < 1) q = 1;
        if (q > nb.size()) q = nb.size();
        const int sreq = nb[q - 1];
        if (sreq > sn_tmp[v]) {
          target[v] = sreq;
          moved++;
        }
      }
    }
    if (moved > 0) {
      std::vector<int> cnt(ns0 + 1, 0);
      for (int v : nd.order) cnt[target[v] + 1]++;
      for (int s = 0; s < ns0; ++s) cnt[s + 1] += cnt[s];
      std::vector<int> order2(nd.order.size());
      std::vector<int> pos(cnt.begin(), cnt.end() - 1);
      for (int v : nd.order) order2[pos[target[v]]++] = v;  // stable inside each front
      nd.order.swap(order2);
      nd.sn_sizes.clear();
      for (int s = 0; s < ns0; ++s)
        if (cnt[s + 1] > cnt[s]) nd.sn_sizes.push_back(cnt[s + 1] - cnt[s]);
    }
  }
  sym.perm = iso;
  sym.perm.insert(sym.perm.end(), nd.order.begin(), nd.order.end());
  if ((int)sym.perm.size() != n) throw std::runtime_error("ordering lost vertices");
  sym.ns = (int)nd.sn_sizes.size();
  sym.sn_ptr.assign(sym.ns + 1, sym.n_iso);
  for (int s = 0; s < sym.ns; ++s) sym.sn_ptr[s + 1] = sym.sn_ptr[s] + nd.sn_sizes[s];
  if (opt.order_last) {
    const unsigned char* fl = opt.order_last;
#pragma omp parallel for schedule(dynamic, 256)
    for (int s = 0; s < sym.ns; ++s)
      std::stable_partition(sym.perm.begin() + sym.sn_ptr[s], sym.perm.begin() + sym.sn_ptr[s + 1],
                            [fl](int v) { return fl[v] == 0; });
  }
  sym.iperm.assign(n, -1);
  for (int i = 0; i < n; ++i) sym.iperm[sym.perm[i]] = i;
  sym.sn_of.assign(n, -1);
  for (int s = 0; s < sym.ns; ++s)
    for (int i = sym.sn_ptr[s]; i < sym.sn_ptr[s + 1]; ++i) sym.sn_of[i] = s;
  double t2 = now();
  sym.seconds[1] = t2 - t1;

  // ---- supernodal symbolic factorisation on the permuted pattern of A + A^T
  const int ns = sym.ns;
  std::vector<int> parent(ns, -1), first_child(ns, -1), next_sib(ns, -1), last_child(ns, -1);
  std::vector<std::vector<int>> st(ns);
  {
    std::vector<int> mark(n, -1);
    for (int s = 0; s < ns; ++s) {
      const int last = sym.sn_ptr[s + 1] - 1;
      std::vector<int>& rows = st[s];
      for (int i = sym.sn_ptr[s]; i <= last; ++i) {
        const int v = sym.perm[i];
        for (long long e = g.xadj[v]; e < g.xadj[v + 1]; ++e) {
          const int j = sym.iperm[g.adj[e]];
          if (j > last && mark[j] != s) {
            mark[j] = s;
            rows.push_back(j);
          }
        }
      }
      for (int c = first_child[s]; c >= 0; c = next_sib[c])
        for (int j : st[c])
          if (j > last && mark[j] != s) {
            mark[j] = s;
            rows.push_back(j);
          }
      std::sort(rows.begin(), rows.end());
      if (!rows.empty()) {
        const int p = sym.sn_of[rows[0]];
        parent[s] = p;
        if (first_child[p] < 0) first_child[p] = s;
        else next_sib[last_child[p]] = s;
        last_child[p] = s;
      }
    }
  }
  sym.st_ptr.assign(ns + 1, 0);
  for (int s = 0; s < ns; ++s) sym.st_ptr[s + 1] = sym.st_ptr[s] + (long long)st[s].size();
  sym.st_idx.resize(sym.st_ptr[ns]);
#pragma omp parallel for schedule(dynamic, 256)
  for (int s = 0; s < ns; ++s) std::copy(st[s].begin(), st[s].end(), sym.st_idx.begin() + sym.st_ptr[s]);
  std::vector<std::vector<int>>().swap(st);

  // ---- assembly tree levels (roots at level 0), children lists
  std::vector<int> level(ns, 0);
  int nlev = 0;
  for (int s = ns - 1; s >= 0; --s) {
    level[s] = parent[s] < 0 ? 0 : level[parent[s]] + 1;
    nlev = std::max(nlev, level[s] + 1);
  }
  sym.nlevels = nlev;
  sym.fronts.assign(ns, Front());
  sym.child_idx.clear();
  for (int s = 0; s < ns; ++s) {
    Front& f = sym.fronts[s];
    f.k = sym.sn_ptr[s + 1] - sym.sn_ptr[s];
    f.r = (int)(sym.st_ptr[s + 1] - sym.st_ptr[s]);
    f.col0 = sym.sn_ptr[s];
    f.st0 = sym.st_ptr[s];
    f.parent = parent[s];
    f.level = level[s];
    f.child0 = (int)sym.child_idx.size();
    f.nchild = 0;
    for (int c = first_child[s]; c >= 0; c = next_sib[c]) {
      sym.child_idx.push_back(c);
      f.nchild++;
    }
  }
  sym.lvl_ptr.assign(nlev + 1, 0);
  for (int s = 0; s < ns; ++s) sym.lvl_ptr[level[s] + 1]++;
  for (int d = 0; d < nlev; ++d) sym.lvl_ptr[d + 1] += sym.lvl_ptr[d];
  sym.lvl_front.resize(ns);
  {
    std::vector<int> pos(sym.lvl_ptr.begin(), sym.lvl_ptr.end() - 1);
    for (int s = 0; s < ns; ++s) sym.lvl_front[pos[level[s]]++] = s;
    for (int d = 0; d < nlev; ++d)
      std::stable_sort(sym.lvl_front.begin() + sym.lvl_ptr[d], sym.lvl_front.begin() + sym.lvl_ptr[d + 1],
                       [&](int a, int b) { return sym.fronts[a].k > sym.fronts[b].k; });
  }

  // ---- storage layout and work counters
  auto align = [](long long x) { return (x + 3) & ~3LL; };
  long long off = 0;
  sym.nnz_lu = sym.n_iso;
  sym.flops = 0.0;
  for (int s = 0; s < ns; ++s) {
    Front& f = sym.fronts[s];
    const long long k = f.k, r = f.r, m = k + r;
    f.p_off = off;
    off = align(off + m * k);
    f.q_off = off;
    if (!opt.symmetric) off = align(off + k * r);   // symmetric: U12 = D L21^T is never stored
    sym.nnz_lu += opt.symmetric ? k * k + k * r : k * k + 2 * k * r;
    sym.flops += opt.symmetric ? (2.0 / 3.0) * k * k * k + 1.0 * k * k * r + 1.0 * k * r * r
                               : (2.0 / 3.0) * k * k * k + 2.0 * k * k * r + 2.0 * k * r * r;
    sym.max_k = std::max(sym.max_k, f.k);
    sym.max_r = std::max(sym.max_r, f.r);
    sym.max_m = std::max(sym.max_m, f.k + f.r);
  }
  sym.diag_off = off;
  off = align(off + sym.n_iso);
  sym.fac_size = off;
  sym.pool_size[0] = sym.pool_size[1] = 0;
  for (int d = 0; d < nlev; ++d) {
    long long c = 0;
    for (int q = sym.lvl_ptr[d]; q < sym.lvl_ptr[d + 1]; ++q) {
      Front& f = sym.fronts[sym.lvl_front[q]];
      f.c_off = c;
      c = align(c + (long long)f.r * f.r);
    }
    sym.pool_size[d & 1] = std::max(sym.pool_size[d & 1], c);
  }
  double t3 = now();
  sym.seconds[2] = t3 - t2;

  // ---- extend-add maps (the value scatter maps depend on the caller's CSR and are built in capi.cu)
  auto local_index = [&](int s, int idx) -> int {
    const Front& f = sym.fronts[s];
    if (idx < f.col0 + f.k) return idx - f.col0;
    const int* b = sym.st_idx.data() + f.st0;
    const int* e = b + f.r;
    const int* it = std::lower_bound(b, e, idx);
    if (it == e || *it != idx) return -1;
    return f.k + (int)(it - b);
  };
  int bad = 0;
  sym.ea_map.resize(sym.st_idx.size());
#pragma omp parallel for schedule(dynamic, 256) reduction(| : bad)
  for (int c = 0; c < ns; ++c) {
    const Front& f = sym.fronts[c];
    if (f.parent < 0) continue;
    for (int t = 0; t < f.r; ++t) {
      const int li = local_index(f.parent, sym.st_idx[f.st0 + t]);
      if (li < 0) bad |= 1;
      sym.ea_map[f.st0 + t] = li;
    }
  }
  if (bad) throw std::runtime_error("child structure not contained in parent front");
  sym.seconds[3] = now() - t3;
}

}  // namespace lsa
