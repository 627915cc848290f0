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
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <numeric>
#include <stdexcept>

namespace lsa {
namespace {

double now() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// nvcc's host pass drops __builtin_prefetch; the instruction is spelled out where the host is x86-64.
static inline void host_prefetch(const void* p) {
#if defined(__x86_64__)
  asm volatile("prefetcht0 %0" : : "m"(*static_cast<const char*>(p)));
#else
  (void)p;
#endif
}

struct Graph {
  int n = 0;
  std::vector<long long> xadj;
  std::unique_ptr<int[]> adj;   // xadj[n] entries, uninitialised until build_graph fills them in parallel
};

// Pattern of A + A^T without the diagonal, sorted and de-duplicated.  Every pass runs over the rows in parallel
// (the transposed entries claim their slots with atomic counters; the per-row sort makes the result independent
// of the order in which they arrive); the scratch array is left uninitialised so that its pages are first touched
// by the threads that fill them.
Graph build_graph(int n, const long long* rowptr, const int* colidx) {
  Graph g;
  g.n = n;
  std::vector<long long> cnt(n + 1, 0);   // cnt[i + 1]: own off-diagonal entries of row i + entries (j, i) of other rows
  std::vector<int> own(n, 0);
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(| : bad)
  for (int i = 0; i < n; ++i) {
    int c = 0;
    for (long long e = rowptr[i]; e < rowptr[i + 1]; ++e) {
      const int j = colidx[e];
      if (j == i) continue;
      if (j < 0 || j >= n) {
        bad |= 1;
        continue;
      }
      ++c;
#pragma omp atomic
      cnt[j + 1]++;
    }
    own[i] = c;
  }
  if (bad) throw std::runtime_error("column index out of range");
  for (int i = 0; i < n; ++i) cnt[i + 1] += cnt[i] + own[i];
  std::unique_ptr<int[]> raw(new int[(size_t)cnt[n]]);
  std::vector<long long> pos(n);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) pos[i] = cnt[i] + own[i];
#pragma omp parallel for schedule(dynamic, 1024)
  for (int i = 0; i < n; ++i) {
    long long o = cnt[i];
    for (long long e = rowptr[i]; e < rowptr[i + 1]; ++e) {
      const int j = colidx[e];
      if (j == i) continue;
      raw[o++] = j;
      long long q;
#pragma omp atomic capture
      q = pos[j]++;
      raw[q] = i;
    }
  }
  std::vector<long long> deg(n + 1, 0);
#pragma omp parallel for schedule(dynamic, 1024)
  for (int i = 0; i < n; ++i) {
    int* b = raw.get() + cnt[i];
    int* e = raw.get() + cnt[i + 1];
    std::sort(b, e);
    deg[i + 1] = std::unique(b, e) - b;
  }
  for (int i = 0; i < n; ++i) deg[i + 1] += deg[i];
  g.xadj = deg;
  g.adj.reset(new int[(size_t)deg[n]]);
#pragma omp parallel for schedule(dynamic, 1024)
  for (int i = 0; i < n; ++i)
    std::copy(raw.get() + cnt[i], raw.get() + cnt[i] + (deg[i + 1] - deg[i]), g.adj.get() + deg[i]);
  return g;
}

struct NDResult {
  std::vector<int> order;     // vertices in elimination order
  std::vector<int> sn_sizes;  // consecutive supernode sizes covering `order`
};

// A dissection node works on a COMPACT copy of its subgraph: local vertex numbers 0..nv-1 in the order of its
// vertex list, adjacency lists in the order of the global graph (sorted by global number) restricted to the
// subgraph.  The six breadth-first sweeps, the cut detection and the matching of a node then touch nv-sized
// arrays only (cache-resident from a few levels below the root), and the children's copies are filtered from the
// parent's, not from the global graph.  Orders of visit are the ones of a walk over the global graph with
// membership marks, so the ordering produced is independent of this storage choice.
struct SubGraph {
  std::vector<int> gid;           // local -> global vertex number
  std::unique_ptr<int[]> xadj;    // nv + 1
  std::unique_ptr<int[]> adj;     // local numbers
  int nv() const { return (int)gid.size(); }
};

class Dissector {
 public:
  Dissector(const Graph& g, const AnalyzeOptions& opt) : g_(g), opt_(opt) {}

  void run(std::vector<int>& verts, NDResult& out) {
#pragma omp parallel
#pragma omp single nowait
    {
      SubGraph top;
      {
        std::vector<int> loc(g_.n, -1);
        const int nv = (int)verts.size();
        for (int i = 0; i < nv; ++i) loc[verts[i]] = i;
        top.gid = verts;
        top.xadj.reset(new int[(size_t)nv + 1]);
        top.xadj[0] = 0;
        std::vector<int> deg(nv);
#pragma omp taskloop grainsize(8192) shared(deg, loc, verts)
        for (int i = 0; i < nv; ++i) {
          const int v = verts[i];
          int c = 0;
          for (long long e = g_.xadj[v]; e < g_.xadj[v + 1]; ++e) c += loc[g_.adj[e]] >= 0;
          deg[i] = c;
        }
        long long tot = 0;
        for (int i = 0; i < nv; ++i) {
          tot += deg[i];
          top.xadj[i + 1] = (int)tot;   // analyze() has checked that the whole graph fits 32-bit offsets
        }
        top.adj.reset(new int[(size_t)tot]);
#pragma omp taskloop grainsize(8192) shared(top, loc, verts)
        for (int i = 0; i < nv; ++i) {
          const int v = verts[i];
          int o = top.xadj[i];
          for (long long e = g_.xadj[v]; e < g_.xadj[v + 1]; ++e) {
            const int u = loc[g_.adj[e]];
            if (u >= 0) top.adj[o++] = u;
          }
        }
      }
      rec(top, 0, out);
    }
  }

 private:
  const Graph& g_;
  const AnalyzeOptions& opt_;

  // Work queue of the breadth-first sweeps: capacity nv + 1 so that the sweep can store a candidate before it
  // knows whether the vertex is new (branch-free inner loop: the store is kept only if the tail advances).
  struct Queue {
    std::vector<int> q;
    size_t tail = 0;
    explicit Queue(int nv) : q((size_t)nv + 1) {}
    void clear() { tail = 0; }
    size_t size() const { return tail; }
    int back() const { return q[tail - 1]; }
  };

  // The queue is known ahead of the vertex being expanded: fetch the adjacency list of the vertex 8 places on (and
  // the offsets of the one 16 places on).  Sweeps whose visiting order does not follow the numbering were bound by
  // the latency of these reads (config 3's root node: 0.42-0.60 s -> 0.28-0.30 s per sweep, 0.20 -> 0.18 s for the
  // sweep that does follow it).
  static inline void prefetch_ahead(const int* q, size_t head, size_t tail, const int* xadj, const int* adj) {
    if (head + 8 < tail) {
      const int* p = adj + xadj[q[head + 8]];
      host_prefetch(p);
      host_prefetch(p + 16);
    }
    if (head + 16 < tail) host_prefetch(xadj + q[head + 16]);
  }

  // BFS over the subgraph from `start`, appended to `queue`; dist must be -1 on entry for the vertices reached.
  static int bfs(const SubGraph& sg, int start, std::vector<int>& dist_v, Queue& queue) {
    int* q = queue.q.data();
    int* dist = dist_v.data();
    size_t head = queue.tail, tail = queue.tail;
    q[tail++] = start;
    dist[start] = 0;
    const int* xadj = sg.xadj.get();
    const int* adj = sg.adj.get();
    while (head < tail) {
      prefetch_ahead(q, head, tail, xadj, adj);
      const int v = q[head++];
      const int dv = dist[v] + 1;
      for (int e = xadj[v]; e < xadj[v + 1]; ++e) {
        const int u = adj[e];
        const int du = dist[u];
        const bool fresh = du < 0;
        q[tail] = u;
        tail += fresh;
        dist[u] = fresh ? dv : du;
      }
    }
    queue.tail = tail;
    return q[tail - 1];
  }

  static void bfs_multi(const SubGraph& sg, const std::vector<int>& sources, std::vector<int>& dist_v, Queue& queue) {
    int* q = queue.q.data();
    int* dist = dist_v.data();
    size_t tail = 0;
    for (int v : sources) {
      dist[v] = 0;
      q[tail++] = v;
    }
    const int* xadj = sg.xadj.get();
    const int* adj = sg.adj.get();
    for (size_t head = 0; head < tail; ++head) {
      prefetch_ahead(q, head, tail, xadj, adj);
      const int v = q[head];
      const int dv = dist[v] + 1;
      for (int e = xadj[v]; e < xadj[v + 1]; ++e) {
        const int u = adj[e];
        const int du = dist[u];
        const bool fresh = du < 0;
        q[tail] = u;
        tail += fresh;
        dist[u] = fresh ? dv : du;
      }
    }
    queue.tail = tail;
  }

  static void emit_leaf(const std::vector<int>& verts, NDResult& out) {
    out.order.insert(out.order.end(), verts.begin(), verts.end());
    out.sn_sizes.push_back((int)verts.size());
  }

  // Hopcroft-Karp maximum matching + Koenig construction on the cut graph (SL x SR); local numbers in and out.
  static std::vector<int> min_vertex_cover(const SubGraph& sg, const std::vector<int>& SL, const std::vector<int>& SR,
                                           const std::vector<unsigned char>& side, std::vector<int>& rid) {
    const int nl = (int)SL.size(), nr = (int)SR.size();
    if (nl == 0 || nr == 0) return {};
    // position of a right-side vertex in SR (scratch, reset afterwards)
    for (int j = 0; j < nr; ++j) rid[SR[j]] = j;
    std::vector<int> xadj(nl + 1, 0), adj;
    for (int i = 0; i < nl; ++i) {
      const int v = SL[i];
      for (int e = sg.xadj[v]; e < sg.xadj[v + 1]; ++e) {
        const int u = sg.adj[e];
        if (side[u] == 1 && rid[u] >= 0) adj.push_back(rid[u]);
      }
      xadj[i + 1] = (int)adj.size();
    }
    for (int j = 0; j < nr; ++j) rid[SR[j]] = -1;
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

  // compact copy of the subgraph induced by `keep` (local numbers of `sg`, in the order the child visits them)
  static void induce(const SubGraph& sg, const std::vector<int>& keep, std::vector<int>& loc, SubGraph& out) {
    const int nc = (int)keep.size();
    out.gid.resize(nc);
    out.xadj.reset(new int[(size_t)nc + 1]);
    out.xadj[0] = 0;
    for (int i = 0; i < nc; ++i) loc[keep[i]] = i;
    const bool par = nc > 200000;
    const int* pxadj = sg.xadj.get();
    const int* padj = sg.adj.get();
    std::vector<int> deg(nc);
#pragma omp taskloop grainsize(8192) shared(deg, loc, keep, out, sg) if (par)
    for (int i = 0; i < nc; ++i) {
      const int v = keep[i];
      out.gid[i] = sg.gid[v];
      int c = 0;
      for (int e = pxadj[v]; e < pxadj[v + 1]; ++e) c += loc[padj[e]] >= 0;
      deg[i] = c;
    }
    for (int i = 0; i < nc; ++i) out.xadj[i + 1] = out.xadj[i] + deg[i];
    out.adj.reset(new int[(size_t)out.xadj[nc]]);
#pragma omp taskloop grainsize(8192) shared(loc, keep, out, sg) if (par)
    for (int i = 0; i < nc; ++i) {
      const int v = keep[i];
      int o = out.xadj[i];
      for (int e = pxadj[v]; e < pxadj[v + 1]; ++e) {
        const int u = loc[padj[e]];
        if (u >= 0) out.adj[o++] = u;
      }
    }
    for (int i = 0; i < nc; ++i) loc[keep[i]] = -1;
  }

  // One dissection step.  false: the node stays one front (emitted to `out`; nothing for an empty node);
  // true: it was split into the compact sub-graphs sgL, sgR and the separator Sg (global numbers), `sg` is released.
  bool split(SubGraph& sg, int depth, NDResult& out, SubGraph& sgL, SubGraph& sgR, std::vector<int>& Sg) {
    const int nv = sg.nv();
    if (nv == 0) return false;
    if (nv <= opt_.leaf_size) {
      emit_leaf(sg.gid, out);
      return false;
    }
    std::vector<int> da(nv, -1), db(nv, -1);
    const bool tr = depth <= 1 && getenv("LSA_TRACE_ANALYZE") != nullptr;
    double tt = tr ? now() : 0.0;
    auto lap = [&](const char* what) {
      if (!tr) return;
      const double t = now();
      fprintf(stderr, "[dissect depth %d, %d vertices] %-18s %7.3f s\n", depth, nv, what, t - tt);
      tt = t;
    };
    Queue queue(nv);
    int far0 = bfs(sg, 0, da, queue);
    lap("first sweep");
    std::vector<int> L, R, S;   // local numbers
    if ((int)queue.size() < nv) {
      // disconnected: distribute the components over two bins, no separator needed
      std::vector<std::pair<int, int>> comps;  // (start, end) in queue
      comps.emplace_back(0, (int)queue.size());
      for (int v = 0; v < nv; ++v)
        if (da[v] < 0) {
          int s = (int)queue.size();
          bfs(sg, v, da, queue);
          comps.emplace_back(s, (int)queue.size());
        }
      std::sort(comps.begin(), comps.end(),
                [](auto& a, auto& b) { return (a.second - a.first) > (b.second - b.first); });
      for (auto& c : comps) {
        auto& bin = (L.size() <= R.size()) ? L : R;
        bin.insert(bin.end(), queue.q.begin() + c.first, queue.q.begin() + c.second);
      }
    } else {
      std::vector<double> key(nv);
      bool have_key = false;
      if (opt_.dim > 0 && opt_.coords) {
        int best = 0;
        double ext = -1;
        for (int a = 0; a < opt_.dim; ++a) {
          double lo = 1e300, hi = -1e300;
          for (int v : sg.gid) {
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
          for (int i = 0; i < nv; ++i) key[i] = opt_.coords[(size_t)sg.gid[i] * opt_.dim + best];
          have_key = true;
        }
      }
      if (!have_key) {
        // two pseudo-peripheral vertices a, b; key = d(a, v) - d(b, v)
        std::fill(da.begin(), da.end(), -1);
        queue.clear();
        int b = bfs(sg, far0, da, queue);
        lap("  sweep from far0");
        queue.clear();
        int a2 = bfs(sg, b, db, queue);
        lap("  sweep from b");
        // one more sweep improves the pair on elongated domains
        std::fill(da.begin(), da.end(), -1);
        queue.clear();
        bfs(sg, a2, da, queue);
        lap("  sweep from a2");
        // da = d(a, .), db = d(b, .).  The bisector of two POINTS is slanted on meshes whose hop metric
        // is anisotropic (structured triangulations); the bisector of the two END CAPS
        //   A = {v : d(b, v) >= (1 - eps) D},  B = {v : d(a, v) >= (1 - eps) D}
        // is perpendicular to the long axis whatever the metric: multi-source BFS from the caps.
        if (opt_.cap_fraction > 0.0) {
          int D = 0;
          for (int v = 0; v < nv; ++v) D = std::max(D, da[v]);
          const int thr = (int)std::floor((1.0 - opt_.cap_fraction) * D);
          std::vector<int> capA, capB;
          for (int v = 0; v < nv; ++v) {
            if (db[v] >= thr) capA.push_back(v);
            if (da[v] >= thr) capB.push_back(v);
          }
          if (!capA.empty() && !capB.empty()) {
            std::fill(da.begin(), da.end(), -1);
            std::fill(db.begin(), db.end(), -1);
            // the two cap sweeps are independent: side by side when the node is large (near the root the other
            // threads of the team have nothing else to do yet)
            if (nv > 100000) {
              Queue queue_b(nv);
#pragma omp task shared(sg, capA, da, queue)
              bfs_multi(sg, capA, da, queue);
              bfs_multi(sg, capB, db, queue_b);
#pragma omp taskwait
            } else {
              bfs_multi(sg, capA, da, queue);
              bfs_multi(sg, capB, db, queue);
            }
            lap("  sweeps from the caps");
          }
        }
        for (int i = 0; i < nv; ++i) key[i] = (double)da[i] - (double)db[i];
      }
      lap("bisection key");
      std::vector<double> tmp(key);
      std::nth_element(tmp.begin(), tmp.begin() + nv / 2, tmp.end());
      const double t = tmp[nv / 2];
      std::vector<double>().swap(tmp);
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
        emit_leaf(sg.gid, out);
        return false;
      }
      std::vector<unsigned char> side(nv);
      for (int i = 0; i < nv; ++i) side[i] = use_le ? (key[i] > t) : (key[i] >= t);
      std::vector<double>().swap(key);
      std::vector<int> SL, SR;
      for (int v = 0; v < nv; ++v) {
        const unsigned char sv = side[v];
        bool cut = false;
        for (int e = sg.xadj[v]; e < sg.xadj[v + 1] && !cut; ++e) cut = side[sg.adj[e]] != sv;
        if (cut) (sv ? SR : SL).push_back(v);
      }
      // minimum vertex separator for this edge cut = minimum vertex cover of the bipartite graph of
      // cut edges between the two boundary layers (Koenig's theorem via Hopcroft-Karp matching).
      // On FE graphs (every element is a clique) this is about half of either boundary layer.
      lap("edge cut");
      std::fill(db.begin(), db.end(), -1);
      S = min_vertex_cover(sg, SL, SR, side, db);
      lap("vertex cover");
      if ((double)S.size() > 0.45 * nv) {
        emit_leaf(sg.gid, out);
        return false;
      }
      for (int v : S) side[v] = 2;
      for (int v = 0; v < nv; ++v) {
        if (side[v] == 0) L.push_back(v);
        else if (side[v] == 1) R.push_back(v);
      }
    }
    std::vector<int>().swap(queue.q);
    std::vector<int>().swap(da);
    // children's compact copies, then this node's adjacency is released
    std::fill(db.begin(), db.end(), -1);
    induce(sg, L, db, sgL);
    induce(sg, R, db, sgR);
    lap("children's graphs");
    Sg.resize(S.size());
    for (size_t i = 0; i < S.size(); ++i) Sg[i] = sg.gid[S[i]];
    sg.adj.reset();
    sg.xadj.reset();
    std::vector<int>().swap(sg.gid);
    return true;
  }

  // Dissection of a sub-graph: order = [order of L][order of R][separator].  The LARGER child is continued in a
  // loop, the smaller one is recursed into (as a task near the root), so the call depth is bounded by log2(n)
  // whatever the splits look like: graphs with hub vertices peel one vertex per step (n steps deep), which must
  // not be n stack frames.  Pieces are concatenated at the end in the order of the plain recursion.
  struct Pending {
    SubGraph sg;            // the smaller child
    NDResult res;           // its ordering
    std::vector<int> sep;   // separator of the step
    bool small_first;       // the smaller child is L: its ordering precedes the continued child's
  };

  void rec(SubGraph& sg0, int depth, NDResult& out) {
    std::deque<Pending> pend;   // stable addresses: tasks hold references into it
    SubGraph cur = std::move(sg0);
    NDResult core;
    for (;; ++depth) {
      const int nv = cur.nv();
      SubGraph sgL, sgR;
      std::vector<int> Sg;
      if (!split(cur, depth, core, sgL, sgR, Sg)) break;
      pend.emplace_back();
      Pending& p = pend.back();
      p.sep.swap(Sg);
      p.small_first = sgL.nv() < sgR.nv();
      p.sg = std::move(p.small_first ? sgL : sgR);
      cur = std::move(p.small_first ? sgR : sgL);
      const bool spawn = nv > 20000;
      Pending* pp = &p;   // the task outlives this iteration: it takes the element's address by value
#pragma omp task firstprivate(pp, depth) if (spawn)
      rec(pp->sg, depth + 1, pp->res);
    }
#pragma omp taskwait
    size_t total = core.order.size();
    for (const Pending& p : pend) total += p.res.order.size() + p.sep.size();
    out.order.reserve(out.order.size() + total);
    auto append = [&](const NDResult& r) {
      out.order.insert(out.order.end(), r.order.begin(), r.order.end());
      out.sn_sizes.insert(out.sn_sizes.end(), r.sn_sizes.begin(), r.sn_sizes.end());
    };
    for (const Pending& p : pend)
      if (p.small_first) append(p.res);
    append(core);
    for (auto it = pend.rbegin(); it != pend.rend(); ++it) {
      if (!it->small_first) append(it->res);
      if (!it->sep.empty()) emit_leaf(it->sep, out);
    }
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
  if (g.xadj[n] > 0x7fffffffLL) throw std::runtime_error("graph too large for the dissection's 32-bit adjacency offsets");
  const bool trace_an = getenv("LSA_TRACE_ANALYZE") != nullptr;
  {
    const double ta = now();
    Dissector d(g, opt);
    d.run(rest, nd);
    if (trace_an) fprintf(stderr, "[analyze] nested dissection            %8.3f s\n", now() - ta);
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
    int moved = 0;
    if (opt.order_last) {
      const int n_ord = (int)nd.order.size();
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : moved)
      for (int q_ord = 0; q_ord < n_ord; ++q_ord) {
        const int v = nd.order[q_ord];
        target[v] = sn_tmp[v];
        if (!opt.order_last[v]) continue;
        std::vector<int> nb;
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
  if (trace_an) fprintf(stderr, "[analyze] graph %.3f s, ordering phase %.3f s\n", sym.seconds[0], sym.seconds[1]);

  // ---- supernodal symbolic factorisation on the permuted pattern of A + A^T
  const int ns = sym.ns;
  std::vector<int> parent(ns, -1), first_child(ns, -1), next_sib(ns, -1), last_child(ns, -1);
  std::vector<std::vector<int>> st(ns);
  // (a) rows a front gets from the matrix itself: every front on its own, in parallel (one mark array per thread)
#pragma omp parallel
  {
    std::vector<int> mark(n, -1);
#pragma omp for schedule(dynamic, 64)
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
    }
  }
  // (b) rows inherited from the children, in elimination order (a front's parent is known once its rows are)
  {
    std::vector<int> mark(n, -1);
    for (int s = 0; s < ns; ++s) {
      const int last = sym.sn_ptr[s + 1] - 1;
      std::vector<int>& rows = st[s];
      if (first_child[s] >= 0) {
        for (int j : rows) mark[j] = s;
        for (int c = first_child[s]; c >= 0; c = next_sib[c])
          for (int j : st[c])
            if (j > last && mark[j] != s) {
              mark[j] = s;
              rows.push_back(j);
            }
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
