// Internal data structures shared by the host symbolic phase and the CUDA numeric phase.
//
// Path covered: the work SLEPc/PETSc/MUMPS do behind `iEpsSolver.solve()`
// (reference Solver/utils.py:268-270): sparse LU of A - sigma M, triangular solves, SpMV with M,
// Krylov-Schur.  Nothing here is derived from those libraries' sources (they are not vendored in
// the reference); the design is a static-structure multifrontal method laid out for one B200.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace lsa {

// Vector for the large per-entry host arrays of the analysis (scatter maps, permuted patterns): resize() leaves
// trivially constructible elements uninitialised, so the pages of a fresh array are first touched by the OpenMP
// threads that fill it instead of being zero-filled by one thread (measured on config 3: 2.8 s of 3.3 s in
// those two stages were page faults of serial zero-fills).  Every user writes all entries after a resize().
template <class T>
struct default_init_allocator : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = default_init_allocator<U>;
  };
  using std::allocator<T>::allocator;
  template <class U, class... Args>
  void construct(U* p, Args&&... args) {
    if constexpr (sizeof...(Args) == 0)
      ::new (static_cast<void*>(p)) U;
    else
      ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...);
  }
};
template <class T>
using hvec = std::vector<T, default_init_allocator<T>>;

// One frontal matrix of the multifrontal LU.  A front with k pivots and r remaining rows is
// stored as three dense column-major pieces:
//   P  (k+r) x k   [F11; F21]  -> after factorisation L11\U11 on top, L21 below   (factor store)
//   Q   k    x r    F12        -> after factorisation U12                          (factor store)
//   C   r    x r    F22        -> Schur complement / contribution block            (level pool)
struct Front {
  long long p_off;  // element offset of P in the factor store
  long long q_off;  // element offset of Q in the factor store
  long long c_off;  // element offset of C in the contribution pool of parity (level & 1)
  long long st0;    // offset of this front's row structure in st_idx / ea_map / cb vector
  int k;            // number of pivots (fully summed variables)
  int r;            // number of remaining (not fully summed) rows
  int col0;         // first permuted column index
  int parent;       // parent front (-1: root)
  int level;        // depth in the assembly tree (roots are level 0)
  int child0;       // first entry in child_idx
  int nchild;       // number of children
  int flags;        // partitioned solve: 0 regular; FRONT_GHOST: root of a sub-tree owned by ANOTHER GPU, present
                    // only as a contribution block (k = 0); FRONT_CUT: root of one of this GPU's sub-trees.
                    // The contribution block of a flagged front lives in the cut pool (c_off refers to it), and
                    // the sweeps hand its contribution vector over through the all-reduce of the replicated rows.
};
enum { FRONT_REGULAR = 0, FRONT_GHOST = 1, FRONT_CUT = 2 };

struct Symbolic {
  int n = 0;        // matrix order
  int n_iso = 0;    // decoupled 1x1 pivots (e.g. Dirichlet identity rows); permuted first
  int ns = 0;       // number of fronts (supernodes)
  int nlevels = 0;
  std::vector<int> perm;    // perm[new] = old
  std::vector<int> iperm;   // iperm[old] = new
  std::vector<int> sn_ptr;  // ns+1 permuted column ranges; sn_ptr[0] == n_iso
  std::vector<int> sn_of;   // n: front owning a permuted column (-1 for decoupled pivots)
  std::vector<long long> st_ptr;  // ns+1
  std::vector<int> st_idx;        // sorted permuted row indices below each front's pivot block
  std::vector<int> ea_map;        // same indexing as st_idx: local row index in the parent front
  std::vector<int> child_idx;     // children of each front, ascending
  std::vector<int> lvl_ptr;       // nlevels+1
  std::vector<int> lvl_front;     // fronts of each level sorted by descending k
  std::vector<Front> fronts;
  hvec<long long> a_dst;   // per entry of the input CSR pattern: destination in the factor store
  long long fac_size = 0;         // elements in the factor store (P and Q of all fronts + decoupled pivots)
  long long diag_off = 0;         // offset of the decoupled pivots in the factor store
  long long pool_size[2] = {0, 0};
  long long nnz_lu = 0;           // sum k^2 + 2 k r + n_iso (algorithmic factor entries)
  double flops = 0.0;             // sum 2/3 k^3 + 2 k^2 r + 2 k r^2 (real-arithmetic flops for T = double)
  int max_k = 0, max_m = 0, max_r = 0;
  double seconds[4] = {0, 0, 0, 0};  // graph, ordering, structure, maps
  bool symmetric = false;         // symmetric factorisation F = L D L^T: no Q (U12) blocks in the factor store
};

// Split of ONE factorisation / solve over the GPUs of a node (stage 1): proportional mapping of the
// assembly tree -- whole sub-trees go to single GPUs (no communication inside), the fronts above the cut ("top")
// are REPLICATED: every GPU factors and sweeps them redundantly on identical data.  Exchange steps: the
// contribution blocks of the sub-tree roots are broadcast once per factorisation; per operator application ONE
// all-reduce over the replicated rows carries both the SpMV partial sums of those rows and the sub-tree roots'
// contribution vectors.
struct Partition {
  int rank = 0, world = 1;
  std::vector<int> owner;       // per GLOBAL front: -1 replicated top, else the owning GPU
  std::vector<int> l2g, g2l;    // local front id <-> global front id (-1: not present on this rank)
  std::vector<int> cut_roots;   // GLOBAL ids of the sub-tree roots with a (top) parent, ascending: broadcast order
  int n_top_levels = 0;         // local levels [0, n_top_levels) hold the top fronts, deeper ones this GPU's sub-trees
  std::vector<int> own_lo, own_hi;   // permuted row ranges [lo, hi) of this GPU's sub-trees
  std::vector<int> top_lo, top_hi;   // replicated rows: decoupled pivots first, then the top fronts
  long long n_top_rows = 0;
  long long cut_pool_size = 0;  // entries of the cut pool (contribution blocks of ALL sub-tree roots)
  double weight_total = 0, weight_top = 0, weight_max = 0, weight_mine = 0;   // work model (see partition.cpp)
  int ns_global = 0;
  long long nnz_lu_global = 0;
  double flops_global = 0;
};

// partition.cpp: `global` -> this rank's local symbolic structures (only the fronts this GPU touches).
void partition(const Symbolic& global, int rank, int world, Symbolic& local, Partition& part);

struct AnalyzeOptions {
  int leaf_size = 64;
  int dim = 0;                   // >0: geometric bisection with `coords` (n x dim, row-major)
  const double* coords = nullptr;
  const unsigned char* order_last = nullptr;  // n flags: order this unknown last inside its front
  int nthreads = 0;
  double cap_fraction = 0.15;     // end-cap thickness (share of the diameter) of the graph bisector; 0: point pair
  double coupled_fraction = 1.0;  // share of a flagged unknown's regular neighbours eliminated before it
  bool symmetric = false;         // lay the factor store out for F = L D L^T (P blocks only: half the entries)
};

// Host symbolic phase: nested dissection, supernode partition, row structures, assembly tree,
// scatter maps.  `rowptr`/`colidx`: CSR pattern of A (union with M), n rows.
void analyze(int n, const long long* rowptr, const int* colidx, const AnalyzeOptions& opt, Symbolic& sym);

}  // namespace lsa
