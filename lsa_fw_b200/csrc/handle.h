// Solver handle: host symbolic data, device plan, device buffers.  One handle = one stream.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/lsa_b200.h"
#include "comm.h"
#include "common.cuh"
#include "lsa_internal.h"

namespace lsa {

// CSR operator in the PERMUTED ordering, values gathered from the caller's entry order.
struct CsrHost {
  std::vector<long long> rowptr;
  hvec<int> colidx;
  hvec<long long> src;  // permuted entry -> original entry
};
struct CsrDev {
  long long nnz = 0;
  long long* rowptr = nullptr;
  int* colidx = nullptr;
  long long* src = nullptr;
  int* rowblk = nullptr;   // row blocks of the streamed SpMV (krylov.cu: spmv_row_blocks), n_rowblk + 1 entries
  long long* blk_e0 = nullptr;   // first entry of every row block (rowptr[rowblk[.]]), n_rowblk + 1 entries
  int n_rowblk = 0;
  int block_entries = 0;   // entries per row block the blocks were cut for (512, 1024 or 2048)
  void* vals = nullptr;  // permuted values, double or z128 (is_complex)
  bool is_complex = false;
};

struct DevStats {
  unsigned long long n_perturbed;
  unsigned long long n_swaps;
  unsigned long long min_piv_bits;  // bit pattern of a non-negative double
  unsigned long long max_piv_bits;
  unsigned long long max_l_bits;    // largest |multiplier| met (element growth monitor)
  int zero_pivot;
  int nonfinite;
};

// Result block written by the Rayleigh-Ritz kernel, read back once per restart.
struct RrInfo {
  int nconv;      // length of the converged leading run
  int keep;       // columns kept for the restart (nconv + l)
  int status;     // 0 ok, 1 QR iteration did not converge
  int pad;
};

// Sweep plan: which kernel family handles a chunk of a tree level (solve.cu)
enum SolveMode { SOLVE_STREAM = 0, SOLVE_INVERTED = 1, SOLVE_STEPS = 2 };
struct SolveChunk {
  int level, first, cnt;    // position in lvl_front
  int maxk, max_r, max_m;   // extents over the chunk's fronts
  int mode;                 // SolveMode
};

struct Launch {
  int kind;   // see factor.cu
  int level;
  int j0;
  int gx;     // tiles
  int pad;
};

struct lsa_handle_impl {
  int n = 0;
  int device = -1;
  cudaStream_t stream = nullptr;
  void* h_stage[2] = {nullptr, nullptr};        // page-locked staging buffers for uploads from pageable memory
  cudaEvent_t ev_stage[2] = {nullptr, nullptr};
  std::string err;

  // ---- host symbolic
  Symbolic sym;
  bool analyzed = false;
  bool has_m = false;
  CsrHost hA, hM, hAt, hMt;     // permuted patterns (transposes built on demand)
  hvec<long long> m_dst;  // like sym.a_dst for the entries of M
  long long nnz_a = 0, nnz_m = 0;

  // ---- device plan
  Front* d_fronts = nullptr;
  int* d_lvl_front = nullptr;
  int num_sms = 148;
  int* d_st_idx = nullptr;
  int* d_ea_map = nullptr;
  int* d_child_idx = nullptr;
  long long* d_a_dst = nullptr;
  long long* d_m_dst = nullptr;
  int* d_perm = nullptr;
  int* d_ipiv = nullptr;
  int* d_gperm = nullptr;   // composed row interchanges of all fronts (gather form)
  DevStats* d_stats = nullptr;
  std::vector<int> lvl_maxchild;   // per level: max number of children of its fronts
  std::vector<int> lvl_maxk;       // per level: largest pivot count
  std::vector<int> lvl_maxrchild;  // per level: largest child contribution block order

  // ---- partitioned solve over the GPUs of a node (stage 1: sub-trees per GPU, replicated top; lsa_internal.h)
  int part_rank = 0, part_world = 1;     // lsa_set_partition, before lsa_analyze
  bool partitioned = false;
  bool part_graphs = true;               // partitioned sweeps (incl. their NCCL all-reduce) replayed from CUDA graphs
  Partition part;
  Comm comm;
  void* d_cut_pool = nullptr;            // contribution blocks of ALL sub-tree roots (own: computed, others: broadcast)
  long long cut_pool_capacity_bytes = 0;
  int* d_top_rows = nullptr;             // replicated rows (decoupled pivots + top fronts), permuted indices
  z128* d_topbuf = nullptr;              // those rows packed: payload of the per-solve all-reduce
  int* d_cut_own = nullptr;              // local ids of this rank's sub-tree roots that hang below a top front
  int n_cut_own = 0;
  std::vector<std::pair<int, int>> upd_ranges, dot_ranges;   // rows this rank maintains / counts in dot products
  double* d_red = nullptr;               // all-reduce payload of a Gram-Schmidt pass: h (<= 256 complex), |w|^2

  // ---- values
  void* d_a_orig = nullptr;  // caller entry order
  void* d_m_orig = nullptr;
  bool a_complex = false, m_complex = false;
  bool have_values = false;
  CsrDev dA, dM, dAt, dMt;
  double a_fro = 0.0, a_max = 0.0, m_max = 0.0;

  // ---- factor
  int scalar = -1;  // lsa_scalar of the current factor, -1: none
  void* d_fac = nullptr;
  long long fac_capacity_bytes = 0;
  void* d_pool[2] = {nullptr, nullptr};
  long long pool_capacity_bytes[2] = {0, 0};
  lsa_factor_stats fstats{};
  z128 f_alpha{1, 0}, f_beta{0, 0};

  // ---- solve / Krylov work space (permuted ordering, complex)
  z128* d_x = nullptr;     // n
  z128* d_w = nullptr;     // n
  z128* d_t = nullptr;     // n
  z128* d_t2 = nullptr;    // n (second sweep vector)
  z128* d_cb = nullptr;    // total structure length: per-front contribution vectors
  z128* d_io = nullptr;    // n staging for host vectors
  z128* d_V = nullptr;     // n x (ncv+1)
  long long V_cols = 0;
  z128* d_S = nullptr;     // (ncv+1) x ncv projected matrix, ld = ncv+1
  z128* d_Q = nullptr;     // ncv x ncv
  z128* d_part = nullptr;  // partial dot products [blocks][128]
  double* d_npart = nullptr;  // partial squared norms
  z128* d_h = nullptr;     // CGS coefficients
  z128* d_brow = nullptr;
  z128* d_U = nullptr;     // M-inner-product mode: U = M V, n x (ncv + 1)
  z128* d_mw = nullptr;    // M w work vector (n)
  z128* d_bn = nullptr;    // w^H M w (one complex number)
  int U_cols = 0;
  z128* d_ywork = nullptr;
  z128* d_r1 = nullptr;    // refinement / residual scratch vectors
  z128* d_r2 = nullptr;
  z128* d_r3 = nullptr;
  z128* d_Xp = nullptr;    // n x ncv Ritz vectors, permuted ordering
  int* d_flag = nullptr;
  int* d_refine = nullptr;    // [0] second Gram-Schmidt pass wanted for the current column, [1] how many were
  double* d_wn2 = nullptr;    // partial |w|^2 before orthogonalisation (refinement criterion)
  double* d_wn2b = nullptr;   // partial |w|^2 after the first pass (fused update + dots kernel)
  bool fuse_ortho = true;     // pass-1 update fused with the pass-2 dot products
  bool ortho_refine_always = false;
  int* d_ipart = nullptr;  // arg-max partial indices
  RrInfo* d_rr = nullptr;
  z128* d_theta = nullptr;
  double* d_resid = nullptr;
  int ncv_alloc = 0;
  z128* d_ns = nullptr;    // attached nullspace: n x ns_count orthonormal columns, permuted ordering
  int ns_count = 0;

  // ---- results
  int nconv = 0;
  std::vector<z128> eigenvalues;   // back-transformed, `which` order
  std::vector<int> eig_order;      // column of d_X for each returned pair
  z128* d_X = nullptr;             // n x nconv eigenvectors, ORIGINAL ordering
  int X_cols = 0;
  lsa_eigs_params last_params{};
  lsa_counters counters{};
  long long launch_count = 0;      // kernels launched (running total)

  // captured solve sweeps (CUDA graphs), keyed by (trans, vector); invalidated by every factorisation
  struct SolveGraph {
    int trans;
    const void* vec;
    cudaGraphExec_t exec;
    int launches;
  };
  std::vector<SolveGraph> solve_graphs;
  std::vector<SolveChunk> solve_plan;   // levels in root-to-leaf order, rebuilt at every factorisation
  int invert_max_k = 8192;              // levels whose pivot blocks are at most this wide get them inverted as a whole
  void* d_inv_scratch = nullptr;        // scratch of the block-inverse merges
  long long inv_scratch_bytes = 0, inv_scratch_entries = 0;
  long long* d_inv_off = nullptr;       // per-front scratch offsets of the current batch
  z128* d_tri_scratch = nullptr;        // partial sums of the split triangular GEMVs
  int* d_tri_tickets = nullptr;         // arrival counters of their row chunks (self-resetting)
  long long tri_slots = 0;
  int tri_span = 512;                   // input entries per CTA of a split triangular GEMV (0: never split)
  bool use_graphs = true;
  bool defer_cb = true;          // cluster up sweeps: contribution rows updated by one wide GEMV after the pivot steps
  bool cluster_slices = true;    // levels with <= 9 fronts: 16-CTA clusters sharing every 128-row block by 8-row slices
  int cluster_max_width = 16;   // CTAs per front in the cluster sweep (16 = non-portable cluster size)
  int cluster_max_rows = 8192;  // fronts taller than this are swept with one grid-wide launch per step (measured optimum)
  bool use_stream = true;     // single-step levels: bulk-copy/mbarrier streamed kernel (complex factors)
  int stream_stages = 0;      // ring depth of the streamed kernel (0 = by level size)
  int stream_flags = 3;       // bit 0: wider tiles for narrow blocks, bit 1: single-copy tiles for contiguous blocks
  int stream_small_rows = 192; // levels whose fronts have at most this many rows use the small-CTA variant
  int stream_min_fronts = 192;// multi-step levels with at least this many fronts are streamed too (one CTA per front)
  bool use_clusters = true;
  double coupled_fraction = 0.5;
  int spmv_block = 1024;      // entries per row block of the streamed SpMV (option "spmv_block": 512, 1024, 2048)
  bool symmetric = false;     // option "symmetric" (before lsa_analyze): F = L D L^T, real FP64, no pivoting, half the factor store
};

}  // namespace lsa

struct lsa_handle : lsa::lsa_handle_impl {};
