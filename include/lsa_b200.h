/* lsa_b200.h -- C ABI of the B200-native shift-and-invert eigensolve backend.
 *
 * The reference (ferdean/lsa-fw) has no FFI of its own: the seam is the Python class API
 * `Solver/eigen.py:48-155` + `Solver/utils.py:190-328`, whose numerical work is done by
 * slepc4py/petsc4py calls.  Each entry point below names the reference call it stands in
 * for; the Python host (lsa_fw_b200/_lib.py, driven by lsa_fw_b200/utils.py) binds them with ctypes.  INTEGRATION.md shows
 * the stub a maintainer of the reference would add.
 *
 * Conventions: every function returns 0 on success, a negative lsa_status otherwise, with a
 * text in lsa_last_error().  Plain pointers + sizes only.  Host or device pointers are both
 * accepted where an `on_device` flag exists (device pointers come from DLPack / torch on the
 * Python side).  One handle = one solver object = one CUDA stream; not thread-safe (neither is
 * the reference: one EPS per Python object, Solver/utils.py:203).
 * Complex numbers are interleaved (re, im) doubles.  Indices are int32, row pointers int64.
 */
#ifndef LSA_B200_H
#define LSA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lsa_handle lsa_handle;

enum lsa_status {
  LSA_OK = 0,
  LSA_ERR_ARG = -1,         /* bad argument / call order                                  */
  LSA_ERR_CUDA = -2,        /* CUDA runtime error or no device (there is NO CPU fallback)  */
  LSA_ERR_ZERO_PIVOT = -3,  /* exactly singular pivot (PETSc.Error "zero pivot" in the reference) */
  LSA_ERR_NONFINITE = -4,   /* NaN/Inf met in the factor or the Krylov basis (eigen2.py:186-189) */
  LSA_ERR_INTERNAL = -5
};

enum lsa_scalar { LSA_F64 = 0, LSA_C128 = 1 };
enum lsa_trans { LSA_OP_N = 0, LSA_OP_T = 1, LSA_OP_H = 2 };
enum lsa_matrix { LSA_MAT_A = 0, LSA_MAT_M = 1 };

/* EPSWhich of SLEPc as exposed by iEpsWhich (Solver/utils.py:152-187). */
enum lsa_which {
  LSA_LARGEST_MAGNITUDE = 1,
  LSA_SMALLEST_MAGNITUDE = 2,
  LSA_LARGEST_REAL = 3,
  LSA_SMALLEST_REAL = 4,
  LSA_LARGEST_IMAGINARY = 5,
  LSA_SMALLEST_IMAGINARY = 6,
  LSA_TARGET_MAGNITUDE = 7,
  LSA_TARGET_REAL = 8,
  LSA_TARGET_IMAGINARY = 9
};

/* Spectral transformation (iSTType, Solver/utils.py:131-149); only these two are built. */
enum lsa_transform { LSA_ST_SHIFT = 0, LSA_ST_SINVERT = 1 };

typedef struct {
  int32_t n, n_decoupled, n_fronts, n_levels;
  int32_t max_pivots, max_front, max_rows, pad;
  int64_t nnz_a, nnz_m;
  int64_t factor_entries;  /* elements allocated for the factor store                         */
  int64_t nnz_lu;          /* algorithmic entries of L+U: sum k^2 + 2 k r (+ decoupled pivots) */
  int64_t pool_entries[2]; /* contribution-block pools                                        */
  int64_t struct_entries;  /* total front row-structure length                                */
  double flops_real;       /* sum 2/3 k^3 + 2 k^2 r + 2 k r^2 ; x4 for complex                */
  double seconds[4];       /* graph, ordering, structure, maps                                */
} lsa_symbolic_info;

typedef struct {
  double seconds;          /* device time of the numeric factorisation (CUDA events)  */
  double flops;            /* real flops executed (flops_real x 4 when complex)        */
  int64_t n_perturbed;     /* tiny pivots replaced (static pivoting, cf. MUMPS CNTL(3) in eigen2.py:135-136) */
  int64_t n_row_swaps;     /* pivot rows exchanged inside fronts                       */
  int32_t scalar;          /* lsa_scalar actually used                                  */
  int32_t n_kernels;       /* kernel launches issued                                    */
  double min_pivot, max_pivot;
  double max_multiplier;   /* largest |l_ij| met: element-growth monitor of the restricted pivoting */
} lsa_factor_stats;

typedef struct {
  int32_t nev, ncv, max_restarts;
  int32_t which;           /* lsa_which, applied to the back-transformed eigenvalue    */
  int32_t transform;       /* lsa_transform                                             */
  int32_t adjoint;         /* 1: left problem (A^H, M^H) at conj(sigma) on the SAME factors
                              (what Sensitivity/__init__.py:246-262 obtains by re-factorising) */
  int32_t purify;          /* singular M: Ritz vectors multiplied once by OP.  1: through the Krylov-Schur
                              relation x = V y + v_next (b.y)/theta (no operator application, what SLEPc's
                              EPSComputeVectors does); 2: one explicit OP apply per vector; 0: off */
  int32_t refine_steps;    /* iterative-refinement steps inside each OP apply           */
  double tol;              /* relative: beta |s_i| <= tol |theta_i|  (EigensolverConfig.atol is
                              handed to SLEPc as its relative tol, Solver/utils.py:236-238) */
  double sigma_re, sigma_im;    /* shift / target                                        */
  uint64_t seed;                /* start vector: splitmix64 Gaussian stream, OP applied once */
  const double* v0;             /* optional host start vector (n complex); NULL = seeded random */
  int32_t b_mode;               /* Hermitian-definite problems (EPS_GHEP, M Hermitian positive (semi)definite):
                                   0: Euclidean inner products, unit 2-norm vectors (general problems);
                                   1: as 0, returned vectors scaled to unit M-norm on the device (x^H M x = 1);
                                   2: M-inner products throughout -- M-orthonormal Krylov basis (symmetric
                                      Lanczos / Krylov-Schur: the projected matrix is Hermitian, as SLEPc's
                                      EPSKrylovSchur does for GHEP), M-orthonormal returned vectors.  Needs M. */
  int32_t pad_;
} lsa_eigs_params;

typedef struct {
  int32_t nconv, n_restarts, n_op_applies, breakdown;
  double seconds;               /* device time of the Krylov-Schur loop (CUDA events)    */
  double seconds_solve, seconds_spmv, seconds_ortho, seconds_rr, seconds_restart;
  int32_t n_kernels;
  int32_t n_reorth;             /* basis columns that needed the second Gram-Schmidt pass */
  int32_t n_arnoldi;            /* Arnoldi expansion steps (one Gram-Schmidt each)         */
  int32_t pad;
  int64_t sum_cols;             /* sum over those steps of the number of basis columns orthogonalised against:
                                   algorithmic bytes of the orthogonalisation =
                                   (1 + n_reorth / n_arnoldi) * (2 sum_cols + 5 n_arnoldi) * n * 16   */
} lsa_eigs_result;

typedef struct {
  double bytes_solve;     /* algorithmic bytes of one fwd+bwd triangular solve   */
  double bytes_spmv_m;    /* algorithmic bytes of one SpMV with M                              */
  double bytes_spmv_a;
  double factor_flops, factor_seconds;
  double solve_seconds;   /* mean device seconds of one fwd+bwd solve                          */
  double spmv_seconds;
  int64_t n_solves, n_spmv;
} lsa_counters;

/* -- lifetime ------------------------------------------------------------------------------
 * lsa_create  <- SLEPc.EPS().create(comm)                       Solver/utils.py:203
 * device >= 0: CUDA device ordinal.  device = -1: symbolic-only handle (no GPU touched); every
 * numeric call on such a handle fails with LSA_ERR_CUDA.                                      */
int lsa_create(int32_t n, int32_t device, lsa_handle** out);
void lsa_destroy(lsa_handle* h);
const char* lsa_last_error(const lsa_handle* h);
const char* lsa_version(void);

/* -- operators ------------------------------------------------------------------------------
 * lsa_analyze  <- eps.setOperators(A.raw, M.raw) + PCSetUp_LU symbolic part
 *                 (Solver/utils.py:216-221; MatGetOrdering/MatLUFactorSymbolic inside PETSc)
 * CSR patterns of A and M as delivered by iPETScMatrix.as_scipy_array() (FEM/utils.py:585-588).
 * m_rowptr == NULL: standard problem (M = I).  coords (n x dim doubles) and order_last (n bytes)
 * are optional hints, NULL allowed.  Runs on the host; reusable across shifts / Reynolds numbers. */
int lsa_analyze(lsa_handle* h, const int64_t* a_rowptr, const int32_t* a_colidx, const int64_t* m_rowptr,
                const int32_t* m_colidx, int32_t leaf_size, int32_t dim, const double* coords,
                const uint8_t* order_last, int32_t nthreads);
/* -- one factorisation / eigensolve split over the GPUs of a node -----------------------------------
 * The reference reaches several ranks through PETSc / SLEPc / MUMPS on PETSc.COMM_WORLD
 * (Solver/utils.py:196-203, README.md:43).  Here: one process per GPU (torchrun); every process creates a handle,
 * calls lsa_set_partition(rank, world) BEFORE lsa_analyze (each rank analyses the whole pattern, then keeps the
 * sub-trees mapped to it plus the replicated top of the assembly tree), and lsa_set_comm with a 128-byte NCCL
 * unique id obtained from lsa_nccl_unique_id on rank 0 and distributed by the caller (torch.distributed).  After
 * that the SAME calls as on one GPU (set_values with the full value arrays, factor, solve, eigs) run collectively:
 * all ranks must issue them in the same order.  NCCL traffic: broadcast of the sub-tree roots' contribution
 * blocks (once per factorisation), one all-reduce over the replicated rows per operator application (SpMV partial
 * sums + sub-tree contribution vectors in one payload), one small all-reduce per Gram-Schmidt pass (coefficients
 * and |w|^2 in one payload).  Vectors passed to / returned by lsa_solve, lsa_spmv and lsa_get_eigenvectors are
 * full-length and identical on every rank.                                                                      */
typedef struct {
  int32_t rank, world;
  int32_t n_fronts_global, n_fronts_local;   /* local: replicated top + own sub-trees + ghost roots            */
  int32_t n_top_fronts, n_top_levels, n_cut_roots, pad;
  int64_t n_replicated_rows, n_own_rows;
  int64_t cut_pool_entries;
  int64_t nnz_lu_global;                     /* lsa_symbolic_info.nnz_lu is this rank's share (top + own)       */
  double flops_real_global;
  double weight_total, weight_top, weight_max_subtrees, weight_mine;   /* work model of the mapping (seconds-like) */
} lsa_partition_info;
int lsa_set_partition(lsa_handle* h, int32_t rank, int32_t world);
int lsa_nccl_load(const char* path_or_null);      /* optional: where libnccl.so.2 lives (default: already loaded / ld path) */
int lsa_nccl_unique_id(void* out128);
int lsa_set_comm(lsa_handle* h, const void* id128);
int lsa_partition_info_get(const lsa_handle* h, lsa_partition_info* out);

/* Tuning knobs, to be set before lsa_analyze / lsa_factor:
 *   "symmetric" (default 0; before lsa_analyze): the pencil is real symmetric and the shift real (GHEP / HEP with
 *       st_pc_type CHOLESKY, Elasticity/utils.py:139-155, tests/benchmark/vibrating_membrane.py): F = L D L^T without
 *       pivoting (tiny pivots replaced, growth monitored), only the P blocks [L11 \ D L11^T; L21] are stored --
 *       nnz_lu = sum k^2 + k r instead of k^2 + 2 k r -- the Schur complements are formed as C -= L21 (D L21^T) with the
 *       scaled, transposed operand built in the GEMM loader, and a solve is  L-sweep, diagonal scaling, L^T-sweep.
 *       scalar must be LSA_F64; lsa_solve gives the same result for every `trans`;
 *   "coupled_fraction" (default 0.5): an unknown with a structurally zero diagonal (pressure) is eliminated
 *       no earlier than the front in which this share of its coupled regular unknowns has been eliminated
 *       (1.0 = all of them: most robust, ~+40-70 % flops in 2-D);
 *   "use_graphs" (default 1): replay the triangular-solve sweeps from CUDA graphs;
 *   "use_stream" (default 1): complex factors, levels with many fronts: streamed sweep kernel (bulk copies into
 *       a shared-memory ring); "stream_min_fronts" (192): multi-step levels with at least this many fronts are
 *       streamed too; "stream_small_rows" (192): levels whose fronts are at most this tall use the small CTA shape;
 *       "stream_stages" (0 = by level size, 2..12): ring depth; "stream_flags" (3): bit 0 wider tiles for narrow
 *       blocks, bit 1 single-copy tiles for contiguous blocks, bit 2 LDGSTS producer (measured slower);
 *   "use_clusters" (default 1): sweep the remaining multi-step levels with one thread-block cluster per front;
 *       "invert_max_k" (8192): levels that are not streamed and whose pivot blocks are at most this wide get those
 *       blocks inverted as a whole after the factorisation (one triangular matrix-vector product per front and
 *       sweep, no dependent 128-pivot steps; 0 = off, wider blocks keep the step kernels below);
 *       "cluster_max_width" (16): CTAs per cluster; "cluster_max_rows" (8192): taller fronts get one grid-wide
 *       launch per 128-pivot step instead; "cluster_slices" (1): levels with <= 9 fronts use 16-CTA clusters that
 *       share every 128-row block by 8-row slices (DSMEM all-gather of the solved entries); "defer_cb" (1): the
 *       contribution rows are updated by one wide GEMV after the pivot steps;
 *       "tri_span" (512, multiple of 32; 0 = never): the triangular matrix-vector products of the few wide fronts of the
 *       tree top split the long rows of the triangle over several CTAs, `tri_span` input entries each (partial sums
 *       combined in a fixed order by the last CTA to arrive: deterministic);
 *   "spmv_block" (1024; 512 / 1024 / 2048): entries per row block of the streamed SpMV (a CTA multiplies the entries
 *       of one block of consecutive rows in storage order and sums the rows out of shared memory);
 *   "partition_graphs" (default 1): partitioned solve: the sweeps, NCCL all-reduce included, are replayed from CUDA
 *       graphs like the single-GPU ones;
 *   "fuse_ortho" (default 1): the update of the first Gram-Schmidt pass and the dot products of the second one in
 *       one kernel (the basis is read three times per column instead of four);
 *   "ortho_refine_always" (default 0): second Gram-Schmidt pass for every basis column instead of SLEPc's
 *       refine-if-needed rule. */
int lsa_set_option(lsa_handle* h, const char* name, double value);
int lsa_symbolic_info_get(const lsa_handle* h, lsa_symbolic_info* out);
/* Copies a named internal array (perm, iperm, sn_ptr, st_ptr, st_idx, ea_map, parent, level, front_k,
 * front_r, p_off, q_off, c_off, a_dst, m_dst, lvl_ptr, lvl_front) to `out`; returns its length in elements
 * or a negative status.  With out == NULL only the length is returned.  Used by the host-logic tests. */
int64_t lsa_symbolic_array(const lsa_handle* h, const char* name, void* out, int64_t capacity_bytes);

/* lsa_set_values <- values of the AIJ matrices handed to eps.setOperators; may be called again
 * with new values on the same pattern (Reynolds sweep, BASELINE config 3).  `a_scalar`/`m_scalar`
 * are lsa_scalar; values are in the ORIGINAL CSR entry order.  on_device: pointers are device memory.
 * m_vals == NULL on a handle that already holds values keeps M (only A changes along a Reynolds sweep). */
int lsa_set_values(lsa_handle* h, const void* a_vals, int32_t a_scalar, const void* m_vals, int32_t m_scalar,
                   int32_t on_device);

/* -- numeric factorisation -------------------------------------------------------------------
 * lsa_factor <- STSetUp_Sinvert (MatAXPY T = A - sigma M) + PCSetUp_LU numeric part, reached from
 *               set_st_type(SINVERT)/set_target/set_st_pc_type(LU) + solve()  (Solver/utils.py:244-270;
 *               explicit in Solver/eigen2.py:109-151).
 * Factors  F = alpha A + beta M  (sinvert: alpha = 1, beta = -sigma; shift on a generalized problem:
 * alpha = 0, beta = 1).  scalar = LSA_F64 requires real values and real alpha/beta.
 * tiny_pivot > 0: pivots smaller than tiny_pivot * max|F| are replaced by that magnitude and counted;
 * tiny_pivot == 0: an exactly zero pivot fails with LSA_ERR_ZERO_PIVOT (PETSc's default behaviour,
 * tests/unit/Solver/test_eigen.py:272-281).                                                        */
int lsa_factor(lsa_handle* h, double alpha_re, double alpha_im, double beta_re, double beta_im, int32_t scalar,
               double tiny_pivot, lsa_factor_stats* stats);

/* lsa_solve <- KSPSolve(preonly + LU) = MatSolve / MatSolveTranspose   (Solver/eigen2.py:178;
 * the adjoint use of Sensitivity/__init__.py:246-262 maps to trans = LSA_OP_H on the same factors).
 * b, x: n complex numbers (interleaved) in the ORIGINAL ordering; b == x allowed.                 */
int lsa_solve(lsa_handle* h, int32_t trans, const double* b, double* x, int32_t refine_steps, int32_t on_device);

/* lsa_spmv <- MatMult / MatMultHermitianTranspose with A or M    (Solver/eigen2.py:174, :52-53). */
int lsa_spmv(lsa_handle* h, int32_t which_matrix, int32_t trans, const double* x, double* y, int32_t on_device);

/* lsa_set_nullspace <- MatSetNullSpace on the operator (FEM/utils.py:604-607, FEM/operators.py:534-545: the
 * constant-pressure vector of an enclosed flow).  `count` ORTHONORMAL vectors (n complex numbers each, original
 * ordering, column after column; count = 0 detaches).  The shifted operator is then singular: the factorisation
 * replaces the vanishing pivot (tiny_pivot > 0), every operator application removes the nullspace component from the
 * right-hand side before the sweeps and from the solution after them (KSP's behaviour with an attached nullspace;
 * explicit in Solver/eigen2.py:171-176), and lsa_solve does the same.                                          */
int lsa_set_nullspace(lsa_handle* h, int32_t count, const double* vecs_c128);

/* lsa_bilinear <- the scalar contractions of the sensitivity analysis, a^H B v with B on the sparsity pattern of A
 * (or M) and caller-supplied values in the ORIGINAL CSR entry order: bi-orthonormalisation a^H M v
 * (Sensitivity/__init__.py:280-287) and d lambda = a^H (dA/dRe) v with a pre-assembled derivative operator in place
 * of the UFL integrals of Sensitivity/__init__.py:354-385.  a, v: n complex numbers, original ordering.     */
int lsa_bilinear(lsa_handle* h, int32_t which_matrix, const void* vals, int32_t scalar, const double* a, const double* v,
                 int32_t on_device, double* out_c128);

/* -- eigensolve ------------------------------------------------------------------------------
 * lsa_eigs <- SLEPc.EPS.solve()  (Solver/utils.py:268-270): Krylov-Schur on OP, Gram-Schmidt with
 * refinement if needed (up to 256 basis columns),
 * Rayleigh-Ritz, locking restart, purification, 2-norm normalisation, `which` ordering.          */
int lsa_eigs(lsa_handle* h, const lsa_eigs_params* p, lsa_eigs_result* out);
/* <- eps.getConverged / getEigenvalue / getEigenvector  (Solver/utils.py:272-297) */
int lsa_get_eigenvalues(const lsa_handle* h, double* out_c128, int32_t capacity);
int lsa_get_eigenvectors(const lsa_handle* h, double* out_c128, int64_t ld, int32_t count, int32_t on_device);
/* ||A x - lambda M x|| / (||A||_F ||x||)  per returned pair, evaluated on the device (north-star bar). */
int lsa_get_residuals(lsa_handle* h, double* out, int32_t capacity);
int lsa_get_counters(const lsa_handle* h, lsa_counters* out);
int lsa_sync(lsa_handle* h);
/* Host helper (no GPU): out[q] = 1 when the diagonal entry of row rows[q] (rows == NULL: row q, nrows = n) of a CSR
 * matrix with SORTED column indices is absent or zero -- the structurally-zero-diagonal (pressure) flags that
 * lsa_analyze takes as `order_last`, without a pass over all nnz entries per solve of a sweep. */
int lsa_host_diag_is_zero(int32_t n, const void* indptr, int32_t indptr_is_64, const int32_t* colidx, const void* vals,
                          int32_t scalar, const int32_t* rows, int32_t nrows, uint8_t* out);
/* Host helper (no GPU): 1 when the two byte ranges are identical, 0 otherwise (a few threads; the exact comparison of
 * a candidate sparsity pattern with a cached one -- ~0.5 GB of indices for config 3 -- that replaces SLEPc's
 * "same nonzero pattern" promise of a sweep). */
int lsa_host_equal(const void* a, const void* b, uint64_t bytes);
/* Page-locked host memory for result buffers (eigenvectors leave the device at PCIe speed instead of through
 * the driver's bounce buffers; VecGetArray of the reference hands out host memory as well, Solver/utils.py:280-297).
 * Plain cudaHostAlloc / cudaFreeHost; no handle needed. */
int lsa_host_alloc(uint64_t bytes, void** ptr);
int lsa_host_free(void* ptr);

/* -- stand-alone kernels exposed for parity tests and roofline measurement -------------------- */
/* Dense Rayleigh-Ritz step (DSSolve/DSSort of SLEPc): Schur form of the m x m matrix S (column-major,
 * ld), ordered by `which` on the back-transformed values.  Runs the SAME kernel the eigensolver uses. */
int lsa_dense_schur(lsa_handle* h, int32_t m, double* S_c128, int32_t ld, double* Q_c128, int32_t which,
                    int32_t transform, double sigma_re, double sigma_im);
/* C -= A B on column-major FP64 (scalar = LSA_F64) or complex (LSA_C128) device matrices with the
 * front-update DMMA kernel; returns device milliseconds in *ms (mean of `reps` launches).          */
int lsa_gemm_bench(lsa_handle* h, int32_t scalar, int32_t m, int32_t n, int32_t k, int32_t reps, double* ms,
                   double* max_abs_err);

#ifdef __cplusplus
}
#endif
#endif /* LSA_B200_H */
