// Declarations of the numeric phases (factor.cu, solve.cu, krylov.cu).
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "handle.h"

namespace lsa {

// Optional per-launch timing of one sweep (LSA_TRACE=1, graphs off): prints level / kernel /
// step / grid / microseconds to stderr.  Debug aid for the latency analysis in profiles/.
struct SweepTrace {
  bool on = false;
  cudaStream_t st = nullptr;
  struct Rec { const char* name; int level, j0, gx, gy; cudaEvent_t ev; };
  std::vector<Rec> recs;
  cudaEvent_t start = nullptr;
  void begin(cudaStream_t s) {
    const char* e = getenv("LSA_TRACE");
    on = e && atoi(e) != 0;
    st = s;
    if (on) {
      cudaEventCreate(&start);
      cudaEventRecord(start, st);
    }
  }
  void mark(const char* name, int level, int j0, int gx, int gy) {
    if (!on) return;
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    cudaEventRecord(ev, st);
    recs.push_back({name, level, j0, gx, gy, ev});
  }
  void end() {
    if (!on) return;
    cudaStreamSynchronize(st);
    cudaEvent_t prev = start;
    double total = 0;
    for (auto& r : recs) {
      float ms = 0;
      cudaEventElapsedTime(&ms, prev, r.ev);
      total += ms;
      fprintf(stderr, "TRACE level %2d %-14s j0 %5d grid %6d x %6d  %9.2f us\n", r.level, r.name, r.j0, r.gx, r.gy, ms * 1e3);
      prev = r.ev;
    }
    fprintf(stderr, "TRACE total %.3f ms over %zu launches\n", total, recs.size());
    for (auto& r : recs) cudaEventDestroy(r.ev);
    cudaEventDestroy(start);
    recs.clear();
  }
};

// factor.cu
template <class T>
void factor_numeric(lsa_handle_impl& h, z128 alpha, z128 beta, double tiny_abs, int* n_kernels);
void gemm_plain(cudaStream_t st, bool cplx, const void* A, long long lda, const void* B, long long ldb, void* C,
                long long ldc, int M, int N, int K);

// solve.cu : x <- F^-1 x (trans = N) or F^-H x (trans = H), in place, permuted ordering, complex vector
template <class T>
void solve_permuted(lsa_handle_impl& h, int trans, z128* x, int* n_kernels);

void plan_solve(lsa_handle_impl& h, int scalar);
void exchange_replicated_rows(lsa_handle_impl& h, z128* x, bool with_cut_contributions);   // partitioned solve
void replicated_rows_to_partial(lsa_handle_impl& h, z128* x);   // solve.cu: level plan of the sweeps, before post_factor

// krylov.cu
void spmv(lsa_handle_impl& h, const CsrDev& M, bool conj_vals, const z128* x, z128* y);
std::vector<int> spmv_row_blocks(int n, const long long* rowptr, int block_entries);
void run_eigs(lsa_handle_impl& h, const lsa_eigs_params& p, lsa_eigs_result& out);
void dense_schur_device(lsa_handle_impl& h, int m, z128* dS, int ld, z128* dQ, int which, int transform, z128 sigma);
void op_solve(lsa_handle_impl& h, int trans, z128* x, int refine_steps);
void drop_solve_graphs(lsa_handle_impl& h);
z128 dot_conj(lsa_handle_impl& h, const z128* a, const z128* w);   // a^H w (single GPU)
void make_full(lsa_handle_impl& h, z128* vec, int ncols, long long ld);   // partitioned solve: complete, identical vectors on every GPU
void permute_gather(cudaStream_t st, const z128* src, z128* dst, const int* perm, int n);   // dst[i] = src[perm[i]]
void permute_scatter(cudaStream_t st, const z128* src, z128* dst, const int* perm, int n);  // dst[perm[i]] = src[i]
void gather_values(cudaStream_t st, const void* orig, bool is_complex, const long long* src, void* out, long long nnz);
void value_norms(lsa_handle_impl& h, const void* vals, bool is_complex, long long nnz, double* fro, double* amax);
void residual_norms(lsa_handle_impl& h, double* out_host, int count);

}  // namespace lsa
