// Declarations of the numeric phases (factor.cu, solve.cu, krylov.cu).
#pragma once
#include <algorithm>

#include "common.cuh"
#include "handle.h"

namespace lsa {

// factor.cu
template <class T>
void factor_numeric(lsa_handle_impl& h, z128 alpha, z128 beta, double tiny_abs, int* n_kernels);
void gemm_plain(cudaStream_t st, bool cplx, const void* A, long long lda, const void* B, long long ldb, void* C,
                long long ldc, int M, int N, int K);

// solve.cu : x <- F^-1 x (trans = N) or F^-H x (trans = H), in place, permuted ordering, complex vector
template <class T>
void solve_permuted(lsa_handle_impl& h, int trans, z128* x, int* n_kernels);

// krylov.cu
void spmv(lsa_handle_impl& h, const CsrDev& M, bool conj_vals, const z128* x, z128* y);
void run_eigs(lsa_handle_impl& h, const lsa_eigs_params& p, lsa_eigs_result& out);
void dense_schur_device(lsa_handle_impl& h, int m, z128* dS, int ld, z128* dQ, int which, int transform, z128 sigma);
void op_solve(lsa_handle_impl& h, int trans, z128* x, int refine_steps);
void drop_solve_graphs(lsa_handle_impl& h);
void permute_gather(cudaStream_t st, const z128* src, z128* dst, const int* perm, int n);   // dst[i] = src[perm[i]]
void permute_scatter(cudaStream_t st, const z128* src, z128* dst, const int* perm, int n);  // dst[perm[i]] = src[i]
void gather_values(cudaStream_t st, const void* orig, bool is_complex, const long long* src, void* out, long long nnz);
void value_norms(lsa_handle_impl& h, const void* vals, bool is_complex, long long nnz, double* fro, double* amax);
void residual_norms(lsa_handle_impl& h, double* out_host, int count);

}  // namespace lsa
