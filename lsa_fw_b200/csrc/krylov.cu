// Krylov-Schur eigensolver on the device: operator application (SpMV + triangular solves),
// block CGS2 orthogonalisation, Rayleigh-Ritz, locking restart, Ritz vectors, purification.
//
// Stands in for `SLEPc.EPS.solve()` with EPS type krylovschur and ST sinvert/shift
// (reference Solver/utils.py:268-270 and the defaults discussed in SURVEY.md section 8a):
//   OP = (A - sigma M)^-1 M          (STApply_Sinvert; explicit in Solver/eigen2.py:164-190)
//   BVOrthogonalizeColumn            -> k_dots / k_reduce_h / k_update   (tall-skinny, HBM bound)
//   DSSolve + DSSort + convergence   -> k_rr (single CTA, rr_core.h)
//   BVMultInPlace (restart)          -> k_basis_gemm
//   EPSComputeVectors + purification -> k_ritz_vectors + k_basis_gemm + one more OP apply
// The host only sequences launches; it reads back one small struct per restart.
#include <cstring>
#include <vector>

#include "factor.cuh"
#include "rr_core.h"

namespace lsa {

// ---------------------------------------------------------------------------------------- SpMV

// y = Op x, CSR, "stream" form: a CTA owns a block of consecutive rows holding at most SPMV_T = 256 U entries (row blocks
// cut on the host when the pattern is uploaded, CsrDev::rowblk).  Phase 1: all threads walk the block's entries in
// storage order -- values and column indices are read fully coalesced, eight independent (value, index, x[index])
// loads per thread in flight -- and leave the products in shared memory.  Phase 2: one thread per row sums its
// segment in storage order (deterministic, independent of the launch geometry).  Rows of the FE pencils hold 0-100
// entries (pressure rows of M are empty), which starves a lanes-per-row kernel; a single row longer than SPMV_T
// entries gets a block of its own and a block-wide reduction.
constexpr int SPMV_ROWS = 1024;   // at most this many rows per block (stretches of empty rows)

template <class VT, bool CONJ, int U>
__global__ void __launch_bounds__(256) k_spmv(const int* __restrict__ rowblk, const long long* __restrict__ blk_e0,
                                              const long long* __restrict__ rowptr, const int* __restrict__ colidx,
                                              const VT* __restrict__ vals, const z128* __restrict__ x,
                                              z128* __restrict__ y) {
  constexpr int SPMV_T = 256 * U;
  __shared__ z128 prod[SPMV_T];
  const int tid = threadIdx.x;
  // a CTA is a chain of dependent memory round trips (block bounds -> entries -> x -> row sums): the block's first
  // entry comes from its own array instead of rowptr[rowblk[.]], and the row bounds of phase 2 are fetched together
  // with the entries -- three round trips instead of five
  const int r0 = rowblk[blockIdx.x], r1 = rowblk[blockIdx.x + 1];
  const long long e0 = blk_e0[blockIdx.x], e1 = blk_e0[blockIdx.x + 1];
  long long rp0 = 0, rp1 = 0;
  if (r0 + tid < r1) {
    rp0 = rowptr[r0 + tid];
    rp1 = rowptr[r0 + tid + 1];
  }
  if (e1 - e0 > SPMV_T) {   // one long row
    z128 acc = mk(0, 0);
    for (long long p = e0 + tid; p < e1; p += 256) acc += cj<CONJ>(vals[p]) * x[colidx[p]];
    prod[tid] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o) prod[tid] += prod[tid + o];
      __syncthreads();
    }
    if (tid == 0) y[r0] = prod[0];
    return;
  }
  const int cnt = (int)(e1 - e0);
  const VT* __restrict__ vb = vals + e0;
  const int* __restrict__ cb = colidx + e0;
  int c[U];
  VT v[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int i = tid + u * 256;
    c[u] = i < cnt ? cb[i] : -1;
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int i = tid + u * 256;
    if (i < cnt) v[u] = vb[i];
  }
  z128 xv[U];
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (c[u] >= 0) xv[u] = x[c[u]];
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (c[u] >= 0) prod[tid + u * 256] = cj<CONJ>(v[u]) * xv[u];
  __syncthreads();
  for (int r = r0 + tid; r < r1; r += 256) {
    if (r != r0 + tid) {
      rp0 = rowptr[r];
      rp1 = rowptr[r + 1];
    }
    const int b = (int)(rp0 - e0), e = (int)(rp1 - e0);
    z128 acc = mk(0, 0);
    for (int p = b; p < e; ++p) acc += prod[p];
    y[r] = acc;
  }
}

// row blocks of a CSR pattern for k_spmv: consecutive rows, at most `block_entries` entries and SPMV_ROWS rows per block
std::vector<int> spmv_row_blocks(int n, const long long* rowptr, int block_entries) {
  std::vector<int> blk;
  blk.push_back(0);
  int r = 0;
  while (r < n) {
    const long long e0 = rowptr[r];
    int q = r + 1;   // a block holds at least one row
    while (q < n && q - r < SPMV_ROWS && rowptr[q + 1] - e0 <= block_entries) ++q;
    blk.push_back(q);
    r = q;
  }
  return blk;
}

template <int U>
static void launch_spmv(cudaStream_t st, const CsrDev& M, bool conj_vals, const z128* x, z128* y) {
  const int blocks = M.n_rowblk;
  if (M.is_complex) {
    if (conj_vals) k_spmv<z128, true, U><<<blocks, 256, 0, st>>>(M.rowblk, M.blk_e0, M.rowptr, M.colidx, (const z128*)M.vals, x, y);
    else k_spmv<z128, false, U><<<blocks, 256, 0, st>>>(M.rowblk, M.blk_e0, M.rowptr, M.colidx, (const z128*)M.vals, x, y);
  } else {
    k_spmv<double, false, U><<<blocks, 256, 0, st>>>(M.rowblk, M.blk_e0, M.rowptr, M.colidx, (const double*)M.vals, x, y);
  }
}

void spmv(lsa_handle_impl& h, const CsrDev& M, bool conj_vals, const z128* x, z128* y) {
  if (M.n_rowblk <= 0) return;
  if (M.block_entries == 512) launch_spmv<2>(h.stream, M, conj_vals, x, y);
  else if (M.block_entries == 1024) launch_spmv<4>(h.stream, M, conj_vals, x, y);
  else if (M.block_entries == 2048) launch_spmv<8>(h.stream, M, conj_vals, x, y);
  else throw std::runtime_error("spmv: unsupported row-block size");
  LSA_LAUNCH_CHECK();
}

// ----------------------------------------------------------------------------- small vector kernels

__global__ void k_perm_gather(const z128* __restrict__ src, z128* __restrict__ dst, const int* __restrict__ perm, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}
__global__ void k_perm_scatter(const z128* __restrict__ src, z128* __restrict__ dst, const int* __restrict__ perm, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[perm[i]] = src[i];
}
void permute_gather(cudaStream_t st, const z128* src, z128* dst, const int* perm, int n) {
  if (n > 0) k_perm_gather<<<cdiv(n, 256), 256, 0, st>>>(src, dst, perm, n);
  LSA_LAUNCH_CHECK();
}
void permute_scatter(cudaStream_t st, const z128* src, z128* dst, const int* perm, int n) {
  if (n > 0) k_perm_scatter<<<cdiv(n, 256), 256, 0, st>>>(src, dst, perm, n);
  LSA_LAUNCH_CHECK();
}

template <class VT>
__global__ void k_gather_vals(const VT* __restrict__ orig, const long long* __restrict__ src, VT* __restrict__ out,
                              long long nnz) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; e < nnz; e += stride) out[e] = orig[src[e]];
}
void gather_values(cudaStream_t st, const void* orig, bool is_complex, const long long* src, void* out, long long nnz) {
  if (nnz == 0) return;
  const int blocks = (int)std::min<long long>((nnz + 255) / 256, 148LL * 16);
  if (is_complex) k_gather_vals<z128><<<blocks, 256, 0, st>>>((const z128*)orig, src, (z128*)out, nnz);
  else k_gather_vals<double><<<blocks, 256, 0, st>>>((const double*)orig, src, (double*)out, nnz);
  LSA_LAUNCH_CHECK();
}

// sum |v|^2 and max |v| of a value array -> part[2*blk], part[2*blk+1]
template <class VT>
__global__ void __launch_bounds__(256) k_value_norms(const VT* __restrict__ v, long long nnz, double* __restrict__ part) {
  double s = 0.0, mx = 0.0;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nnz; e += (long long)gridDim.x * blockDim.x) {
    const double a2 = abs2(v[e]);
    s += a2;
    mx = fmax(mx, a2);
  }
  __shared__ double ss[256], sm[256];
  ss[threadIdx.x] = s;
  sm[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      ss[threadIdx.x] += ss[threadIdx.x + o];
      sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    part[2 * blockIdx.x] = ss[0];
    part[2 * blockIdx.x + 1] = sm[0];
  }
}
void value_norms(lsa_handle_impl& h, const void* vals, bool is_complex, long long nnz, double* fro, double* amax) {
  *fro = 0.0;
  *amax = 0.0;
  if (nnz == 0) return;
  const int blocks = (int)std::min<long long>((nnz + 255) / 256, 256);
  double* d_part = nullptr;
  LSA_CUDA(cudaMalloc(&d_part, sizeof(double) * 2 * blocks));
  if (is_complex) k_value_norms<z128><<<blocks, 256, 0, h.stream>>>((const z128*)vals, nnz, d_part);
  else k_value_norms<double><<<blocks, 256, 0, h.stream>>>((const double*)vals, nnz, d_part);
  std::vector<double> part(2 * blocks);
  LSA_CUDA(cudaMemcpyAsync(part.data(), d_part, sizeof(double) * 2 * blocks, cudaMemcpyDeviceToHost, h.stream));
  LSA_CUDA(cudaStreamSynchronize(h.stream));
  cudaFree(d_part);
  double s = 0, mx = 0;
  for (int b = 0; b < blocks; ++b) {
    s += part[2 * b];
    mx = std::max(mx, part[2 * b + 1]);
  }
  *fro = std::sqrt(s);
  *amax = std::sqrt(mx);
}

// splitmix64 -> Box-Muller standard normal start vector (real part; imaginary part zero)
__global__ void k_randn(z128* __restrict__ x, int n, unsigned long long seed) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(2 * i + 1);
  auto next = [&]() {
    z += 0x9E3779B97F4A7C15ULL;
    unsigned long long r = z;
    r = (r ^ (r >> 30)) * 0xBF58476D1CE4E5B9ULL;
    r = (r ^ (r >> 27)) * 0x94D049BB133111EBULL;
    return r ^ (r >> 31);
  };
  const double u1 = ((next() >> 11) + 1.0) * (1.0 / 9007199254740993.0);
  const double u2 = (next() >> 11) * (1.0 / 9007199254740992.0);
  x[i] = mk(sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2), 0.0);
}

// y = a x + y   /  y = x * s
__global__ void k_axpy(int n, z128 a, const z128* __restrict__ x, z128* __restrict__ y) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += a * x[i];
}
__global__ void k_copy(int n, const z128* __restrict__ x, z128* __restrict__ y) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = x[i];
}
__global__ void k_sub(int n, const z128* __restrict__ a, const z128* __restrict__ b, z128* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] - b[i];
}

// ------------------------------------------------------------------------------------- CGS2 kernels

static constexpr int DOT_CG = 8;      // column groups (warps per block)

// part[blk * ldp + c] = sum over the block's rows of conj(V[i, c]) w[i],  c0 <= c < min(j, c0 + 8 NQ)
// wn2 (optional): wn2[blk] = sum over the block's rows of |w[i]|^2.  skip (optional): *skip == 0 -> nothing to do
// (second Gram-Schmidt pass that the refinement criterion did not ask for).
// Warp = column group (columns c0 + warp + 8 q), lanes along rows (512-byte coalesced reads of a column); every
// thread issues the U x NQ column loads of U row slabs before it uses any of them, so narrow bases (few columns
// per thread) still keep ~16 independent 16-byte loads per thread in flight.
template <int NQ, int U>
__global__ void __launch_bounds__(256) k_dots(int n, int j, int c0, const z128* __restrict__ V, long long ldv,
                                              const z128* __restrict__ w, z128* __restrict__ part, int ldp,
                                              int rows_per_block, double* __restrict__ wn2, const int* __restrict__ skip) {
  if (skip && *skip == 0) return;
  const int lane = threadIdx.x & 31, cg = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min((long long)n, r0 + rows_per_block);
  z128 acc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q] = mk(0, 0);
  double wacc = 0.0;
  for (long long i = r0 + lane; i < r1; i += 32 * U) {
    z128 wi[U], v[U][NQ];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ii = i + 32 * u;
      const bool ok = ii < r1;
      wi[u] = ok ? w[ii] : mk(0, 0);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int c = c0 + cg + q * DOT_CG;
        v[u][q] = (ok && c < j) ? V[ii + c * ldv] : mk(0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      wacc += abs2(wi[u]);
#pragma unroll
      for (int q = 0; q < NQ; ++q) acc[q] += conj_(v[u][q]) * wi[u];
    }
  }
  if (wn2 && cg == 0) {
    for (int o = 16; o > 0; o >>= 1) wacc += __shfl_xor_sync(0xffffffffu, wacc, o);
    if (lane == 0) wn2[blockIdx.x] = wacc;
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int c = c0 + cg + q * DOT_CG;
    if (c < j) {
      z128 a = acc[q];
      for (int o = 16; o > 0; o >>= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
        a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
      }
      if (lane == 0) part[(long long)blockIdx.x * ldp + c] = a;
    }
  }
}

// Fused "update of pass 1 + dots of pass 2" of the two-pass Gram-Schmidt: w <- w - V h over the block's rows and,
// with the SAME tile of V still in registers, the partial products conj(V)^T w_new of the second pass -- the basis
// is read three times per column instead of four when the second pass is taken (it is for > 99 % of the columns of
// the shift-and-invert runs measured here), and the second pass' dot products come for free when it is not.
// Same thread layout as k_dots (warp = column group, lanes = rows); the row sums of the update cross the eight
// warps through shared memory, one slab of 32 U rows at a time.  j <= 8 NQ (all columns in one block).
// part[blk, c] = partial h2[c];  wn2[blk] = partial |w_new|^2 (norm after the first pass).
template <int NQ, int U>
__global__ void __launch_bounds__(256) k_update_dots(int n, int j, const z128* __restrict__ V, long long ldv,
                                                     const z128* __restrict__ h, z128* __restrict__ w,
                                                     z128* __restrict__ part, int ldp, int rows_per_block,
                                                     double* __restrict__ wn2) {
  __shared__ z128 hs[8 * NQ];
  __shared__ z128 red[DOT_CG][U][33];
  __shared__ z128 wsh[U][32];
  const int lane = threadIdx.x & 31, cg = threadIdx.x >> 5;
  if ((int)threadIdx.x < 8 * NQ) hs[threadIdx.x] = (int)threadIdx.x < j ? h[threadIdx.x] : mk(0, 0);
  __syncthreads();
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min((long long)n, r0 + rows_per_block);
  z128 acc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q] = mk(0, 0);
  double wacc = 0.0;
  for (long long i0 = r0; i0 < r1; i0 += 32 * U) {
    z128 v[U][NQ];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ii = i0 + lane + 32 * u;
      const bool ok = ii < r1;
      z128 s = mk(0, 0);
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int c = cg + q * DOT_CG;
        v[u][q] = (ok && c < j) ? V[ii + c * ldv] : mk(0, 0);
        s += v[u][q] * hs[c];
      }
      red[cg][u][lane] = s;
    }
    __syncthreads();
    if (cg < U) {   // warp u finishes slab u: w_new = w - sum over the column groups
      const long long ii = i0 + lane + 32 * cg;
      z128 wn = mk(0, 0);
      if (ii < r1) {
        z128 s = red[0][cg][lane];
#pragma unroll
        for (int g = 1; g < DOT_CG; ++g) s += red[g][cg][lane];
        wn = w[ii] - s;
        w[ii] = wn;
        wacc += abs2(wn);
      }
      wsh[cg][lane] = wn;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const z128 wn = wsh[u][lane];
#pragma unroll
      for (int q = 0; q < NQ; ++q) acc[q] += conj_(v[u][q]) * wn;
    }
  }
  // |w_new|^2 of the block: the U finishing warps each hold a share
  __shared__ double wred[8];
  for (int o = 16; o > 0; o >>= 1) wacc += __shfl_xor_sync(0xffffffffu, wacc, o);
  if (lane == 0) wred[cg] = wacc;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int c = cg + q * DOT_CG;
    if (c < j) {
      z128 a = acc[q];
      for (int o = 16; o > 0; o >>= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
        a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
      }
      if (lane == 0) part[(long long)blockIdx.x * ldp + c] = a;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sacc = 0.0;
    for (int g = 0; g < U && g < 8; ++g) sacc += wred[g];
    wn2[blockIdx.x] = sacc;
  }
}

static void launch_update_dots(cudaStream_t st, int nblk, int n, int j, const z128* V, long long ldv, const z128* h, z128* w,
                               z128* part, int ldp, int rows_per_block, double* wn2) {
  if (j <= 16) k_update_dots<2, 8><<<nblk, 256, 0, st>>>(n, j, V, ldv, h, w, part, ldp, rows_per_block, wn2);
  else if (j <= 32) k_update_dots<4, 4><<<nblk, 256, 0, st>>>(n, j, V, ldv, h, w, part, ldp, rows_per_block, wn2);
  else if (j <= 64) k_update_dots<8, 2><<<nblk, 256, 0, st>>>(n, j, V, ldv, h, w, part, ldp, rows_per_block, wn2);
  else if (j <= 96) k_update_dots<12, 2><<<nblk, 256, 0, st>>>(n, j, V, ldv, h, w, part, ldp, rows_per_block, wn2);
  else k_update_dots<16, 1><<<nblk, 256, 0, st>>>(n, j, V, ldv, h, w, part, ldp, rows_per_block, wn2);
}

// All partial dot products of w against V[:, 0:j]: chunks of up to 128 columns, columns per thread by chunk width.
static void launch_dots(cudaStream_t st, int nblk, int n, int j, const z128* V, long long ldv, const z128* w, z128* part,
                        int ldp, int rows_per_block, double* wn2, const int* skip) {
  for (int c0 = 0; c0 < j; c0 += 128) {
    const int cols = std::min(128, j - c0);
    double* wn = c0 == 0 ? wn2 : nullptr;
    if (cols <= 16) k_dots<2, 8><<<nblk, 256, 0, st>>>(n, j, c0, V, ldv, w, part, ldp, rows_per_block, wn, skip);
    else if (cols <= 32) k_dots<4, 4><<<nblk, 256, 0, st>>>(n, j, c0, V, ldv, w, part, ldp, rows_per_block, wn, skip);
    else if (cols <= 64) k_dots<8, 2><<<nblk, 256, 0, st>>>(n, j, c0, V, ldv, w, part, ldp, rows_per_block, wn, skip);
    else if (cols <= 96) k_dots<12, 2><<<nblk, 256, 0, st>>>(n, j, c0, V, ldv, w, part, ldp, rows_per_block, wn, skip);
    else k_dots<16, 1><<<nblk, 256, 0, st>>>(n, j, c0, V, ldv, w, part, ldp, rows_per_block, wn, skip);
  }
}

// h[c] = sum_blk part[blk, c]; S column update: scol[c] = (accumulate ? scol[c] : 0) + h[c].
// One warp per column.
__global__ void __launch_bounds__(256) k_reduce_h(int j, int nblk, const z128* __restrict__ part, int ldp,
                                                  z128* __restrict__ h, z128* __restrict__ scol, int accumulate,
                                                  const int* __restrict__ skip) {
  if (skip && *skip == 0) return;
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= j) return;
  z128 a = mk(0, 0);
  for (int b = lane; b < nblk; b += 32) a += part[(long long)b * ldp + c];
  for (int o = 16; o > 0; o >>= 1) {
    a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
    a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
  }
  if (lane == 0) {
    h[c] = a;
    scol[c] = accumulate ? scol[c] + a : a;
  }
}

// w -= V[:, 0:j] h ; optionally npart[blk] = sum |w_new|^2 over the block's rows
__global__ void __launch_bounds__(256) k_update(int n, int j, const z128* __restrict__ V, long long ldv,
                                                const z128* __restrict__ h, z128* __restrict__ w,
                                                double* __restrict__ npart, const int* __restrict__ skip) {
  if (skip && *skip == 0) return;
  __shared__ z128 hs[256];
  __shared__ double red[8];
  if ((int)threadIdx.x < j) hs[threadIdx.x] = h[threadIdx.x];
  __syncthreads();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double nn = 0.0;
  if (i < n) {
    z128 acc = w[i];
    const z128* v = V + i;
#pragma unroll 8
    for (int c = 0; c < j; ++c) acc -= v[c * ldv] * hs[c];
    w[i] = acc;
    nn = abs2(acc);
  }
  if (npart) {
    for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = nn;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0;
      for (int q = 0; q < 8; ++q) s += red[q];
      npart[blockIdx.x] = s;
    }
  }
}

// Reorthogonalisation criterion of the "refine if needed" Gram-Schmidt (SLEPc's default,
// BV_ORTHOG_REFINE_IFNEEDED with eta = 1/sqrt(2)): a second pass only if the first one removed more than half
// of the squared norm.  flag = 1: refine.  Fixed-order reductions (deterministic).
__global__ void __launch_bounds__(256) k_refine_flag(int nb_before, const double* __restrict__ wn2, int nb_after,
                                                     const double* __restrict__ npart, int always,
                                                     int* __restrict__ flag, int* __restrict__ count) {
  __shared__ double ra[256], rb[256];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < nb_before; i += 256) a += wn2[i];
  for (int i = threadIdx.x; i < nb_after; i += 256) b += npart[i];
  ra[threadIdx.x] = a;
  rb[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      ra[threadIdx.x] += ra[threadIdx.x + o];
      rb[threadIdx.x] += rb[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int f = always || !(rb[0] >= 0.5 * ra[0]);   // also true for NaN
    *flag = f;
    if (f && count) atomicAdd(count, 1);
  }
}

// ---- partitioned solve (row-sharded basis): the Gram-Schmidt coefficients of a pass and |w|^2 travel in ONE
// all-reduce payload  red = [h_0 .. h_{j-1} (complex), |w|^2];  the norm after the update comes from Pythagoras,
// |w - V h|^2 = |w|^2 - |h|^2 (what SLEPc's BV does for its refinement criterion): no second reduction per pass.
__global__ void __launch_bounds__(256) k_reduce_red(int j, int nblk, const z128* __restrict__ part, int ldp,
                                                    const double* __restrict__ wn2, double* __restrict__ red,
                                                    const int* __restrict__ skip) {
  if (skip && *skip == 0) return;
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c > j) return;
  if (c == j) {   // one extra warp: |w|^2 over this rank's rows
    double a = 0.0;
    for (int b = lane; b < nblk; b += 32) a += wn2[b];
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[2 * j] = a;
    return;
  }
  z128 a = mk(0, 0);
  for (int b = lane; b < nblk; b += 32) a += part[(long long)b * ldp + c];
  for (int o = 16; o > 0; o >>= 1) {
    a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
    a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
  }
  if (lane == 0) {
    red[2 * c] = a.x;
    red[2 * c + 1] = a.y;
  }
}
// After the all-reduce: h <- red, S column, norm estimate and (pass 1) the refinement decision.
//   pass 1: norm2 = max(|w|^2 - |h|^2, 0);  refine = always || !(norm2 >= 0.5 |w|^2)
//   pass 2 (only when refine was set): norm2 = max(|w'|^2 - |h2|^2, 0) with |w'|^2 measured exactly in the pass
__global__ void __launch_bounds__(256) k_apply_red(int j, const double* __restrict__ red, z128* __restrict__ h,
                                                   z128* __restrict__ scol, int pass, int always,
                                                   double* __restrict__ norm2, int* __restrict__ flag,
                                                   int* __restrict__ count) {
  if (pass == 2 && *flag == 0) return;
  __shared__ double sh[256];
  double hn = 0.0;
  for (int c = threadIdx.x; c < j; c += 256) {
    const z128 v = mk(red[2 * c], red[2 * c + 1]);
    h[c] = v;
    if (scol) scol[c] = pass == 2 ? scol[c] + v : v;
    hn += abs2(v);
  }
  sh[threadIdx.x] = hn;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double wn = red[2 * j];
    const double est = wn - sh[0];
    *norm2 = (est == est) ? fmax(est, 0.0) : est;   // NaN stays NaN (reported as non-finite)
    if (pass == 1) {
      const int f = always || !(est >= 0.5 * wn);
      *flag = f;
      if (f && count) atomicAdd(count, 1);
    }
  }
}
// red[0] = sum of partials (single block)
__global__ void __launch_bounds__(256) k_sum_partials(const double* __restrict__ part, int n, int stride, double* __restrict__ out) {
  __shared__ double sh[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) a += part[(long long)i * stride];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// npart[blk] = sum |w|^2 over the block's rows
__global__ void __launch_bounds__(256) k_norm2_part(int n, const z128* __restrict__ w, double* __restrict__ npart) {
  __shared__ double red[8];
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double nn = i < n ? abs2(w[i]) : 0.0;
  for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = nn;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int q = 0; q < 8; ++q) s += red[q];
    npart[blockIdx.x] = s;
  }
}

// beta = sqrt(sum npart); out = w / beta.  Every block re-reduces the partials in the same order
// (deterministic).  Block 0 stores beta (to *beta_out, and to *s_entry as a complex number) and
// flags a breakdown (beta tiny relative to hnorm_ref) in flag[0] and a non-finite norm in flag[1].
__global__ void __launch_bounds__(256) k_normalize(int n, const z128* __restrict__ w, z128* __restrict__ out,
                                                   const double* __restrict__ npart, int nparts,
                                                   double* __restrict__ beta_out, z128* __restrict__ s_entry,
                                                   const z128* __restrict__ hcol, int hlen, int* __restrict__ flag,
                                                   int step, const double* __restrict__ npart_alt = nullptr,
                                                   int nparts_alt = 0, const int* __restrict__ use_alt_if_zero = nullptr) {
  __shared__ double red[256];
  // fused Gram-Schmidt: when the second pass was NOT taken (*use_alt_if_zero == 0) the norm is the one measured
  // after the first pass (npart_alt)
  if (use_alt_if_zero && *use_alt_if_zero == 0) {
    npart = npart_alt;
    nparts = nparts_alt;
  }
  double s = 0.0;
  for (int b = threadIdx.x; b < nparts; b += blockDim.x) s += npart[b];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const double beta = sqrt(red[0]);
  __shared__ int s_break;
  if (threadIdx.x == 0) {
    double hn = 0.0;
    for (int c = 0; c < hlen; ++c) hn += abs2(hcol[c]);
    hn = sqrt(hn);
    // NaN / Inf in the new direction or in its Gram-Schmidt coefficients is NOT a breakdown: it is reported
    // through flag[1] and ends the solve with LSA_ERR_NONFINITE
    const bool bad = !(beta == beta) || isinf(beta) || !(hn == hn) || isinf(hn);
    const bool brk = bad || !(beta > 1e-13 * fmax(hn, 1e-300));
    s_break = brk;
    if (blockIdx.x == 0) {
      if (beta_out) *beta_out = beta;
      if (s_entry) *s_entry = mk(brk ? 0.0 : beta, 0.0);
      if (brk && flag) atomicMin(flag, step);
      if (bad && flag) flag[1] = 1;
    }
  }
  __syncthreads();
  const double inv = s_break ? 0.0 : 1.0 / beta;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = w[i] * inv;
}

// ------------------------------------------------------------------------------ basis GEMM (restart)

// Out[:, c] = sum_p V[:, p] Qm[p, c],  p < mp, c < nk.  In-place safe (Out may alias V) because each
// block stages its 32 rows of V in shared memory before writing.
__global__ void __launch_bounds__(256) k_basis_gemm(int n, int mp, int nk, const z128* __restrict__ V, long long ldv,
                                                    const z128* __restrict__ Qm, int ldq, z128* __restrict__ Out,
                                                    long long ldo) {
  extern __shared__ unsigned char smem_raw[];
  z128* vs = reinterpret_cast<z128*>(smem_raw);  // [mp][33]
  const int lane = threadIdx.x & 31, cg = threadIdx.x >> 5;
  const long long row = (long long)blockIdx.x * 32 + lane;
  for (int p = cg; p < mp; p += 8) vs[p * 33 + lane] = row < n ? V[row + p * ldv] : mk(0, 0);
  __syncthreads();
  for (int c0 = cg; c0 < nk; c0 += 8 * 4) {
    z128 acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = mk(0, 0);
    for (int p = 0; p < mp; ++p) {
      const z128 v = vs[p * 33 + lane];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = c0 + q * 8;
        if (c < nk) acc[q] += v * Qm[p + (long long)c * ldq];
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = c0 + q * 8;
      if (c < nk && row < n) Out[row + c * ldo] = acc[q];
    }
  }
}

// ------------------------------------------------------------------------------------ Rayleigh-Ritz

__global__ void __launch_bounds__(128) k_rr(z128* S, z128* Q, RrParams p, z128* theta, double* resid, z128* brow,
                                            z128* ywork, RrInfo* info, int use_smem) {
  __shared__ double s_rot_c[264];
  __shared__ z128 s_rot_s[264];
  __shared__ z128 s_vec[264];
  __shared__ double s_key[264];
  __shared__ int s_flag[4];
  __shared__ RrOut s_out;
  extern __shared__ unsigned char rr_dyn[];
  RrWork w{s_rot_c, s_rot_s, s_vec, s_key, s_flag};
  const int m = p.m, tid = threadIdx.x, nt = blockDim.x;
  if (use_smem) {
    // the projected matrix and the accumulated transformation live in shared memory for the whole
    // Schur iteration (global round trips were 18 ms per call at m = 80, see profiles/)
    z128* Ss = reinterpret_cast<z128*>(rr_dyn);
    z128* Qs = Ss + (size_t)(m + 1) * m;
    for (int e = tid; e < (m + 1) * m; e += nt) Ss[e] = S[(e % (m + 1)) + (long long)(e / (m + 1)) * p.ld];
    __syncthreads();
    RrParams ps = p;
    ps.ld = m + 1;
    ps.ldq = m;
    rr_full(Ss, Qs, ps, theta, resid, brow, ywork, &s_out, tid, nt, w);
    __syncthreads();
    for (int e = tid; e < (m + 1) * m; e += nt) S[(e % (m + 1)) + (long long)(e / (m + 1)) * p.ld] = Ss[e];
    for (int e = tid; e < m * m; e += nt) Q[(e % m) + (long long)(e / m) * p.ldq] = Qs[e];
  } else {
    rr_full(S, Q, p, theta, resid, brow, ywork, &s_out, tid, nt, w);
  }
  __syncthreads();
  if (tid == 0) {
    info->nconv = s_out.nconv;
    info->keep = s_out.keep;
    info->status = s_out.status;
    info->pad = 0;
  }
}

// Schur form only (lsa_dense_schur): S (m x m, ld), sorted by `which`.
__global__ void __launch_bounds__(128) k_schur_only(z128* S, int ld, z128* Q, int m, RrParams p, RrInfo* info) {
  __shared__ double s_rot_c[264];
  __shared__ z128 s_rot_s[264];
  __shared__ z128 s_vec[264];
  __shared__ double s_key[264];
  __shared__ int s_flag[4];
  RrWork w{s_rot_c, s_rot_s, s_vec, s_key, s_flag};
  const int st = rr_schur(S, ld, Q, m, m, 0, threadIdx.x, blockDim.x, w);
  rr_sort(S, ld, Q, m, m, 0, p, threadIdx.x, blockDim.x, w);
  if (threadIdx.x == 0) {
    info->status = st;
    info->nconv = 0;
    info->keep = 0;
  }
}

void dense_schur_device(lsa_handle_impl& h, int m, z128* dS, int ld, z128* dQ, int which, int transform, z128 sigma) {
  if (m > 256) throw std::runtime_error("dense Schur kernel supports m <= 256");
  RrParams p{};
  p.m = m; p.ld = ld; p.ldq = m; p.nconv = 0; p.nev = m; p.which = which; p.transform = transform;
  p.sigma = sigma; p.tol = 0; p.last = 1; p.beta_scale = 1.0;
  RrInfo* d_info;
  LSA_CUDA(cudaMalloc(&d_info, sizeof(RrInfo)));
  k_schur_only<<<1, 128, 0, h.stream>>>(dS, ld, dQ, m, p, d_info);
  LSA_LAUNCH_CHECK();
  RrInfo info;
  LSA_CUDA(cudaMemcpyAsync(&info, d_info, sizeof(RrInfo), cudaMemcpyDeviceToHost, h.stream));
  LSA_CUDA(cudaStreamSynchronize(h.stream));
  cudaFree(d_info);
  if (info.status != 0) throw std::runtime_error("dense Schur: QR iteration did not converge");
}

// Eigenvectors of the leading nc x nc triangle of S: column c of Y (nc x nc, ld = ldy), unit 2-norm.
// With brow != nullptr Y gets one more row, Y[nc, c] = (b . y_c) / theta_c  (b = coupling row of the Krylov-Schur
// relation  OP V = V S + v_next b^T):  V y + v_next (b . y) / theta  =  OP (V y) / theta, i.e. the purified Ritz
// vector (one multiplication by OP, which removes components outside range(OP) when M is singular) at no
// extra operator application -- what EPSComputeVectors does in SLEPc's Krylov-Schur.
__global__ void k_ritz_vectors(const z128* __restrict__ S, int ld, int nc, z128* __restrict__ Y, int ldy,
                               const z128* __restrict__ brow) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nc) return;
  z128* y = Y + (long long)c * ldy;
  for (int i = c + 1; i < nc + (brow ? 1 : 0); ++i) y[i] = mk(0, 0);
  // brow unused: pass a zero row through y itself is not possible, so replicate the substitution here
  const z128 tkk = S[c + (long long)c * ld];
  double smax = 0.0;
  for (int i = 0; i <= c; ++i) smax = fmax(smax, abs1(S[i + (long long)i * ld]));
  const double smin = fmax(smax * 2.220446049250313e-16, 1e-300);
  y[c] = mk(1, 0);
  for (int i = c - 1; i >= 0; --i) {
    z128 acc = mk(0, 0);
    for (int j = i + 1; j <= c; ++j) acc += S[i + (long long)j * ld] * y[j];
    z128 d = S[i + (long long)i * ld] - tkk;
    if (abs1(d) < smin) d = mk(smin, 0);
    y[i] = mk(0, 0) - acc / d;
  }
  double nrm2 = 0.0;
  for (int i = 0; i <= c; ++i) nrm2 += abs2(y[i]);
  const double inv = 1.0 / sqrt(nrm2);
  for (int i = 0; i <= c; ++i) y[i] = y[i] * inv;
  if (brow) {
    z128 dot = mk(0, 0);
    for (int i = 0; i <= c; ++i) dot += brow[i] * y[i];
    y[nc] = (tkk.x != 0.0 || tkk.y != 0.0) ? dot / tkk : mk(0, 0);
  }
}

// ---- phase normalisation of a Ritz vector: rotate so that its largest component is real positive
// (eigenvectors are defined up to a phase; fixing it on the device makes the output deterministic and
// saves a host pass over n x nev numbers).  Two-stage arg-max with index tie-break, then the rotation
// is fused into the un-permutation scatter.
__global__ void __launch_bounds__(256) k_absmax_part(int n, const z128* __restrict__ x, double* __restrict__ pv,
                                                     int* __restrict__ pi) {
  __shared__ double sv[256];
  __shared__ int si[256];
  double best = -1.0;
  int bi = 0x7fffffff;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double a = abs2(x[i]);
    if (a > best) {
      best = a;
      bi = (int)i;
    }
  }
  sv[threadIdx.x] = best;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      const double v = sv[threadIdx.x + o];
      const int j = si[threadIdx.x + o];
      if (v > sv[threadIdx.x] || (v == sv[threadIdx.x] && j < si[threadIdx.x])) {
        sv[threadIdx.x] = v;
        si[threadIdx.x] = j;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    pv[blockIdx.x] = sv[0];
    pi[blockIdx.x] = si[0];
  }
}
__global__ void k_absmax_final(int nparts, const double* __restrict__ pv, const int* __restrict__ pi,
                               const z128* __restrict__ x, z128* __restrict__ phase) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double best = -1.0;
  int bi = 0;
  for (int b = 0; b < nparts; ++b)
    if (pv[b] > best || (pv[b] == best && pi[b] < bi)) {
      best = pv[b];
      bi = pi[b];
    }
  const z128 v = x[bi];
  const double a = absz(v);
  *phase = a > 0.0 ? conj_(v) * (1.0 / a) : mk(1, 0);
}
__global__ void k_perm_scatter_scaled(const z128* __restrict__ src, z128* __restrict__ dst, const int* __restrict__ perm,
                                      int n, const z128* __restrict__ phase) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[perm[i]] = src[i] * (*phase);
}

// Rows per block of the Gram-Schmidt dot kernels: the grid is a whole number of waves of resident blocks.  The
// register footprint (and with it the number of resident blocks per SM) depends on the column variant; a grid
// that is not a multiple of the resident capacity ends in an almost empty wave (578 blocks on 444 slots: 2.6 TB/s,
// profiles/r2e_ncu_full_ortho_summary.txt).  Measured on config 3 (profiles/r2g_*): 2 waves 0.123 s of
// Gram-Schmidt per step, 1 wave 0.141 s, 3 waves 0.131 s.  LSA_GS_WAVES overrides for experiments.
static int gs_rows_per_block(int n, int j, int num_sms) {
  static int slots_per_sm[5] = {0, 0, 0, 0, 0};
  const int v = j <= 16 ? 0 : j <= 32 ? 1 : j <= 64 ? 2 : j <= 96 ? 3 : 4;
  if (slots_per_sm[v] == 0) {
    int a = 1, b = 1;
    switch (v) {
      case 0: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_dots<2, 8>, 256, 0); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_update_dots<2, 8>, 256, 0); break;
      case 1: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_dots<4, 4>, 256, 0); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_update_dots<4, 4>, 256, 0); break;
      case 2: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_dots<8, 2>, 256, 0); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_update_dots<8, 2>, 256, 0); break;
      case 3: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_dots<12, 2>, 256, 0); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_update_dots<12, 2>, 256, 0); break;
      default: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_dots<16, 1>, 256, 0); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_update_dots<16, 1>, 256, 0); break;
    }
    slots_per_sm[v] = std::max(1, std::min(a, b));
  }
  long long slots = (long long)slots_per_sm[v] * num_sms;
  double waves = 2.0;
  if (const char* e = getenv("LSA_GS_WAVES")) waves = atof(e);
  slots = (long long)(waves * (double)slots);
  slots = std::max<long long>(1, std::min<long long>(slots, 1000));
  const long long rpb = (((long long)n + slots - 1) / slots + 255) / 256 * 256;
  return (int)std::max<long long>(1024, rpb);
}

// ------------------------------------------------------------------ row ranges (partitioned solve)
//
// Single GPU: one range [0, n).  Partitioned: `upd_ranges` = rows this GPU maintains (replicated rows + its
// sub-trees), `dot_ranges` = rows it counts in dot products (its sub-trees; rank 0 also the replicated rows).
typedef std::vector<std::pair<int, int>> Ranges;

static int blocks_of(const Ranges& rg, int per_block) {
  int b = 0;
  for (auto& r : rg) b += cdiv(r.second - r.first, per_block);
  return b;
}
// npart[..] = partial |w|^2 over `rg`; returns the number of partials
static int norm2_partials(lsa_handle_impl& h, const Ranges& rg, const z128* w, double* npart) {
  int boff = 0;
  for (auto& r : rg) {
    const int len = r.second - r.first, nb = cdiv(len, 256);
    if (nb > 0) k_norm2_part<<<nb, 256, 0, h.stream>>>(len, w + r.first, npart + boff);
    boff += nb;
  }
  return boff;
}
// out = w / |w| over the maintained rows; the norm is global (all-reduced when partitioned)
static void normalize_vector(lsa_handle_impl& h, const z128* w, z128* out, double* beta_out, z128* s_entry,
                             const z128* hcol, int hlen, int* flag, int step) {
  cudaStream_t st = h.stream;
  const double* np = h.d_npart;
  int nparts = norm2_partials(h, h.dot_ranges, w, h.d_npart);
  if (h.partitioned) {
    k_sum_partials<<<1, 256, 0, st>>>(h.d_npart, nparts, 1, h.d_red + 516);
    comm_allreduce_sum(h.comm, h.d_red + 516, 1, st);
    np = h.d_red + 516;
    nparts = 1;
  }
  for (auto& r : h.upd_ranges) {
    const int len = r.second - r.first;
    if (len > 0) k_normalize<<<cdiv(len, 256), 256, 0, st>>>(len, w + r.first, out + r.first, np, nparts, beta_out, s_entry, hcol, hlen, flag, step);
  }
  LSA_LAUNCH_CHECK();
}
static void basis_gemm(lsa_handle_impl& h, int mp, int nk, const z128* V, long long ldv, const z128* Qm, int ldq, z128* Out,
                       long long ldo) {
  const size_t smem = sizeof(z128) * (size_t)mp * 33;
  for (auto& r : h.upd_ranges) {
    const int len = r.second - r.first;
    if (len > 0) k_basis_gemm<<<cdiv(len, 32), 256, smem, h.stream>>>(len, mp, nk, V + r.first, ldv, Qm, ldq, Out + r.first, ldo);
  }
  LSA_LAUNCH_CHECK();
}
// vec (n x ncols, ld) complete and identical on every GPU: rows outside dot_ranges zeroed, then summed
void make_full(lsa_handle_impl& h, z128* vec, int ncols, long long ld) {
  if (!h.partitioned) return;
  for (int c = 0; c < ncols; ++c) {
    int prev = 0;
    for (size_t q = 0; q <= h.dot_ranges.size(); ++q) {
      const int lo = q < h.dot_ranges.size() ? h.dot_ranges[q].first : h.n;
      if (lo > prev) LSA_CUDA(cudaMemsetAsync(vec + (long long)c * ld + prev, 0, sizeof(z128) * (size_t)(lo - prev), h.stream));
      if (q < h.dot_ranges.size()) prev = h.dot_ranges[q].second;
    }
    comm_allreduce_sum(h.comm, (double*)(vec + (long long)c * ld), 2 * (size_t)h.n, h.stream);
  }
}

// w <- w - V[:, 0:jj] (V^H w), classical Gram-Schmidt with a second pass if needed (or `force_two`), then
// out = w / |w|.  scol (optional): column of the projected matrix that receives the coefficients and, at
// scol[jj], the norm.  Device-side decisions only; the host enqueues the same launches either way.
static void orthonormalize(lsa_handle_impl& h, z128* V, long long ldv, int jj, z128* w, z128* out, z128* scol,
                           int rows_per_block, int ldp, bool force_two, int* flag, int step) {
  cudaStream_t st = h.stream;
  const int always = (h.ortho_refine_always || force_two) ? 1 : 0;
  auto dots = [&](const int* skip, bool want_wn2) {
    int boff = 0;
    for (auto& r : h.dot_ranges) {
      const int len = r.second - r.first, nb = cdiv(len, rows_per_block);
      if (nb > 0) launch_dots(st, nb, len, jj, V + r.first, ldv, w + r.first, h.d_part + (long long)boff * ldp, ldp, rows_per_block,
                              want_wn2 ? h.d_wn2 + boff : nullptr, skip);
      boff += nb;
    }
    return boff;
  };
  auto update = [&](const int* skip, double* npart) {
    int boff = 0;
    for (auto& r : h.upd_ranges) {
      const int len = r.second - r.first, nb = cdiv(len, 256);
      if (nb > 0) k_update<<<nb, 256, 0, st>>>(len, jj, V + r.first, ldv, h.d_h, w + r.first, npart ? npart + boff : nullptr, skip);
      boff += nb;
    }
    return boff;
  };
  if (!h.partitioned) {
    const int nblk = dots(nullptr, true);
    k_reduce_h<<<cdiv(jj, 8), 256, 0, st>>>(jj, nblk, h.d_part, ldp, h.d_h, scol ? scol : h.d_brow, 0, nullptr);
    const int len = h.upd_ranges[0].second - h.upd_ranges[0].first;
    if (jj <= 128 && h.fuse_ortho) {
      // pass-1 update fused with the pass-2 dot products (V read once for both); norm after pass 1 -> d_wn2b
      launch_update_dots(st, nblk, len, jj, V, ldv, h.d_h, w, h.d_part, ldp, rows_per_block, h.d_wn2b);
      k_refine_flag<<<1, 256, 0, st>>>(nblk, h.d_wn2, nblk, h.d_wn2b, always, h.d_refine, h.d_refine + 1);
      k_reduce_h<<<cdiv(jj, 8), 256, 0, st>>>(jj, nblk, h.d_part, ldp, h.d_h, scol ? scol : h.d_brow, 1, h.d_refine);
      const int nb = update(h.d_refine, h.d_npart);
      k_normalize<<<cdiv(len, 256), 256, 0, st>>>(len, w, out, h.d_npart, nb, nullptr, scol ? scol + jj : nullptr, scol,
                                                   scol ? jj : 0, flag, step, h.d_wn2b, nblk, h.d_refine);
    } else {
      const int nb = update(nullptr, h.d_npart);
      k_refine_flag<<<1, 256, 0, st>>>(nblk, h.d_wn2, nb, h.d_npart, always, h.d_refine, h.d_refine + 1);
      dots(h.d_refine, false);
      k_reduce_h<<<cdiv(jj, 8), 256, 0, st>>>(jj, nblk, h.d_part, ldp, h.d_h, scol ? scol : h.d_brow, 1, h.d_refine);
      update(h.d_refine, h.d_npart);
      k_normalize<<<cdiv(len, 256), 256, 0, st>>>(len, w, out, h.d_npart, nb, nullptr, scol ? scol + jj : nullptr, scol,
                                                   scol ? jj : 0, flag, step);
    }
  } else {
    double* red = h.d_red;
    double* norm2 = h.d_red + 514;
    for (int pass = 1; pass <= 2; ++pass) {
      const int* skip = pass == 2 ? h.d_refine : nullptr;
      const int nblk = dots(skip, true);
      k_reduce_red<<<cdiv(jj + 1, 8), 256, 0, st>>>(jj, nblk, h.d_part, ldp, h.d_wn2, red, skip);
      // every rank enqueues the all-reduce of both passes (the decision lives on the device); a pass that was
      // not needed reduces stale numbers that k_apply_red then ignores
      comm_allreduce_sum(h.comm, red, 2 * (size_t)jj + 1, st);
      k_apply_red<<<1, 256, 0, st>>>(jj, red, h.d_h, scol, pass, always, norm2, h.d_refine, h.d_refine + 1);
      update(skip, nullptr);
    }
    for (auto& r : h.upd_ranges) {
      const int len = r.second - r.first;
      if (len > 0) k_normalize<<<cdiv(len, 256), 256, 0, st>>>(len, w + r.first, out + r.first, norm2, 1, nullptr,
                                                                scol ? scol + jj : nullptr, scol, scol ? jj : 0, flag, step);
    }
  }
  LSA_LAUNCH_CHECK();
}

// beta^2 = w^H M w (real part) -> h.d_bn, with M w left in h.d_mw.  Single GPU.
static void m_norm2(lsa_handle_impl& h, const z128* w) {
  const int n = h.n;
  if (h.has_m) spmv(h, h.dM, false, w, h.d_mw);
  else k_copy<<<cdiv(n, 256), 256, 0, h.stream>>>(n, w, h.d_mw);
  const int rpb = std::max(1024, (int)(((long long)n + 591) / 592 + 31) / 32 * 32);
  const int nblk = cdiv(n, rpb);
  launch_dots(h.stream, nblk, n, 1, w, n, h.d_mw, h.d_part, 256, rpb, nullptr, nullptr);
  k_reduce_h<<<1, 256, 0, h.stream>>>(1, nblk, h.d_part, 256, h.d_bn, h.d_bn + 1, 0, nullptr);
}

// M-inner-product Gram-Schmidt (Hermitian-definite problems, b_mode = 2): U = M V is kept next to V, so
// V^H M w = U^H w costs no extra product with M.  w <- w - V (U^H w), twice (the refinement criterion of the
// Euclidean path compares 2-norms and does not carry over; two passes are what SLEPc's symmetric Krylov-Schur
// ends up doing on shift-and-invert operators anyway), then beta = sqrt(w^H M w), out_v = w / beta,
// out_u = M w / beta.
static void orthonormalize_m(lsa_handle_impl& h, const z128* V, const z128* U, long long ldv, int jj, z128* w, z128* out_v,
                             z128* out_u, z128* scol, int rows_per_block, int ldp, int* flag, int step) {
  cudaStream_t st = h.stream;
  const int n = h.n;
  const int nblk = cdiv(n, rows_per_block);
  for (int pass = 0; pass < 2; ++pass) {
    launch_dots(st, nblk, n, jj, U, ldv, w, h.d_part, ldp, rows_per_block, nullptr, nullptr);
    k_reduce_h<<<cdiv(jj, 8), 256, 0, st>>>(jj, nblk, h.d_part, ldp, h.d_h, scol ? scol : h.d_brow, pass, nullptr);
    k_update<<<cdiv(n, 256), 256, 0, st>>>(n, jj, V, ldv, h.d_h, w, nullptr, nullptr);
  }
  m_norm2(h, w);
  const double* b2 = reinterpret_cast<const double*>(h.d_bn);
  k_normalize<<<cdiv(n, 256), 256, 0, st>>>(n, w, out_v, b2, 1, nullptr, scol ? scol + jj : nullptr, scol, scol ? jj : 0, flag, step);
  k_normalize<<<cdiv(n, 256), 256, 0, st>>>(n, h.d_mw, out_u, b2, 1, nullptr, nullptr, nullptr, 0, nullptr, 0);
  LSA_LAUNCH_CHECK();
}

// a^H w on the device (single GPU): the tall-skinny dot kernel with a one-column "basis"
z128 dot_conj(lsa_handle_impl& h, const z128* a, const z128* w) {
  const int n = h.n;
  const int rows_per_block = std::max(1024, (int)(((long long)n + 591) / 592 + 31) / 32 * 32);
  const int nblk = cdiv(n, rows_per_block);
  launch_dots(h.stream, nblk, n, 1, a, n, w, h.d_part, 256, rows_per_block, nullptr, nullptr);
  k_reduce_h<<<1, 256, 0, h.stream>>>(1, nblk, h.d_part, 256, h.d_h, h.d_brow, 0, nullptr);
  LSA_LAUNCH_CHECK();
  z128 res;
  LSA_CUDA(cudaMemcpyAsync(&res, h.d_h, sizeof(z128), cudaMemcpyDeviceToHost, h.stream));
  LSA_CUDA(cudaStreamSynchronize(h.stream));
  return res;
}

// ------------------------------------------------------------------------------------- OP and driver

void drop_solve_graphs(lsa_handle_impl& h) {
  for (auto& g : h.solve_graphs) cudaGraphExecDestroy(g.exec);
  h.solve_graphs.clear();
}

// One triangular-solve sweep is a fixed sequence of ~10 launches per tree level; it is captured once
// per (trans, vector) into a CUDA graph and replayed, which removes the per-launch CPU cost from the
// latency-bound 2-D cases (a tracing compiler is not involved: plain stream capture of our kernels).
static void solve_dispatch(lsa_handle_impl& h, int trans, z128* x, int* nk) {
  auto run = [&](int* count) {
    if (h.scalar == LSA_C128) solve_permuted<z128>(h, trans, x, count);
    else solve_permuted<double>(h, trans, x, count);
  };
  int local = 0;
  // partitioned solve: the sweep contains an NCCL all-reduce; NCCL kernels are captured like any other launch
  // (the communicator was warmed up outside any capture in lsa_set_comm).  "partition_graphs" = 0 enqueues directly.
  if (!h.use_graphs || (h.partitioned && !h.part_graphs)) {
    run(&local);
  } else {
    lsa_handle_impl::SolveGraph* found = nullptr;
    for (auto& g : h.solve_graphs)
      if (g.trans == trans && g.vec == (const void*)x) found = &g;
    if (!found) {
      cudaGraph_t graph = nullptr;
      LSA_CUDA(cudaStreamBeginCapture(h.stream, cudaStreamCaptureModeThreadLocal));
      try {
        run(&local);
      } catch (...) {
        cudaStreamEndCapture(h.stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      LSA_CUDA(cudaStreamEndCapture(h.stream, &graph));
      cudaGraphExec_t exec = nullptr;
      LSA_CUDA(cudaGraphInstantiate(&exec, graph, 0));
      cudaGraphDestroy(graph);
      h.solve_graphs.push_back({trans, (const void*)x, exec, local});
      found = &h.solve_graphs.back();
    }
    LSA_CUDA(cudaGraphLaunch(found->exec, h.stream));
    local = found->launches;
  }
  h.launch_count += local;
  if (nk) *nk += local;
}

// x <- x - N (N^H x): removes the attached nullspace (orthonormal columns N) from a vector, one Gram-Schmidt pass
// with the basis kernels
static void remove_nullspace(lsa_handle_impl& h, z128* x) {
  const int n = h.n, j = h.ns_count;
  const int rows_per_block = std::max(1024, (int)(((long long)n + 591) / 592 + 31) / 32 * 32);
  const int nblk = cdiv(n, rows_per_block);
  launch_dots(h.stream, nblk, n, j, h.d_ns, n, x, h.d_part, 256, rows_per_block, nullptr, nullptr);
  k_reduce_h<<<cdiv(j, 8), 256, 0, h.stream>>>(j, nblk, h.d_part, 256, h.d_h, h.d_brow, 0, nullptr);
  k_update<<<cdiv(n, 256), 256, 0, h.stream>>>(n, j, h.d_ns, n, h.d_h, x, nullptr, nullptr);
  LSA_LAUNCH_CHECK();
}

// x <- F^-1 x (or F^-H x) with optional iterative refinement against F = alpha A + beta M.
void op_solve(lsa_handle_impl& h, int trans, z128* x, int refine_steps) {
  int nk = 0;
  const int n = h.n;
  if (h.ns_count > 0) {
    // singular F with a known nullspace: compatible right-hand side in, nullspace-free solution out
    remove_nullspace(h, x);
    solve_dispatch(h, trans, x, &nk);
    remove_nullspace(h, x);
    return;
  }
  if (refine_steps <= 0) {
    solve_dispatch(h, trans, x, &nk);
    return;
  }
  if (h.partitioned) throw ArgError("iterative refinement (refine_steps > 0) is not available in the partitioned solve");
  // keep b in d_w2, iterate x_{k+1} = x_k + F^-1 (b - F x_k)
  z128* b = h.d_r1;
  z128* r = h.d_r2;
  z128* t = h.d_r3;
  const int blocks = cdiv(n, 256);
  k_copy<<<blocks, 256, 0, h.stream>>>(n, x, b);
  solve_dispatch(h, trans, x, &nk);
  const bool H = trans == LSA_OP_H;
  for (int it = 0; it < refine_steps; ++it) {
    // r = b - (alpha A + beta M) x     (H: conj(alpha) A^H + conj(beta) M^H)
    k_copy<<<blocks, 256, 0, h.stream>>>(n, b, r);
    const z128 al = H ? conj_(h.f_alpha) : h.f_alpha, be = H ? conj_(h.f_beta) : h.f_beta;
    if (al.x != 0.0 || al.y != 0.0) {
      spmv(h, H ? h.dAt : h.dA, H, x, t);
      k_axpy<<<blocks, 256, 0, h.stream>>>(n, mk(0, 0) - al, t, r);
    }
    if (h.has_m && (be.x != 0.0 || be.y != 0.0)) {
      spmv(h, H ? h.dMt : h.dM, H, x, t);
      k_axpy<<<blocks, 256, 0, h.stream>>>(n, mk(0, 0) - be, t, r);
    }
    solve_dispatch(h, trans, r, &nk);
    k_axpy<<<blocks, 256, 0, h.stream>>>(n, mk(1, 0), r, x);
  }
  LSA_LAUNCH_CHECK();
}

struct EventTimer {
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
  cudaStream_t st;
  explicit EventTimer(cudaStream_t s) : st(s) {}
  size_t begin() {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, st);
    ev.emplace_back(a, b);
    return ev.size() - 1;
  }
  void end(size_t i) { cudaEventRecord(ev[i].second, st); }
  double total_seconds() {
    double s = 0;
    for (auto& e : ev) {
      float ms = 0;
      cudaEventElapsedTime(&ms, e.first, e.second);
      s += ms * 1e-3;
      cudaEventDestroy(e.first);
      cudaEventDestroy(e.second);
    }
    ev.clear();
    return s;
  }
};

// w = OP v   (all vectors in the permuted ordering)
static void apply_op(lsa_handle_impl& h, const lsa_eigs_params& p, const z128* v, z128* w, EventTimer& t_spmv,
                     EventTimer& t_solve, const z128* mv = nullptr) {
  const bool adj = p.adjoint != 0;
  const int n = h.n, blocks = cdiv(n, 256);
  if (p.transform == LSA_ST_SINVERT) {
    size_t e = t_spmv.begin();
    if (mv) k_copy<<<blocks, 256, 0, h.stream>>>(n, mv, w);   // M v is at hand (M-inner-product mode keeps U = M V)
    else if (h.has_m) spmv(h, adj ? h.dMt : h.dM, adj, v, w);
    else {
      k_copy<<<blocks, 256, 0, h.stream>>>(n, v, w);
      replicated_rows_to_partial(h, w);   // the sweep expects replicated rows that SUM to the value over the GPUs
    }
    t_spmv.end(e);
    e = t_solve.begin();
    op_solve(h, adj ? LSA_OP_H : LSA_OP_N, w, p.refine_steps);
    t_solve.end(e);
  } else {
    // shift: OP = M^-1 A - sigma I   (standard problem: A - sigma I)
    size_t e = t_spmv.begin();
    spmv(h, adj ? h.dAt : h.dA, adj, v, w);
    t_spmv.end(e);
    if (h.has_m) {
      e = t_solve.begin();
      op_solve(h, adj ? LSA_OP_H : LSA_OP_N, w, p.refine_steps);
      t_solve.end(e);
    } else if (h.partitioned) {
      exchange_replicated_rows(h, w, false);   // no sweep follows: sum the partial products of the replicated rows here
    }
    z128 sg = mk(p.sigma_re, p.sigma_im);
    if (adj) sg = conj_(sg);
    if (sg.x != 0.0 || sg.y != 0.0) k_axpy<<<blocks, 256, 0, h.stream>>>(n, mk(0, 0) - sg, v, w);
  }
  LSA_LAUNCH_CHECK();
}

void run_eigs(lsa_handle_impl& h, const lsa_eigs_params& p, lsa_eigs_result& out) {
  const int n = h.n;
  cudaStream_t st = h.stream;
  const int ncv = std::max(1, std::min(p.ncv, n));
  const int nev = std::max(1, std::min(p.nev, n));
  const int ld = ncv + 1;
  const int blocks = cdiv(n, 256);
  z128* V = h.d_V;
  const long long ldv = n;
  z128* S = h.d_S;
  z128* Q = h.d_Q;
  const int rows_per_block = std::max(1024, (int)(((long long)n + 591) / 592 + 31) / 32 * 32);
  const int nblk = cdiv(n, rows_per_block);
  const int ldp = 256;
  z128 sigma = mk(p.sigma_re, p.sigma_im);
  if (p.adjoint) sigma = conj_(sigma);

  {
    static PerDeviceOnce gemm_attr;
    if (gemm_attr.first())
      LSA_CUDA(cudaFuncSetAttribute(k_basis_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, 257 * 33 * (int)sizeof(z128)));
  }
  EventTimer t_all(st), t_spmv(st), t_solve(st), t_ortho(st), t_rr(st), t_restart(st);
  const long long launches0 = h.launch_count;
  const size_t e_all = t_all.begin();

  LSA_CUDA(cudaMemsetAsync(S, 0, sizeof(z128) * (size_t)ld * ncv, st));
  LSA_CUDA(cudaMemsetAsync(h.d_refine, 0, 2 * sizeof(int), st));
  int h_flag[2] = {0x7fffffff, 0};   // [0] first Arnoldi step that broke down, [1] NaN / Inf met
  LSA_CUDA(cudaMemcpyAsync(h.d_flag, h_flag, 2 * sizeof(int), cudaMemcpyHostToDevice, st));

  // ---- start vector: random (or caller supplied), pushed through OP once (range of OP; M singular)
  if (p.v0) {
    LSA_CUDA(cudaMemcpyAsync(h.d_io, p.v0, sizeof(z128) * (size_t)n, cudaMemcpyHostToDevice, st));
    permute_gather(st, h.d_io, h.d_x, h.d_perm, n);
  } else {
    k_randn<<<blocks, 256, 0, st>>>(h.d_x, n, p.seed);
  }
  int n_applies = 0, n_arnoldi = 0;
  long long sum_cols = 0;
  const bool bmode = p.b_mode == 2;   // M-inner products (Hermitian-definite problem)
  z128* U = h.d_U;
  apply_op(h, p, h.d_x, h.d_w, t_spmv, t_solve);
  n_applies++;
  // (flag: a NaN / Inf start vector or first operator application is reported through flag[1]; the step number is
  // out of range, so a vanishing norm here does not register as an Arnoldi breakdown)
  if (bmode) {
    m_norm2(h, h.d_w);
    const double* b2 = reinterpret_cast<const double*>(h.d_bn);
    k_normalize<<<blocks, 256, 0, st>>>(n, h.d_w, V, b2, 1, nullptr, nullptr, nullptr, 0, h.d_flag, 0x7fffffff);
    k_normalize<<<blocks, 256, 0, st>>>(n, h.d_mw, U, b2, 1, nullptr, nullptr, nullptr, 0, nullptr, 0);
  } else {
    normalize_vector(h, h.d_w, V, nullptr, nullptr, nullptr, 0, h.d_flag, 0x7fffffff);
  }

  int nconv = 0, keep = 0, restarts = 0, breakdown = 0, m_last = ncv;
  bool invariant = false;
  RrInfo info{};
  while (true) {
    restarts++;
    int m = ncv;
    // ---- Arnoldi expansion with CGS2
    for (int j = keep; j < ncv; ++j) {
      apply_op(h, p, V + (long long)j * ldv, h.d_w, t_spmv, t_solve, bmode ? U + (long long)j * ldv : nullptr);
      n_applies++;
      const size_t e = t_ortho.begin();
      const int jj = j + 1;  // orthogonalise against columns 0..j
      n_arnoldi++;
      sum_cols += jj;
      z128* scol = S + (long long)j * ld;
      // classical Gram-Schmidt, second pass only when the criterion asks for it (decided on the device: the
      // pass-2 kernels return at once when the flag is clear)
      if (bmode)
        orthonormalize_m(h, V, U, ldv, jj, h.d_w, V + (long long)(j + 1) * ldv, U + (long long)(j + 1) * ldv, scol,
                         gs_rows_per_block(n, jj, h.num_sms), ldp, h.d_flag, j);
      else
        orthonormalize(h, V, ldv, jj, h.d_w, V + (long long)(j + 1) * ldv, scol, gs_rows_per_block(n, jj, h.num_sms), ldp, false, h.d_flag, j);
      h.launch_count += 9;  // spmv + 8 orthogonalisation kernels
      t_ortho.end(e);
    }
    // ---- breakdown check (one small read-back per restart)
    LSA_CUDA(cudaMemcpyAsync(h_flag, h.d_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    LSA_CUDA(cudaStreamSynchronize(st));
    if (h_flag[1]) throw NonFiniteError("NaN/Inf met in the Krylov basis (non-finite triangular solve or SpMV result)");
    if (h_flag[0] < ncv) {
      m = h_flag[0] + 1;
      breakdown = 1;
    }
    // an exhausted Krylov space (breakdown, or m = n) is an exactly invariant subspace
    invariant = (h_flag[0] < ncv) || m >= n;
    m_last = m;
    const double beta_scale = invariant ? 0.0 : 1.0;
    // ---- Rayleigh-Ritz
    RrParams rp{};
    rp.m = m; rp.ld = ld; rp.ldq = ncv; rp.nconv = nconv; rp.nev = nev; rp.which = p.which;
    rp.transform = p.transform; rp.tol = p.tol; rp.sigma = sigma; rp.beta_scale = beta_scale;
    rp.last = (restarts >= p.max_restarts) || m >= n;
    size_t e = t_rr.begin();
    {
      const size_t rr_smem = sizeof(z128) * ((size_t)(m + 1) * m + (size_t)m * m);
      const int use_smem = rr_smem <= 210 * 1024;
      static PerDeviceOnce rr_attr;
      if (use_smem && rr_attr.first())
        LSA_CUDA(cudaFuncSetAttribute(k_rr, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
      k_rr<<<1, 128, use_smem ? rr_smem : 0, st>>>(S, Q, rp, h.d_theta, h.d_resid, h.d_brow, h.d_ywork, h.d_rr, use_smem);
    }
    LSA_LAUNCH_CHECK();
    h.launch_count += 3;  // rr + restart gemm + copy
    t_rr.end(e);
    LSA_CUDA(cudaMemcpyAsync(&info, h.d_rr, sizeof(RrInfo), cudaMemcpyDeviceToHost, st));
    LSA_CUDA(cudaStreamSynchronize(st));
    if (info.status != 0) throw std::runtime_error("Rayleigh-Ritz: QR iteration did not converge");
    const int nconv_old = nconv;
    nconv = info.nconv;
    keep = info.keep;
    // ---- restart: V[:, nconv_old:keep] = V[:, nconv_old:m] Q[nconv_old:m, nconv_old:keep]
    e = t_restart.begin();
    const int mp = m - nconv_old, nk = keep - nconv_old;
    if (nk > 0) {
      basis_gemm(h, mp, nk, V + (long long)nconv_old * ldv, ldv, Q + nconv_old + (long long)nconv_old * ncv, ncv,
                 V + (long long)nconv_old * ldv, ldv);
      if (bmode)   // U = M V follows the same rotation
        basis_gemm(h, mp, nk, U + (long long)nconv_old * ldv, ldv, Q + nconv_old + (long long)nconv_old * ncv, ncv,
                   U + (long long)nconv_old * ldv, ldv);
    }
    const bool done = rp.last || nconv >= nev;
    if (!done && invariant) {
      // the invariant subspace found so far is locked (keep == m); continue from a fresh random
      // direction orthogonal to it (SLEPc does the same after a breakdown)
      z128* vnew = V + (long long)keep * ldv;
      k_randn<<<blocks, 256, 0, st>>>(h.d_w, n, p.seed + 7919ULL * (unsigned long long)restarts);
      if (bmode) orthonormalize_m(h, V, U, ldv, keep, h.d_w, vnew, U + (long long)keep * ldv, nullptr, gs_rows_per_block(n, keep, h.num_sms), ldp, nullptr, 0);
      else orthonormalize(h, V, ldv, keep, h.d_w, vnew, nullptr, gs_rows_per_block(n, keep, h.num_sms), ldp, true, nullptr, 0);
      h_flag[0] = 0x7fffffff;
      LSA_CUDA(cudaMemcpyAsync(h.d_flag, h_flag, 2 * sizeof(int), cudaMemcpyHostToDevice, st));
      LSA_LAUNCH_CHECK();
    } else if (!done && keep != m) {
      k_copy<<<blocks, 256, 0, st>>>(n, V + (long long)m * ldv, V + (long long)keep * ldv);
      if (bmode) k_copy<<<blocks, 256, 0, st>>>(n, U + (long long)m * ldv, U + (long long)keep * ldv);
    }
    t_restart.end(e);
    if (done) break;
  }

  // ---- Ritz vectors of the converged block, purification, normalisation, un-permutation
  h.nconv = nconv;
  h.eigenvalues.assign(nconv, mk(0, 0));
  h.eig_order.assign(nconv, 0);
  if (nconv > 0) {
    if (nconv > h.X_cols) {
      if (h.d_X) cudaFree(h.d_X);
      LSA_CUDA(cudaMalloc(&h.d_X, sizeof(z128) * (size_t)n * nconv));
      h.X_cols = nconv;
    }
    // purify = 1: purification through the Krylov-Schur relation (no operator application): the next basis
    // vector v_{m+1} joins the Ritz-vector product with coefficient (b . y) / theta.  purify = 2: explicit OP apply.
    const bool free_purify = p.purify == 1 && !invariant && nconv < ncv;
    if (free_purify && nconv != m_last)
      k_copy<<<blocks, 256, 0, st>>>(n, V + (long long)m_last * ldv, V + (long long)nconv * ldv);
    k_ritz_vectors<<<cdiv(nconv, 64), 64, 0, st>>>(S, ld, nconv, Q, ncv, free_purify ? h.d_brow : nullptr);
    const int mpx = nconv + (free_purify ? 1 : 0);
    // Xp (permuted) staged in the tail of the basis is not possible in general -> use d_Xp
    basis_gemm(h, mpx, nconv, V, ldv, Q, ncv, h.d_Xp, n);
    // partitioned solve: from here on complete vectors, identical on every GPU (the result is handed out whole)
    make_full(h, h.d_Xp, nconv, n);
    for (int i = 0; i < nconv; ++i) {
      z128* xi = h.d_Xp + (long long)i * n;
      if (p.purify == 2) {
        apply_op(h, p, xi, h.d_w, t_spmv, t_solve);
        n_applies++;
        k_norm2_part<<<blocks, 256, 0, st>>>(n, h.d_w, h.d_npart);
        k_normalize<<<blocks, 256, 0, st>>>(n, h.d_w, xi, h.d_npart, blocks, nullptr, nullptr, nullptr, 0, nullptr, 0);
      } else if (p.b_mode == 0) {
        k_norm2_part<<<blocks, 256, 0, st>>>(n, xi, h.d_npart);
        k_normalize<<<blocks, 256, 0, st>>>(n, xi, xi, h.d_npart, blocks, nullptr, nullptr, nullptr, 0, nullptr, 0);
      }
      if (p.b_mode != 0) {
        // Hermitian-definite problem: unit M-norm, x^H M x = 1 (SLEPc's normalisation for EPS_GHEP)
        m_norm2(h, xi);
        k_normalize<<<blocks, 256, 0, st>>>(n, xi, xi, reinterpret_cast<const double*>(h.d_bn), 1, nullptr, nullptr, nullptr, 0, nullptr, 0);
      }
      {
        const int nb = std::min(blocks, 256);
        k_absmax_part<<<nb, 256, 0, st>>>(n, xi, h.d_npart, h.d_ipart);
        k_absmax_final<<<1, 32, 0, st>>>(nb, h.d_npart, h.d_ipart, xi, h.d_h);
        k_perm_scatter_scaled<<<blocks, 256, 0, st>>>(xi, h.d_X + (long long)i * n, h.d_perm, n, h.d_h);
      }
    }
    LSA_LAUNCH_CHECK();
    std::vector<z128> theta(nconv);
    LSA_CUDA(cudaMemcpyAsync(theta.data(), h.d_theta, sizeof(z128) * nconv, cudaMemcpyDeviceToHost, st));
    LSA_CUDA(cudaStreamSynchronize(st));
    RrParams kp{};
    kp.which = p.which; kp.transform = p.transform; kp.sigma = sigma;
    std::vector<double> key(nconv);
    for (int i = 0; i < nconv; ++i) {
      h.eigenvalues[i] = rr_back(kp, theta[i]);
      key[i] = rr_key(kp, theta[i]);
      h.eig_order[i] = i;
    }
    std::stable_sort(h.eig_order.begin(), h.eig_order.end(), [&](int a, int b) { return key[a] < key[b]; });
    std::vector<z128> sorted(nconv);
    for (int i = 0; i < nconv; ++i) sorted[i] = h.eigenvalues[h.eig_order[i]];
    h.eigenvalues = sorted;
  }
  t_all.end(e_all);
  int n_reorth = 0;
  LSA_CUDA(cudaMemcpyAsync(&n_reorth, h.d_refine + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  LSA_CUDA(cudaStreamSynchronize(st));
  out.n_reorth = n_reorth;
  out.n_arnoldi = n_arnoldi;
  out.sum_cols = sum_cols;
  out.nconv = nconv;
  out.n_restarts = restarts;
  out.n_op_applies = n_applies;
  out.breakdown = breakdown;
  out.seconds = t_all.total_seconds();
  out.seconds_solve = t_solve.total_seconds();
  out.seconds_spmv = t_spmv.total_seconds();
  out.seconds_ortho = t_ortho.total_seconds();
  out.seconds_rr = t_rr.total_seconds();
  out.seconds_restart = t_restart.total_seconds();
  out.n_kernels = (int)(h.launch_count - launches0);
}

// ||A x - lambda M x|| / (||A||_F ||x||) for every returned pair (original ordering)
__global__ void __launch_bounds__(256) k_resid_part(int n, const z128* __restrict__ ax, const z128* __restrict__ mx,
                                                    z128 lam, const z128* __restrict__ x, double* __restrict__ part) {
  __shared__ double r1[8], r2[8];
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double a = 0, b = 0;
  if (i < n) {
    a = abs2(ax[i] - lam * mx[i]);
    b = abs2(x[i]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) {
    r1[threadIdx.x >> 5] = a;
    r2[threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s1 = 0, s2 = 0;
    for (int q = 0; q < 8; ++q) {
      s1 += r1[q];
      s2 += r2[q];
    }
    part[2 * blockIdx.x] = s1;
    part[2 * blockIdx.x + 1] = s2;
  }
}

void residual_norms(lsa_handle_impl& h, double* out_host, int count) {
  const int n = h.n, blocks = cdiv(n, 256);
  const bool adj = h.last_params.adjoint != 0;
  double* d_part = nullptr;
  LSA_CUDA(cudaMalloc(&d_part, sizeof(double) * 2 * (blocks + (int)h.dot_ranges.size())));
  for (int i = 0; i < count && i < h.nconv; ++i) {
    const int col = h.eig_order[i];
    // work in the permuted ordering: x_p = gather(x)
    permute_gather(h.stream, h.d_X + (long long)col * n, h.d_x, h.d_perm, n);
    spmv(h, adj ? h.dAt : h.dA, adj, h.d_x, h.d_w);
    if (h.has_m) spmv(h, adj ? h.dMt : h.dM, adj, h.d_x, h.d_r1);
    else {
      k_copy<<<blocks, 256, 0, h.stream>>>(n, h.d_x, h.d_r1);
      replicated_rows_to_partial(h, h.d_r1);
    }
    if (h.partitioned) {   // each GPU multiplied the columns it owns: complete the replicated rows
      exchange_replicated_rows(h, h.d_w, false);
      exchange_replicated_rows(h, h.d_r1, false);
    }
    int boff = 0;
    for (auto& r : h.dot_ranges) {
      const int len = r.second - r.first, nb = cdiv(len, 256);
      if (nb > 0) k_resid_part<<<nb, 256, 0, h.stream>>>(len, h.d_w + r.first, h.d_r1 + r.first, h.eigenvalues[i], h.d_x + r.first, d_part + 2 * boff);
      boff += nb;
    }
    k_sum_partials<<<1, 256, 0, h.stream>>>(d_part, boff, 2, h.d_red + 516);
    k_sum_partials<<<1, 256, 0, h.stream>>>(d_part + 1, boff, 2, h.d_red + 517);
    LSA_LAUNCH_CHECK();
    comm_allreduce_sum(h.comm, h.d_red + 516, 2, h.stream);
    double s12[2] = {0, 0};
    LSA_CUDA(cudaMemcpyAsync(s12, h.d_red + 516, sizeof(s12), cudaMemcpyDeviceToHost, h.stream));
    LSA_CUDA(cudaStreamSynchronize(h.stream));
    out_host[i] = std::sqrt(s12[0]) / (std::max(h.a_fro, 1e-300) * std::sqrt(std::max(s12[1], 1e-300)));
  }
  cudaFree(d_part);
}

}  // namespace lsa
