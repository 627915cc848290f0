// Numeric multifrontal LU of  F = alpha A + beta M  on one B200, real or complex FP64.
//
// Stands in for PETSc's MatLUFactorNumeric / MUMPS' numeric phase reached from
// `iEpsSolver.set_st_pc_type(LU)` + `solve()` (reference Solver/utils.py:261-270; explicit in
// Solver/eigen2.py:109-151).  Structure is static (restricted partial pivoting inside the pivot
// block of each front + tiny-pivot replacement), so the whole factorisation is a fixed sequence
// of batched launches over assembly-tree levels:
//
//   scatter      F entries -> front panels                      (HBM bound, bytes only)
//   extend_add   children contribution blocks -> parent front   (HBM bound, smem-staged maps)
//   panel_lu     nb-wide pivot panel, pivot search in rows [j, k)
//   swap_trsm    row interchanges + U-row block (unit-lower solve, smem-staged tiles)
//   trsm_cols    L21 panel (upper solve from the right)
//   gemm         trailing update / Schur complement  C -= A B   (FP64 DMMA tensor pipe)
//
// The GEMM is the one dense contraction; it runs on the FP64 tensor pipe with
// mma.sync.m8n8k4.f64 (SASS: DMMA) -- tcgen05 has no f64 kind on sm_100a.  Complex products use the
// real embedding  C~ = A^ B~  on the interleaved storage (A^ built on the fly in the fragment
// loader), i.e. 4 real DMMAs per complex multiply-add and no planar copies.
#include "factor.cuh"

namespace lsa {

static constexpr int NB = 32;   // panel width
static constexpr int PANEL_SMEM_CAP = 192 * 1024;  // shared-memory budget of the panel kernel
static constexpr int OB = 128;  // outer block: trailing updates beyond it are deferred and done with K = 128

// ------------------------------------------------------------------------------------------ scatter

template <class T, class VT>
__global__ void k_scatter(T* __restrict__ fac, const long long* __restrict__ dst, const VT* __restrict__ vals,
                          long long nnz, z128 coef, int accumulate) {
  long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; e < nnz; e += stride) {
    if (dst[e] < 0) continue;   // partitioned solve: the entry belongs to a front of another GPU
    z128 v;
    if constexpr (sizeof(VT) == 16) v = coef * vals[e];
    else v = coef * (double)vals[e];
    T t = scalar_traits<T>::from(v);
    T* p = fac + dst[e];
    if (accumulate) *p = *p + t;
    else *p = t;
  }
}

template <class T>
__global__ void k_decoupled_pivots(T* __restrict__ diag, int n_iso, double tiny_abs, DevStats* st) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_iso) return;
  T d = diag[i];
  double a = abs1(d);
  if (!(a == a) || isinf(a)) {
    st->nonfinite = 1;
    return;
  }
  if (a <= tiny_abs) {
    if (tiny_abs == 0.0) {
      st->zero_pivot = 1;
      return;
    }
    diag[i] = scalar_traits<T>::from(mk(tiny_abs, 0.0));
    atomicAdd(&st->n_perturbed, 1ULL);
    a = tiny_abs;
  }
  atomicMin(&st->min_piv_bits, (unsigned long long)__double_as_longlong(a));
  atomicMax(&st->max_piv_bits, (unsigned long long)__double_as_longlong(a));
}

// --------------------------------------------------------------------------------------- extend-add

// grid: (column groups, fronts of the level); block: 256 threads.
// Adds the contribution block of the `slot`-th child of each parent into the parent's P / Q / C.
// Children are processed slot by slot in separate launches, so no two blocks ever add to the same
// destination concurrently: deterministic, no atomics.
template <class T>
__global__ void __launch_bounds__(256, 4) k_extend_add(const Front* __restrict__ fronts, const int* __restrict__ lvl_front, int first,
                             const int* __restrict__ child_idx, const int* __restrict__ ea_map, int slot,
                             T* __restrict__ fac, const T* __restrict__ pool_child, T* __restrict__ pool_parent,
                             T* __restrict__ pool_cut, int symmetric) {
  constexpr int CHUNK = 2048;
  __shared__ int s_map[CHUNK];
  const Front p = fronts[lvl_front[first + blockIdx.y]];
  if (slot >= p.nchild) return;
  const Front c = fronts[child_idx[p.child0 + slot]];
  const int rc = c.r;
  if (rc == 0) return;
  const int* map = ea_map + c.st0;
  // flagged fronts (sub-tree roots of the partitioned solve) keep their contribution block in the cut pool
  const T* cb = (c.flags ? pool_cut : pool_child) + c.c_off;
  T* P = fac + p.p_off;
  T* Q = fac + p.q_off;
  T* C = (p.flags ? pool_cut : pool_parent) + p.c_off;
  const long long kp = p.k, rp = p.r, mp = kp + rp;
  if (symmetric) {
    // Symmetric factorisation: only the lower triangle of a contribution block is computed (tile rows >= tile columns)
    // and only entries on / below the diagonal are read.  The map is increasing, so they land on / below the parent's
    // diagonal: L21 and the lower triangle of C directly; inside the pivot block (factored as a full square) the
    // mirror image is written as well.  U12 is not stored.
    for (int base = 0; base < rc; base += CHUNK) {
      const int len = min(CHUNK, rc - base);
      __syncthreads();
      for (int t = threadIdx.x; t < len; t += blockDim.x) s_map[t] = map[base + t];
      __syncthreads();
      for (int b = blockIdx.x; b < base + len; b += gridDim.x) {
        const long long jp = map[b];
        const T* col = cb + (long long)b * rc + base;
        for (int t = max(b - base, 0) + threadIdx.x; t < len; t += blockDim.x) {
          const long long ip = s_map[t];
          const T v = col[t];
          if (jp < kp) {
            T* d = P + jp * mp + ip;
            *d = *d + v;
            if (ip < kp && ip != jp) {
              T* u = P + ip * mp + jp;
              *u = *u + v;
            }
          } else {
            T* d = C + (ip - kp) + (jp - kp) * rp;
            *d = *d + v;
          }
        }
      }
    }
    return;
  }
  for (int base = 0; base < rc; base += CHUNK) {
    const int len = min(CHUNK, rc - base);
    __syncthreads();
    for (int t = threadIdx.x; t < len; t += blockDim.x) s_map[t] = map[base + t];
    __syncthreads();
    // four child columns per pass: their loads (child entry + parent entry each) are issued together and the four
    // read-modify-writes completed afterwards -- the destinations are distinct (the map is injective), which the
    // compiler cannot know; one column at a time every entry was a dependent L2 round trip
    for (int b0 = blockIdx.x; b0 < rc; b0 += 4 * gridDim.x) {
      long long jp[4];
      const T* col[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int b = b0 + u * gridDim.x;
        ok[u] = b < rc;
        jp[u] = ok[u] ? map[b] : 0;
        col[u] = cb + (long long)(ok[u] ? b : 0) * rc + base;
      }
      for (int t = threadIdx.x; t < len; t += blockDim.x) {
        const long long ip = s_map[t];
        T* d[4];
        T v[4], o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          d[u] = jp[u] < kp ? P + jp[u] * mp + ip : ip < kp ? Q + ip + (jp[u] - kp) * kp : C + (ip - kp) + (jp[u] - kp) * rp;
          if (ok[u]) {
            v[u] = col[u][t];
            o[u] = *d[u];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (ok[u]) *d[u] = o[u] + v[u];
      }
    }
  }
}

// ----------------------------------------------------------------------------------------- panel LU

struct ArgMax {
  double v;
  int i;
};
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

// One CTA per front.  Factors columns [j0, j0+jb) of the pivot rows [j0, k) of P with partial
// pivoting restricted to those rows; interchanges are applied inside the panel only (the rest of
// the row is swapped by k_swap_trsm).  The panel is staged in shared memory whenever it fits
// (`smem_bytes`), so the 32 dependent column steps run at shared-memory latency; taller panels are
// processed in place in global memory (L2).
// (A register-resident variant -- thread = row, 32 entries in registers, two barriers per column step -- was
// measured in round 2 (profiles/r2s_*, r2t_*): 78 us per 32-column panel at the tree top against 65 us here, and 3.5 x
// slower on the leaf levels, where its 255 registers leave two CTAs per SM; a column step is a ~2 us chain of
// shuffle / barrier / FP64-division latencies either way.)
template <class T>
__global__ void __launch_bounds__(256) k_panel_lu(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                                  int first, int j0, T* __restrict__ fac, int* __restrict__ ipiv,
                                                  int* __restrict__ wperm, double tiny_abs, DevStats* st, int smem_bytes,
                                                  int ob0, int nopivot) {
  const Front f = fronts[lvl_front[first + blockIdx.x]];
  const int k = f.k;
  if (k <= j0) return;
  const long long m = (long long)f.k + f.r;
  const int jb = min(NB, k - j0), pk = k - j0;
  // pivot candidates: rows of the current outer block only.  Rows below it have not yet received
  // the deferred rank-128 updates in the columns right of the block, so they must not be swapped in.
  const int pcand = min(pk, ob0 + OB - j0);
  // ... and ONLY those rows are factored here (<= 128 x 32, always fits in shared memory); the rows
  // below the outer block get their L entries from k_trsm_cols, in parallel over many CTAs.
  T* G = fac + f.p_off + j0 + (long long)j0 * m;  // panel origin in global memory
  extern __shared__ unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  const bool use_smem = (long long)pcand * jb * (long long)sizeof(T) <= (long long)smem_bytes;
  __shared__ ArgMax s_red[8];
  __shared__ int s_piv;
  __shared__ T s_inv;
  __shared__ T s_cinv[OB];     // reciprocals of the pivot candidates of the current column
  __shared__ int s_orig[OB];   // which row of the pivot window (before this panel's interchanges) sits at each position
  for (int i = threadIdx.x; i < OB; i += blockDim.x) s_orig[i] = i;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  T* base = G;
  long long ld = m;
  double lmax = 0.0;  // largest multiplier (growth monitor)
  if (use_smem) {
    for (int e = tid; e < pcand * jb; e += blockDim.x) {
      const int il = e % pcand, cl = e / pcand;
      sm[il + (long long)cl * pcand] = G[il + (long long)cl * m];
    }
    __syncthreads();
    base = sm;
    ld = pcand;
  }
  for (int jl = 0; jl < jb; ++jl) {
    // 1. pivot search in column jl, local rows [jl, pk)
    ArgMax best{-1.0, 0x7fffffff};
    // nopivot (symmetric factorisation): the diagonal entry is the pivot; only its size is checked
    for (int i = jl + tid; i < (nopivot ? jl + 1 : pcand); i += blockDim.x) {
      const T v = base[i + jl * ld];
      double a = abs1(v);
      if (!(a == a)) a = INFINITY;  // propagate NaN as "largest" so it is detected below
      best = better(best, ArgMax{a, i});
      // every candidate's reciprocal, computed while the search reduces (independent instructions): the divisions
      // leave the single-thread section between the barriers below
      s_cinv[i] = recip(v);
    }
    for (int o = 16; o > 0; o >>= 1) {
      ArgMax other{__shfl_down_sync(0xffffffffu, best.v, o), __shfl_down_sync(0xffffffffu, best.i, o)};
      best = better(best, other);
    }
    if (lane == 0) s_red[wid] = best;
    __syncthreads();
    if (tid == 0) {
      ArgMax b = s_red[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) b = better(b, s_red[w]);
      int piv = b.i;
      double a = b.v;
      bool replaced = false;
      if (isinf(a)) st->nonfinite = 1;
      if (a <= tiny_abs) {
        replaced = true;
        if (tiny_abs == 0.0) {
          st->zero_pivot = 1;
          a = 1.0;  // keep going with a harmless value; the host reports the error
          base[piv + jl * ld] = scalar_traits<T>::one();
        } else {
          // static pivoting: keep the diagonal candidate, replace by tiny_abs with its phase
          piv = jl;
          T d = base[jl + jl * ld];
          double ad = absz(d);
          T repl = ad > 0.0 ? d * (tiny_abs / ad) : scalar_traits<T>::from(mk(tiny_abs, 0.0));
          base[jl + jl * ld] = repl;
          atomicAdd(&st->n_perturbed, 1ULL);
          a = tiny_abs;
        }
      }
      atomicMin(&st->min_piv_bits, (unsigned long long)__double_as_longlong(a));
      atomicMax(&st->max_piv_bits, (unsigned long long)__double_as_longlong(a));
      if (piv != jl) atomicAdd(&st->n_swaps, 1ULL);
      ipiv[f.col0 + j0 + jl] = j0 + piv;  // front-local row index
      s_piv = piv;
      const int o = s_orig[jl];
      s_orig[jl] = s_orig[piv];
      s_orig[piv] = o;
      s_inv = replaced ? recip(base[piv + jl * ld]) : s_cinv[piv];   // (the pivot is still in its old row)
    }
    __syncthreads();
    // 2. interchange inside the panel
    const int piv = s_piv;
    if (piv != jl && tid < jb) {
      T a = base[jl + tid * ld], b = base[piv + tid * ld];
      base[jl + tid * ld] = b;
      base[piv + tid * ld] = a;
    }
    __syncthreads();
    // 3. scale the column and rank-1 update of the remaining panel columns
    const T inv = s_inv;
    for (int i = jl + 1 + tid; i < pcand; i += blockDim.x) {
      const T l = base[i + jl * ld] * inv;
      lmax = fmax(lmax, abs1(l));
      base[i + jl * ld] = l;
      // (loads of four columns issued together: written as one statement per column, every load would wait for the
      // previous column's store -- the compiler cannot tell that they never alias)
      int c = jl + 1;
      for (; c + 4 <= jb; c += 4) {
        const T p0 = base[jl + c * ld], p1 = base[jl + (c + 1) * ld], p2 = base[jl + (c + 2) * ld], p3 = base[jl + (c + 3) * ld];
        const T a0 = base[i + c * ld], a1 = base[i + (c + 1) * ld], a2 = base[i + (c + 2) * ld], a3 = base[i + (c + 3) * ld];
        base[i + c * ld] = a0 - l * p0;
        base[i + (c + 1) * ld] = a1 - l * p1;
        base[i + (c + 2) * ld] = a2 - l * p2;
        base[i + (c + 3) * ld] = a3 - l * p3;
      }
      for (; c < jb; ++c) base[i + c * ld] = base[i + c * ld] - l * base[jl + c * ld];
    }
    __syncthreads();
  }
  if (use_smem) {
    for (int e = tid; e < pcand * jb; e += blockDim.x) {
      const int il = e % pcand, cl = e / pcand;
      G[il + (long long)cl * m] = sm[il + (long long)cl * pcand];
    }
  }
  // the interchanges of this panel composed (front-local indices; k_swap_trsm moves the rest of those rows with it)
  for (int i = tid; i < pcand; i += blockDim.x) wperm[f.col0 + j0 + i] = j0 + s_orig[i];
  for (int o = 16; o > 0; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  if (lane == 0 && lmax > 0.0) atomicMax(&st->max_l_bits, (unsigned long long)__double_as_longlong(lmax));
}

// ------------------------------------------------------------------------------ row swaps + U rows

// After panel [j0, j0+jb): apply its interchanges to every other column of the k pivot rows
// (left part of P, right part of P, all of Q) and solve the unit-lower block system for the columns
// to the right (U12 rows).  grid: (column groups of 64, fronts); block 128 threads.
// The interchanges arrive COMPOSED (k_panel_lu: `wperm[p]` = the row that position p of the <= 128-row pivot window
// holds afterwards), so a column is permuted with independent loads followed by independent stores -- one warp per
// column, lanes along the window rows -- instead of 32 dependent read-modify-write swaps per thread.  The jb x 64
// tile of U rows is staged through shared memory so that global traffic is coalesced.
template <class T>
__global__ void __launch_bounds__(128) k_swap_trsm(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                                   int first, int j0, int ob0, T* __restrict__ fac,
                                                   const int* __restrict__ wperm, int symmetric) {
  const Front f = fronts[lvl_front[first + blockIdx.y]];
  const int k = f.k, r = f.r;
  if (k <= j0) return;
  const int jb = min(NB, k - j0), j1 = j0 + jb;
  const long long m = (long long)k + r;
  // logical column space: [0, j0) left of panel | [j1, k) right of panel in P | [0, r) of Q
  const int n_left = j0, n_right = k - j1, ncols = n_left + n_right + (symmetric ? 0 : r);   // symmetric: no Q
  constexpr int TC = 64;   // columns per CTA: 50 KB of shared memory and ~128 registers -> four CTAs per SM (with 128
                           // columns and eight columns per gather round it was two, and the SM issued 35 % of the time)
  const int cbase = blockIdx.x * TC;
  if (cbase >= ncols) return;
  const int W = min(k, ob0 + OB) - j0;   // rows of the pivot window
  T* P = fac + f.p_off;
  T* Q = fac + f.q_off;
  extern __shared__ unsigned char smem_raw[];
  T* s_L = reinterpret_cast<T*>(smem_raw);  // NB x NB, column-major, unit lower
  T* s_tile = s_L + NB * NB;                // TC columns x (NB+1)
  __shared__ int s_src[OB];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int e = tid; e < NB * NB; e += 128) {
    const int i = e % NB, c = e / NB;
    s_L[e] = (i < jb && c < jb && i > c) ? P[(j0 + i) + (long long)(j0 + c) * m] : scalar_traits<T>::zero();
  }
  s_src[tid] = tid < W ? wperm[f.col0 + j0 + tid] : j0 + tid;
  __syncthreads();
  auto column = [&](int g) -> T* {
    if (g < n_left) return P + (long long)g * m;
    if (g < n_left + n_right) return P + (long long)(j1 + g - n_left) * m;
    return Q + (long long)(g - n_left - n_right) * k;
  };
  // ---- interchanges; the U rows of the right columns go to the tile.  A warp takes CB columns per round: all their
  // loads are in flight together (one column per round is a chain of 32 dependent L2 round trips per warp)
  int src[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) src[q] = s_src[lane + 32 * q];
  constexpr int CB = 4;
  for (int cb = wid * CB; cb < TC; cb += 4 * CB) {
    if (cbase + cb >= ncols) break;
    T v[CB][4];
#pragma unroll
    for (int u = 0; u < CB; ++u) {
      const int g = cbase + cb + u;
      const bool right = g >= n_left;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = lane + 32 * q;
        v[u][q] = scalar_traits<T>::zero();
        if (g < ncols && i < W && (src[q] != j0 + i || (q == 0 && right && lane < jb))) v[u][q] = column(g)[src[q]];
      }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < CB; ++u) {
      const int g = cbase + cb + u;
      if (g >= ncols) continue;
      const bool right = g >= n_left;
      T* col = column(g);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int i = lane + 32 * q;
        if (q == 0 && right && lane < jb) continue;   // written back after the triangular solve
        if (i < W && src[q] != j0 + i) col[j0 + i] = v[u][q];
      }
      if (right) s_tile[(cb + u) * (NB + 1) + lane] = lane < jb ? v[u][0] : scalar_traits<T>::zero();
    }
  }
  // whole block: either all "left" columns (nothing more to do) or some right columns
  const bool block_has_right = cbase + TC - 1 >= n_left;
  if (!block_has_right) return;
  __syncthreads();
  const int lc = cbase + tid;
  if (tid < TC && lc < ncols && lc >= n_left) {
    T* u = s_tile + tid * (NB + 1);
#pragma unroll 1
    for (int t = 1; t < jb; ++t) {
      // four partial sums: the chain of dependent multiply-adds per row is a quarter as long
      T a0 = u[t], a1 = scalar_traits<T>::zero(), a2 = a1, a3 = a1;
      int s = 0;
      for (; s + 4 <= t; s += 4) {
        a0 = a0 - s_L[t + s * NB] * u[s];
        a1 = a1 - s_L[t + (s + 1) * NB] * u[s + 1];
        a2 = a2 - s_L[t + (s + 2) * NB] * u[s + 2];
        a3 = a3 - s_L[t + (s + 3) * NB] * u[s + 3];
      }
      for (; s < t; ++s) a0 = a0 - s_L[t + s * NB] * u[s];
      u[t] = (a0 + a1) + (a2 + a3);
    }
  }
  __syncthreads();
  for (int c = wid; c < TC; c += 4) {
    const int g = cbase + c;
    if (g < ncols && g >= n_left && lane < jb) column(g)[j0 + lane] = s_tile[c * (NB + 1) + lane];
  }
}

// ----------------------------------------------------------------------------------- L21 panel

// Rows below the current outer block (remaining pivot rows and contribution rows) of the panel
// columns: X U_jj = B.  One thread per row; grid (row groups of 128, fronts).
template <class T>
__global__ void __launch_bounds__(128) k_trsm_cols(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                                   int first, int j0, int ob0, T* __restrict__ fac, DevStats* st) {
  const Front f = fronts[lvl_front[first + blockIdx.y]];
  const int k = f.k;
  if (k <= j0) return;
  const int rbase = min(ob0 + OB, k);           // first row below the outer block
  const int r = (int)((long long)k + f.r - rbase);  // rows to process: rest of the pivot rows + CB rows
  if (r <= 0) return;
  if (blockIdx.x * 128 >= r) return;
  const int jb = min(NB, k - j0);
  const long long m = (long long)k + f.r;
  T* P = fac + f.p_off;
  __shared__ T s_U[NB * NB];  // column-major upper triangle, diagonal holds reciprocals
  for (int e = threadIdx.x; e < NB * NB; e += 128) {
    const int i = e % NB, c = e / NB;
    T v = scalar_traits<T>::zero();
    if (i < jb && c < jb && i <= c) {
      v = P[(j0 + i) + (long long)(j0 + c) * m];
      if (i == c) v = recip(v);
    }
    s_U[e] = v;
  }
  __syncthreads();
  const int row = blockIdx.x * 128 + threadIdx.x;
  if (row >= r) return;
  T* x = P + rbase + row + (long long)j0 * m;
  // all 32 loads first, all stores last: with a load and a store per column the row is a chain of 32 dependent
  // memory round trips (the compiler cannot move a load above the previous column's store)
  T xs[NB];
  double lmax = 0.0;
#pragma unroll
  for (int c = 0; c < NB; ++c) xs[c] = c < jb ? x[(long long)c * m] : scalar_traits<T>::zero();
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    if (c < jb) {
      T acc = xs[c];
#pragma unroll
      for (int s = 0; s < c; ++s) acc = acc - xs[s] * s_U[s + c * NB];
      xs[c] = acc * s_U[c + c * NB];
      lmax = fmax(lmax, abs1(xs[c]));
    }
  }
#pragma unroll
  for (int c = 0; c < NB; ++c)
    if (c < jb) x[(long long)c * m] = xs[c];
  atomicMax(&st->max_l_bits, (unsigned long long)__double_as_longlong(lmax));
}

// --------------------------------------------------------------------------------- FP64 DMMA GEMM

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// C(Mr x N) -= A^(Mr x Kr) B(Kr x N), everything in the "real view" of column-major storage.
//   real    : A^ = A                                   (lda_r = lda)
//   complex : rows/ld doubled (interleaved re,im);  A^[2i+a, 2k+b] = a==b ? Re A[i,k] : (a ? +Im : -Im)
// CTA tile 64 x 64, 4 warps as 2 x 2, warp tile 32 x 32 = 4 x 4 m8n8k4 fragments, K chunk 16.
// AMASK / BMASK: the operand is a triangular matrix stored together with the other triangle of something else
// (inverted pivot blocks: L11^-1 strictly below the diagonal, U11^-1 on and above it): 1 = unit lower, 2 = upper,
// applied while loading (element coordinates relative to the operand's origin).
// EPI: 0  C -= acc ;  1  C = -acc ;  2  C = acc.
// BT (real only): the B operand is given TRANSPOSED and row-scaled, B[kk, n] = Bsrc[n + kk * ldb] * dg[kk * dg_stride]
// -- the Schur complement of the symmetric factorisation, C -= L21 (D L21^T), without ever forming D L21^T.
template <bool CPLX, int AMASK = 0, int BMASK = 0, int EPI = 0, int BT = 0>
__device__ __forceinline__ void gemm_tile(const double* __restrict__ A, long long lda_r, const double* __restrict__ B,
                                          long long ldb_r, double* __restrict__ C, long long ldc_r, int Mr, int N,
                                          int Kr, int m0, int n0, const double* __restrict__ dg = nullptr,
                                          long long dg_stride = 0) {
  constexpr int BM = 64, BN = 64, KC = 16;
  constexpr int LDA_S = BM + 4;               // conflict-free fragment reads
  constexpr int LDB_S = KC + 4;
  constexpr int KA = CPLX ? KC / 2 : KC;      // stored A columns per chunk
  constexpr int A_PER_T = BM * KA / 128;      // 8 real, 4 complex
  // one block of shared memory: the operand double buffers during the K loop, then (complex) the 64 x 64 tile of
  // results on its way to a coalesced read-modify-write of C
  constexpr int A_SZ = KA * LDA_S, B_SZ = BN * LDB_S;
  constexpr int LDC_S = BM + 2;
  constexpr int SMEM_D = 2 * (A_SZ + B_SZ) > BN * LDC_S ? 2 * (A_SZ + B_SZ) : BN * LDC_S;
  __shared__ __align__(16) double smem_all[SMEM_D];
  double(*As)[A_SZ] = reinterpret_cast<double(*)[A_SZ]>(smem_all);
  double(*Bs)[B_SZ] = reinterpret_cast<double(*)[B_SZ]>(smem_all + 2 * A_SZ);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int wm = wid & 1, wn = wid >> 1;
  const int g = lane >> 2, tg = lane & 3;

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // global -> register staging
  const int a_row = tid & 63, a_kg = tid >> 6;  // 2 groups of stored columns
  const int b_n = tid >> 1, b_kh = (tid & 1) * 8;
  double ra[A_PER_T], rb[8];
  double rd[BT ? 8 : 1];   // BT: diagonal scale factors, applied when the chunk goes to shared memory (after the MMA loop,
                           // so that the loads stay in flight behind the arithmetic)
  // K range of the tile: a triangular operand (mask) is zero outside it -- lower B: k >= n0, upper B: k < n0 + BN,
  // lower A: k < m0 + BM, upper A: k >= m0 (real-view indices; complex rows / K come in pairs, so the same bounds
  // scaled by the view factor) -- which halves the work of the block-inverse merges
  constexpr int SV = CPLX ? 2 : 1;
  int k_begin = 0, k_end = Kr;
  if (BMASK == 1) k_begin = max(k_begin, n0 * SV);
  if (BMASK == 2) k_end = min(k_end, (n0 + BN) * SV);
  if (AMASK == 1) k_end = min(k_end, m0 + BM);
  if (AMASK == 2) k_begin = max(k_begin, m0);
  const int kc_first = k_begin / KC;
  const int nchunks = max(kc_first, (k_end + KC - 1) / KC);

  auto load_global = [&](int kc) {
    const int k0 = kc * KC;  // real-view K offset
    const int ka0 = CPLX ? k0 / 2 : k0;
    const int ka_lim = CPLX ? Kr / 2 : Kr;
    const bool rok = (m0 + a_row) < Mr;
#pragma unroll
    for (int q = 0; q < A_PER_T; ++q) {
      const int ka = a_kg * A_PER_T + q;
      double v = (rok && (ka0 + ka) < ka_lim) ? __ldg(A + (m0 + a_row) + (long long)(ka0 + ka) * lda_r) : 0.0;
      if (AMASK != 0) {
        const int ci = CPLX ? (m0 + a_row) >> 1 : (m0 + a_row), ck = ka0 + ka;   // element (ci, ck) of the operand
        if (AMASK == 1) v = ci > ck ? v : (ci == ck && (!CPLX || ((m0 + a_row) & 1) == 0) && rok) ? 1.0 : 0.0;
        else v = ci <= ck ? v : 0.0;
      }
      ra[q] = v;
    }
    const bool nok = (n0 + b_n) < N;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int kk = k0 + b_kh + q;
      double v;
      if (BT) {
        v = (nok && kk < Kr) ? __ldg(B + (n0 + b_n) + (long long)kk * ldb_r) : 0.0;
        rd[q] = kk < Kr ? __ldg(dg + (long long)kk * dg_stride) : 0.0;
      } else v = (nok && kk < Kr) ? __ldg(B + kk + (long long)(n0 + b_n) * ldb_r) : 0.0;
      if (BMASK != 0) {
        const int ck = CPLX ? kk >> 1 : kk, cn = n0 + b_n;
        if (BMASK == 1) v = ck > cn ? v : (ck == cn && (!CPLX || (kk & 1) == 0) && nok && kk < Kr) ? 1.0 : 0.0;
        else v = ck <= cn ? v : 0.0;
      }
      rb[q] = v;
    }
  };
  auto store_smem = [&](int buf) {
#pragma unroll
    for (int q = 0; q < A_PER_T; ++q) As[buf][(a_kg * A_PER_T + q) * LDA_S + a_row] = ra[q];
#pragma unroll
    for (int q = 0; q < 8; ++q) Bs[buf][b_n * LDB_S + b_kh + q] = BT ? rb[q] * rd[q] : rb[q];
  };

  if (kc_first < nchunks) {
    load_global(kc_first);
    store_smem(kc_first & 1);
  }
  __syncthreads();
  for (int kc = kc_first; kc < nchunks; ++kc) {
    const int cur = kc & 1;
    if (kc + 1 < nchunks) load_global(kc + 1);
    const double* as = As[cur];
    const double* bs = Bs[cur];
#pragma unroll
    for (int ks = 0; ks < KC / 4; ++ks) {
      double af[4], bf[4];
      const int kk = ks * 4 + tg;  // real-view k inside the chunk
#pragma unroll
      for (int mf = 0; mf < 4; ++mf) {
        const int R = wm * 32 + mf * 8 + g;
        if (CPLX) {
          const int a = R & 1, b = kk & 1;
          const double v = as[(kk >> 1) * LDA_S + (R - a) + (a != b)];
          af[mf] = (a == 0 && b == 1) ? -v : v;
        } else {
          af[mf] = as[kk * LDA_S + R];
        }
      }
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) bf[nf] = bs[(wn * 32 + nf * 8 + g) * LDB_S + kk];
#pragma unroll
      for (int mf = 0; mf < 4; ++mf)
#pragma unroll
        for (int nf = 0; nf < 4; ++nf) dmma884(acc[mf][nf][0], acc[mf][nf][1], af[mf], bf[nf]);
    }
    if (kc + 1 < nchunks) store_smem(cur ^ 1);
    __syncthreads();
  }
  // ---- epilogue
  if (CPLX) {
    // Complex: rows of the real view come in (re, im) pairs, C and ldc are 16-byte aligned.  The m8n8k4 accumulator
    // layout gives every thread two ADJACENT COLUMNS of one row -- scattered 8-byte accesses in column-major C (the
    // rank-32 updates and the small-K Schur complements are bound by exactly this traffic).  The tile goes through
    // shared memory instead and is applied with one 16-byte read-modify-write per thread and instruction, a warp
    // covering 512 contiguous bytes of a column.  (The loop above ended with a block barrier: the buffers are free.)
    double* Cs = smem_all;
#pragma unroll
    for (int mf = 0; mf < 4; ++mf) {
      const int row = wm * 32 + mf * 8 + g;
#pragma unroll
      for (int nf = 0; nf < 4; ++nf) {
        const int col = wn * 32 + nf * 8 + tg * 2;
        Cs[col * LDC_S + row] = acc[mf][nf][0];
        Cs[(col + 1) * LDC_S + row] = acc[mf][nf][1];
      }
    }
    __syncthreads();
    const int rp = (tid & 31) * 2, cq = tid >> 5;
#pragma unroll 4
    for (int i = 0; i < BN / 4; ++i) {
      const int col = cq + 4 * i;
      if (m0 + rp < Mr && n0 + col < N) {   // Mr is even: the pair is inside or outside together
        double2* p = reinterpret_cast<double2*>(C + (m0 + rp) + (long long)(n0 + col) * ldc_r);
        const double2 a = *reinterpret_cast<const double2*>(Cs + col * LDC_S + rp);
        double2 c;
        if (EPI == 0) {
          c = *p;
          c.x -= a.x;
          c.y -= a.y;
        } else if (EPI == 1) {
          c.x = -a.x;
          c.y = -a.y;
        } else {
          c = a;
        }
        *p = c;
      }
    }
    __syncthreads();   // the buffers are reused by the caller's next tile (persistent callers) -- and by nobody else
    return;
  }
#pragma unroll
  for (int mf = 0; mf < 4; ++mf) {
    const int R = m0 + wm * 32 + mf * 8 + g;
    if (R >= Mr) continue;
#pragma unroll
    for (int nf = 0; nf < 4; ++nf) {
      const int c0 = n0 + wn * 32 + nf * 8 + tg * 2;
      if (c0 < N) {
        double* p = C + R + (long long)c0 * ldc_r;
        *p = EPI == 0 ? *p - acc[mf][nf][0] : EPI == 1 ? -acc[mf][nf][0] : acc[mf][nf][0];
      }
      if (c0 + 1 < N) {
        double* p = C + R + (long long)(c0 + 1) * ldc_r;
        *p = EPI == 0 ? *p - acc[mf][nf][1] : EPI == 1 ? -acc[mf][nf][1] : acc[mf][nf][1];
      }
    }
  }
}

// Two-level blocking of the partial LU: panels of NB = 32 columns inside outer blocks of OB = 128.
// mode 0: after panel [j0, j1) of the outer block [ob0, ob1): rank-32 updates of what the next
//         panels of THIS outer block need
//           region 1  P[j1:m,   j1:ob1] -= P[j1:m,   j0:j1] * P[j0:j1, j1:ob1]     (panel columns)
//           region 2  P[j1:ob1, ob1:k]  -= P[j1:ob1, j0:j1] * P[j0:j1, ob1:k]      (U rows, right of the block)
//           region 3  Q[j1:ob1, 0:r]    -= P[j1:ob1, j0:j1] * Q[j0:j1, 0:r]        (U12 rows)
// mode 2: after the whole outer block: the bulk of the flops as rank-128 updates
//           region 1  P[ob1:m, ob1:k]   -= P[ob1:m, ob0:ob1] * P[ob0:ob1, ob1:k]
//           region 2  Q[ob1:k, 0:r]     -= P[ob1:k, ob0:ob1] * Q[ob0:ob1, 0:r]
// mode 1: Schur complement  C[0:r, 0:r] -= P[k:m, 0:k] * Q[0:k, 0:r]
// grid: (tiles, fronts of the level)
template <class T>
__global__ void __launch_bounds__(128) k_front_gemm(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                                    int first, int j0, int ob0, int mode, T* __restrict__ fac,
                                                    T* __restrict__ pool, T* __restrict__ pool_cut, int symmetric) {
  constexpr bool CPLX = scalar_traits<T>::is_complex;
  constexpr int S = CPLX ? 2 : 1;
  const Front f = fronts[lvl_front[first + blockIdx.y]];
  const int k = f.k, r = f.r;
  const long long m = (long long)k + r;
  T* P = fac + f.p_off;
  T* Q = fac + f.q_off;
  int t = blockIdx.x;
  auto tiles = [](long long rows, long long cols) { return (int)(((rows * S + 63) / 64) * ((cols + 63) / 64)); };
  if (mode == 0) {
    if (k <= j0) return;
    const int jb = min(NB, k - j0), j1 = j0 + jb, ob1 = min(ob0 + OB, k);
    const int R1 = (int)(m - j1), C1 = ob1 - j1, R2 = ob1 - j1, C2 = k - ob1, C3 = r;
    const double* A = (const double*)(P + j1 + (long long)j0 * m);
    int nt = tiles(R1, C1);
    if (t < nt) {
      const int tm1 = (R1 * S + 63) / 64;
      gemm_tile<CPLX>(A, m * S, (const double*)(P + j0 + (long long)j1 * m), m * S,
                      (double*)(P + j1 + (long long)j1 * m), m * S, R1 * S, C1, jb * S, (t % tm1) * 64, (t / tm1) * 64);
      return;
    }
    t -= nt;
    const int tm2 = (R2 * S + 63) / 64;
    if (tm2 == 0) return;
    nt = tiles(R2, C2);
    if (t < nt) {
      gemm_tile<CPLX>(A, m * S, (const double*)(P + j0 + (long long)ob1 * m), m * S,
                      (double*)(P + j1 + (long long)ob1 * m), m * S, R2 * S, C2, jb * S, (t % tm2) * 64, (t / tm2) * 64);
      return;
    }
    t -= nt;
    nt = tiles(R2, C3);
    if (t < nt && !symmetric)   // symmetric: no U12 rows to update
      gemm_tile<CPLX>(A, m * S, (const double*)(Q + j0), (long long)k * S, (double*)(Q + j1), (long long)k * S, R2 * S, C3,
                      jb * S, (t % tm2) * 64, (t / tm2) * 64);
  } else if (mode == 2) {
    const int ob1 = ob0 + OB;
    if (k <= ob1) return;
    const int R1 = (int)(m - ob1), C1 = k - ob1, R2 = k - ob1, C2 = r;
    const double* A = (const double*)(P + ob1 + (long long)ob0 * m);
    int nt = tiles(R1, C1);
    if (t < nt) {
      const int tm1 = (R1 * S + 63) / 64;
      gemm_tile<CPLX>(A, m * S, (const double*)(P + ob0 + (long long)ob1 * m), m * S,
                      (double*)(P + ob1 + (long long)ob1 * m), m * S, R1 * S, C1, OB * S, (t % tm1) * 64, (t / tm1) * 64);
      return;
    }
    t -= nt;
    nt = tiles(R2, C2);
    if (t < nt && !symmetric) {
      const int tm2 = (R2 * S + 63) / 64;
      gemm_tile<CPLX>(A, m * S, (const double*)(Q + ob0), (long long)k * S, (double*)(Q + ob1), (long long)k * S, R2 * S, C2,
                      OB * S, (t % tm2) * 64, (t / tm2) * 64);
    }
  } else {
    if (r == 0 || k == 0) return;
    T* C = (f.flags ? pool_cut : pool) + f.c_off;
    const int tm1 = (r * S + 63) / 64;
    if (t < tiles(r, r))
      gemm_tile<CPLX>((const double*)(P + k), m * S, (const double*)Q, (long long)k * S, (double*)C, (long long)r * S,
                      r * S, r, k * S, (t % tm1) * 64, (t / tm1) * 64);
  }
}

// ------------------------------------------------------ merging inverted diagonal blocks (post-factor)
//
// Two adjacent HB x HB diagonal blocks of a pivot block whose triangular inverses are known (A at j0, D at
// jd = j0 + HB; each holds L^-1 strictly below and U^-1 on/above its diagonal) are merged into the inverse of
// the 2 HB block in place:   lower  X = -D_l^-1 C A_l^-1   (C = P[jd.., j0..]),   upper  Y = -A_u^-1 B D_u^-1.
// Each product is two GEMMs through a scratch block T:  phase 1  T = C A_l^-1 / T = B D_u^-1,
// phase 2  X = -D_l^-1 T / Y = -A_u^-1 T.  grid: (tiles, block pairs, fronts of the batch).
// Extends the 32 -> 64 -> 128 merges of solve.cu to whole pivot blocks: a front then needs ONE triangular
// matrix-vector product per sweep instead of a chain of dependent 128-pivot steps.
template <class T, int PHASE, bool UPPER>
__global__ void __launch_bounds__(128) k_inv_merge(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                                   int first, const long long* __restrict__ scr_off, int HB,
                                                   T* __restrict__ fac, T* __restrict__ scratch) {
  constexpr bool CPLX = scalar_traits<T>::is_complex;
  constexpr int S = CPLX ? 2 : 1;
  const Front f = fronts[lvl_front[first + blockIdx.z]];
  const int k = f.k;
  const int j0 = blockIdx.y * 2 * HB, jd = j0 + HB;
  if (jd >= k) return;
  const int hd = min(HB, k - jd);
  const long long m = (long long)k + f.r;
  T* P = fac + f.p_off;
  // scratch slot of pair p: [2 p HB^2, ...): T (hd x HB, ld = hd) then T2 (HB x hd, ld = HB) -- 2 hd HB entries, so a
  // front needs at most k HB <= k^2 / 2 entries at any merge level (post_factor sizes the regions accordingly)
  T* Tm = scratch + scr_off[blockIdx.z] + (long long)blockIdx.y * 2 * HB * HB + (UPPER ? (long long)hd * HB : 0);
  const long long ldt = UPPER ? HB : hd;
  const int Mrows = UPPER ? HB : hd, Ncols = UPPER ? hd : HB;
  const int tm = (Mrows * S + 63) / 64, tn = (Ncols + 63) / 64;
  const int t = blockIdx.x;
  if (t >= tm * tn) return;
  const int m0 = (t % tm) * 64, n0 = (t / tm) * 64;
  const double* Ablk = (const double*)(P + j0 + (long long)j0 * m);
  const double* Dblk = (const double*)(P + jd + (long long)jd * m);
  double* Cblk = (double*)(P + jd + (long long)j0 * m);   // below the diagonal
  double* Bblk = (double*)(P + j0 + (long long)jd * m);   // above the diagonal
  if (PHASE == 1) {
    if (!UPPER) gemm_tile<CPLX, 0, 1, 2>(Cblk, m * S, Ablk, m * S, (double*)Tm, ldt * S, hd * S, HB, HB * S, m0, n0);
    else gemm_tile<CPLX, 0, 2, 2>(Bblk, m * S, Dblk, m * S, (double*)Tm, ldt * S, HB * S, hd, hd * S, m0, n0);
  } else {
    if (!UPPER) gemm_tile<CPLX, 1, 0, 1>(Dblk, m * S, (const double*)Tm, ldt * S, Cblk, m * S, hd * S, HB, hd * S, m0, n0);
    else gemm_tile<CPLX, 2, 0, 1>(Ablk, m * S, (const double*)Tm, ldt * S, Bblk, m * S, HB * S, hd, HB * S, m0, n0);
  }
}

template <class T>
void launch_inv_merge(cudaStream_t st, const Front* fronts, const int* lvl_front, int first, int cnt, const long long* scr_off,
                      int HB, int maxk, T* fac, T* scratch) {
  constexpr int S = scalar_traits<T>::is_complex ? 2 : 1;
  const int pairs = cdiv(maxk, 2 * HB);
  const int tiles = cdiv((long long)HB * S, 64) * cdiv(HB, 64);
  const dim3 grid(tiles, pairs, cnt);
  k_inv_merge<T, 1, false><<<grid, 128, 0, st>>>(fronts, lvl_front, first, scr_off, HB, fac, scratch);
  k_inv_merge<T, 1, true><<<grid, 128, 0, st>>>(fronts, lvl_front, first, scr_off, HB, fac, scratch);
  k_inv_merge<T, 2, false><<<grid, 128, 0, st>>>(fronts, lvl_front, first, scr_off, HB, fac, scratch);
  k_inv_merge<T, 2, true><<<grid, 128, 0, st>>>(fronts, lvl_front, first, scr_off, HB, fac, scratch);
  LSA_LAUNCH_CHECK();
}
template void launch_inv_merge<double>(cudaStream_t, const Front*, const int*, int, int, const long long*, int, int, double*, double*);
template void launch_inv_merge<z128>(cudaStream_t, const Front*, const int*, int, int, const long long*, int, int, z128*, z128*);

// Schur complement of the symmetric factorisation (real FP64): C -= L21 (D L21^T) with the transposed, diagonally
// scaled B operand taken straight from L21 and diag(U11) in the tile loader.  grid: (tiles, fronts of the level).
__global__ void __launch_bounds__(128) k_front_schur_sym(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                                         int first, double* __restrict__ fac, double* __restrict__ pool,
                                                         double* __restrict__ pool_cut) {
  const Front f = fronts[lvl_front[first + blockIdx.y]];
  const int k = f.k, r = f.r;
  if (r == 0 || k == 0) return;
  const long long m = (long long)k + r;
  const double* P = fac + f.p_off;
  double* C = (f.flags ? pool_cut : pool) + f.c_off;
  // lower triangle of tiles only: t -> (tm, tn), tn <= tm
  const int tm1 = (r + 63) / 64;
  const int t = blockIdx.x;
  if (t >= tm1 * (tm1 + 1) / 2) return;
  int tm = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
  while (tm * (tm + 1) / 2 > t) --tm;
  while ((tm + 1) * (tm + 2) / 2 <= t) ++tm;
  const int tn = t - tm * (tm + 1) / 2;
  gemm_tile<false, 0, 0, 0, 1>(P + k, m, P + k, m, C, (long long)r, r, r, k, tm * 64, tn * 64, P, m + 1);
}

// stand-alone GEMM used by lsa_gemm_bench (same tile code as the front updates)
template <bool CPLX>
__global__ void __launch_bounds__(128) k_gemm_plain(const double* A, long long lda_r, const double* B, long long ldb_r,
                                                    double* C, long long ldc_r, int Mr, int N, int Kr) {
  const int tm1 = (Mr + 63) / 64;
  const int tm = blockIdx.x % tm1, tn = blockIdx.x / tm1;
  gemm_tile<CPLX>(A, lda_r, B, ldb_r, C, ldc_r, Mr, N, Kr, tm * 64, tn * 64);
}

void gemm_plain(cudaStream_t st, bool cplx, const void* A, long long lda, const void* B, long long ldb, void* C,
                long long ldc, int M, int N, int K) {
  const int S = cplx ? 2 : 1;
  const int tiles = cdiv((long long)M * S, 64) * cdiv(N, 64);
  if (cplx)
    k_gemm_plain<true><<<tiles, 128, 0, st>>>((const double*)A, lda * 2, (const double*)B, ldb * 2, (double*)C,
                                               ldc * 2, M * 2, N, K * 2);
  else
    k_gemm_plain<false><<<tiles, 128, 0, st>>>((const double*)A, lda, (const double*)B, ldb, (double*)C, ldc, M, N, K);
  LSA_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------- driver

template <class T, class VT>
static void launch_scatter(cudaStream_t st, T* fac, const long long* dst, const void* vals, long long nnz, z128 coef,
                           int accumulate) {
  if (nnz == 0) return;
  const int blocks = (int)std::min<long long>((nnz + 255) / 256, 148LL * 16);
  k_scatter<T, VT><<<blocks, 256, 0, st>>>(fac, dst, (const VT*)vals, nnz, coef, accumulate);
  LSA_LAUNCH_CHECK();
}

template <class T>
void factor_numeric(lsa_handle_impl& h, z128 alpha, z128 beta, double tiny_abs, int* n_kernels) {
  const Symbolic& sym = h.sym;
  cudaStream_t st = h.stream;
  T* fac = (T*)h.d_fac;
  T* pool[2] = {(T*)h.d_pool[0], (T*)h.d_pool[1]};
  T* pool_cut = (T*)h.d_cut_pool;
  const int symm = sym.symmetric ? 1 : 0;   // F = L D L^T: diagonal pivots, no Q blocks (real FP64 only, checked by the caller)
  int launches = 0;
  SweepTrace tr;
  tr.begin(st);
  LSA_CUDA(cudaMemsetAsync(fac, 0, sym.fac_size * sizeof(T), st));
  if (h.partitioned && h.part.cut_pool_size > 0)
    LSA_CUDA(cudaMemsetAsync(pool_cut, 0, h.part.cut_pool_size * sizeof(T), st));
  DevStats init{};
  init.min_piv_bits = (unsigned long long)0x7ff0000000000000ULL;  // +inf
  init.max_piv_bits = 0ULL;
  LSA_CUDA(cudaMemcpyAsync(h.d_stats, &init, sizeof(DevStats), cudaMemcpyHostToDevice, st));

  // ---- assemble F = alpha A + beta M into the front panels (never materialised as a matrix)
  const bool use_a = alpha.x != 0.0 || alpha.y != 0.0;
  const bool use_m = h.has_m && (beta.x != 0.0 || beta.y != 0.0);
  if (use_a) {
    if (h.a_complex) {
      if constexpr (scalar_traits<T>::is_complex) launch_scatter<T, z128>(st, fac, h.d_a_dst, h.d_a_orig, h.nnz_a, alpha, 0);
    } else {
      launch_scatter<T, double>(st, fac, h.d_a_dst, h.d_a_orig, h.nnz_a, alpha, 0);
    }
    launches++;
  }
  if (use_m) {
    if (h.m_complex) {
      if constexpr (scalar_traits<T>::is_complex) launch_scatter<T, z128>(st, fac, h.d_m_dst, h.d_m_orig, h.nnz_m, beta, 1);
    } else {
      launch_scatter<T, double>(st, fac, h.d_m_dst, h.d_m_orig, h.nnz_m, beta, 1);
    }
    launches++;
  }
  if (sym.n_iso > 0) {
    k_decoupled_pivots<T><<<cdiv(sym.n_iso, 256), 256, 0, st>>>(fac + sym.diag_off, sym.n_iso, tiny_abs, h.d_stats);
    LSA_LAUNCH_CHECK();
    launches++;
  }

  const size_t swap_smem = (size_t)(NB * NB + 64 * (NB + 1)) * sizeof(T);
  static PerDeviceOnce attr_set;   // per instantiation (T)
  if (attr_set.first()) {
    LSA_CUDA(cudaFuncSetAttribute(k_swap_trsm<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)swap_smem));
    LSA_CUDA(cudaFuncSetAttribute(k_panel_lu<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM_CAP));
  }
  constexpr int S = scalar_traits<T>::is_complex ? 2 : 1;
  constexpr int YMAX = 32768;

  for (int d = sym.nlevels - 1; d >= 0; --d) {
    const int lbeg = sym.lvl_ptr[d], lend = sym.lvl_ptr[d + 1];
    const int cnt_all = lend - lbeg;
    if (h.partitioned && d == h.part.n_top_levels - 1) {
      // ---- all sub-trees are factored: their roots' contribution blocks go to every GPU (NCCL broadcast from the
      // owner, all roots in one group), then the replicated top is factored redundantly on identical data
      constexpr int SR = scalar_traits<T>::is_complex ? 2 : 1;
      comm_group_begin();
      for (int gs : h.part.cut_roots) {
        const Front& f = sym.fronts[h.part.g2l[gs]];
        comm_bcast(h.comm, (double*)(pool_cut + f.c_off), (size_t)f.r * f.r * SR, h.part.owner[gs], st);
      }
      comm_group_end();
      tr.mark("bcast_cut", d, 0, (int)h.part.cut_roots.size(), 1);
    }
    // contribution blocks of this level start from zero
    long long pool_used = 0;
    for (int q = lbeg; q < lend; ++q) {
      const Front& f = sym.fronts[sym.lvl_front[q]];
      if (f.flags == FRONT_REGULAR) pool_used = std::max(pool_used, f.c_off + (long long)f.r * f.r);
    }
    if (pool_used > 0) LSA_CUDA(cudaMemsetAsync(pool[d & 1], 0, pool_used * sizeof(T), st));
    for (int y0 = 0; y0 < cnt_all; y0 += YMAX) {
      const int cnt = std::min(YMAX, cnt_all - y0);
      const int first = lbeg + y0;
      // level-chunk extents (fronts are sorted by descending k inside a level)
      int maxk = 0, maxchild = 0, max_rc = 0;
      for (int q = first; q < first + cnt; ++q) {
        const Front& f = sym.fronts[sym.lvl_front[q]];
        maxk = std::max(maxk, f.k);
        maxchild = std::max(maxchild, f.nchild);
        for (int c = 0; c < f.nchild; ++c) max_rc = std::max(max_rc, sym.fronts[sym.child_idx[f.child0 + c]].r);
      }
      // ---- extend-add, one child slot per launch
      for (int slot = 0; slot < maxchild; ++slot) {
        const int gx = std::max(1, std::min(64, max_rc / 8));
        k_extend_add<T><<<dim3(gx, cnt), 256, 0, st>>>(h.d_fronts, h.d_lvl_front, first, h.d_child_idx, h.d_ea_map, slot,
                                                       fac, pool[(d + 1) & 1], pool[d & 1], pool_cut, symm);
        LSA_LAUNCH_CHECK();
        tr.mark("extend_add", d, slot, gx, cnt);
        launches++;
      }
      // ---- blocked partial LU of every front of the chunk (fronts with k > j0 form a prefix)
      auto tiles = [](long long rows, long long cols) { return ((rows * S + 63) / 64) * ((cols + 63) / 64); };
      for (int ob0 = 0; ob0 < maxk; ob0 += OB) {
        for (int j0 = ob0; j0 < std::min(ob0 + OB, maxk); j0 += NB) {
          int act = 0, gx_cols = 1, gx_rows = 0;
          long long gx_tiles = 0;
          for (int q = first; q < first + cnt; ++q) {
            const Front& f = sym.fronts[sym.lvl_front[q]];
            if (f.k <= j0) break;
            act++;
            const int jb = std::min(NB, f.k - j0), j1 = j0 + jb, ob1 = std::min(ob0 + OB, f.k);
            const long long m = (long long)f.k + f.r;
            gx_cols = std::max(gx_cols, cdiv((long long)j0 + (f.k - j1) + f.r, 64));
            gx_rows = std::max(gx_rows, cdiv(m - ob1, 128));
            gx_tiles = std::max(gx_tiles, tiles(m - j1, ob1 - j1) + tiles(ob1 - j1, f.k - ob1) + tiles(ob1 - j1, f.r));
          }
          if (act == 0) break;
          {
            // the level's fronts are sorted by descending k: the first one has the tallest panel
            const int max_pk = std::min(sym.fronts[sym.lvl_front[first]].k, ob0 + OB) - j0;
            const long long want = (long long)max_pk * NB * (long long)sizeof(T);
            const int panel_smem = (int)std::min<long long>(want, PANEL_SMEM_CAP);
            // the pivot candidates are the <= 128 rows of the current outer block: 128 threads cover them (one
            // row each in the scaling / rank-1 step), with half the warps to synchronise per column step
            // (32 / 64 threads for the <= 32 / <= 64-row panels of the leaf-side levels measured slower: 33.3 vs 31.5 ms)
            k_panel_lu<T><<<act, 128, panel_smem, st>>>(h.d_fronts, h.d_lvl_front, first, j0, fac, h.d_ipiv, h.d_gperm, tiny_abs,
                                                        h.d_stats, panel_smem, ob0, symm);
          }
          LSA_LAUNCH_CHECK();
          tr.mark("panel_lu", d, j0, act, 1);
          k_swap_trsm<T><<<dim3(gx_cols, act), 128, swap_smem, st>>>(h.d_fronts, h.d_lvl_front, first, j0, ob0, fac, h.d_gperm, symm);
          LSA_LAUNCH_CHECK();
          tr.mark("swap_trsm", d, j0, gx_cols, act);
          launches += 2;
          if (gx_rows > 0) {
            k_trsm_cols<T><<<dim3(gx_rows, act), 128, 0, st>>>(h.d_fronts, h.d_lvl_front, first, j0, ob0, fac, h.d_stats);
            LSA_LAUNCH_CHECK();
            tr.mark("trsm_cols", d, j0, gx_rows, act);
            launches++;
          }
          if (gx_tiles > 0) {
            k_front_gemm<T><<<dim3((unsigned)gx_tiles, act), 128, 0, st>>>(h.d_fronts, h.d_lvl_front, first, j0, ob0, 0, fac,
                                                                             pool[d & 1], pool_cut, symm);
            LSA_LAUNCH_CHECK();
            tr.mark("gemm_inner", d, j0, (int)gx_tiles, act);
            launches++;
          }
        }
        // deferred rank-128 update of everything right of / below the outer block
        int act2 = 0;
        long long gx2 = 0;
        for (int q = first; q < first + cnt; ++q) {
          const Front& f = sym.fronts[sym.lvl_front[q]];
          if (f.k <= ob0 + OB) break;
          act2++;
          const long long m = (long long)f.k + f.r, ob1 = ob0 + OB;
          gx2 = std::max(gx2, tiles(m - ob1, f.k - ob1) + tiles(f.k - ob1, f.r));
        }
        if (act2 > 0 && gx2 > 0) {
          k_front_gemm<T><<<dim3((unsigned)gx2, act2), 128, 0, st>>>(h.d_fronts, h.d_lvl_front, first, 0, ob0, 2, fac, pool[d & 1], pool_cut, symm);
          LSA_LAUNCH_CHECK();
          tr.mark("gemm_outer", d, ob0, (int)gx2, act2);
          launches++;
        }
      }
      // ---- Schur complement of every front of the chunk
      int gx_schur = 0;
      for (int q = first; q < first + cnt; ++q) {
        const Front& f = sym.fronts[sym.lvl_front[q]];
        if (f.r > 0 && f.k > 0) {
          const long long tr_ = cdiv(f.r, 64);
          gx_schur = (int)std::max<long long>(gx_schur, symm ? tr_ * (tr_ + 1) / 2 : (long long)cdiv((long long)f.r * S, 64) * tr_);
        }
      }
      if (gx_schur > 0 && symm) {
        if constexpr (!scalar_traits<T>::is_complex)
          k_front_schur_sym<<<dim3(gx_schur, cnt), 128, 0, st>>>(h.d_fronts, h.d_lvl_front, first, fac, pool[d & 1], pool_cut);
        LSA_LAUNCH_CHECK();
        tr.mark("schur_sym", d, 0, gx_schur, cnt);
        launches++;
      } else if (gx_schur > 0) {
        k_front_gemm<T><<<dim3(gx_schur, cnt), 128, 0, st>>>(h.d_fronts, h.d_lvl_front, first, 0, 0, 1, fac, pool[d & 1], pool_cut, symm);
        LSA_LAUNCH_CHECK();
        tr.mark("gemm_schur", d, 0, gx_schur, cnt);
        launches++;
      }
    }
  }
  tr.end();
  if (n_kernels) *n_kernels = launches;
}

template void factor_numeric<double>(lsa_handle_impl&, z128, z128, double, int*);
template void factor_numeric<z128>(lsa_handle_impl&, z128, z128, double, int*);

}  // namespace lsa
