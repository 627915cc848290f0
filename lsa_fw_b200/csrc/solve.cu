// Supernodal forward / backward triangular solves on the multifrontal factors, incl. the
// conjugate-transposed solve used for adjoint modes.
//
// Stands in for PETSc's MatSolve / MatSolveTranspose behind `KSP preonly + PC lu`, executed once per
// Krylov step by SLEPc's STApply (reference call site Solver/utils.py:268-270; explicit in
// Solver/eigen2.py:174-178).  The adjoint eigensolve of Sensitivity/__init__.py:246-262 re-factors
// (A^H - conj(sigma) M^H); here it is the `trans = H` sweep over the SAME factors.
//
// The work is HBM bound by nature: one pass over L (forward) and one over U (backward), i.e.
// nnz(L+U) * sizeof(T) bytes per solve plus the index vectors.  Launches are batched per assembly tree level;
// within a front the pivot block advances in 128-wide steps whose diagonal blocks were inverted explicitly
// after the factorisation (32 -> 64 -> 128), so a step is one small GEMV instead of a scalar substitution.
// Which kernel sweeps a level depends on its shape (plan_solve / solve_impl at the end of the file):
//   many fronts (>= 96, or all fronts <= 128 pivots), complex   k_front_stream   bulk copies into a smem ring
//   few fronts, pivot block <= invert_max_k                     k_tri_gemv       the WHOLE pivot block is inverted
//                                                                explicitly after the factorisation (merges
//                                                                128 -> 256 -> ... as DMMA GEMMs): one triangular
//                                                                matrix-vector product, no dependent steps
//   larger pivot blocks: <= 9 tall fronts                       k_sweep_slices   16-CTA clusters, 8-row slices
//                        10 .. 95 multi-step fronts             k_sweep_cluster  1-16 CTAs per front, by chunk
//                        taller than cluster_max_rows           k_step           one launch per 128-pivot step
//   rows outside the pivot block                                k_up_off / k_down_off (wide GEMVs)
//   children -> parent contributions                            k_up_gather
//
//   trans = N :  up sweep   y_top = L11^-1 P x_top ;  contrib  -= L21 y_top
//                down sweep y_top = U11^-1 (y_top - U12 y_anc)
//   trans = H :  up sweep   y_top = U11^-H x_top   ;  contrib  -= U12^H y_top
//                down sweep y_top = L11^-H (y_top - L21^H y_anc) ; x = P^T y
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cooperative_groups.h>

#include "factor.cuh"

namespace lsa {

static constexpr int SB = 128;  // pivot-block step of the solve kernels = order of the inverted diagonal blocks
static constexpr int IB = 32;   // inverted diagonal block order (= factorisation panel width)

// ------------------------------------------------------------------------------- post-factor set-up

// Composes the row interchanges of every front into one gather permutation:
// (P x)[col0 + i] = x[gperm[col0 + i]].
__global__ void k_compose_perm(const Front* __restrict__ fronts, int ns, const int* __restrict__ ipiv,
                               int* __restrict__ gperm) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ns) return;
  const Front f = fronts[s];
  int* g = gperm + f.col0;
  for (int j = 0; j < f.k; ++j) g[j] = f.col0 + j;
  for (int j = 0; j < f.k; ++j) {
    const int p = ipiv[f.col0 + j];
    if (p != j) {
      const int a = g[j];
      g[j] = g[p];
      g[p] = a;
    }
  }
}

__global__ void k_iota(int* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

// Inverts the 32 x 32 diagonal blocks of L11 (unit lower) and U11 (upper) in place.
// grid: (diagonal blocks, fronts); block: 32 threads, thread c computes column c of both inverses.
template <class T>
__global__ void __launch_bounds__(32) k_invert_diag(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                                    int first, T* __restrict__ fac) {
  const Front f = fronts[lvl_front[first + blockIdx.y]];
  const int j0 = blockIdx.x * IB;
  if (j0 >= f.k) return;
  const int nb = min(IB, f.k - j0);
  const long long m = (long long)f.k + f.r;
  T* D = fac + f.p_off + j0 + (long long)j0 * m;
  __shared__ T s_a[IB][IB + 1];  // original block
  __shared__ T s_x[IB][IB + 1];  // s_x[c][i]: column c of an inverse
  const int c = threadIdx.x;
  for (int j = 0; j < nb; ++j)
    if (c < nb) s_a[c][j] = D[c + (long long)j * m];  // s_a[row][col]
  __syncwarp();
  if (c < nb) {
    // unit lower inverse, column c
    for (int i = c + 1; i < nb; ++i) {
      T acc = s_a[i][c];
      for (int s = c + 1; s < i; ++s) acc = acc + s_a[i][s] * s_x[c][s];
      s_x[c][i] = scalar_traits<T>::zero() - acc;
    }
  }
  __syncwarp();
  for (int j = 0; j < nb; ++j)
    if (c < nb && c > j) D[c + (long long)j * m] = s_x[j][c];
  __syncwarp();
  // reciprocals of the diagonal once (one per lane) instead of up to 31 dependent divisions per column
  __shared__ T s_rd[IB];
  if (c < nb) s_rd[c] = recip(s_a[c][c]);
  __syncwarp();
  if (c < nb) {
    // upper inverse, column c
    s_x[c][c] = s_rd[c];
    for (int i = c - 1; i >= 0; --i) {
      T acc = scalar_traits<T>::zero();
      for (int s = i + 1; s <= c; ++s) acc = acc + s_a[i][s] * s_x[c][s];
      s_x[c][i] = scalar_traits<T>::zero() - acc * s_rd[i];
    }
  }
  __syncwarp();
  for (int j = 0; j < nb; ++j)
    if (c < nb && c <= j) D[c + (long long)j * m] = s_x[j][c];
}

// Merges the inverses of two adjacent HB x HB diagonal blocks (A above-left, D below-right) into the
// inverse of the 2HB block, in place:   lower  X = -D^-1 C A^-1 ,   upper  X = -A^-1 B D^-1 .
// Launched for HB = 32 and HB = 64, so that 128 x 128 diagonal blocks of L11 (unit lower) and U11
// (upper) end up explicitly inverted: the solve sweeps then need ONE small GEMV per 128 pivots
// instead of a chain of dependent substitutions.   grid: (block pairs, fronts); block 256.
template <class T, int HB>
__global__ void __launch_bounds__(256) k_merge_inv(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                                   int first, T* __restrict__ fac) {
  const Front f = fronts[lvl_front[first + blockIdx.y]];
  const int j0 = blockIdx.x * 2 * HB, jd = j0 + HB;
  if (jd >= f.k) return;
  const int hd = min(HB, f.k - jd);
  const long long m = (long long)f.k + f.r;
  T* P = fac + f.p_off;
  extern __shared__ unsigned char smem_raw[];
  T* Tm = reinterpret_cast<T*>(smem_raw);  // HB x HB scratch, column-major
  const int tid = threadIdx.x;
  // ---- lower part
  for (int e = tid; e < hd * HB; e += 256) {
    const int p = e % hd, c = e / hd;
    T acc = P[(jd + p) + (long long)(j0 + c) * m];  // q == c: unit diagonal of A^-1
    for (int q = c + 1; q < HB; ++q) acc = acc + P[(jd + p) + (long long)(j0 + q) * m] * P[(j0 + q) + (long long)(j0 + c) * m];
    Tm[p + c * HB] = acc;
  }
  __syncthreads();
  for (int e = tid; e < hd * HB; e += 256) {
    const int i = e % hd, c = e / hd;
    T acc = Tm[i + c * HB];
    for (int p = 0; p < i; ++p) acc = acc + P[(jd + i) + (long long)(jd + p) * m] * Tm[p + c * HB];
    P[(jd + i) + (long long)(j0 + c) * m] = scalar_traits<T>::zero() - acc;
  }
  __syncthreads();
  // ---- upper part
  for (int e = tid; e < HB * hd; e += 256) {
    const int p = e % HB, c = e / HB;
    T acc = scalar_traits<T>::zero();
    for (int q = 0; q <= c; ++q) acc = acc + P[(j0 + p) + (long long)(jd + q) * m] * P[(jd + q) + (long long)(jd + c) * m];
    Tm[p + c * HB] = acc;
  }
  __syncthreads();
  for (int e = tid; e < HB * hd; e += 256) {
    const int i = e % HB, c = e / HB;
    T acc = scalar_traits<T>::zero();
    for (int p = i; p < HB; ++p) acc = acc + P[(j0 + i) + (long long)(j0 + p) * m] * Tm[p + c * HB];
    P[(j0 + i) + (long long)(jd + c) * m] = scalar_traits<T>::zero() - acc;
  }
}

template <class T>
void launch_inv_merge(cudaStream_t st, const Front* fronts, const int* lvl_front, int first, int cnt, const long long* scr_off,
                      int HB, int maxk, T* fac, T* scratch);   // factor.cu

// scratch entries the block-inverse merges of a k-pivot front need (see k_inv_merge): max over the merge levels
// HB = 128, 256, ... < k of  2 HB^2 floor((k - HB) / (2 HB)) + 2 HB hd_last
static long long inv_region(long long k) {
  long long need = 0;
  for (long long HB = SB; HB < k; HB *= 2) {
    const long long pairs = (k + 2 * HB - 1) / (2 * HB);
    long long last = -1;
    for (long long p = pairs - 1; p >= 0 && last < 0; --p)
      if (p * 2 * HB + HB < k) last = p;
    if (last < 0) continue;
    const long long hd = std::min(HB, k - (last * 2 * HB + HB));
    need = std::max(need, last * 2 * HB * HB + 2 * hd * HB);
  }
  return (need + 3) & ~3LL;
}

template <class T>
void post_factor(lsa_handle_impl& h, int* n_kernels) {
  const Symbolic& sym = h.sym;
  cudaStream_t st = h.stream;
  SweepTrace tr;
  tr.begin(st);
  if (h.n > 0) k_iota<<<cdiv(h.n, 256), 256, 0, st>>>(h.d_gperm, h.n);
  LSA_LAUNCH_CHECK();
  if (sym.ns > 0) {
    k_compose_perm<<<cdiv(sym.ns, 64), 64, 0, st>>>(h.d_fronts, sym.ns, h.d_ipiv, h.d_gperm);
    LSA_LAUNCH_CHECK();
    // 32 x 32 diagonal blocks inverted, merged pairwise up to 128 x 128: per level chunk of the sweep plan (fronts of a
    // level are sorted by descending k, so the grids are tight; batches by front number mixed the 1 444-pivot root
    // with the leaves and launched 1.5 M CTAs per batch that had nothing to do)
    for (const SolveChunk& c : h.solve_plan) {
      const int s0 = c.first, cnt = c.cnt, maxk = c.maxk;
      if (maxk <= 0) continue;
      k_invert_diag<T><<<dim3(cdiv(maxk, IB), cnt), 32, 0, st>>>(h.d_fronts, h.d_lvl_front, s0, (T*)h.d_fac);
      LSA_LAUNCH_CHECK();
      tr.mark("inv_32", c.level, s0, cdiv(maxk, IB), cnt);
      if (maxk > 32) {
        k_merge_inv<T, 32><<<dim3(cdiv(maxk, 64), cnt), 256, 32 * 32 * sizeof(T), st>>>(h.d_fronts, h.d_lvl_front, s0, (T*)h.d_fac);
        LSA_LAUNCH_CHECK();
        tr.mark("inv_64", c.level, s0, cdiv(maxk, 64), cnt);
      }
      if (maxk > 64) {
        static PerDeviceOnce attr_done;   // per instantiation (T)
        if (attr_done.first())
          LSA_CUDA(cudaFuncSetAttribute(k_merge_inv<T, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(64 * 64 * sizeof(T))));
        k_merge_inv<T, 64><<<dim3(cdiv(maxk, 128), cnt), 256, 64 * 64 * sizeof(T), st>>>(h.d_fronts, h.d_lvl_front, s0, (T*)h.d_fac);
        LSA_LAUNCH_CHECK();
        tr.mark("inv_128", c.level, s0, cdiv(maxk, 128), cnt);
      }
      if (n_kernels) (*n_kernels) += 3;
    }
    // ---- levels swept with k_tri_gemv: merge on, 128 -> 256 -> ... -> whole pivot block (GEMMs on the DMMA
    // pipe through a scratch block per pair; the fronts of a level chunk are batched as far as the scratch
    // buffer reaches; fronts are sorted by descending k, those with k > 128 form a prefix)
    for (const SolveChunk& c : h.solve_plan) {
      if (c.mode != SOLVE_INVERTED || c.maxk <= SB) continue;
      int q = c.first;
      const int qend = c.first + c.cnt;
      while (q < qend && sym.fronts[sym.lvl_front[q]].k > SB) {
        std::vector<long long> off;
        long long used = 0;
        int q1 = q;
        while (q1 < qend && sym.fronts[sym.lvl_front[q1]].k > SB) {
          const long long kk = sym.fronts[sym.lvl_front[q1]].k;
          // scratch region of a front: pair p of merge level HB starts at 2 p HB^2 and uses 2 hd HB entries (hd = HB
          // for all but the last pair), i.e. at most k HB <= k^2 / 2 + k entries at any level (inv_region)
          if (!off.empty() && (used + inv_region(kk) > h.inv_scratch_entries || off.size() >= 32768)) break;
          off.push_back(used);
          used += inv_region(kk);
          ++q1;
        }
        LSA_CUDA(cudaMemcpyAsync(h.d_inv_off, off.data(), off.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
        const int maxk = sym.fronts[sym.lvl_front[q]].k;
        for (int HB = SB; HB < maxk; HB *= 2) {
          launch_inv_merge<T>(st, h.d_fronts, h.d_lvl_front, q, q1 - q, h.d_inv_off, HB, maxk, (T*)h.d_fac, (T*)h.d_inv_scratch);
          if (n_kernels) (*n_kernels) += 4;
          tr.mark("inv_merge", c.level, HB, cdiv(maxk, 2 * HB), q1 - q);
        }
        // the offsets buffer is reused by the next batch: pageable copies are staged before the call returns,
        // and the launches above are stream ordered behind them
        q = q1;
      }
    }
  }
  if (n_kernels) (*n_kernels) += 2;
  tr.end();
}
template void post_factor<double>(lsa_handle_impl&, int*);
template void post_factor<z128>(lsa_handle_impl&, int*);

// --------------------------------------------------------------------------------- decoupled pivots

template <class T, bool H>
__global__ void k_solve_decoupled(const T* __restrict__ diag, int n_iso, const z128* __restrict__ x, z128* __restrict__ y) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_iso) y[i] = recip(cj<H>(diag[i])) * x[i];
}

// ----------------------------------------------------------------------------------------- up sweep

// One CTA per front: collect the children's contribution vectors, then move the pivot rows into
// the work vector y (with the front's row permutation when PERM).  The extend-add map of ONE child is injective,
// so a child's entries are added without any ordering inside the CTA; children follow each other in a fixed
// order (deterministic) with a block barrier in between.  NT = 1024 for the few tall fronts of the tree top
// (3 passes over a 3 000-row child instead of 12), 256 for the wide levels.
template <bool PERM, int NT>
__global__ void __launch_bounds__(NT) k_up_gather(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                                  int first, const int* __restrict__ child_idx,
                                                  const int* __restrict__ ea_map, const int* __restrict__ gperm,
                                                  z128* __restrict__ x, z128* __restrict__ y, z128* __restrict__ cb) {
  const Front p = fronts[lvl_front[first + blockIdx.x]];
  z128* cbp = cb + p.st0;
  for (int t = threadIdx.x; t < p.r; t += NT) cbp[t] = mk(0, 0);
  __syncthreads();
  for (int q = 0; q < p.nchild; ++q) {
    const Front c = fronts[child_idx[p.child0 + q]];
    // partitioned solve: the contribution vector of a sub-tree root reaches the replicated rows through the
    // all-reduce (k_cut_scatter), whichever GPU owns it
    if (c.flags != FRONT_REGULAR) continue;
    const int* map = ea_map + c.st0;
    const z128* cbc = cb + c.st0;
    for (int t = threadIdx.x; t < c.r; t += NT) {
      const int ip = map[t];
      const z128 v = cbc[t];
      if (ip < p.k) x[p.col0 + ip] += v;
      else cbp[ip - p.k] += v;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < p.k; i += NT) y[p.col0 + i] = x[PERM ? gperm[p.col0 + i] : p.col0 + i];
}

// Same job for the wide levels near the leaves (tens of thousands of fronts with a few dozen rows each): one WARP
// per front, eight fronts per CTA, warp barriers only.  With one 256-thread CTA per front those levels ran at
// < 1 TB/s (a CTA's worth of barriers and scheduling for ~100 useful entries).
template <bool PERM>
__global__ void __launch_bounds__(256) k_up_gather_warp(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                                        int first, int cnt, const int* __restrict__ child_idx,
                                                        const int* __restrict__ ea_map, const int* __restrict__ gperm,
                                                        z128* __restrict__ x, z128* __restrict__ y, z128* __restrict__ cb) {
  const int lane = threadIdx.x & 31;
  const int fi = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (fi >= cnt) return;
  const Front p = fronts[lvl_front[first + fi]];
  z128* cbp = cb + p.st0;
  for (int t = lane; t < p.r; t += 32) cbp[t] = mk(0, 0);
  __syncwarp();
  for (int q = 0; q < p.nchild; ++q) {
    const Front c = fronts[child_idx[p.child0 + q]];
    if (c.flags != FRONT_REGULAR) continue;
    const int* map = ea_map + c.st0;
    const z128* cbc = cb + c.st0;
    for (int t = lane; t < c.r; t += 32) {
      const int ip = map[t];
      const z128 v = cbc[t];
      if (ip < p.k) x[p.col0 + ip] += v;
      else cbp[ip - p.k] += v;
    }
    __syncwarp();
  }
  for (int i = lane; i < p.k; i += 32) y[p.col0 + i] = x[PERM ? gperm[p.col0 + i] : p.col0 + i];
}

// --------------------------------------------------------------------------------------- down sweep

// y_top -= Off * anc[ancestor rows].  N: Off = U12 = Q (k x r);  H: Off = L21^H, L21 = P[k:m, 0:k].
// `anc` holds the FINAL values of the ancestors: the work vector itself for N; for H the output
// vector, because each front applies its own P^T as soon as its pivot block is solved.
// grid: (pivot-row chunks of 64, fronts); block 256.
template <class T, bool H, int NW>
__global__ void __launch_bounds__(NW * 32) k_down_off(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                                   int first, const int* __restrict__ st_idx, const T* __restrict__ fac,
                                                   const z128* anc, z128* y) {
  constexpr int ROWS = 32;   // pivot rows per CTA; NW warps = NW column groups (N) / NW rows per pass (H)
  constexpr int CHUNK = NW * 32;
  const Front f = fronts[lvl_front[first + blockIdx.y]];
  const int k = f.k, r = f.r;
  const int r0 = blockIdx.x * ROWS;
  if (r0 >= k || r == 0) return;
  const long long m = (long long)k + r;
  const T* P = fac + f.p_off;
  const T* Q = fac + f.q_off;
  const int* idx = st_idx + f.st0;
  __shared__ z128 xs[CHUNK];
  __shared__ z128 red[NW][ROWS + 1];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  z128 acc = mk(0, 0);
  const int rowN = r0 + lane;  // N: this thread's pivot row
  z128 accH[ROWS / NW];        // H: rows wid, wid + NW, ...
#pragma unroll
  for (int q = 0; q < ROWS / NW; ++q) accH[q] = mk(0, 0);
  for (int c0 = 0; c0 < r; c0 += CHUNK) {
    const int len = min(CHUNK, r - c0);
    __syncthreads();
    if (tid < len) xs[tid] = anc[idx[c0 + tid]];
    __syncthreads();
    if (!H) {
      if (rowN < k) {
        const T* a = Q + rowN + (long long)c0 * k;
#pragma unroll 8
        for (int c = wid; c < len; c += NW) acc += a[(long long)c * k] * xs[c];
      }
    } else {
      // the rows of a warp (wid, wid + NW, ...) are read together: ROWS / NW independent streams per lane
      for (int cb0 = 0; cb0 < len; cb0 += 32) {
        const int c = cb0 + lane;
        if (c < len) {
          const z128 xc = xs[c];
#pragma unroll
          for (int q = 0; q < ROWS / NW; ++q) {
            const int rowH = r0 + wid + q * NW;
            if (rowH < k) accH[q] += conj_(P[k + c0 + c + (long long)rowH * m]) * xc;
          }
        }
      }
    }
  }
  if (!H) {
    red[wid][lane] = acc;
    __syncthreads();
    if (tid < ROWS && r0 + tid < k) {
      z128 s = red[0][tid];
#pragma unroll 8
      for (int q = 1; q < NW; ++q) s += red[q][tid];
      y[f.col0 + r0 + tid] -= s;
    }
  } else {
#pragma unroll
    for (int q = 0; q < ROWS / NW; ++q) {
      z128 a = accH[q];
      for (int o = 16; o > 0; o >>= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
        a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
      }
      const int rowH = r0 + wid + q * NW;
      if (lane == 0 && rowH < k) y[f.col0 + rowH] -= a;
    }
  }
}

// L2 prefetch of a contiguous run of factor entries, one `prefetch.global.L2` per 128-byte line (plain LSU
// instructions: the bulk form `cp.async.bulk.prefetch.L2` runs on the uniform datapath, one lane at a time, and
// cost ~0.1 us per segment when measured here).  The factor data never depends on the sweep's dependency chain,
// so the panel of step s+1 is pulled into L2 while step s is being solved.
template <class T>
__device__ __forceinline__ void prefetch_l2(const T* p, int count) {
  const unsigned long long a1 = (unsigned long long)(p + count);
  for (unsigned long long a = (unsigned long long)p & ~127ull; a < a1; a += 128)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a) : "memory");
}

// cb -= Off * z_front  once all pivots of the front are solved (up sweep of the tall fronts: the rows of the
// contribution vector do not take part in the dependent chain of 128-pivot steps, so they are updated by ONE
// wide GEMV afterwards instead of step by step).  N: Off = L21 = P[k:m, 0:k];  H: Off = U12^H, U12 = Q (k x r).
// grid: (chunks of 32 contribution rows, fronts); block NW * 32.
template <class T, bool H, int NW>
__global__ void __launch_bounds__(NW * 32) k_up_off(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                                 int first, const T* __restrict__ fac, const z128* __restrict__ zsol,
                                                 z128* __restrict__ cb) {
  constexpr int ROWS = 32;
  constexpr int CHUNK = NW * 32;
  const Front f = fronts[lvl_front[first + blockIdx.y]];
  const int k = f.k, r = f.r;
  const int r0 = blockIdx.x * ROWS;
  if (r0 >= r) return;
  const long long m = (long long)k + r;
  const T* P = fac + f.p_off;
  const T* Q = fac + f.q_off;
  __shared__ z128 xs[CHUNK];
  __shared__ z128 red[NW][ROWS + 1];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  z128 acc = mk(0, 0);
  const int rowN = r0 + lane;
  z128 accH[ROWS / NW];
#pragma unroll
  for (int q = 0; q < ROWS / NW; ++q) accH[q] = mk(0, 0);
  for (int c0 = 0; c0 < k; c0 += CHUNK) {
    const int len = min(CHUNK, k - c0);
    __syncthreads();
    if (tid < len) xs[tid] = zsol[f.col0 + c0 + tid];
    __syncthreads();
    if (!H) {
      if (rowN < r) {
        const T* a = P + k + rowN + (long long)c0 * m;
#pragma unroll 8
        for (int c = wid; c < len; c += NW) acc += a[(long long)c * m] * xs[c];
      }
    } else {
      for (int cb0 = 0; cb0 < len; cb0 += 32) {
        const int c = cb0 + lane;
        if (c < len) {
          const z128 xc = xs[c];
#pragma unroll
          for (int q = 0; q < ROWS / NW; ++q) {
            const int j = r0 + wid + q * NW;
            if (j < r) accH[q] += conj_(Q[c0 + c + (long long)j * k]) * xc;
          }
        }
      }
    }
  }
  if (!H) {
    red[wid][lane] = acc;
    __syncthreads();
    if (tid < ROWS && r0 + tid < r) {
      z128 sacc = red[0][tid];
#pragma unroll 8
      for (int q = 1; q < NW; ++q) sacc += red[q][tid];
      cb[f.st0 + r0 + tid] -= sacc;
    }
  } else {
#pragma unroll
    for (int q = 0; q < ROWS / NW; ++q) {
      z128 a = accH[q];
      for (int o = 16; o > 0; o >>= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
        a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
      }
      const int j = r0 + wid + q * NW;
      if (lane == 0 && j < r) cb[f.st0 + j] -= a;
    }
  }
}

__device__ __forceinline__ z128 ld_cg(const z128* p) {  // L2 read (data written by a peer CTA of the cluster)
  const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
  return mk(v.x, v.y);
}

// --------------------------------------------------------------- whole pivot block inverted: one GEMV
//
// Fronts whose pivot block D = P[0:k, 0:k] was inverted as a whole after the factorisation (post_factor: L11^-1
// strictly below the diagonal with an implied unit diagonal, U11^-1 on and above it) need no dependent 128-pivot
// steps: the triangular solve of the front is ONE matrix-vector product, every row independent.
//   up,N    z_i = y_i + sum_{c<i}  D[i,c] y_c            up,H    z_i = sum_{c<=i} conj(D[c,i]) y_c
//   down,N  z_i =       sum_{c>=i} D[i,c] y_c            down,H  z_i = y_i + sum_{c>i} conj(D[c,i]) y_c
// N reduces along rows (thread = row, warps = column groups, partial sums through shared memory), H along columns
// (warp = 32 / NW output entries, lanes along the contiguous column).  grid: (chunks of 32 pivots, fronts).
// Levels with very few fronts (the root: k / 32 = 46 CTAs on 148 SMs) split every row chunk's input range over
// gridDim.z CTAs: each writes its partial sums to a scratch slot, the LAST one to arrive (atomic ticket) adds the
// partials in split order -- deterministic -- and finishes the outputs; the ticket resets itself.
template <class T, bool H, bool UP, int NW>
__global__ void __launch_bounds__(NW * 32) k_tri_gemv(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                                   int first, const T* __restrict__ fac, const z128* __restrict__ in,
                                                   z128* __restrict__ out, z128* __restrict__ scratch,
                                                   int* __restrict__ tickets, int span) {
  constexpr int ROWS = 32;
  constexpr int CHUNK = NW * 32;
  const Front f = fronts[lvl_front[first + blockIdx.y]];
  const int k = f.k;
  const int r0 = blockIdx.x * ROWS;
  if (r0 >= k) return;
  const long long m = (long long)k + f.r;
  const T* P = fac + f.p_off;
  __shared__ z128 xs[CHUNK];
  __shared__ z128 red[NW][ROWS + 1];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int rowN = r0 + lane;
  z128 acc = mk(0, 0);
  z128 accH[ROWS / NW];
#pragma unroll
  for (int q = 0; q < ROWS / NW; ++q) accH[q] = mk(0, 0);
  // index range of the input entries this CTA's outputs depend on, and this split's share of it: spans of `span`
  // entries, so that only the long rows of the triangle are split (a row chunk near the apex stays one CTA) and no
  // CTA streams more than 32 x span entries
  int c_lo = UP ? 0 : r0, c_hi = UP ? min(k, r0 + ROWS) : k;
  const int nsplit = gridDim.z > 1 ? min((int)gridDim.z, (c_hi - c_lo + span - 1) / span) : 1;
  if ((int)blockIdx.z >= nsplit) return;
  if (nsplit > 1) {
    c_lo = c_lo + (int)blockIdx.z * span;
    if ((int)blockIdx.z + 1 < nsplit) c_hi = min(c_hi, c_lo + span);
  }
  for (int c0 = c_lo; c0 < c_hi; c0 += CHUNK) {
    const int len = min(CHUNK, c_hi - c0);
    __syncthreads();
    if (tid < len) xs[tid] = in[f.col0 + c0 + tid];
    __syncthreads();
    if (!H) {
      if (rowN < k) {
        const T* a = P + rowN + (long long)c0 * m;
#pragma unroll 8
        for (int c = wid; c < len; c += NW) {
          const int cc = c0 + c;
          if (UP ? cc < rowN : cc >= rowN) acc += a[(long long)c * m] * xs[c];
        }
      }
    } else {
      for (int cb0 = 0; cb0 < len; cb0 += 32) {
        const int c = cb0 + lane;
        if (c < len) {
          const int cc = c0 + c;
          const z128 xc = xs[c];
#pragma unroll
          for (int q = 0; q < ROWS / NW; ++q) {
            const int i = r0 + wid + q * NW;
            if (i < k && (UP ? cc <= i : cc > i)) accH[q] += conj_(P[cc + (long long)i * m]) * xc;
          }
        }
      }
    }
  }
  // ---- the CTA's 32 partial results -> red[0][0..31]
  if (!H) {
    red[wid][lane] = acc;
    __syncthreads();
    if (tid < ROWS) {
      z128 sum = red[0][tid];
#pragma unroll 8
      for (int q = 1; q < NW; ++q) sum += red[q][tid];
      red[0][tid] = sum;
    }
  } else {
#pragma unroll
    for (int q = 0; q < ROWS / NW; ++q) {
      z128 a = accH[q];
      for (int o = 16; o > 0; o >>= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
        a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
      }
      if (lane == 0) red[0][wid + q * NW] = a;
    }
  }
  __syncthreads();
  z128 sum = tid < ROWS ? red[0][tid] : mk(0, 0);
  if (nsplit > 1) {
    const long long slot = ((long long)blockIdx.y * gridDim.x + blockIdx.x);
    const int zmax = gridDim.z;
    z128* mine = scratch + (slot * zmax + blockIdx.z) * ROWS;
    if (tid < ROWS) mine[tid] = sum;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const int t = atomicAdd(tickets + slot, 1);
      s_last = t == nsplit - 1;
      if (s_last) tickets[slot] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid < ROWS) {
      sum = mk(0, 0);
      for (int z = 0; z < nsplit; ++z) sum += ld_cg(scratch + (slot * zmax + z) * ROWS + tid);
    }
  }
  if (tid < ROWS && r0 + tid < k) {
    // unit diagonals: L11^-1 (up, N) and L11^-H (down, H)
    if (UP != H) sum += in[f.col0 + r0 + tid];
    out[f.col0 + r0 + tid] = sum;
  }
}

// ---------------------------------------------------------------------------- fused sweep steps
//
// One launch per 128-pivot step: every CTA first applies the explicitly inverted diagonal block to
// the step's slice of the input vector (a 128 x 128 triangular GEMV, redundantly per CTA: the block
// is L2 resident after the first read) and then updates its own 128 rows with the result.  CTA 0 also
// stores the solved slice to `out`.  Input and output vectors are distinct (in: pre-solve values,
// out: solved values), so no CTA can read a slice that another one has already overwritten.
//
//   up   (UP = true ):  z = Op in[j0:j1) ;  rows below:  in[row] / cb[row-k]  -=  Off[row, :] z
//   down (UP = false):  z = Op in[j0:j1) ;  rows above:  in[row]              -=  Off[row, :] z
//        Op / Off:   up,N: L^-1 / L      up,H: (U^-1)^H / U^H     down,N: U^-1 / U     down,H: (L^-1)^H / L^H
template <class T, bool H, bool UP, int NT>
__global__ void __launch_bounds__(NT) k_step(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                               int first, int j0, const T* __restrict__ fac, z128* in, z128* out,
                                               z128* __restrict__ cb) {
  const Front f = fronts[lvl_front[first + blockIdx.y]];
  const int k = f.k;
  if (k <= j0) return;
  const int len = min(SB, k - j0), j1 = j0 + len;
  const long long m = (long long)k + f.r;
  const int nrows = UP ? (int)(m - j1) : j0;          // rows to update
  const int r0 = blockIdx.x * SB;
  if (blockIdx.x > 0 && r0 >= nrows) return;
  const T* P = fac + f.p_off;
  const T* Q = fac + f.q_off;
  const T* D = P + j0 + (long long)j0 * m;            // diagonal block origin
  __shared__ z128 ys[SB];
  __shared__ z128 zs[SB];
  constexpr int CG = NT / SB;      // column groups of the row-contiguous (N) phases
  constexpr int NWARP = NT / 32;   // warps: one row each in the column-contiguous (H) phases
  __shared__ z128 part[CG][SB];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid < len) ys[tid] = in[f.col0 + j0 + tid];
  // panel of the next step (the next launch) -> L2, same chunk position
  const int nj0 = UP ? j0 + SB : j0 - SB;
  if (NT > SB && nj0 >= 0 && nj0 < k && tid >= NT - SB) {
    const int i = tid - (NT - SB);
    const int plen = min(SB, k - nj0), pj1 = nj0 + plen;
    const int pn = UP ? (int)(m - pj1) : nj0;
    if (blockIdx.x == 0 && i < plen) prefetch_l2(P + nj0 + (long long)(nj0 + i) * m, plen);
    if (r0 < pn) {
      const int rows = min(SB, pn - r0), row0 = (UP ? pj1 : 0) + r0;
      if (!H) {
        if (i < plen) prefetch_l2(P + row0 + (long long)(nj0 + i) * m, rows);
      } else if (i < rows) {
        const int row = row0 + i;
        prefetch_l2((!UP || row < k) ? P + nj0 + (long long)row * m : Q + nj0 + (long long)(row - k) * k, plen);
      }
    }
  }
  __syncthreads();
  // ---- phase 1: z = Op ys
  if (!H) {
    const int i = tid & (SB - 1), cg = tid >> 7;
    z128 acc = mk(0, 0);
    if (i < len) {
      const T* row = D + i;
      if (UP) {
#pragma unroll 8
        for (int c = cg; c < i; c += CG) acc += row[(long long)c * m] * ys[c];
      } else {
#pragma unroll 8
        for (int c = i + cg; c < len; c += CG) acc += row[(long long)c * m] * ys[c];
      }
    }
    part[cg][i] = acc;
    __syncthreads();
    if (tid < len) {
      z128 sum = part[0][tid];
#pragma unroll
      for (int q = 1; q < CG; ++q) sum += part[q][tid];
      zs[tid] = UP ? sum + ys[tid] : sum;  // unit diagonal of L^-1
    }
  } else {
    for (int i = wid; i < len; i += NWARP) {
      const T* col = D + (long long)i * m;
      z128 acc = mk(0, 0);
      if (UP) {
        for (int c = lane; c <= i; c += 32) acc += conj_(col[c]) * ys[c];
      } else {
        for (int c = i + 1 + lane; c < len; c += 32) acc += conj_(col[c]) * ys[c];
      }
      for (int o = 16; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      }
      if (lane == 0) zs[i] = UP ? acc : acc + ys[i];  // unit diagonal of L^-H
    }
  }
  __syncthreads();
  if (blockIdx.x == 0 && tid < len) out[f.col0 + j0 + tid] = zs[tid];
  if (r0 >= nrows) return;
  // ---- phase 2: update this CTA's rows with z
  if (!H) {
    const int rr = tid & (SB - 1), cg = tid >> 7;
    const int row = (UP ? j1 : 0) + r0 + rr;  // front-local row
    z128 acc = mk(0, 0);
    if (r0 + rr < nrows) {
      const T* a = P + row + (long long)j0 * m;
#pragma unroll 8
      for (int c = cg; c < len; c += CG) acc += a[(long long)c * m] * zs[c];
    }
    __syncthreads();  // part[] is being reused
    part[cg][rr] = acc;
    __syncthreads();
    if (tid < SB && r0 + tid < nrows) {
      z128 sum = part[0][tid];
#pragma unroll
      for (int q = 1; q < CG; ++q) sum += part[q][tid];
      const int rw = (UP ? j1 : 0) + r0 + tid;
      if (!UP || rw < k) in[f.col0 + rw] -= sum;
      else cb[f.st0 + (rw - k)] -= sum;
    }
  } else {
#pragma unroll
    for (int q = 0; q < SB / NWARP; ++q) {
      const int rr = wid + q * NWARP;
      if (r0 + rr >= nrows) break;
      const int row = (UP ? j1 : 0) + r0 + rr;
      // column `row` of U (up) / of L (down), entries of the step's rows: contiguous in memory
      const T* u = (!UP || row < k) ? P + j0 + (long long)row * m : Q + j0 + (long long)(row - k) * k;
      z128 acc = mk(0, 0);
      for (int c = lane; c < len; c += 32) acc += conj_(u[c]) * zs[c];
      for (int o = 16; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      }
      if (lane == 0) {
        if (!UP || row < k) in[f.col0 + row] -= acc;
        else cb[f.st0 + (row - k)] -= acc;
      }
    }
  }
}


// ---------------------------------------------------------------------- cluster sweep (big fronts)
//
// Levels whose fronts need several dependent 128-pivot steps are swept with ONE launch per level and
// direction: every front is owned by a thread-block cluster of C CTAs (C = 1, 2, 4, 8 by front height),
// the steps run inside the kernel and are separated by the hardware cluster barrier (release/acquire at
// cluster scope) instead of kernel boundaries.  Per step every CTA redundantly applies the inverted
// diagonal block (L2 resident) and updates its share of the rows; rank 0 stores the solved slice.
template <class T, bool H, bool UP, int C>
__global__ void __launch_bounds__(1024) k_sweep_cluster(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                                        int first, const T* __restrict__ fac, z128* in, z128* out,
                                                        z128* cb, int pivots_only) {
  namespace cg = cooperative_groups;
  constexpr int NT = 1024, CG = NT / SB, NWARP = NT / 32;
  const int rank = C > 1 ? (int)cg::this_cluster().block_rank() : 0;
  const Front f = fronts[lvl_front[first + blockIdx.x / C]];
  const int k = f.k;
  const long long m = (long long)k + f.r;
  const T* P = fac + f.p_off;
  const T* Q = fac + f.q_off;
  __shared__ z128 ys[SB];
  __shared__ z128 zs[SB];
  __shared__ z128 part[CG][SB];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nsteps = (k + SB - 1) / SB;
  const long long mtop = (UP && pivots_only) ? (long long)k : m;   // rows of the step-by-step updates (up)
  // panel of step sp -> L2: the diagonal block (one rank) and this rank's row chunks
  auto prefetch_step = [&](int sp) {
    if (sp >= nsteps) return;
    const int pj0 = (UP ? sp : nsteps - 1 - sp) * SB;
    const int plen = min(SB, k - pj0), pj1 = pj0 + plen;
    const int pn = UP ? (int)(mtop - pj1) : pj0;
    if (rank == sp % C && tid < plen) prefetch_l2(P + pj0 + (long long)(pj0 + tid) * m, plen);
    const int q = tid >> 7, i = tid & (SB - 1);   // 8 chunks of 128 segments per pass
    for (int r0 = (rank + q * C) * SB; r0 < pn; r0 += 8 * C * SB) {
      const int rows = min(SB, pn - r0);
      const int row0 = (UP ? pj1 : 0) + r0;
      if (!H) {
        if (i < plen) prefetch_l2(P + row0 + (long long)(pj0 + i) * m, rows);
      } else if (i < rows) {
        const int row = row0 + i;
        prefetch_l2((!UP || row < k) ? P + pj0 + (long long)row * m : Q + pj0 + (long long)(row - k) * k, plen);
      }
    }
  };
  prefetch_step(0);
  for (int s = 0; s < nsteps; ++s) {
    prefetch_step(s + 1);
    const int j0 = (UP ? s : nsteps - 1 - s) * SB;
    const int len = min(SB, k - j0), j1 = j0 + len;
    const int nrows = UP ? (int)(mtop - j1) : j0;
    const T* D = P + j0 + (long long)j0 * m;
    if (tid < len) {
      // values written by peer CTAs in the previous step: read through L2
      ys[tid] = ld_cg(in + f.col0 + j0 + tid);
    }
    __syncthreads();
    // ---- phase 1: z = Op ys
    if (!H) {
      const int i = tid & (SB - 1), cgi = tid >> 7;
      z128 acc = mk(0, 0);
      if (i < len) {
        const T* row = D + i;
        if (UP) {
#pragma unroll 8
          for (int c = cgi; c < i; c += CG) acc += row[(long long)c * m] * ys[c];
        } else {
#pragma unroll 8
          for (int c = i + cgi; c < len; c += CG) acc += row[(long long)c * m] * ys[c];
        }
      }
      part[cgi][i] = acc;
      __syncthreads();
      if (tid < len) {
        z128 sum = part[0][tid];
#pragma unroll
        for (int q = 1; q < CG; ++q) sum += part[q][tid];
        zs[tid] = UP ? sum + ys[tid] : sum;
      }
    } else {
      // one warp per 4 columns of the block (i = wid + 32 q): the four column reads are issued together
      constexpr int NQ = SB / NWARP;
      z128 acc[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) acc[q] = mk(0, 0);
#pragma unroll
      for (int cb0 = 0; cb0 < SB; cb0 += 32) {
        const int c = cb0 + lane;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int i = wid + q * NWARP;
          const bool use = i < len && c < len && (UP ? c <= i : c > i);
          if (use) acc[q] += conj_(D[(long long)i * m + c]) * ys[c];
        }
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        for (int o = 16; o > 0; o >>= 1) {
          acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o);
          acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
        }
        const int i = wid + q * NWARP;
        if (lane == 0 && i < len) zs[i] = UP ? acc[q] : acc[q] + ys[i];
      }
    }
    __syncthreads();
    if (rank == 0 && tid < len) out[f.col0 + j0 + tid] = zs[tid];
    // ---- phase 2: this CTA's row chunks
    for (int r0 = rank * SB; r0 < nrows; r0 += C * SB) {
      if (!H) {
        const int rr = tid & (SB - 1), cgi = tid >> 7;
        const int row = (UP ? j1 : 0) + r0 + rr;
        z128 acc = mk(0, 0);
        if (r0 + rr < nrows) {
          const T* a = P + row + (long long)j0 * m;
#pragma unroll 8
          for (int c = cgi; c < len; c += CG) acc += a[(long long)c * m] * zs[c];
        }
        __syncthreads();
        part[cgi][rr] = acc;
        __syncthreads();
        if (tid < SB && r0 + tid < nrows) {
          z128 sum = part[0][tid];
#pragma unroll
          for (int q = 1; q < CG; ++q) sum += part[q][tid];
          const int rw = (UP ? j1 : 0) + r0 + tid;
          z128* dst = (!UP || rw < k) ? in + f.col0 + rw : cb + f.st0 + (rw - k);
          *dst = ld_cg(dst) - sum;
        }
      } else {
        // one warp per 4 rows of the chunk: all reads of the four rows in flight together, then the four
        // read-modify-writes from four lanes at once
        constexpr int NQ = SB / NWARP;
        const T* u[NQ];
        z128 acc[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int rr = wid + q * NWARP;
          const int row = (UP ? j1 : 0) + r0 + rr;
          const bool ok = r0 + rr < nrows;
          u[q] = !ok ? nullptr : (!UP || row < k) ? P + j0 + (long long)row * m : Q + j0 + (long long)(row - k) * k;
          acc[q] = mk(0, 0);
        }
#pragma unroll
        for (int cb0 = 0; cb0 < SB; cb0 += 32) {
          const int c = cb0 + lane;
          if (c < len) {
            const z128 zc = zs[c];
#pragma unroll
            for (int q = 0; q < NQ; ++q)
              if (u[q]) acc[q] += conj_(u[q][c]) * zc;
          }
        }
        z128 mine = mk(0, 0);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          for (int o = 16; o > 0; o >>= 1) {
            acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o);
            acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
          }
          if (lane == q) mine = acc[q];
        }
        if (lane < NQ) {
          const int rr = wid + lane * NWARP;
          if (r0 + rr < nrows) {
            const int row = (UP ? j1 : 0) + r0 + rr;
            z128* dst = (!UP || row < k) ? in + f.col0 + row : cb + f.st0 + (row - k);
            *dst = ld_cg(dst) - mine;
          }
        }
      }
    }
    // ---- all updates of this step visible to the whole cluster before the next slice is read
    if (C > 1) cg::this_cluster().sync();
    else __syncthreads();
  }
}

template <class T, bool H, bool UP, int C>
static void launch_sweep_cluster(cudaStream_t st, int cnt, const Front* fronts, const int* lvl_front, int first, const T* fac,
                                 z128* in, z128* out, z128* cb, int pivots_only) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(C * cnt), 1, 1);
  cfg.blockDim = dim3(1024, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (C > 8) {
    static PerDeviceOnce allowed;   // per instantiation
    if (allowed.first())
      LSA_CUDA(cudaFuncSetAttribute(k_sweep_cluster<T, H, UP, C>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  }
  LSA_CUDA(cudaLaunchKernelEx(&cfg, k_sweep_cluster<T, H, UP, C>, fronts, lvl_front, first, fac, in, out, cb, pivots_only));
}

// ------------------------------------------------- cluster sweep with row slices (few, tall fronts: the tree top)
//
// The top levels hold 1 ... 9 fronts; a 128-pivot step there is a chain of dependent 128 x 128 block products.
// On one SM each of them is bound by that SM's L2 port (256 KB ~ 2 us).  Here the 16 CTAs of a cluster share
// EVERY 128-row block of a front: CTA r owns the entries with (index mod 128) / 8 == r of the pivot and the
// contribution vector, in all blocks -- a static, perfectly balanced ownership in which the critical block
// of a step (the next pivot block) is 8 rows per CTA.  Per step:
//   B  every CTA applies its 8 rows (N) / columns (H) of the inverted diagonal block to the gathered pivot
//      values y_s, and pushes its 8 entries of z_s into the shared memory of all 16 CTAs (DSMEM) -> barrier;
//   A  every CTA updates its entries of the next pivot block with z_s and pushes them to all CTAs (the next
//      y) -> barrier arrive; then its entries of all other blocks, eight blocks per pass -> barrier wait.
// N: block products with one coalesced load per thread (8 rows x 4 columns per warp), partial sums over the
// 32 warps through shared memory.  H: one warp per target entry, contiguous 2 KB reads, shuffles only.
template <class T, bool H, bool UP>
__global__ void __launch_bounds__(1024) k_sweep_slices(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                                       int first, const T* __restrict__ fac, z128* in, z128* out,
                                                       z128* cb) {
  namespace cg = cooperative_groups;
  constexpr int C = 16, SL = SB / C, NB = 8, NWARP = 32;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const Front f = fronts[lvl_front[first + blockIdx.x / C]];
  const int k = f.k;
  const long long m = (long long)k + f.r;
  const T* P = fac + f.p_off;
  const T* Q = fac + f.q_off;
  __shared__ z128 ybuf[2][SB];
  __shared__ z128 zbuf[2][SB];
  __shared__ z128 part[NWARP][NB * SL];
  __shared__ z128 loc[SL];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int nsteps = (k + SB - 1) / SB;
  const int row8 = lane & (SL - 1), csub = lane >> 3;   // N layout inside a warp: 8 rows x 4 columns

  // entry v of the front's vector: pivot entries live in `in`, the others in the contribution vector
  auto slot = [&](int v) -> z128* { return (!UP || v < k) ? in + f.col0 + v : cb + f.st0 + (v - k); };
  // push loc[0:SL) (this CTA's entries of a 128-block) into buf[rank * SL + q] of every CTA of the cluster
  auto broadcast = [&](z128* buf, int len) {
    if (tid < SB) {
      const int dstc = tid >> 3, q = tid & (SL - 1);
      if (rank * SL + q < len) cluster.map_shared_rank(buf, dstc)[rank * SL + q] = loc[q];
    }
  };

  // ---- B: entries rank*SL .. of  z = Op(D_s) y
  auto solve_slice = [&](int s, const z128* ys, z128* zdst) {
    const int j0 = s * SB, len = min(SB, k - j0);
    const T* D = P + j0 + (long long)j0 * m;
    if (!H) {
      const int row = rank * SL + row8, col = wid * 4 + csub;
      z128 p = mk(0, 0);
      if (row < len && col < len && (UP ? col < row : col >= row)) p = D[row + (long long)col * m] * ys[col];
      for (int o = 8; o < 32; o <<= 1) {
        p.x += __shfl_xor_sync(0xffffffffu, p.x, o);
        p.y += __shfl_xor_sync(0xffffffffu, p.y, o);
      }
      if (lane < SL) part[wid][lane] = p;
      __syncthreads();
      if (tid < SL) {
        z128 sum = part[0][tid];
#pragma unroll
        for (int w = 1; w < NWARP; ++w) sum += part[w][tid];
        const int i = rank * SL + tid;
        if (i < len) {
          const z128 z = UP ? sum + ys[i] : sum;
          loc[tid] = z;
          out[f.col0 + j0 + i] = z;
        }
      }
    } else if (wid < SL) {
      const int i = rank * SL + wid;
      z128 acc = mk(0, 0);
      if (i < len) {
        const T* col = D + (long long)i * m;
#pragma unroll
        for (int c0 = 0; c0 < SB; c0 += 32) {
          const int c = c0 + lane;
          if (c < len && (UP ? c <= i : c > i)) acc += conj_(col[c]) * ys[c];
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      }
      if (lane == 0 && i < len) {
        const z128 z = UP ? acc : acc + ys[i];
        loc[wid] = z;
        out[f.col0 + j0 + i] = z;
      }
    }
    __syncthreads();
    broadcast(zdst, len);
  };

  // ---- A (critical): this CTA's entries [vlo, vhi) of block g  -=  Off(., step columns) z, all warps on the one
  //      block (one load per thread); the new values are pushed to every CTA as the next pivot values.
  auto update_critical = [&](int j0, int len, int g, int vlo, int vhi, const z128* zs, z128* ynext, int ylen) {
    if (!H) {
      const int col = wid * 4 + csub;
      const int v = g * SB + rank * SL + row8;
      const bool ok = v >= vlo && v < vhi;
      z128 old = mk(0, 0);
      if (tid < SL) {   // value to be updated: read together with the block entries
        const int vv = g * SB + rank * SL + tid;
        if (vv >= vlo && vv < vhi) old = *slot(vv);
      }
      z128 p = mk(0, 0);
      if (col < len && ok) p = P[v + (long long)(j0 + col) * m] * zs[col];
      for (int o = 8; o < 32; o <<= 1) {
        p.x += __shfl_xor_sync(0xffffffffu, p.x, o);
        p.y += __shfl_xor_sync(0xffffffffu, p.y, o);
      }
      if (lane < SL) part[wid][lane] = p;
      __syncthreads();
      if (tid < SL) {
        z128 sum = part[0][tid];
#pragma unroll
        for (int w = 1; w < NWARP; ++w) sum += part[w][tid];
        const int vv = g * SB + rank * SL + tid;
        if (vv >= vlo && vv < vhi) {
          const z128 nv = old - sum;
          *slot(vv) = nv;
          loc[tid] = nv;
        }
      }
    } else if (wid < SL) {
      const int v = g * SB + rank * SL + wid;
      const bool ok = v >= vlo && v < vhi;
      z128 acc = mk(0, 0), old = mk(0, 0);
      if (ok) {
        if (lane == 0) old = *slot(v);
        const T* bcol = (!UP || v < k) ? P + j0 + (long long)v * m : Q + j0 + (long long)(v - k) * k;
#pragma unroll
        for (int c0 = 0; c0 < SB; c0 += 32) {
          const int c = c0 + lane;
          if (c < len) acc += conj_(bcol[c]) * zs[c];
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
      }
      if (lane == 0 && ok) {
        const z128 nv = old - acc;
        *slot(v) = nv;
        loc[wid] = nv;
      }
    }
    __syncthreads();
    broadcast(ynext, ylen);
  };

  // ---- A (the rest): this CTA's entries of the blocks [ga, gb) restricted to [vlo, vhi), block `skip` left out.
  //      N: eight blocks per pass with all warps (a warp per block was measured slower: 32 dependent-ish loads
  //      per lane); H: one warp per entry, two entries in flight.
  auto update_rest = [&](int j0, int len, int ga, int gb, int skip, int vlo, int vhi, const z128* zs) {
    if (!H) {
      // eight blocks per pass, all warps on them: one load per thread and block (8 rows x 4 columns per warp),
      // partial sums over the 32 warps through shared memory
      const int col = wid * 4 + csub;
      for (int g0 = ga; g0 < gb; g0 += NB) {
        const int cnt = min(NB, gb - g0);
        z128 old = mk(0, 0);
        bool mine = false;
        if (tid < cnt * SL) {   // the values to be updated: read together with the block entries
          const int g = g0 + (tid >> 3), v = g * SB + rank * SL + (tid & (SL - 1));
          mine = g != skip && v >= vlo && v < vhi;
          if (mine) old = *slot(v);
        }
        z128 acc[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) acc[u] = mk(0, 0);
        if (col < len) {
          const z128 zc = zs[col];
          const T* a = P + (long long)(j0 + col) * m;
#pragma unroll
          for (int u = 0; u < NB; ++u) {
            const int v = (g0 + u) * SB + rank * SL + row8;
            if (u < cnt && g0 + u != skip && v >= vlo && v < vhi) acc[u] += a[v] * zc;
          }
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
          for (int o = 8; o < 32; o <<= 1) {
            acc[u].x += __shfl_xor_sync(0xffffffffu, acc[u].x, o);
            acc[u].y += __shfl_xor_sync(0xffffffffu, acc[u].y, o);
          }
          if (lane < SL && u < cnt) part[wid][u * SL + lane] = acc[u];
        }
        __syncthreads();
        if (mine) {
          z128 sum = part[0][tid];
#pragma unroll 8
          for (int w = 1; w < NWARP; ++w) sum += part[w][tid];
          const int v = (g0 + (tid >> 3)) * SB + rank * SL + (tid & (SL - 1));
          *slot(v) = old - sum;
        }
        __syncthreads();
      }
    } else {
      const int nt = (gb - ga) * SL;
      for (int t0 = 0; t0 < nt; t0 += 2 * NWARP) {
        int v[2];
        const T* b[2];
        z128 acc[2], old[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int t = t0 + wid + q * NWARP;
          const int g = ga + (t >> 3);
          v[q] = g * SB + rank * SL + (t & (SL - 1));
          const bool ok = t < nt && g != skip && v[q] >= vlo && v[q] < vhi;
          b[q] = !ok ? nullptr : (!UP || v[q] < k) ? P + j0 + (long long)v[q] * m : Q + j0 + (long long)(v[q] - k) * k;
          acc[q] = mk(0, 0);
          old[q] = (ok && lane == 0) ? *slot(v[q]) : mk(0, 0);
        }
#pragma unroll
        for (int c0 = 0; c0 < SB; c0 += 32) {
          const int c = c0 + lane;
          if (c < len) {
            const z128 zc = zs[c];
#pragma unroll
            for (int q = 0; q < 2; ++q)
              if (b[q]) acc[q] += conj_(b[q][c]) * zc;
          }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          for (int o = 16; o > 0; o >>= 1) {
            acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o);
            acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
          }
          if (lane == 0 && b[q]) *slot(v[q]) = old[q] - acc[q];
        }
      }
    }
    __syncthreads();
  };

  auto step_of = [&](int i) { return UP ? i : nsteps - 1 - i; };
  // everything this CTA reads at step index i -> L2 (one step ahead: the factor does not depend on the vectors)
  auto prefetch_step = [&](int i) {
    if (i >= nsteps) return;
    const int s = step_of(i);
    const int j0 = s * SB, len = min(SB, k - j0), j1 = j0 + len;
    const T* D = P + j0 + (long long)j0 * m;
    const int gl = UP ? j1 / SB : 0, gh = UP ? nsteps : s;
    const int vlo = UP ? j1 : 0, vhi = UP ? k : j0;
    if (!H) {
      if (tid < len) prefetch_l2(D + rank * SL + (long long)tid * m, SL);
      for (int e = tid; e < (gh - gl) * SB; e += 1024) {
        const int g = gl + (e >> 7), col = e & (SB - 1);
        const int v0 = max(vlo, g * SB + rank * SL), v1 = min(vhi, g * SB + rank * SL + SL);
        if (col < len && v1 > v0) prefetch_l2(P + v0 + (long long)(j0 + col) * m, v1 - v0);
      }
    } else {
      constexpr int LINES = SB * (int)sizeof(T) / 128 + 1;   // 128-byte lines of one contiguous run (+1: not aligned)
      for (int e = tid; e < SL * LINES; e += 1024) {
        const int i2 = rank * SL + e / LINES, l = e % LINES;
        if (i2 < len) {
          const char* b = (const char*)(D + (long long)i2 * m);
          const char* a = (const char*)((unsigned long long)b & ~127ull) + 128 * l;
          if (a < b + len * sizeof(T)) asm volatile("prefetch.global.L2 [%0];" ::"l"(a) : "memory");
        }
      }
      for (int e = tid; e < (gh - gl) * SL * LINES; e += 1024) {
        const int t = e / LINES, l = e % LINES;
        const int v = (gl + (t >> 3)) * SB + rank * SL + (t & (SL - 1));
        if (v >= vlo && v < vhi) {
          const char* b = (const char*)((!UP || v < k) ? P + j0 + (long long)v * m : Q + j0 + (long long)(v - k) * k);
          const char* a = (const char*)((unsigned long long)b & ~127ull) + 128 * l;
          if (a < b + len * sizeof(T)) asm volatile("prefetch.global.L2 [%0];" ::"l"(a) : "memory");
        }
      }
    }
  };
  // pivot values of the first block: nobody has touched them in this launch
  {
    const int s = step_of(0), j0 = s * SB, len = min(SB, k - j0);
    if (tid < len) ybuf[0][tid] = in[f.col0 + j0 + tid];
    prefetch_step(0);
    __syncthreads();
  }
  for (int i = 0; i < nsteps; ++i) {
    const int s = step_of(i);
    const int j0 = s * SB, len = min(SB, k - j0), j1 = j0 + len;
    z128* zs = zbuf[i & 1];
    prefetch_step(i + 1);
    solve_slice(s, ybuf[i & 1], zs);
    cluster.sync();                                   // z_s complete in every CTA
    const bool more = i + 1 < nsteps;
    // up: pivot entries only -- the contribution rows are not on the dependent chain, k_up_off updates them
    // with one wide GEMV once z is complete
    const int gl = UP ? j1 / SB : 0, gh = UP ? nsteps : s;       // blocks with entries to update: [gl, gh)
    const int vlo = UP ? j1 : 0, vhi = UP ? k : j0;
    const int gc = UP ? s + 1 : s - 1;                            // next pivot block
    int skip = -1;
    if (more) {
      const int jn = gc * SB, lenn = min(SB, k - jn);
      // entries of block gc that are pivots of the next step
      update_critical(j0, len, gc, max(vlo, jn), min(vhi, jn + lenn), zs, ybuf[(i + 1) & 1], lenn);
      cluster.barrier_arrive();
      skip = gc;
    }
    update_rest(j0, len, gl, gh, skip, vlo, vhi, zs);
    if (more) cluster.barrier_wait();                 // y_{next} complete in every CTA
  }
}

template <class T, bool H, bool UP>
static void launch_sweep_slices(cudaStream_t st, int cnt, const Front* fronts, const int* lvl_front, int first,
                                const T* fac, z128* in, z128* out, z128* cb) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(16 * cnt), 1, 1);
  cfg.blockDim = dim3(1024, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 16;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  static PerDeviceOnce allowed;   // per instantiation
  if (allowed.first())
    LSA_CUDA(cudaFuncSetAttribute(k_sweep_slices<T, H, UP>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  LSA_CUDA(cudaLaunchKernelEx(&cfg, k_sweep_slices<T, H, UP>, fronts, lvl_front, first, fac, in, out, cb));
}

// Cluster width of a level: as many SMs per front as the level leaves free (one 1024-thread CTA per SM),
// but no more than the front has 128-row chunks to hand out.
static int cluster_width(int cnt, int max_rows, int num_sms, int max_width) {
  int c = 1;
  while (c * 2 <= max_width && cnt * c * 2 <= num_sms) c *= 2;
  const int chunks = std::max(1, cdiv(max_rows, SB));
  while (c > 1 && c / 2 >= chunks) c /= 2;
  return c;
}

template <class T, bool H, bool UP>
static void sweep_cluster(cudaStream_t st, int csize, int cnt, const Front* fronts, const int* lvl_front, int first,
                          const T* fac, z128* in, z128* out, z128* cb, int pivots_only = 0) {
  switch (csize) {
    case 1: launch_sweep_cluster<T, H, UP, 1>(st, cnt, fronts, lvl_front, first, fac, in, out, cb, pivots_only); break;
    case 2: launch_sweep_cluster<T, H, UP, 2>(st, cnt, fronts, lvl_front, first, fac, in, out, cb, pivots_only); break;
    case 4: launch_sweep_cluster<T, H, UP, 4>(st, cnt, fronts, lvl_front, first, fac, in, out, cb, pivots_only); break;
    case 8: launch_sweep_cluster<T, H, UP, 8>(st, cnt, fronts, lvl_front, first, fac, in, out, cb, pivots_only); break;
    default: launch_sweep_cluster<T, H, UP, 16>(st, cnt, fronts, lvl_front, first, fac, in, out, cb, pivots_only); break;
  }
}

// ------------------------------------------------------------ streamed sweep (bulk copies + mbarrier ring)
//
// Levels that hold many fronts are swept with ONE persistent launch per level and direction in which the
// factor entries are STREAMED through shared memory: every CTA owns a ring of stages, a producer warp walks
// the tile list of the CTA's fronts (round robin over the level) and fills the ring with `cp.async.bulk`
// copies completing on mbarriers, running ahead of the consumers across operation and front boundaries
// (the factor never depends on the right-hand side); four consumer warps apply each tile out of shared
// memory.  With plain loads every front was a chain of dependent memory round trips and the SMs idled.
//
// A front is a short list of block operations on its column-major storage (D_s = explicitly inverted
// 128 x 128 diagonal block s: unit lower L^-1 below, U^-1 on/above the diagonal; j0 = 128 s):
//
//   up,N   per step s:  z_s = y_s + strict_lower(D_s) y_s   ;  rows below:  y / cb -= P[j1:m, j0:j1] z_s
//   up,H   per step s:  z_s = upper(D_s)^H y_s              ;  y -= P[j0:j1, j1:k]^H z_s ; cb -= Q[j0:j1, :]^H z_s
//   down,N y -= Q anc ; per step s (last first):  x_s = upper(D_s) y_s        ;  y -= P[0:j0, j0:j1] x_s
//   down,H y -= P[k:m, :]^H anc ; per step s:     x_s = y_s + strict_lower(D_s)^H y_s ;  y -= P[j0:j1, 0:j0]^H x_s
//
// N operations reduce along rows of the block (thread = row), H operations along columns (thread = column,
// conjugated entries).  Tiles hold up to 128 rows; narrower blocks get proportionally more columns per tile
// (8 ... 64) so that a tile stays ~16 KB, and a block that is contiguous in memory (Q with few pivots) is
// fetched with a single bulk copy per tile.  Complex factors only (16-byte entries keep every copy aligned).
namespace stream {
constexpr int TR = 128;             // max tile rows
constexpr int MAX_STAGES = 12;
constexpr int TILE = 1088;          // entries per stage: 64 columns x (16 + 1) rows is the largest layout
constexpr int STREAM_MAX_SMEM = 200 * 1024;
enum Mask { NONE = 0, STRICT_LOWER = 1, UPPER = 2 };
enum Emit { Z_UNIT = 0, Z_PLAIN = 1, SUB_SPLIT = 2, SUB_Y = 3 };
struct Op {
  const z128* src;   // entry (0, 0) of the block
  long long ld;
  int R, C;          // rows, columns
  int mask;          // triangular part that is used (block-local coordinates)
  int vsel, voff;    // vector the block is applied to: 0 = ys, 1 = zs, 2 = vbuf; offset
  int emit, out0;    // what happens to the results; offset of result 0 in the front's pivot numbering
};
__device__ __forceinline__ unsigned sa(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
struct Geo {
  int rp_log2, tca_log2, tca, ldt, nrb, ncc, contig;
};
// flags: bit 0 = narrow blocks get wider tiles, bit 1 = single-copy tiles for contiguous blocks,
// bit 2 = the producer copies with 16-byte cp.async (LDGSTS) instead of bulk copies
__device__ __forceinline__ Geo geometry(const Op& b, bool hmode, int flags) {
  Geo g;
  const int rb = min(b.R, TR);
  g.rp_log2 = !(flags & 1) ? 7 : rb <= 16 ? 4 : rb <= 32 ? 5 : rb <= 64 ? 6 : 7;
  g.tca_log2 = 10 - g.rp_log2;
  g.tca = 1 << g.tca_log2;
  g.nrb = (b.R + TR - 1) / TR;
  g.ncc = (b.C + g.tca - 1) >> g.tca_log2;
  // whole columns back to back in memory: one copy per tile (row stride = R; column-wise reads tolerate a 2-way
  // bank conflict, R = 2 mod 4, but not the 4/8-way ones of R = 0 mod 4)
  g.contig = (flags & 2) && b.mask == NONE && b.ld == (long long)b.R && g.nrb == 1 && (!hmode || (b.R & 3));
  g.ldt = g.contig ? b.R : (1 << g.rp_log2) + 1;
  return g;
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sa(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sa(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(sa(b)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(z128* dst, const z128* src, unsigned bytes, unsigned long long* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa(dst)),
               "l"(src), "r"(bytes), "r"(sa(b))
               : "memory");
}
template <int NCONS>
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory"); }
// rows [lo, hi) of block column c that a tile covering rows [r0, r1) has to hold
__device__ __forceinline__ void seg(int mask, int c, int r0, int r1, int& lo, int& hi) {
  lo = r0;
  hi = r1;
  if (mask == STRICT_LOWER) lo = max(r0, c + 1);
  if (mask == UPPER) hi = min(r1, c + 1);
  if (hi < lo) hi = lo;
}
__device__ __forceinline__ bool keep(int mask, int row, int c) {
  return mask == NONE || (mask == STRICT_LOWER ? row > c : row <= c);
}
// acc += a * b  /  acc += conj(a) * b  as four FMAs (the operator form compiles to 4 multiplies + 2 adds)
__device__ __forceinline__ void cmac(z128& acc, const z128 a, const z128 b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ void cmac_conj(z128& acc, const z128 a, const z128 b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.x = fma(a.y, b.y, acc.x);
  acc.y = fma(-a.y, b.x, acc.y);
}
template <bool H, bool UP>
__device__ __forceinline__ int op_count(const Front& f) {
  const int S = (f.k + SB - 1) / SB;
  return UP ? (H ? 3 * S : 2 * S) : 1 + 2 * S;
}
// operation i of a front (same enumeration in the producer and the consumers)
template <bool H, bool UP>
__device__ __forceinline__ Op get_op(const Front& f, const z128* fac, int i) {
  const int k = f.k, r = f.r;
  const long long m = (long long)k + r;
  const z128* P = fac + f.p_off;
  const z128* Q = fac + f.q_off;
  const int S = (k + SB - 1) / SB;
  Op o;
  if (UP) {
    const int per = H ? 3 : 2;
    const int s = i / per, kind = i % per;
    const int j0 = s * SB, len = min(SB, k - j0), j1 = j0 + len;
    if (kind == 0) {
      o = Op{P + j0 + (long long)j0 * m, m, len, len, H ? UPPER : STRICT_LOWER, 0, j0, H ? Z_PLAIN : Z_UNIT, j0};
    } else if (!H) {
      o = Op{P + j1 + (long long)j0 * m, m, (int)(m - j1), len, NONE, 1, j0, SUB_SPLIT, j1};
    } else if (kind == 1) {
      o = Op{P + j0 + (long long)j1 * m, m, len, k - j1, NONE, 1, j0, SUB_SPLIT, j1};
    } else {
      o = Op{Q + j0, (long long)k, len, r, NONE, 1, j0, SUB_SPLIT, k};
    }
  } else {
    if (i == 0) {
      if (!H) o = Op{Q, (long long)k, k, r, NONE, 2, 0, SUB_Y, 0};
      else o = Op{P + k, m, r, k, NONE, 2, 0, SUB_Y, 0};
    } else {
      const int s = S - 1 - (i - 1) / 2, kind = (i - 1) % 2;
      const int j0 = s * SB, len = min(SB, k - j0);
      if (kind == 0) {
        o = Op{P + j0 + (long long)j0 * m, m, len, len, H ? STRICT_LOWER : UPPER, 0, j0, H ? Z_UNIT : Z_PLAIN, j0};
      } else if (!H) {
        o = Op{P + (long long)j0 * m, m, j0, len, NONE, 1, j0, SUB_Y, 0};
      } else {
        o = Op{P + j0, m, len, j0, NONE, 1, j0, SUB_Y, 0};
      }
    }
  }
  return o;
}
}  // namespace stream

template <bool H, bool UP, int NCONS, int NPROD>
__global__ void __launch_bounds__(NCONS + 32 * NPROD) k_front_stream(const Front* __restrict__ fronts,
                                                                   const int* __restrict__ lvl_front, int first, int cnt,
                                                                   const int* __restrict__ st_idx,
                                                                   const z128* __restrict__ fac, const z128* vin,
                                                                   z128* vout, z128* cb, const z128* anc, int kmax, int rmax,
                                                                   int nstages, int flags) {
  using namespace stream;
  constexpr int ITER = 1024 / NCONS;   // entries of a tile per consumer thread
  extern __shared__ __align__(16) unsigned char stream_smem[];
  // dynamic shared memory: ring | ys x 3 | zs | vb x 2 | idx x 3 (down sweep)
  z128* stages = reinterpret_cast<z128*>(stream_smem);
  z128* ysb = stages + (size_t)nstages * TILE;   // pivot-row values of fronts n, n+1, n+2            [3][kmax]
  z128* zs = ysb + 3 * kmax;                     // solved values of the current front                [kmax]
  z128* vbb = zs + kmax;                         // ancestor values (down) / old cb values (up)        [2][rmax]
  int* idxb = reinterpret_cast<int*>(vbb + 2 * rmax);   // ancestor row indices (down)                 [3][rmax]
  __shared__ z128 part[2][NCONS];
  __shared__ __align__(16) Front s_hdr[5];
  __shared__ int s_fidx[5];
  __shared__ __align__(8) unsigned long long full[MAX_STAGES], empty[MAX_STAGES];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full[s], (flags & 4) ? 32 : 1);
      mbar_init(&empty[s], NCONS / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nmine = blockIdx.x < cnt ? (cnt - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // fronts of this CTA
  if (wid >= NCONS / 32) {
    // ------------------------------------------------------------------ producer warps
    const int pw = wid - NCONS / 32;   // tiles with (tile number mod NPROD) == pw are this warp's
    int turn = 0;
    unsigned s = 0, use = 0;   // ring position
    Front fn = nmine > 0 ? fronts[lvl_front[first + blockIdx.x]] : Front{};
    for (int n = 0; n < nmine; ++n) {
      const Front f = fn;
      if (n + 1 < nmine) fn = fronts[lvl_front[first + blockIdx.x + (n + 1) * gridDim.x]];   // in flight during this front
      const int nops = op_count<H, UP>(f);
      for (int oi = 0; oi < nops; ++oi) {
        const Op b = get_op<H, UP>(f, fac, oi);
        const Geo g = geometry(b, H, flags);
        const int nt = g.nrb * g.ncc;
        int rb = 0, cc = 0;
        for (int t = 0; t < nt; ++t) {
          const int r0 = rb * TR, r1 = min(b.R, r0 + TR);
          const int c0 = cc * g.tca, nc = min(g.tca, b.C - c0);
          z128* dst = stages + (size_t)s * TILE;
          const bool mine = turn == pw;
          if (++turn == NPROD) turn = 0;
          if (!mine) {
            if (++s == (unsigned)nstages) { s = 0; ++use; }
            if (H) { if (++rb == g.nrb) { rb = 0; ++cc; } }
            else { if (++cc == g.ncc) { cc = 0; ++rb; } }
            continue;
          }
          if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
          if (flags & 4) {
            // LDGSTS path: the warp copies the tile in 16-byte pieces (512 B per instruction, coalesced along
            // the rows of a column) and every lane posts an arrive-on-completion to the stage's barrier
            if (g.contig) {
              const int total = nc * b.R;
              const z128* sp = b.src + (long long)c0 * b.ld;
              for (int e = lane; e < total; e += 32) cp_async16_cg(dst + e, sp + e);
            } else {
              for (int j = 0; j < nc; ++j) {
                int lo, hi;
                seg(b.mask, c0 + j, r0, r1, lo, hi);
                const z128* sp = b.src + (long long)(c0 + j) * b.ld;
                z128* dp = dst + j * g.ldt - r0;
                for (int row = lo + lane; row < hi; row += 32) cp_async16_cg(dp + row, sp + row);
              }
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(sa(&full[s])) : "memory");
          } else if (g.contig) {
            if (lane == 0) {
              const unsigned bytes = (unsigned)(nc * b.R) * (unsigned)sizeof(z128);
              mbar_expect_tx(&full[s], bytes);
              bulk_g2s(dst, b.src + (long long)c0 * b.ld, bytes, &full[s]);
            }
          } else if (b.mask == NONE) {
            // full-height columns: every copy has the same size
            const unsigned bytes = (unsigned)(r1 - r0) * (unsigned)sizeof(z128);
            if (lane == 0) mbar_expect_tx(&full[s], bytes * (unsigned)nc);
            __syncwarp();
            for (int j = lane; j < nc; j += 32)
              bulk_g2s(dst + j * g.ldt, b.src + r0 + (long long)(c0 + j) * b.ld, bytes, &full[s]);
          } else {
            int lo[2] = {0, 0}, hi[2] = {0, 0};
            unsigned total = 0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int j = lane + 32 * q;
              if (j < nc) seg(b.mask, c0 + j, r0, r1, lo[q], hi[q]);
              total += (unsigned)(hi[q] - lo[q]) * (unsigned)sizeof(z128);
            }
            for (int o2 = 16; o2 > 0; o2 >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o2);
            if (lane == 0) mbar_expect_tx(&full[s], total);
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int j = lane + 32 * q;
              if (hi[q] > lo[q])
                bulk_g2s(dst + j * g.ldt + (lo[q] - r0), b.src + lo[q] + (long long)(c0 + j) * b.ld,
                         (unsigned)(hi[q] - lo[q]) * (unsigned)sizeof(z128), &full[s]);
            }
          }
          if (++s == (unsigned)nstages) { s = 0; ++use; }
          if (H) { if (++rb == g.nrb) { rb = 0; ++cc; } }
          else { if (++cc == g.ncc) { cc = 0; ++rb; } }
        }
      }
    }
    return;
  }
  // -------------------------------------------------------------------- consumer warps
  // Header pipeline (all cp.async, landed by the wait + barrier at the top of each front): front index 4 fronts
  // ahead, Front record 3 ahead, pivot-row values and ancestor indices 2 ahead, ancestor / old cb values 1 ahead.
  auto fi_of = [&](int n) { return first + blockIdx.x + n * (int)gridDim.x; };
  auto stage_idx = [&](int n) {   // S4
    if (n < nmine && tid == 0) cp_async4(&s_fidx[n % 5], lvl_front + fi_of(n));
  };
  auto stage_hdr = [&](int n) {   // S3
    if (n < nmine && tid < 4) cp_async16(reinterpret_cast<char*>(&s_hdr[n % 5]) + 16 * tid,
                                         reinterpret_cast<const char*>(fronts + s_fidx[n % 5]) + 16 * tid);
  };
  auto stage_vec = [&](int n) {   // S2
    if (n >= nmine) return;
    const Front& hf = s_hdr[n % 5];
    z128* yd = ysb + (n % 3) * kmax;
    for (int i = tid; i < hf.k; i += NCONS) cp_async16(yd + i, vin + hf.col0 + i);
    if (!UP) {
      int* id = idxb + (n % 3) * rmax;
      for (int j = tid; j < hf.r; j += NCONS) cp_async4(id + j, st_idx + hf.st0 + j);
    }
  };
  auto stage_anc = [&](int n) {   // S1
    if (n >= nmine) return;
    const Front& hf = s_hdr[n % 5];
    z128* vd = vbb + (n & 1) * rmax;
    if (UP) {
      for (int j = tid; j < hf.r; j += NCONS) cp_async16(vd + j, cb + hf.st0 + j);
    } else {
      const int* id = idxb + (n % 3) * rmax;
      for (int j = tid; j < hf.r; j += NCONS) cp_async16(vd + j, anc + id[j]);
    }
  };
  auto land = [&]() {
    asm volatile("cp.async.wait_all;" ::: "memory");
    cons_sync<NCONS>();
  };
  // warm-up: bring fronts 0 .. 3 to their pipeline positions
  stage_idx(0); stage_idx(1); stage_idx(2); stage_idx(3);
  land();
  stage_hdr(0); stage_hdr(1); stage_hdr(2);
  land();
  stage_vec(0); stage_vec(1);
  land();
  stage_anc(0);
  unsigned grp = 0;          // output groups that went through part[]
  unsigned s = 0, use = 0;   // ring position
  for (int n = 0; n < nmine; ++n) {
    land();   // front n complete in shared memory; the stages issued one front ago have landed as well
    stage_idx(n + 4);
    stage_hdr(n + 3);
    stage_vec(n + 2);
    stage_anc(n + 1);
    const Front f = s_hdr[n % 5];
    const int k = f.k;
    z128* ys = ysb + (n % 3) * kmax;
    z128* vbuf = vbb + (n & 1) * rmax;
    const int nops = op_count<H, UP>(f);
    for (int oi = 0; oi < nops; ++oi) {
      const Op b = get_op<H, UP>(f, fac, oi);
      const Geo g = geometry(b, H, flags);
      const z128* v = (b.vsel == 0 ? ys : b.vsel == 1 ? zs : vbuf) + b.voff;
      const int nt = g.nrb * g.ncc;
      // N: thread = (row, column group), ITER columns of the tile each; H: thread = (column, row group), ITER rows each
      const int sh = H ? g.tca_log2 : g.rp_log2;
      const int a_idx = tid & ((1 << sh) - 1);   // column (H) / row (N) inside the tile
      const int a_grp = tid >> sh;
      const int ngrp = NCONS >> sh;
      const int nout = 1 << sh;                  // results per output group
      z128 acc = mk(0, 0), acc2 = mk(0, 0);
      int rb = 0, cc = 0;
      for (int t = 0; t < nt; ++t) {
        const int r0 = rb * TR, c0 = cc * g.tca;
        mbar_wait(&full[s], use & 1);
        const z128* tile = stages + (size_t)s * TILE;
        if (!H) {
          const int row = r0 + a_idx;
          if (row < b.R) {
            const z128* tp = tile + a_idx + a_grp * g.ldt;
            const z128* vp = v + c0 + a_grp;
            const int st = ngrp * g.ldt;
            if (b.mask == NONE && c0 + a_grp + ngrp * (ITER - 1) < b.C) {
              // all of this thread's columns exist: no predicates, two accumulators
#pragma unroll
              for (int q = 0; q < ITER; q += 2) {
                cmac(acc, tp[q * st], vp[q * ngrp]);
                cmac(acc2, tp[(q + 1) * st], vp[(q + 1) * ngrp]);
              }
            } else {
#pragma unroll
              for (int q = 0; q < ITER; ++q) {
                const int c = c0 + a_grp + ngrp * q;
                if (c < b.C && keep(b.mask, row, c)) cmac(acc, tp[q * st], vp[q * ngrp]);
              }
            }
          }
        } else {
          const int c = c0 + a_idx;
          if (c < b.C) {
            const z128* tp = tile + a_grp * ITER + a_idx * g.ldt;
            const z128* vp = v + r0 + a_grp * ITER;
            if (b.mask == NONE && r0 + a_grp * ITER + ITER <= b.R) {
#pragma unroll
              for (int q = 0; q < ITER; q += 2) {
                cmac_conj(acc, tp[q], vp[q]);
                cmac_conj(acc2, tp[q + 1], vp[q + 1]);
              }
            } else {
#pragma unroll
              for (int q = 0; q < ITER; ++q) {
                const int row = r0 + a_grp * ITER + q;
                if (row < b.R && keep(b.mask, row, c)) cmac_conj(acc, tp[q], vp[q]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == (unsigned)nstages) { s = 0; ++use; }
        // ---- end of an output group: row block (N) / column chunk (H)
        bool last;
        if (H) { last = ++rb == g.nrb; if (last) { rb = 0; ++cc; } }
        else { last = ++cc == g.ncc; if (last) { cc = 0; ++rb; } }
        if (!last) continue;
        z128 total = acc + acc2;
        acc = mk(0, 0);
        acc2 = mk(0, 0);
        if (H && g.tca < 32) {
          // row groups of a column sit tca lanes apart: fold them inside the warp, then one partial per warp
          z128* pp = part[grp & 1];
          ++grp;
          for (int o2 = g.tca; o2 < 32; o2 <<= 1) {
            total.x += __shfl_xor_sync(0xffffffffu, total.x, o2);
            total.y += __shfl_xor_sync(0xffffffffu, total.y, o2);
          }
          if (lane < g.tca) pp[wid * g.tca + lane] = total;
          cons_sync<NCONS>();
          if (tid < nout) {
            total = pp[tid];
#pragma unroll
            for (int q = 1; q < NCONS / 32; ++q) total += pp[tid + q * g.tca];
          }
        } else if (ngrp > 1) {
          z128* pp = part[grp & 1];
          ++grp;
          pp[tid] = total;
          cons_sync<NCONS>();
          if (tid < nout) {
            total = pp[tid];
            for (int q = 1; q < ngrp; ++q) total += pp[tid + q * nout];
          }
        }
        const int li = (H ? c0 : r0) + tid;   // block-local index of this thread's result
        if (tid < nout && li < (H ? b.C : b.R)) {
          const int t_i = b.out0 + li;
          if (b.emit == Z_UNIT || b.emit == Z_PLAIN) {
            const z128 z = b.emit == Z_UNIT ? ys[t_i] + total : total;
            zs[t_i] = z;
            vout[f.col0 + t_i] = z;
          } else if (b.emit == SUB_Y || t_i < k) {
            ys[t_i] -= total;
          } else {
            // cb -= total on the copy fetched ahead of time: no dependent global read here.  A row is updated
            // once per step, so the running value is kept in shared memory across the steps.
            const z128 nv = vbuf[t_i - k] - total;
            vbuf[t_i - k] = nv;
            cb[f.st0 + (t_i - k)] = nv;
          }
        }
      }
      cons_sync<NCONS>();   // results of this operation visible to the next one
    }
  }
}

template <bool H, bool UP, int NCONS, int NPROD>
static bool launch_front_stream_t(lsa_handle_impl& h, cudaStream_t st, int cnt, const int* lvl_front, int first, int maxk,
                                int maxr, const z128* fac, const z128* vin, z128* vout, z128* cb, const z128* anc) {
  using namespace stream;
  const int kmax = (maxk + 7) / 8 * 8, rmax = (maxr + 7) / 8 * 8;
  const size_t fixed = sizeof(z128) * (4 * (size_t)kmax + 2 * (size_t)rmax) + 3 * sizeof(int) * (size_t)rmax;
  // ring depth: levels with many fronts want many CTAs per SM (the per-front latencies overlap across CTAs),
  // levels with few fronts want one deep ring per SM (a CTA streams ~ depth x 16 KB per memory round trip)
  constexpr int NTHREADS = NCONS + 32 * NPROD;
  const int max_per_sm = NCONS == 128 ? 5 : 3;   // register file: 64 registers x threads per CTA
  const int per_sm = std::max(1, std::min(max_per_sm, cdiv(cnt, h.num_sms)));
  int nstages = h.stream_stages > 0 ? h.stream_stages : std::max(2, std::min(MAX_STAGES, 10 / per_sm));
  while (nstages > 2 && fixed + sizeof(z128) * (size_t)nstages * TILE > (size_t)STREAM_MAX_SMEM) --nstages;
  const size_t smem = fixed + sizeof(z128) * (size_t)nstages * TILE;
  if (smem > (size_t)STREAM_MAX_SMEM) return false;
  static PerDeviceOnce attr_done;   // per instantiation
  static int occ_cache[2] = {0, 0};  // (shared-memory size, CTAs per SM): the same for every B200 of a node
  if (attr_done.first())
    LSA_CUDA(cudaFuncSetAttribute(k_front_stream<H, UP, NCONS, NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_MAX_SMEM));
  int occ = 0;
  if (occ_cache[0] == (int)smem) occ = occ_cache[1];
  else {
    LSA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_front_stream<H, UP, NCONS, NPROD>, NTHREADS, smem));
    occ_cache[0] = (int)smem;
    occ_cache[1] = occ;
  }
  const int grid = std::max(1, std::min(cnt, std::max(1, occ) * h.num_sms));
  k_front_stream<H, UP, NCONS, NPROD><<<grid, NTHREADS, smem, st>>>(h.d_fronts, lvl_front, first, cnt, h.d_st_idx, fac, vin, vout, cb, anc,
                                                      kmax, rmax, nstages, h.stream_flags);
  LSA_LAUNCH_CHECK();
  return true;
}

// Small fronts (leaf-like levels): 4 consumer warps + 1 producer, up to 5 CTAs per SM, the per-front latencies
// overlap across CTAs.  Larger fronts: 8 consumer warps + 2 producers per CTA (per-CTA throughput matters).
template <bool H, bool UP>
static bool launch_front_stream(lsa_handle_impl& h, cudaStream_t st, int cnt, const int* lvl_front, int first, int maxk,
                                int maxr, int max_m, const z128* fac, const z128* vin, z128* vout, z128* cb,
                                const z128* anc) {
  const bool small = h.stream_small_rows > 0 && max_m <= h.stream_small_rows;
  if (small) return launch_front_stream_t<H, UP, 128, 1>(h, st, cnt, lvl_front, first, maxk, maxr, fac, vin, vout, cb, anc);
  return launch_front_stream_t<H, UP, 256, 2>(h, st, cnt, lvl_front, first, maxk, maxr, fac, vin, vout, cb, anc);
}

__global__ void k_unpermute(const z128* __restrict__ y, z128* __restrict__ x, const int* __restrict__ gperm, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[gperm[i]] = y[i];
}

// H down sweep: x[pivot rows of the level's fronts] = P^T y  (one CTA per front)
__global__ void __launch_bounds__(256) k_level_unpermute(const Front* __restrict__ fronts, const int* __restrict__ lvl_front,
                                                         int first, const int* __restrict__ gperm,
                                                         const z128* __restrict__ y, z128* __restrict__ x) {
  const Front f = fronts[lvl_front[first + blockIdx.x]];
  for (int i = threadIdx.x; i < f.k; i += blockDim.x) x[gperm[f.col0 + i]] = y[f.col0 + i];
}

// ------------------------------------------------------ partitioned solve: exchange over the replicated rows
//
// After the up sweep of a GPU's own sub-trees every sub-tree root holds a contribution vector for rows of the
// replicated top.  Contributions are additive on their way up the tree, so they are added straight to the
// right-hand-side entries of their final rows (x[st_idx[...]]); then ONE all-reduce over the replicated rows
// sums, in the same payload, the sub-tree contributions of all GPUs and the partial SpMV results of those rows
// (each GPU multiplied only the columns it owns).  One CTA, the roots one after the other (a root's row set is
// injective, two roots may share rows): deterministic.
__global__ void __launch_bounds__(1024) k_cut_scatter(const Front* __restrict__ fronts, const int* __restrict__ cut_own,
                                                      int ncut, const int* __restrict__ st_idx,
                                                      const z128* __restrict__ cb, z128* __restrict__ x) {
  for (int q = 0; q < ncut; ++q) {
    const Front c = fronts[cut_own[q]];
    for (int t = threadIdx.x; t < c.r; t += blockDim.x) x[st_idx[c.st0 + t]] += cb[c.st0 + t];
    __syncthreads();
  }
}
__global__ void k_pack_rows(const z128* __restrict__ x, const int* __restrict__ rows, int nrows, z128* __restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) buf[i] = x[rows[i]];
}
__global__ void k_unpack_rows(z128* __restrict__ x, const int* __restrict__ rows, int nrows, const z128* __restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) x[rows[i]] = buf[i];
}
__global__ void k_zero_rows(z128* __restrict__ x, const int* __restrict__ rows, int nrows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) x[rows[i]] = mk(0, 0);
}
// A vector whose replicated rows hold the SAME (complete) values on every GPU becomes one whose replicated rows
// sum to those values over the GPUs: the form the sweeps expect on entry.
void replicated_rows_to_partial(lsa_handle_impl& h, z128* x) {
  if (!h.partitioned || h.part.rank == 0 || h.part.n_top_rows == 0) return;
  k_zero_rows<<<cdiv(h.part.n_top_rows, 256), 256, 0, h.stream>>>(x, h.d_top_rows, (int)h.part.n_top_rows);
  LSA_LAUNCH_CHECK();
}

// x[replicated rows] <- sum over the GPUs (after adding this GPU's sub-tree contributions)
void exchange_replicated_rows(lsa_handle_impl& h, z128* x, bool with_cut_contributions) {
  cudaStream_t st = h.stream;
  const int nrows = (int)h.part.n_top_rows;
  if (with_cut_contributions && h.n_cut_own > 0) {
    k_cut_scatter<<<1, 1024, 0, st>>>(h.d_fronts, h.d_cut_own, h.n_cut_own, h.d_st_idx, h.d_cb, x);
    LSA_LAUNCH_CHECK();
  }
  if (nrows == 0) return;
  k_pack_rows<<<cdiv(nrows, 256), 256, 0, st>>>(x, h.d_top_rows, nrows, h.d_topbuf);
  comm_allreduce_sum(h.comm, (double*)h.d_topbuf, 2 * (size_t)nrows, st);
  k_unpack_rows<<<cdiv(nrows, 256), 256, 0, st>>>(x, h.d_top_rows, nrows, h.d_topbuf);
  LSA_LAUNCH_CHECK();
}

// splits of a chunk's triangular GEMV (gridDim.z; a CTA handles `tri_span` input entries of its 32 rows): only for the
// few wide fronts of the tree top, where one CTA per 32-row chunk leaves SMs idle while the longest rows of the
// triangle (32 x k entries) bound the launch
static int tri_splits(const lsa_handle_impl& h, const SolveChunk& c) {
  const long long ctas = (long long)cdiv(c.maxk, 32) * c.cnt;
  if (c.maxk <= 512 || ctas > 3LL * h.num_sms || h.tri_span <= 0) return 1;
  return std::min(8, cdiv(c.maxk, h.tri_span));
}

template <class T, bool H, bool UP>
static void launch_tri_gemv(lsa_handle_impl& h, cudaStream_t st, const SolveChunk& c, const int* d_lvl_front, const T* fac,
                            const z128* in, z128* out) {
  int ns = tri_splits(h, c);
  if ((long long)cdiv(c.maxk, 32) * c.cnt > h.tri_slots) ns = 1;   // (option changed after the plan was made)
  const dim3 grid(cdiv(c.maxk, 32), c.cnt, ns);
  const int span = std::max(32, h.tri_span);
  if (c.maxk <= 512)
    k_tri_gemv<T, H, UP, 8><<<grid, 256, 0, st>>>(h.d_fronts, d_lvl_front, c.first, fac, in, out, h.d_tri_scratch, h.d_tri_tickets, span);
  else
    k_tri_gemv<T, H, UP, 32><<<grid, 1024, 0, st>>>(h.d_fronts, d_lvl_front, c.first, fac, in, out, h.d_tri_scratch, h.d_tri_tickets, span);
  LSA_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------- driver

// Level plan of the sweeps (host, once per factorisation): chunks of <= 32768 fronts of one level, each with the
// kernel family that sweeps it.  post_factor reads it to know which pivot blocks to invert as a whole.
void plan_solve(lsa_handle_impl& h, int scalar) {
  using namespace stream;
  const Symbolic& sym = h.sym;
  constexpr int YMAX = 32768;
  h.solve_plan.clear();
  long long scratch = 0, max_region = 0;
  for (int d = 0; d < sym.nlevels; ++d) {
    const int lbeg = sym.lvl_ptr[d], cnt_all = sym.lvl_ptr[d + 1] - lbeg;
    for (int y0 = 0; y0 < cnt_all; y0 += YMAX) {
      SolveChunk c{};
      c.level = d;
      c.first = lbeg + y0;
      c.cnt = std::min(YMAX, cnt_all - y0);
      long long sumk2 = 0;
      for (int q = c.first; q < c.first + c.cnt; ++q) {
        const Front& f = sym.fronts[sym.lvl_front[q]];
        c.maxk = std::max(c.maxk, f.k);
        c.max_r = std::max(c.max_r, f.r);
        c.max_m = std::max(c.max_m, f.k + f.r);
        if (f.k > SB) sumk2 += inv_region(f.k);
      }
      const int kmax = (c.maxk + 7) / 8 * 8, rmax = (c.max_r + 7) / 8 * 8;
      const size_t fixed = sizeof(z128) * (4 * (size_t)kmax + 2 * (size_t)rmax) + 3 * sizeof(int) * (size_t)rmax;
      const bool stream_fits = fixed + sizeof(z128) * 2 * (size_t)TILE <= (size_t)STREAM_MAX_SMEM;
      if (scalar == LSA_C128 && h.use_stream && stream_fits && (c.maxk <= SB || c.cnt >= h.stream_min_fronts))
        c.mode = SOLVE_STREAM;
      else if (c.maxk <= h.invert_max_k && inv_region(c.maxk) <= (4LL << 30) / (scalar == LSA_C128 ? 16 : 8))
        c.mode = SOLVE_INVERTED;   // (whole-block inverses need a scratch region per front: capped at 4 GiB)
      else
        c.mode = SOLVE_STEPS;
      if (c.mode == SOLVE_INVERTED && c.maxk > SB) {
        scratch = std::max(scratch, std::min<long long>(sumk2, 1LL << 26));
        max_region = std::max(max_region, inv_region(c.maxk));
      }
      h.solve_plan.push_back(c);
    }
  }
  // scratch of the whole-block inversions: every batch holds at least one front
  scratch = std::max(scratch, max_region);
  const long long bytes = scratch * (scalar == LSA_C128 ? 16 : 8);
  if (bytes > h.inv_scratch_bytes) {
    if (h.d_inv_scratch) cudaFree(h.d_inv_scratch);
    h.d_inv_scratch = nullptr;
    LSA_CUDA(cudaMalloc(&h.d_inv_scratch, bytes));
    h.inv_scratch_bytes = bytes;
  }
  h.inv_scratch_entries = scratch;
  if (!h.d_inv_off) LSA_CUDA(cudaMalloc(&h.d_inv_off, sizeof(long long) * YMAX));
  // scratch + tickets of the split triangular GEMVs (tri_splits)
  long long slots = 0;
  for (const SolveChunk& c : h.solve_plan)
    if (c.mode == SOLVE_INVERTED && tri_splits(h, c) > 1) slots = std::max(slots, (long long)cdiv(c.maxk, 32) * c.cnt);
  if (slots > h.tri_slots) {
    if (h.d_tri_scratch) cudaFree(h.d_tri_scratch);
    if (h.d_tri_tickets) cudaFree(h.d_tri_tickets);
    h.d_tri_scratch = nullptr;
    h.d_tri_tickets = nullptr;
    LSA_CUDA(cudaMalloc(&h.d_tri_scratch, sizeof(z128) * (size_t)slots * 8 * 32));
    LSA_CUDA(cudaMalloc(&h.d_tri_tickets, sizeof(int) * (size_t)slots));
    LSA_CUDA(cudaMemsetAsync(h.d_tri_tickets, 0, sizeof(int) * (size_t)slots, h.stream));
    h.tri_slots = slots;
  }
}

// z[pivot rows] *= 1 / d  between the two sweeps of the symmetric factorisation (the inverted diagonal blocks
// carry 1 / u_ii on their diagonal)
template <class T>
__global__ void __launch_bounds__(128) k_scale_diag(const Front* __restrict__ fronts, int ns, const T* __restrict__ fac,
                                                    z128* __restrict__ z) {
  const int s = blockIdx.x;
  if (s >= ns) return;
  const Front f = fronts[s];
  const long long m = (long long)f.k + f.r;
  const T* P = fac + f.p_off;
  for (int i = threadIdx.x; i < f.k; i += blockDim.x) z[f.col0 + i] = P[i + (long long)i * m] * z[f.col0 + i];
}

// HU / HD: conjugate-transposed (adjoint) kernels in the up / down sweep.  <false, false>: F^-1; <true, true>: F^-H;
// <false, true> (real symmetric factorisation F = L D L^T): L-sweep up, scaling by D^-1, L^T-sweep down.
template <class T, bool HU, bool HD>
static void solve_impl(lsa_handle_impl& h, z128* x, int* n_kernels) {
  const Symbolic& sym = h.sym;
  cudaStream_t st = h.stream;
  const T* fac = (const T*)h.d_fac;
  z128* y = h.d_t;    // up sweep: pre-solve values;  down sweep: final values
  z128* z = h.d_t2;   // up sweep: solved values;     down sweep: pre-solve values
  z128* cb = h.d_cb;
  int launches = 0;
  SweepTrace tr;
  tr.begin(st);
  const std::vector<int>& lvl_front = sym.lvl_front;
  const int* d_lvl_front = h.d_lvl_front;
  auto up_off = [&](const SolveChunk& c) {
    if (c.max_r <= 0) return;
    if (c.maxk > 512) k_up_off<T, HU, 32><<<dim3(cdiv(c.max_r, 32), c.cnt), 1024, 0, st>>>(h.d_fronts, d_lvl_front, c.first, fac, z, cb);
    else k_up_off<T, HU, 8><<<dim3(cdiv(c.max_r, 32), c.cnt), 256, 0, st>>>(h.d_fronts, d_lvl_front, c.first, fac, z, cb);
    LSA_LAUNCH_CHECK();
    tr.mark("up_off", c.level, 0, cdiv(c.max_r, 32), c.cnt);
    launches++;
  };
  // ---- up sweep: deepest level first
  bool exchanged = !h.partitioned;
  for (int ci = (int)h.solve_plan.size() - 1; ci >= -1; --ci) {
    if (!exchanged && (ci < 0 || h.solve_plan[ci].level < h.part.n_top_levels)) {
      // partitioned solve: this GPU's sub-trees are done -> all-reduce over the replicated rows, then the top
      exchange_replicated_rows(h, x, true);
      tr.mark("exchange", h.part.n_top_levels, 0, (int)h.part.n_top_rows, 1);
      launches += 3;
      exchanged = true;
    }
    if (ci < 0) break;
    const SolveChunk& c = h.solve_plan[ci];
    const int d = c.level, cnt = c.cnt, first = c.first, maxk = c.maxk, max_r = c.max_r, max_m = c.max_m;
    if (cnt <= 2 * h.num_sms)
      k_up_gather<!HU, 1024><<<cnt, 1024, 0, st>>>(h.d_fronts, d_lvl_front, first, h.d_child_idx, h.d_ea_map, h.d_gperm, x, y, cb);
    else if (max_m <= 512 && cnt >= 16 * h.num_sms)
      k_up_gather_warp<!HU><<<cdiv(cnt, 8), 256, 0, st>>>(h.d_fronts, d_lvl_front, first, cnt, h.d_child_idx, h.d_ea_map, h.d_gperm, x, y, cb);
    else
      k_up_gather<!HU, 256><<<cnt, 256, 0, st>>>(h.d_fronts, d_lvl_front, first, h.d_child_idx, h.d_ea_map, h.d_gperm, x, y, cb);
    LSA_LAUNCH_CHECK();
    tr.mark("up_gather", d, 0, cnt, 1);
    launches++;
    if constexpr (scalar_traits<T>::is_complex) {
      if (c.mode == SOLVE_STREAM) {
        if (!launch_front_stream<HU, true>(h, st, cnt, d_lvl_front, first, maxk, max_r, max_m, fac, y, z, cb, nullptr))
          throw std::runtime_error("streamed sweep does not fit although the plan says so");
        tr.mark("up_stream", d, 0, cnt, 1);
        launches++;
        continue;
      }
    }
    if (c.mode == SOLVE_INVERTED) {
      launch_tri_gemv<T, HU, true>(h, st, c, d_lvl_front, fac, y, z);
      tr.mark("up_tri", d, 0, cdiv(maxk, 32), cnt);
      launches++;
      up_off(c);
      continue;
    }
    // ---- pivot blocks too large to invert as a whole: chains of 128-pivot steps
    const int csize = cluster_width(cnt, max_m, h.num_sms, h.cluster_max_width);
    if (maxk > SB && h.use_clusters && h.cluster_slices && cnt * 16 <= h.num_sms) {
      launch_sweep_slices<T, HU, true>(st, cnt, h.d_fronts, d_lvl_front, first, fac, y, z, cb);
      tr.mark("up_slices", d, 16, 16 * cnt, 1);
      launches++;
      up_off(c);
      continue;
    }
    // tall fronts need the whole GPU per step; up to `cluster_max_rows` rows a cluster of <= 8 SMs keeps up
    // and saves the launches (inside a CUDA graph the two are within 3 % of each other, profiles/r1f_*)
    if (maxk > SB && max_m <= h.cluster_max_rows && h.use_clusters) {
      // contribution rows deferred to one wide GEMV
      const int defer = (h.defer_cb && max_r > 0) ? 1 : 0;
      const int cs = defer ? cluster_width(cnt, maxk, h.num_sms, h.cluster_max_width) : csize;
      sweep_cluster<T, HU, true>(st, cs, cnt, h.d_fronts, d_lvl_front, first, fac, y, z, cb, defer);
      tr.mark("up_cluster", d, cs, cs * cnt, 1);
      launches++;
      if (defer) up_off(c);
      continue;
    }
    for (int j0 = 0; j0 < maxk; j0 += SB) {
      int act = 0, max_rows = 0;
      for (int q = first; q < first + cnt; ++q) {
        const Front& f = sym.fronts[lvl_front[q]];
        if (f.k <= j0) break;
        act++;
        max_rows = std::max(max_rows, f.k + f.r - std::min(f.k, j0 + SB));
      }
      const int gx = std::max(1, cdiv(max_rows, SB));
      if (maxk > SB) k_step<T, HU, true, 1024><<<dim3(gx, act), 1024, 0, st>>>(h.d_fronts, d_lvl_front, first, j0, fac, y, z, cb);
      else k_step<T, HU, true, 256><<<dim3(gx, act), 256, 0, st>>>(h.d_fronts, d_lvl_front, first, j0, fac, y, z, cb);
      LSA_LAUNCH_CHECK();
      tr.mark("up_step", d, j0, gx, act);
      launches++;
    }
  }
  if (HU != HD) {
    // symmetric factorisation: D^-1 between the L-sweep and the L^T-sweep
    if (sym.ns > 0) k_scale_diag<T><<<sym.ns, 128, 0, st>>>(h.d_fronts, sym.ns, fac, z);
    LSA_LAUNCH_CHECK();
    launches++;
  }
  // decoupled 1 x 1 pivots (after the exchange of a partitioned solve: their right-hand sides are replicated rows)
  if (sym.n_iso > 0) {
    k_solve_decoupled<T, HD><<<cdiv(sym.n_iso, 256), 256, 0, st>>>(fac + sym.diag_off, sym.n_iso, x, y);
    LSA_LAUNCH_CHECK();
    launches++;
  }
  // ---- down sweep: roots first
  for (size_t ci = 0; ci < h.solve_plan.size(); ++ci) {
    const SolveChunk& c = h.solve_plan[ci];
    const int d = c.level, cnt = c.cnt, first = c.first, maxk = c.maxk, maxr = c.max_r, max_mk = c.max_m;
    bool streamed = false;
    if constexpr (scalar_traits<T>::is_complex) {
      if (c.mode == SOLVE_STREAM) {
        if (!launch_front_stream<HD, false>(h, st, cnt, d_lvl_front, first, maxk, maxr, maxk + maxr, fac, z, y, cb, HD ? x : y))
          throw std::runtime_error("streamed sweep does not fit although the plan says so");
        tr.mark("down_stream", d, 0, cnt, 1);
        launches++;
        streamed = true;
      }
    }
    if (!streamed && maxr > 0) {
      if (maxr > 512) k_down_off<T, HD, 32><<<dim3(cdiv(maxk, 32), cnt), 1024, 0, st>>>(h.d_fronts, d_lvl_front, first, h.d_st_idx, fac, HD ? x : y, z);
      else k_down_off<T, HD, 8><<<dim3(cdiv(maxk, 32), cnt), 256, 0, st>>>(h.d_fronts, d_lvl_front, first, h.d_st_idx, fac, HD ? x : y, z);
      LSA_LAUNCH_CHECK();
      tr.mark("down_off", d, 0, cdiv(maxk, 32), cnt);
      launches++;
    }
    if (streamed) {
    } else if (c.mode == SOLVE_INVERTED) {
      launch_tri_gemv<T, HD, false>(h, st, c, d_lvl_front, fac, z, y);
      tr.mark("down_tri", d, 0, cdiv(maxk, 32), cnt);
      launches++;
    } else if (maxk > SB && h.use_clusters && h.cluster_slices && cnt * 16 <= h.num_sms) {
      launch_sweep_slices<T, HD, false>(st, cnt, h.d_fronts, d_lvl_front, first, fac, z, y, cb);
      tr.mark("down_slices", d, 16, 16 * cnt, 1);
      launches++;
    } else if (maxk > SB && max_mk <= h.cluster_max_rows && h.use_clusters) {
      const int csize = cluster_width(cnt, maxk, h.num_sms, h.cluster_max_width);
      sweep_cluster<T, HD, false>(st, csize, cnt, h.d_fronts, d_lvl_front, first, fac, z, y, cb);
      tr.mark("down_cluster", d, csize, csize * cnt, 1);
      launches++;
    } else {
      for (int j0 = ((maxk - 1) / SB) * SB; j0 >= 0; j0 -= SB) {
        int act = 0;
        for (int q = first; q < first + cnt; ++q) {
          if (sym.fronts[lvl_front[q]].k <= j0) break;
          act++;
        }
        if (act == 0) continue;
        const int gx = std::max(1, cdiv(j0, SB));
        if (maxk > SB) k_step<T, HD, false, 1024><<<dim3(gx, act), 1024, 0, st>>>(h.d_fronts, d_lvl_front, first, j0, fac, z, y, cb);
        else k_step<T, HD, false, 256><<<dim3(gx, act), 256, 0, st>>>(h.d_fronts, d_lvl_front, first, j0, fac, z, y, cb);
        LSA_LAUNCH_CHECK();
        tr.mark("down_step", d, j0, gx, act);
        launches++;
      }
    }
    if (HD) {
      k_level_unpermute<<<cnt, 256, 0, st>>>(h.d_fronts, d_lvl_front, first, h.d_gperm, y, x);
      LSA_LAUNCH_CHECK();
      launches++;
    }
  }
  if (HD) {
    if (sym.n_iso > 0) k_unpermute<<<cdiv(sym.n_iso, 256), 256, 0, st>>>(y, x, h.d_gperm, sym.n_iso);
    LSA_LAUNCH_CHECK();
  } else {
    LSA_CUDA(cudaMemcpyAsync(x, y, (size_t)h.n * sizeof(z128), cudaMemcpyDeviceToDevice, st));
  }
  launches++;
  tr.end();
  if (n_kernels) *n_kernels += launches;
}

template <class T>
void solve_permuted(lsa_handle_impl& h, int trans, z128* x, int* n_kernels) {
  if (h.sym.symmetric) {
    // real symmetric F = L D L^T: F^-1 = F^-T = F^-H, one sweep pair with the L blocks only
    if constexpr (!scalar_traits<T>::is_complex) solve_impl<T, false, true>(h, x, n_kernels);
    else throw std::runtime_error("symmetric factorisation with a complex factor");
  } else if (trans == LSA_OP_H) solve_impl<T, true, true>(h, x, n_kernels);
  else solve_impl<T, false, false>(h, x, n_kernels);
}
template void solve_permuted<double>(lsa_handle_impl&, int, z128*, int*);
template void solve_permuted<z128>(lsa_handle_impl&, int, z128*, int*);

}  // namespace lsa
