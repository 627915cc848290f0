// C ABI of the B200 shift-and-invert eigensolve backend (see include/lsa_b200.h).
#include <omp.h>
#include <algorithm>
#include <cstring>
#include <memory>
#include <numeric>

#include "factor.cuh"

namespace lsa {

template <class T>
void post_factor(lsa_handle_impl& h, int* n_kernels);

namespace {

template <class T>
T* dalloc(size_t count) {
  T* p = nullptr;
  if (count == 0) count = 1;
  LSA_CUDA(cudaMalloc(&p, count * sizeof(T)));
  return p;
}
template <class T, class A>
T* dupload(const std::vector<T, A>& v, cudaStream_t st) {
  T* p = dalloc<T>(v.size());
  if (!v.empty()) LSA_CUDA(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  return p;
}
template <class T>
void dfree(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

void free_csr(CsrDev& c) {
  dfree(c.rowptr);
  dfree(c.colidx);
  dfree(c.src);
  dfree(c.rowblk);
  dfree(c.blk_e0);
  c.n_rowblk = 0;
  if (c.vals) cudaFree(c.vals);
  c.vals = nullptr;
  c.nnz = 0;
}

void free_device(lsa_handle_impl& h) {
  drop_solve_graphs(h);
  dfree(h.d_fronts); dfree(h.d_lvl_front); dfree(h.d_st_idx); dfree(h.d_ea_map); dfree(h.d_child_idx);
  dfree(h.d_a_dst); dfree(h.d_m_dst); dfree(h.d_perm); dfree(h.d_ipiv); dfree(h.d_gperm); dfree(h.d_stats);
  if (h.d_a_orig) cudaFree(h.d_a_orig);
  if (h.d_m_orig) cudaFree(h.d_m_orig);
  h.d_a_orig = h.d_m_orig = nullptr;
  free_csr(h.dA); free_csr(h.dM); free_csr(h.dAt); free_csr(h.dMt);
  if (h.d_fac) cudaFree(h.d_fac);
  h.d_fac = nullptr;
  h.fac_capacity_bytes = 0;
  if (h.d_cut_pool) cudaFree(h.d_cut_pool);
  h.d_cut_pool = nullptr;
  h.cut_pool_capacity_bytes = 0;
  dfree(h.d_top_rows); dfree(h.d_topbuf); dfree(h.d_cut_own); dfree(h.d_red);
  if (h.d_inv_scratch) cudaFree(h.d_inv_scratch);
  h.d_inv_scratch = nullptr;
  h.inv_scratch_bytes = h.inv_scratch_entries = 0;
  dfree(h.d_inv_off);
  dfree(h.d_tri_scratch); dfree(h.d_tri_tickets);
  h.tri_slots = 0;
  h.solve_plan.clear();
  for (int q = 0; q < 2; ++q) {
    if (h.d_pool[q]) cudaFree(h.d_pool[q]);
    h.d_pool[q] = nullptr;
    h.pool_capacity_bytes[q] = 0;
  }
  dfree(h.d_x); dfree(h.d_w); dfree(h.d_t); dfree(h.d_t2); dfree(h.d_cb); dfree(h.d_io); dfree(h.d_V); dfree(h.d_S); dfree(h.d_Q);
  dfree(h.d_U); dfree(h.d_mw); dfree(h.d_bn); h.U_cols = 0;
  dfree(h.d_part); dfree(h.d_npart); dfree(h.d_h); dfree(h.d_brow); dfree(h.d_ywork); dfree(h.d_r1); dfree(h.d_r2);
  dfree(h.d_r3); dfree(h.d_Xp); dfree(h.d_ns); dfree(h.d_flag); dfree(h.d_refine); dfree(h.d_wn2); dfree(h.d_wn2b); dfree(h.d_ipart); dfree(h.d_rr); dfree(h.d_theta); dfree(h.d_resid); dfree(h.d_X);
  h.ns_count = 0;
  h.V_cols = 0; h.X_cols = 0; h.ncv_alloc = 0; h.scalar = -1; h.have_values = false;
}

// permuted CSR view of a matrix given in the caller's ordering.
// `cls` (partitioned solve, per PERMUTED index: 2 = row of this GPU's sub-trees, 1 = replicated row, 0 = another
// GPU's row): rows of class 2 keep all entries, rows of class 1 keep the columns of class 2 (and, on rank 0 only,
// of class 1), rows of class 0 keep nothing -- so that the replicated rows of y = Op x SUM to the full product
// over the GPUs (one all-reduce) while every GPU only needs the entries of x it maintains.
void build_permuted(int n, const int64_t* rowptr, const int32_t* colidx, const Symbolic& sym, CsrHost& out,
                    const std::vector<char>* cls = nullptr, bool top_top = true) {
  auto keep = [&](int pi, int pj) {
    if (!cls) return true;
    const char ci = (*cls)[pi], cj = (*cls)[pj];
    if (ci == 2) return true;
    if (ci == 1) return cj == 2 || (cj == 1 && top_top);
    return false;
  };
  out.rowptr.assign(n + 1, 0);
  for (int i = 0; i < n; ++i) {
    const int v = sym.perm[i];
    long long cnt = rowptr[v + 1] - rowptr[v];
    if (cls) {
      cnt = 0;
      for (long long e = rowptr[v]; e < rowptr[v + 1]; ++e) cnt += keep(i, sym.iperm[colidx[e]]);
    }
    out.rowptr[i + 1] = out.rowptr[i] + cnt;
  }
  const long long nnz = out.rowptr[n];
  out.colidx.resize(nnz);   // uninitialised (hvec): first touched by the threads of the fill below
  out.src.resize(nnz);
#pragma omp parallel
  {
    std::vector<std::pair<int, long long>> tmp;
#pragma omp for schedule(dynamic, 512)
    for (int i = 0; i < n; ++i) {
      const int v = sym.perm[i];
      tmp.clear();
      for (long long e = rowptr[v]; e < rowptr[v + 1]; ++e)
        if (keep(i, sym.iperm[colidx[e]])) tmp.emplace_back(sym.iperm[colidx[e]], e);
      std::sort(tmp.begin(), tmp.end());
      long long o = out.rowptr[i];
      for (auto& t : tmp) {
        out.colidx[o] = t.first;
        out.src[o] = t.second;
        ++o;
      }
    }
  }
}

void build_transpose(int n, const CsrHost& a, CsrHost& t) {
  const long long nnz = a.rowptr[n];
  t.rowptr.assign(n + 1, 0);
  for (long long e = 0; e < nnz; ++e) t.rowptr[a.colidx[e] + 1]++;
  for (int i = 0; i < n; ++i) t.rowptr[i + 1] += t.rowptr[i];
  t.colidx.resize(nnz);
  t.src.resize(nnz);
  std::vector<long long> pos(t.rowptr.begin(), t.rowptr.end() - 1);
  for (int i = 0; i < n; ++i)
    for (long long e = a.rowptr[i]; e < a.rowptr[i + 1]; ++e) {
      const long long o = pos[a.colidx[e]]++;
      t.colidx[o] = i;
      t.src[o] = a.src[e];
    }
}

void cut_row_blocks(lsa_handle_impl& h, const CsrHost& c, CsrDev& d) {
  dfree(d.rowblk);
  dfree(d.blk_e0);
  const std::vector<int> blk = spmv_row_blocks(h.n, c.rowptr.data(), h.spmv_block);
  std::vector<long long> e0(blk.size());
  for (size_t i = 0; i < blk.size(); ++i) e0[i] = c.rowptr[blk[i]];
  d.rowblk = dupload(blk, h.stream);
  d.blk_e0 = dupload(e0, h.stream);
  LSA_CUDA(cudaStreamSynchronize(h.stream));   // `blk` is pageable and goes out of scope
  d.n_rowblk = (int)blk.size() - 1;
  d.block_entries = h.spmv_block;
}

void upload_csr(lsa_handle_impl& h, const CsrHost& c, CsrDev& d, bool is_complex) {
  free_csr(d);
  d.nnz = (long long)c.colidx.size();
  d.rowptr = dupload(c.rowptr, h.stream);
  d.colidx = dupload(c.colidx, h.stream);
  d.src = dupload(c.src, h.stream);
  cut_row_blocks(h, c, d);
  d.is_complex = is_complex;
  LSA_CUDA(cudaMalloc(&d.vals, std::max<size_t>(16, (size_t)d.nnz * (is_complex ? 16 : 8))));
}

void refresh_values(lsa_handle_impl& h, CsrDev& d, const void* orig, bool is_complex) {
  if (d.is_complex != is_complex) {
    if (d.vals) cudaFree(d.vals);
    LSA_CUDA(cudaMalloc(&d.vals, std::max<size_t>(16, (size_t)d.nnz * (is_complex ? 16 : 8))));
    d.is_complex = is_complex;
  }
  gather_values(h.stream, orig, is_complex, d.src, d.vals, d.nnz);
}

void ensure_transposes(lsa_handle_impl& h) {
  if (h.dAt.rowptr) return;
  build_transpose(h.n, h.hA, h.hAt);
  upload_csr(h, h.hAt, h.dAt, h.a_complex);
  if (h.has_m) {
    build_transpose(h.n, h.hM, h.hMt);
    upload_csr(h, h.hMt, h.dMt, h.m_complex);
  }
  if (h.have_values) {
    refresh_values(h, h.dAt, h.d_a_orig, h.a_complex);
    if (h.has_m) refresh_values(h, h.dMt, h.d_m_orig, h.m_complex);
  }
}

void ensure_krylov(lsa_handle_impl& h, int ncv) {
  if (ncv <= h.ncv_alloc) return;
  dfree(h.d_V); dfree(h.d_S); dfree(h.d_Q); dfree(h.d_Xp); dfree(h.d_ywork); dfree(h.d_theta); dfree(h.d_resid);
  dfree(h.d_brow);
  const size_t n = h.n;
  h.d_V = dalloc<z128>(n * (size_t)(ncv + 1));
  h.d_Xp = dalloc<z128>(n * (size_t)ncv);
  h.d_S = dalloc<z128>((size_t)(ncv + 1) * ncv);
  h.d_Q = dalloc<z128>((size_t)ncv * ncv);
  h.d_ywork = dalloc<z128>((size_t)ncv * ncv);
  h.d_theta = dalloc<z128>(ncv);
  h.d_resid = dalloc<double>(ncv);
  h.d_brow = dalloc<z128>(ncv + 1);
  h.ncv_alloc = ncv;
}

__global__ void k_conj_inplace(z128* x, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i].y = -x[i].y;
}

template <class T>
__global__ void k_fill_random(T* p, long long n, unsigned long long seed) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (unsigned long long)(i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  p[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

template <bool CPLX>
__global__ void k_gemm_naive(const double* A, const double* B, double* C, int M, int N, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= M || j >= N) return;
  if (CPLX) {
    const z128* a = (const z128*)A;
    const z128* b = (const z128*)B;
    z128 acc = mk(0, 0);
    for (int k = 0; k < K; ++k) acc += a[i + (long long)k * M] * b[k + (long long)j * K];
    z128* c = (z128*)C;
    c[i + (long long)j * M] -= acc;
  } else {
    double acc = 0;
    for (int k = 0; k < K; ++k) acc += A[i + (long long)k * M] * B[k + (long long)j * K];
    C[i + (long long)j * M] -= acc;
  }
}

__global__ void k_maxdiff(const double* a, const double* b, long long n, double* out) {
  __shared__ double s[256];
  double m = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmax(m, fabs(a[i] - b[i]));
  s[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] = fmax(s[threadIdx.x], s[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = s[0];
}

int fail(lsa_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

#define LSA_API_BEGIN try {
#define LSA_API_END(h)                                             \
  }                                                                \
  catch (const lsa::CudaError& e) { return fail(h, LSA_ERR_CUDA, e.what()); } \
  catch (const lsa::NonFiniteError& e) { return fail(h, LSA_ERR_NONFINITE, e.what()); } \
  catch (const lsa::ArgError& e) { return fail(h, LSA_ERR_ARG, e.what()); } \
  catch (const std::bad_alloc&) { return fail(h, LSA_ERR_INTERNAL, "host out of memory"); } \
  catch (const std::exception& e) { return fail(h, LSA_ERR_INTERNAL, e.what()); }

int need_device(lsa_handle* h) {
  if (h->device < 0) return fail(h, LSA_ERR_CUDA, "handle was created without a CUDA device; there is no CPU fallback");
  LSA_API_BEGIN
  LSA_CUDA(cudaSetDevice(h->device));
  LSA_API_END(h)
  return 0;
}

}  // namespace
}  // namespace lsa

using namespace lsa;

extern "C" {

const char* lsa_version(void) { return "lsa_b200 0.1 (sm_100a)"; }

int lsa_create(int32_t n, int32_t device, lsa_handle** out) {
  if (!out || n < 0) return LSA_ERR_ARG;
  lsa_handle* h = new lsa_handle();
  h->n = n;
  h->device = device;
  if (device >= 0) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device >= count) {
      delete h;
      *out = nullptr;
      return LSA_ERR_CUDA;
    }
    cudaSetDevice(device);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete h;
      *out = nullptr;
      return LSA_ERR_CUDA;
    }
  }
  *out = h;
  return LSA_OK;
}

void lsa_destroy(lsa_handle* h) {
  if (!h) return;
  if (h->device >= 0) {
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_device(*h);
    comm_destroy(h->comm);
    for (int q = 0; q < 2; ++q) {
      if (h->h_stage[q]) cudaFreeHost(h->h_stage[q]);
      if (h->ev_stage[q]) cudaEventDestroy(h->ev_stage[q]);
    }
    if (h->stream) cudaStreamDestroy(h->stream);
  }
  delete h;
}

const char* lsa_last_error(const lsa_handle* h) { return h ? h->err.c_str() : "null handle"; }

int lsa_analyze(lsa_handle* h, const int64_t* a_rowptr, const int32_t* a_colidx, const int64_t* m_rowptr,
                const int32_t* m_colidx, int32_t leaf_size, int32_t dim, const double* coords,
                const uint8_t* order_last, int32_t nthreads) {
  if (!h || !a_rowptr || !a_colidx) return LSA_ERR_ARG;
  LSA_API_BEGIN
  const int n = h->n;
  const bool trace_an = getenv("LSA_TRACE_ANALYZE") != nullptr;
  double t_an = omp_get_wtime();
  auto mark = [&](const char* what) {
    if (!trace_an) return;
    const double t = omp_get_wtime();
    fprintf(stderr, "[lsa_analyze] %-28s %8.3f s\n", what, t - t_an);
    t_an = t;
  };
  h->has_m = m_rowptr != nullptr;
  h->nnz_a = a_rowptr[n];
  h->nnz_m = h->has_m ? m_rowptr[n] : 0;
  // union pattern for the graph
  std::vector<long long> urow(n + 1, 0);
  std::unique_ptr<int[]> ucol;
  if (h->has_m) {
    ucol.reset(new int[(size_t)(h->nnz_a + h->nnz_m)]);
#pragma omp parallel for schedule(static)
    for (int i = 0; i <= n; ++i) urow[i] = a_rowptr[i] + m_rowptr[i];
#pragma omp parallel for schedule(dynamic, 2048)
    for (int i = 0; i < n; ++i) {
      int* o = std::copy(a_colidx + a_rowptr[i], a_colidx + a_rowptr[i + 1], ucol.get() + urow[i]);
      std::copy(m_colidx + m_rowptr[i], m_colidx + m_rowptr[i + 1], o);
    }
  }
  AnalyzeOptions opt;
  opt.leaf_size = leaf_size > 0 ? leaf_size : 64;
  opt.dim = coords ? dim : 0;
  opt.coords = coords;
  opt.order_last = order_last;
  opt.nthreads = nthreads;
  opt.coupled_fraction = h->coupled_fraction;
  opt.symmetric = h->symmetric;
  mark("union pattern");
  if (const char* e = getenv("LSA_COUPLED_FRACTION")) opt.coupled_fraction = atof(e);
  if (const char* e = getenv("LSA_CAP_FRACTION")) opt.cap_fraction = atof(e);
  if (h->has_m) analyze(n, urow.data(), ucol.get(), opt, h->sym);
  else analyze(n, (const long long*)a_rowptr, a_colidx, opt, h->sym);
  mark("analyze (graph, ND, symbolic)");
  // ---- partitioned solve: every rank analyses the whole pattern (deterministic), then keeps its part
  h->partitioned = h->part_world > 1;
  std::vector<char> cls;
  h->upd_ranges.clear();
  h->dot_ranges.clear();
  if (h->partitioned) {
    Symbolic global;
    global = std::move(h->sym);
    partition(global, h->part_rank, h->part_world, h->sym, h->part);
    cls.assign(n, 0);
    const Partition& pt = h->part;
    for (size_t q = 0; q < pt.top_lo.size(); ++q) std::fill(cls.begin() + pt.top_lo[q], cls.begin() + pt.top_hi[q], (char)1);
    for (size_t q = 0; q < pt.own_lo.size(); ++q) std::fill(cls.begin() + pt.own_lo[q], cls.begin() + pt.own_hi[q], (char)2);
    // maintained rows (replicated + own) and the rows this rank counts in dot products, as merged sorted ranges
    for (int i = 0; i < n;) {
      if (cls[i] == 0) { ++i; continue; }
      int j = i;
      while (j < n && cls[j] != 0) ++j;
      h->upd_ranges.emplace_back(i, j);
      i = j;
    }
    for (int i = 0; i < n;) {
      const bool mine = cls[i] == 2 || (cls[i] == 1 && pt.rank == 0);
      if (!mine) { ++i; continue; }
      int j = i;
      while (j < n && (cls[j] == 2 || (cls[j] == 1 && pt.rank == 0))) ++j;
      h->dot_ranges.emplace_back(i, j);
      i = j;
    }
  } else {
    h->part = Partition();
    h->upd_ranges.emplace_back(0, n);
    h->dot_ranges.emplace_back(0, n);
  }
  // scatter maps for the caller's entry order of A and M (the union map is recomputed per matrix)
  const Symbolic& sym = h->sym;
  auto local_index = [&](int s, int idx) -> int {
    const Front& f = sym.fronts[s];
    if (idx < f.col0 + f.k) return idx - f.col0;
    const int* b = sym.st_idx.data() + f.st0;
    const int* e = b + f.r;
    const int* it = std::lower_bound(b, e, idx);
    return (it == e || *it != idx) ? -1 : f.k + (int)(it - b);
  };
  auto build_dst = [&](const int64_t* rowptr, const int32_t* colidx, hvec<long long>& dst) {
    dst.resize(rowptr[n]);
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 512) reduction(| : bad)
    for (int i = 0; i < n; ++i) {
      const int pi = sym.iperm[i];
      for (long long e = rowptr[i]; e < rowptr[i + 1]; ++e) {
        const int pj = sym.iperm[colidx[e]];
        if (pi < sym.n_iso || pj < sym.n_iso) {
          if (pi != pj) bad |= 1;
          dst[e] = sym.diag_off + pi;
          continue;
        }
        const int s = sym.sn_of[std::min(pi, pj)];
        if (s < 0) {   // partitioned solve: the entry is assembled on another GPU
          dst[e] = -1;
          continue;
        }
        const Front& f = sym.fronts[s];
        const int lr = local_index(s, pi), lc = local_index(s, pj);
        if (lr < 0 || lc < 0) {
          bad |= 1;
          continue;
        }
        const long long m = f.k + f.r;
        if (sym.symmetric && lc >= f.k) {   // U12 entry: its mirror image lands in L21, nothing is stored for it
          dst[e] = -1;
          continue;
        }
        dst[e] = lc < f.k ? f.p_off + lr + (long long)lc * m : f.q_off + lr + (long long)(lc - f.k) * f.k;
      }
    }
    if (bad) throw std::runtime_error("front structure does not cover the matrix pattern");
  };
  mark("partition");
  build_dst(a_rowptr, a_colidx, h->sym.a_dst);
  if (h->has_m) build_dst(m_rowptr, m_colidx, h->m_dst);
  mark("scatter maps");
  const std::vector<char>* pcls = h->partitioned ? &cls : nullptr;
  build_permuted(n, a_rowptr, a_colidx, sym, h->hA, pcls, h->part.rank == 0);
  if (h->has_m) build_permuted(n, m_rowptr, m_colidx, sym, h->hM, pcls, h->part.rank == 0);
  mark("permuted CSR");
  h->hAt = CsrHost();
  h->hMt = CsrHost();
  h->analyzed = true;

  if (h->device >= 0) {
    LSA_CUDA(cudaSetDevice(h->device));
    free_device(*h);
    cudaStream_t st = h->stream;
    h->d_fronts = dupload(sym.fronts, st);
    h->d_lvl_front = dupload(sym.lvl_front, st);
    {
      cudaDeviceProp prop;
      LSA_CUDA(cudaGetDeviceProperties(&prop, h->device));
      h->num_sms = prop.multiProcessorCount;
    }
    h->d_st_idx = dupload(sym.st_idx, st);
    h->d_ea_map = dupload(sym.ea_map, st);
    h->d_child_idx = dupload(sym.child_idx, st);
    h->d_a_dst = dupload(sym.a_dst, st);
    if (h->has_m) h->d_m_dst = dupload(h->m_dst, st);
    h->d_perm = dupload(sym.perm, st);
    h->d_ipiv = dalloc<int>(n);
    h->d_gperm = dalloc<int>(n);
    h->d_stats = dalloc<DevStats>(1);
    h->d_x = dalloc<z128>(n); h->d_w = dalloc<z128>(n); h->d_t = dalloc<z128>(n); h->d_t2 = dalloc<z128>(n); h->d_io = dalloc<z128>(n);
    h->d_r1 = dalloc<z128>(n); h->d_r2 = dalloc<z128>(n); h->d_r3 = dalloc<z128>(n);
    h->d_cb = dalloc<z128>(sym.st_idx.size());
    h->d_part = dalloc<z128>((size_t)1024 * 256);
    h->d_npart = dalloc<double>((size_t)cdiv(n, 256) + 1);
    h->d_h = dalloc<z128>(512);
    h->d_flag = dalloc<int>(2);
    h->d_refine = dalloc<int>(2);
    h->d_wn2 = dalloc<double>(1024);
    h->d_wn2b = dalloc<double>(1024);
    h->d_ipart = dalloc<int>(256);
    h->d_rr = dalloc<RrInfo>(1);
    h->d_red = dalloc<double>(2 * 256 + 16);
    if (h->partitioned) {
      const Partition& pt = h->part;
      std::vector<int> rows;
      for (size_t q = 0; q < pt.top_lo.size(); ++q)
        for (int i = pt.top_lo[q]; i < pt.top_hi[q]; ++i) rows.push_back(i);
      h->d_top_rows = dupload(rows, st);
      h->d_topbuf = dalloc<z128>(rows.size());
      std::vector<int> cut;
      for (int gs : pt.cut_roots)
        if (pt.owner[gs] == pt.rank) cut.push_back(pt.g2l[gs]);
      h->n_cut_own = (int)cut.size();
      h->d_cut_own = dupload(cut, st);
    }
    upload_csr(*h, h->hA, h->dA, false);
    if (h->has_m) upload_csr(*h, h->hM, h->dM, false);
    LSA_CUDA(cudaStreamSynchronize(st));
    mark("device structures");
  }
  LSA_API_END(h)
  return LSA_OK;
}

int lsa_set_partition(lsa_handle* h, int32_t rank, int32_t world) {
  if (!h || world < 1 || rank < 0 || rank >= world) return LSA_ERR_ARG;
  if (h->analyzed) return fail(h, LSA_ERR_ARG, "lsa_set_partition must precede lsa_analyze");
  h->part_rank = rank;
  h->part_world = world;
  return LSA_OK;
}

int lsa_nccl_load(const char* path) {
  try {
    nccl_load(path);
  } catch (const std::exception&) {
    return LSA_ERR_INTERNAL;
  }
  return LSA_OK;
}

int lsa_nccl_unique_id(void* out128) {
  if (!out128) return LSA_ERR_ARG;
  try {
    nccl_unique_id(out128);
  } catch (const std::exception&) {
    return LSA_ERR_INTERNAL;
  }
  return LSA_OK;
}

int lsa_set_comm(lsa_handle* h, const void* id128) {
  if (!h || !id128) return LSA_ERR_ARG;
  if (int rc = need_device(h)) return rc;
  LSA_API_BEGIN
  comm_destroy(h->comm);
  comm_init(h->comm, id128, h->part_rank, h->part_world);
  // first collective outside any stream capture: NCCL sets its channels up lazily
  double* warm = dalloc<double>(8);
  LSA_CUDA(cudaMemsetAsync(warm, 0, 8 * sizeof(double), h->stream));
  comm_allreduce_sum(h->comm, warm, 8, h->stream);
  LSA_CUDA(cudaStreamSynchronize(h->stream));
  cudaFree(warm);
  LSA_API_END(h)
  return LSA_OK;
}

int lsa_partition_info_get(const lsa_handle* h, lsa_partition_info* out) {
  if (!h || !out || !h->analyzed) return LSA_ERR_ARG;
  std::memset(out, 0, sizeof(*out));
  const Partition& p = h->part;
  out->rank = h->part_rank;
  out->world = h->part_world;
  out->n_fronts_global = h->partitioned ? p.ns_global : h->sym.ns;
  out->n_fronts_local = h->sym.ns;
  out->n_top_levels = p.n_top_levels;
  out->n_cut_roots = (int32_t)p.cut_roots.size();
  for (int o : p.owner) out->n_top_fronts += o == -1;
  out->n_replicated_rows = p.n_top_rows;
  for (size_t q = 0; q < p.own_lo.size(); ++q) out->n_own_rows += p.own_hi[q] - p.own_lo[q];
  out->cut_pool_entries = p.cut_pool_size;
  out->nnz_lu_global = h->partitioned ? p.nnz_lu_global : h->sym.nnz_lu;
  out->flops_real_global = h->partitioned ? p.flops_global : h->sym.flops;
  out->weight_total = p.weight_total;
  out->weight_top = p.weight_top;
  out->weight_max_subtrees = p.weight_max;
  out->weight_mine = p.weight_mine;
  return LSA_OK;
}

int lsa_set_option(lsa_handle* h, const char* name, double value) {
  if (!h || !name) return LSA_ERR_ARG;
  const std::string nm(name);
  if (nm == "symmetric") {
    if (h->analyzed) return fail(h, LSA_ERR_ARG, "option symmetric must be set before lsa_analyze");
    h->symmetric = value != 0.0;
  } else if (nm == "coupled_fraction") {
    if (!(value >= 0.0 && value <= 1.0)) return fail(h, LSA_ERR_ARG, "coupled_fraction must lie in [0, 1]");
    h->coupled_fraction = value;
  } else if (nm == "use_graphs") {
    h->use_graphs = value != 0.0;
  } else if (nm == "use_clusters") {
    h->use_clusters = value != 0.0;
  } else if (nm == "ortho_refine_always") {
    h->ortho_refine_always = value != 0.0;
  } else if (nm == "partition_graphs") {
    h->part_graphs = value != 0.0;
  } else if (nm == "fuse_ortho") {
    h->fuse_ortho = value != 0.0;
  } else if (nm == "use_stream") {
    h->use_stream = value != 0.0;
  } else if (nm == "stream_min_fronts") {
    h->stream_min_fronts = (int)value;
  } else if (nm == "stream_small_rows") {
    h->stream_small_rows = (int)value;
  } else if (nm == "stream_stages") {
    if (value < 0 || value > 12 || value == 1) return fail(h, LSA_ERR_ARG, "stream_stages must be 0 (automatic) or 2 .. 12");
    h->stream_stages = (int)value;
  } else if (nm == "stream_flags") {
    h->stream_flags = (int)value & 7;
  } else if (nm == "spmv_block") {
    if (value != 512 && value != 1024 && value != 2048) return fail(h, LSA_ERR_ARG, "spmv_block must be 512, 1024 or 2048");
    h->spmv_block = (int)value;
    LSA_API_BEGIN
    if (h->dA.rowptr) cut_row_blocks(*h, h->hA, h->dA);
    if (h->dM.rowptr) cut_row_blocks(*h, h->hM, h->dM);
    if (h->dAt.rowptr) cut_row_blocks(*h, h->hAt, h->dAt);
    if (h->dMt.rowptr) cut_row_blocks(*h, h->hMt, h->dMt);
    LSA_API_END(h)
  } else if (nm == "tri_span") {
    if (value < 0 || value > 65536) return fail(h, LSA_ERR_ARG, "tri_span must lie in [0, 65536]");
    h->tri_span = (int)value / 32 * 32;
  } else if (nm == "invert_max_k") {
    if (value < 0 || value > 65536) return fail(h, LSA_ERR_ARG, "invert_max_k must lie in [0, 65536]");
    h->invert_max_k = (int)value;
  } else if (nm == "defer_cb") {
    h->defer_cb = value != 0.0;
  } else if (nm == "cluster_slices") {
    h->cluster_slices = value != 0.0;
  } else if (nm == "cluster_max_rows") {
    h->cluster_max_rows = (int)value;
  } else if (nm == "cluster_max_width") {
    if (value != 1 && value != 2 && value != 4 && value != 8 && value != 16)
      return fail(h, LSA_ERR_ARG, "cluster_max_width must be 1, 2, 4, 8 or 16");
    h->cluster_max_width = (int)value;
  } else {
    return fail(h, LSA_ERR_ARG, "unknown option " + nm);
  }
  return LSA_OK;
}

int lsa_symbolic_info_get(const lsa_handle* h, lsa_symbolic_info* out) {
  if (!h || !out || !h->analyzed) return LSA_ERR_ARG;
  const Symbolic& s = h->sym;
  std::memset(out, 0, sizeof(*out));
  out->n = s.n; out->n_decoupled = s.n_iso; out->n_fronts = s.ns; out->n_levels = s.nlevels;
  out->max_pivots = s.max_k; out->max_front = s.max_m; out->max_rows = s.max_r;
  out->nnz_a = h->nnz_a; out->nnz_m = h->nnz_m;
  out->factor_entries = s.fac_size; out->nnz_lu = s.nnz_lu;
  out->pool_entries[0] = s.pool_size[0]; out->pool_entries[1] = s.pool_size[1];
  out->struct_entries = (int64_t)s.st_idx.size();
  out->flops_real = s.flops;
  for (int q = 0; q < 4; ++q) out->seconds[q] = s.seconds[q];
  return LSA_OK;
}

int64_t lsa_symbolic_array(const lsa_handle* h, const char* name, void* out, int64_t capacity_bytes) {
  if (!h || !name || !h->analyzed) return LSA_ERR_ARG;
  const Symbolic& s = h->sym;
  const std::string nm(name);
  auto give = [&](const void* data, size_t count, size_t elem) -> int64_t {
    if (out) {
      if ((int64_t)(count * elem) > capacity_bytes) return LSA_ERR_ARG;
      std::memcpy(out, data, count * elem);
    }
    return (int64_t)count;
  };
  auto from_fronts = [&](auto getter, size_t elem) -> int64_t {
    if (elem == 4) {
      std::vector<int> v(s.ns);
      for (int i = 0; i < s.ns; ++i) v[i] = (int)getter(s.fronts[i]);
      return give(v.data(), v.size(), 4);
    }
    std::vector<long long> v(s.ns);
    for (int i = 0; i < s.ns; ++i) v[i] = (long long)getter(s.fronts[i]);
    return give(v.data(), v.size(), 8);
  };
  if (nm == "perm") return give(s.perm.data(), s.perm.size(), 4);
  if (nm == "iperm") return give(s.iperm.data(), s.iperm.size(), 4);
  if (nm == "sn_ptr") return give(s.sn_ptr.data(), s.sn_ptr.size(), 4);
  if (nm == "st_ptr") return give(s.st_ptr.data(), s.st_ptr.size(), 8);
  if (nm == "st_idx") return give(s.st_idx.data(), s.st_idx.size(), 4);
  if (nm == "ea_map") return give(s.ea_map.data(), s.ea_map.size(), 4);
  if (nm == "lvl_ptr") return give(s.lvl_ptr.data(), s.lvl_ptr.size(), 4);
  if (nm == "lvl_front") return give(s.lvl_front.data(), s.lvl_front.size(), 4);
  if (nm == "a_dst") return give(s.a_dst.data(), s.a_dst.size(), 8);
  if (nm == "m_dst") return give(h->m_dst.data(), h->m_dst.size(), 8);
  if (nm == "front_col0") return from_fronts([](const Front& f) { return f.col0; }, 4);
  if (nm == "front_flags") return from_fronts([](const Front& f) { return f.flags; }, 4);
  if (nm == "child_idx") return give(s.child_idx.data(), s.child_idx.size(), 4);
  if (nm == "child0") return from_fronts([](const Front& f) { return f.child0; }, 4);
  if (nm == "nchild") return from_fronts([](const Front& f) { return f.nchild; }, 4);
  if (nm == "owner") return give(h->part.owner.data(), h->part.owner.size(), 4);
  if (nm == "g2l") return give(h->part.g2l.data(), h->part.g2l.size(), 4);
  if (nm == "l2g") return give(h->part.l2g.data(), h->part.l2g.size(), 4);
  if (nm == "cut_roots") return give(h->part.cut_roots.data(), h->part.cut_roots.size(), 4);
  if (nm == "top_lo") return give(h->part.top_lo.data(), h->part.top_lo.size(), 4);
  if (nm == "top_hi") return give(h->part.top_hi.data(), h->part.top_hi.size(), 4);
  if (nm == "own_lo") return give(h->part.own_lo.data(), h->part.own_lo.size(), 4);
  if (nm == "own_hi") return give(h->part.own_hi.data(), h->part.own_hi.size(), 4);
  if (nm == "parent") return from_fronts([](const Front& f) { return f.parent; }, 4);
  if (nm == "level") return from_fronts([](const Front& f) { return f.level; }, 4);
  if (nm == "front_k") return from_fronts([](const Front& f) { return f.k; }, 4);
  if (nm == "front_r") return from_fronts([](const Front& f) { return f.r; }, 4);
  if (nm == "p_off") return from_fronts([](const Front& f) { return f.p_off; }, 8);
  if (nm == "q_off") return from_fronts([](const Front& f) { return f.q_off; }, 8);
  if (nm == "c_off") return from_fronts([](const Front& f) { return f.c_off; }, 8);
  return LSA_ERR_ARG;
}

// Pageable host memory -> device: the driver stages such copies through one bounce buffer at ~10 GB/s.  Here a few
// threads copy 16 MiB pieces into two page-locked staging buffers while the previous piece is on the wire.
static bool is_page_locked(const void* p) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

static void upload_pageable(lsa_handle_impl& h, void* dst, const void* src, size_t len) {
  constexpr size_t PIECE = 16u << 20;
  if (!h.h_stage[0]) {
    for (int q = 0; q < 2; ++q) {
      LSA_CUDA(cudaHostAlloc(&h.h_stage[q], PIECE, cudaHostAllocDefault));
      LSA_CUDA(cudaEventCreateWithFlags(&h.ev_stage[q], cudaEventDisableTiming));
    }
  }
  const int nth = std::max(1, std::min(omp_get_max_threads(), 8));
  size_t off = 0;
  for (int it = 0; off < len; ++it, off += PIECE) {
    const int q = it & 1;
    const size_t n = std::min(PIECE, len - off);
    if (it >= 2) LSA_CUDA(cudaEventSynchronize(h.ev_stage[q]));   // the piece sent from this buffer two rounds ago has left
    const size_t part = ((n + nth - 1) / nth + 63) & ~(size_t)63;
#pragma omp parallel for schedule(static) num_threads(nth)
    for (int t = 0; t < nth; ++t) {
      const size_t b = std::min(n, (size_t)t * part), e = std::min(n, b + part);
      if (e > b) std::memcpy((char*)h.h_stage[q] + b, (const char*)src + off + b, e - b);
    }
    LSA_CUDA(cudaMemcpyAsync((char*)dst + off, h.h_stage[q], n, cudaMemcpyHostToDevice, h.stream));
    LSA_CUDA(cudaEventRecord(h.ev_stage[q], h.stream));
  }
}

int lsa_set_values(lsa_handle* h, const void* a_vals, int32_t a_scalar, const void* m_vals, int32_t m_scalar,
                   int32_t on_device) {
  if (!h || !h->analyzed || !a_vals) return LSA_ERR_ARG;
  if (h->has_m && !m_vals && !h->have_values) return fail(h, LSA_ERR_ARG, "M values missing");
  if (int rc = need_device(h)) return rc;
  LSA_API_BEGIN
  cudaStream_t st = h->stream;
  const auto kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  auto put = [&](void*& dst, const void* src, long long nnz, bool cplx, bool& flag) {
    if (dst && flag != cplx) {
      cudaFree(dst);
      dst = nullptr;
    }
    const size_t bytes = std::max<size_t>(16, (size_t)nnz * (cplx ? 16 : 8));
    if (!dst) LSA_CUDA(cudaMalloc(&dst, bytes));
    flag = cplx;
    const size_t len = (size_t)nnz * (cplx ? 16 : 8);
    if (nnz == 0) return;
    if (!on_device && len >= (size_t)(32u << 20) && !is_page_locked(src)) upload_pageable(*h, dst, src, len);
    else LSA_CUDA(cudaMemcpyAsync(dst, src, len, kind, st));
  };
  // m_vals == NULL on a handle that already holds values: M is kept (a Reynolds sweep changes A only)
  const bool new_m = h->has_m && m_vals;
  put(h->d_a_orig, a_vals, h->nnz_a, a_scalar == LSA_C128, h->a_complex);
  if (new_m) put(h->d_m_orig, m_vals, h->nnz_m, m_scalar == LSA_C128, h->m_complex);
  refresh_values(*h, h->dA, h->d_a_orig, h->a_complex);
  if (new_m) refresh_values(*h, h->dM, h->d_m_orig, h->m_complex);
  if (h->dAt.rowptr) {
    refresh_values(*h, h->dAt, h->d_a_orig, h->a_complex);
    if (new_m) refresh_values(*h, h->dMt, h->d_m_orig, h->m_complex);
  }
  value_norms(*h, h->d_a_orig, h->a_complex, h->nnz_a, &h->a_fro, &h->a_max);
  double dummy = 0;
  if (new_m) value_norms(*h, h->d_m_orig, h->m_complex, h->nnz_m, &dummy, &h->m_max);
  h->have_values = true;
  h->scalar = -1;  // previous factors are stale
  LSA_CUDA(cudaStreamSynchronize(st));
  LSA_API_END(h)
  return LSA_OK;
}

int lsa_factor(lsa_handle* h, double alpha_re, double alpha_im, double beta_re, double beta_im, int32_t scalar,
               double tiny_pivot, lsa_factor_stats* stats) {
  if (!h || !h->have_values) return LSA_ERR_ARG;
  if (int rc = need_device(h)) return rc;
  if (scalar == LSA_F64 && (alpha_im != 0.0 || beta_im != 0.0 || h->a_complex || (h->has_m && h->m_complex)))
    return fail(h, LSA_ERR_ARG, "real factorisation requested for complex data or a complex shift");
  if (h->sym.symmetric && scalar != LSA_F64)
    return fail(h, LSA_ERR_ARG, "the symmetric factorisation (option symmetric) is built for real FP64 only");
  LSA_API_BEGIN
  const Symbolic& sym = h->sym;
  cudaStream_t st = h->stream;
  const size_t esz = scalar == LSA_C128 ? 16 : 8;
  const long long need = sym.fac_size * (long long)esz;
  if (need > h->fac_capacity_bytes) {
    if (h->d_fac) cudaFree(h->d_fac);
    h->d_fac = nullptr;
    LSA_CUDA(cudaMalloc(&h->d_fac, std::max<long long>(need, 16)));
    h->fac_capacity_bytes = need;
  }
  for (int q = 0; q < 2; ++q) {
    const long long pn = sym.pool_size[q] * (long long)esz;
    if (pn > h->pool_capacity_bytes[q]) {
      if (h->d_pool[q]) cudaFree(h->d_pool[q]);
      h->d_pool[q] = nullptr;
      LSA_CUDA(cudaMalloc(&h->d_pool[q], std::max<long long>(pn, 16)));
      h->pool_capacity_bytes[q] = pn;
    }
  }
  if (h->partitioned) {
    if (!h->comm.nccl_comm) return fail(h, LSA_ERR_ARG, "partitioned handle: call lsa_set_comm before lsa_factor");
    const long long cn = h->part.cut_pool_size * (long long)esz;
    if (cn > h->cut_pool_capacity_bytes) {
      if (h->d_cut_pool) cudaFree(h->d_cut_pool);
      h->d_cut_pool = nullptr;
      LSA_CUDA(cudaMalloc(&h->d_cut_pool, std::max<long long>(cn, 16)));
      h->cut_pool_capacity_bytes = cn;
    }
  }
  const z128 alpha = mk(alpha_re, alpha_im), beta = mk(beta_re, beta_im);
  const double fmax_est = absz(alpha) * h->a_max + (h->has_m ? absz(beta) * h->m_max : 0.0);
  const double tiny_abs = tiny_pivot * fmax_est;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int nk = 0;
  cudaEventRecord(e0, st);
  h->scalar = -1;
  drop_solve_graphs(*h);
  if (const char* e = getenv("LSA_CLUSTER_MAX_ROWS")) h->cluster_max_rows = atoi(e);
  if (const char* e = getenv("LSA_DEFER_CB")) h->defer_cb = atoi(e) != 0;
  if (const char* e = getenv("LSA_CLUSTER_SLICES")) h->cluster_slices = atoi(e) != 0;
  if (const char* e = getenv("LSA_CLUSTER_MAX_WIDTH")) h->cluster_max_width = atoi(e);
  if (const char* e = getenv("LSA_NO_STREAM")) h->use_stream = atoi(e) == 0;
  if (const char* e = getenv("LSA_STREAM_MIN_FRONTS")) h->stream_min_fronts = atoi(e);
  if (const char* e = getenv("LSA_STREAM_SMALL_ROWS")) h->stream_small_rows = atoi(e);
  if (const char* e = getenv("LSA_STREAM_STAGES")) h->stream_stages = std::max(0, std::min(12, atoi(e)));
  if (const char* e = getenv("LSA_STREAM_FLAGS")) h->stream_flags = atoi(e) & 7;
  if (const char* e = getenv("LSA_NO_CLUSTERS")) {
    if (atoi(e) != 0) h->use_clusters = false;
  }
  if (const char* e = getenv("LSA_NO_GRAPHS")) {
    if (atoi(e) != 0) h->use_graphs = false;
  }
  if (const char* e = getenv("LSA_INVERT_MAX_K")) h->invert_max_k = atoi(e);
  if (const char* e = getenv("LSA_TRI_SPAN")) h->tri_span = std::max(0, atoi(e)) / 32 * 32;
  if (const char* e = getenv("LSA_PARTITION_GRAPHS")) h->part_graphs = atoi(e) != 0;
  plan_solve(*h, scalar);
  if (scalar == LSA_C128) {
    factor_numeric<z128>(*h, alpha, beta, tiny_abs, &nk);
    post_factor<z128>(*h, &nk);
  } else {
    factor_numeric<double>(*h, alpha, beta, tiny_abs, &nk);
    post_factor<double>(*h, &nk);
  }
  cudaEventRecord(e1, st);
  DevStats ds{};
  LSA_CUDA(cudaMemcpyAsync(&ds, h->d_stats, sizeof(DevStats), cudaMemcpyDeviceToHost, st));
  LSA_CUDA(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  lsa_factor_stats fs{};
  fs.seconds = ms * 1e-3;
  fs.flops = sym.flops * (scalar == LSA_C128 ? 4.0 : 1.0);
  fs.n_perturbed = (int64_t)ds.n_perturbed;
  fs.n_row_swaps = (int64_t)ds.n_swaps;
  fs.scalar = scalar;
  fs.n_kernels = nk;
  std::memcpy(&fs.min_pivot, &ds.min_piv_bits, 8);
  std::memcpy(&fs.max_pivot, &ds.max_piv_bits, 8);
  std::memcpy(&fs.max_multiplier, &ds.max_l_bits, 8);
  h->fstats = fs;
  if (stats) *stats = fs;
  if (ds.nonfinite) return fail(h, LSA_ERR_NONFINITE, "non-finite value met during the factorisation");
  if (ds.zero_pivot) return fail(h, LSA_ERR_ZERO_PIVOT, "zero pivot: the shifted operator is singular");
  h->scalar = scalar;
  h->f_alpha = alpha;
  h->f_beta = beta;
  h->counters.factor_flops = fs.flops;
  h->counters.factor_seconds = fs.seconds;
  LSA_API_END(h)
  return LSA_OK;
}

int lsa_solve(lsa_handle* h, int32_t trans, const double* b, double* x, int32_t refine_steps, int32_t on_device) {
  if (!h || !b || !x) return LSA_ERR_ARG;
  if (h->scalar < 0) return fail(h, LSA_ERR_ARG, "lsa_solve called without a valid factorisation");
  if (int rc = need_device(h)) return rc;
  LSA_API_BEGIN
  const int n = h->n;
  cudaStream_t st = h->stream;
  const z128* src = (const z128*)b;
  if (!on_device) {
    LSA_CUDA(cudaMemcpyAsync(h->d_io, b, sizeof(z128) * (size_t)n, cudaMemcpyHostToDevice, st));
    src = h->d_io;
  }
  permute_gather(st, src, h->d_x, h->d_perm, n);
  replicated_rows_to_partial(*h, h->d_x);   // partitioned solve: b is complete on every rank
  if (trans != LSA_OP_N && refine_steps > 0) ensure_transposes(*h);
  if (trans == LSA_OP_T) {
    k_conj_inplace<<<cdiv(n, 256), 256, 0, st>>>(h->d_x, n);
    op_solve(*h, LSA_OP_H, h->d_x, refine_steps);
    k_conj_inplace<<<cdiv(n, 256), 256, 0, st>>>(h->d_x, n);
  } else {
    op_solve(*h, trans, h->d_x, refine_steps);
  }
  make_full(*h, h->d_x, 1, n);
  if (on_device) {
    permute_scatter(st, h->d_x, (z128*)x, h->d_perm, n);
  } else {
    permute_scatter(st, h->d_x, h->d_io, h->d_perm, n);
    LSA_CUDA(cudaMemcpyAsync(x, h->d_io, sizeof(z128) * (size_t)n, cudaMemcpyDeviceToHost, st));
  }
  LSA_CUDA(cudaStreamSynchronize(st));
  h->counters.n_solves++;
  LSA_API_END(h)
  return LSA_OK;
}

int lsa_spmv(lsa_handle* h, int32_t which_matrix, int32_t trans, const double* x, double* y, int32_t on_device) {
  if (!h || !x || !y || !h->have_values) return LSA_ERR_ARG;
  if (which_matrix == LSA_MAT_M && !h->has_m) return fail(h, LSA_ERR_ARG, "no M operator");
  if (int rc = need_device(h)) return rc;
  LSA_API_BEGIN
  const int n = h->n;
  cudaStream_t st = h->stream;
  const z128* src = (const z128*)x;
  if (!on_device) {
    LSA_CUDA(cudaMemcpyAsync(h->d_io, x, sizeof(z128) * (size_t)n, cudaMemcpyHostToDevice, st));
    src = h->d_io;
  }
  permute_gather(st, src, h->d_x, h->d_perm, n);
  if (trans != LSA_OP_N) ensure_transposes(*h);
  const CsrDev& mat = which_matrix == LSA_MAT_A ? (trans == LSA_OP_N ? h->dA : h->dAt) : (trans == LSA_OP_N ? h->dM : h->dMt);
  spmv(*h, mat, trans == LSA_OP_H, h->d_x, h->d_w);
  if (h->partitioned)   // rows of other GPUs are empty here, replicated rows hold partial sums: one sum completes both
    comm_allreduce_sum(h->comm, (double*)h->d_w, 2 * (size_t)n, st);
  if (on_device) {
    permute_scatter(st, h->d_w, (z128*)y, h->d_perm, n);
  } else {
    permute_scatter(st, h->d_w, h->d_io, h->d_perm, n);
    LSA_CUDA(cudaMemcpyAsync(y, h->d_io, sizeof(z128) * (size_t)n, cudaMemcpyDeviceToHost, st));
  }
  LSA_CUDA(cudaStreamSynchronize(st));
  h->counters.n_spmv++;
  LSA_API_END(h)
  return LSA_OK;
}

int lsa_set_nullspace(lsa_handle* h, int32_t count, const double* vecs_c128) {
  if (!h || !h->analyzed || count < 0 || count > 16 || (count > 0 && !vecs_c128)) return LSA_ERR_ARG;
  if (int rc = need_device(h)) return rc;
  if (h->partitioned && count > 0) return fail(h, LSA_ERR_ARG, "nullspace projection is not available on a partitioned handle");
  LSA_API_BEGIN
  const int n = h->n;
  dfree(h->d_ns);
  h->ns_count = 0;
  if (count > 0) {
    h->d_ns = dalloc<z128>((size_t)n * count);
    for (int c = 0; c < count; ++c) {
      LSA_CUDA(cudaMemcpyAsync(h->d_io, (const z128*)vecs_c128 + (size_t)c * n, sizeof(z128) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
      permute_gather(h->stream, h->d_io, h->d_ns + (size_t)c * n, h->d_perm, n);
    }
    LSA_CUDA(cudaStreamSynchronize(h->stream));
    h->ns_count = count;
  }
  LSA_API_END(h)
  return LSA_OK;
}

int lsa_bilinear(lsa_handle* h, int32_t which_matrix, const void* vals, int32_t scalar, const double* a, const double* v,
                 int32_t on_device, double* out_c128) {
  if (!h || !vals || !a || !v || !out_c128 || !h->analyzed) return LSA_ERR_ARG;
  if (which_matrix == LSA_MAT_M && !h->has_m) return fail(h, LSA_ERR_ARG, "no M operator");
  if (int rc = need_device(h)) return rc;
  if (h->partitioned) return fail(h, LSA_ERR_ARG, "lsa_bilinear is not available on a partitioned handle");
  LSA_API_BEGIN
  const int n = h->n;
  cudaStream_t st = h->stream;
  const bool cplx = scalar == LSA_C128;
  const CsrDev& pat = which_matrix == LSA_MAT_A ? h->dA : h->dM;
  const size_t esz = cplx ? 16 : 8;
  void* d_orig = nullptr;
  void* d_perm_vals = nullptr;
  z128* d_a = nullptr;
  try {
    const void* src_vals = vals;
    if (!on_device) {
      LSA_CUDA(cudaMalloc(&d_orig, std::max<size_t>(16, (size_t)pat.nnz * esz)));
      LSA_CUDA(cudaMemcpyAsync(d_orig, vals, (size_t)pat.nnz * esz, cudaMemcpyHostToDevice, st));
      src_vals = d_orig;
    }
    LSA_CUDA(cudaMalloc(&d_perm_vals, std::max<size_t>(16, (size_t)pat.nnz * esz)));
    gather_values(st, src_vals, cplx, pat.src, d_perm_vals, pat.nnz);
    CsrDev tmp = pat;          // same pattern, the caller's values
    tmp.vals = d_perm_vals;
    tmp.is_complex = cplx;
    d_a = dalloc<z128>(n);
    const z128* av = (const z128*)a;
    const z128* vv = (const z128*)v;
    if (!on_device) {
      LSA_CUDA(cudaMemcpyAsync(h->d_io, a, sizeof(z128) * (size_t)n, cudaMemcpyHostToDevice, st));
      permute_gather(st, h->d_io, d_a, h->d_perm, n);
      LSA_CUDA(cudaMemcpyAsync(h->d_io, v, sizeof(z128) * (size_t)n, cudaMemcpyHostToDevice, st));
      permute_gather(st, h->d_io, h->d_x, h->d_perm, n);
    } else {
      permute_gather(st, av, d_a, h->d_perm, n);
      permute_gather(st, vv, h->d_x, h->d_perm, n);
    }
    spmv(*h, tmp, false, h->d_x, h->d_w);
    z128 res = dot_conj(*h, d_a, h->d_w);   // a^H (B v)
    out_c128[0] = res.x;
    out_c128[1] = res.y;
  } catch (...) {
    if (d_orig) cudaFree(d_orig);
    if (d_perm_vals) cudaFree(d_perm_vals);
    if (d_a) cudaFree(d_a);
    throw;
  }
  if (d_orig) cudaFree(d_orig);
  cudaFree(d_perm_vals);
  cudaFree(d_a);
  LSA_API_END(h)
  return LSA_OK;
}

int lsa_eigs(lsa_handle* h, const lsa_eigs_params* p, lsa_eigs_result* out) {
  if (!h || !p || !out || !h->have_values) return LSA_ERR_ARG;
  if (int rc = need_device(h)) return rc;
  const bool needs_factor = p->transform == LSA_ST_SINVERT || h->has_m;
  if (needs_factor && h->scalar < 0) return fail(h, LSA_ERR_ARG, "lsa_eigs needs lsa_factor first");
  if (p->which < 1 || p->which > 9) return fail(h, LSA_ERR_ARG, "unsupported `which`");
  if (std::min(p->ncv, h->n) > 256)
    return fail(h, LSA_ERR_ARG, "ncv > 256: the device orthogonalisation / Rayleigh-Ritz kernels hold at most 256 basis columns");
  if (p->b_mode < 0 || p->b_mode > 2) return fail(h, LSA_ERR_ARG, "b_mode must be 0, 1 or 2");
  if (p->b_mode == 2 && (!h->has_m || h->partitioned || p->adjoint))
    return fail(h, LSA_ERR_ARG, "b_mode = 2 (M-inner products) needs M, a single GPU and the direct problem");
  LSA_API_BEGIN
  const int ncv = std::max(1, std::min(p->ncv, h->n));
  ensure_krylov(*h, ncv);
  if (p->b_mode == 2 && h->U_cols < ncv + 1) {
    dfree(h->d_U);
    h->d_U = dalloc<z128>((size_t)h->n * (size_t)(ncv + 1));
    h->U_cols = ncv + 1;
  }
  if (p->b_mode != 0 && !h->d_mw) {
    h->d_mw = dalloc<z128>((size_t)h->n);
    h->d_bn = dalloc<z128>(2);
  }
  if (p->adjoint || p->refine_steps > 0) ensure_transposes(*h);
  h->last_params = *p;
  h->last_params.v0 = nullptr;
  std::memset(out, 0, sizeof(*out));
  run_eigs(*h, *p, *out);
  LSA_API_END(h)
  return LSA_OK;
}

int lsa_get_eigenvalues(const lsa_handle* h, double* out_c128, int32_t capacity) {
  if (!h || !out_c128) return LSA_ERR_ARG;
  const int cnt = std::min<int>(capacity, h->nconv);
  for (int i = 0; i < cnt; ++i) {
    out_c128[2 * i] = h->eigenvalues[i].x;
    out_c128[2 * i + 1] = h->eigenvalues[i].y;
  }
  return cnt;
}

int lsa_get_eigenvectors(const lsa_handle* hc, double* out_c128, int64_t ld, int32_t count, int32_t on_device) {
  lsa_handle* h = const_cast<lsa_handle*>(hc);
  if (!h || !out_c128 || ld < h->n) return LSA_ERR_ARG;
  if (int rc = need_device(h)) return rc;
  LSA_API_BEGIN
  const int cnt = std::min<int>(count, h->nconv);
  for (int i = 0; i < cnt; ++i) {
    const z128* src = h->d_X + (long long)h->eig_order[i] * h->n;
    LSA_CUDA(cudaMemcpyAsync((z128*)out_c128 + (long long)i * ld, src, sizeof(z128) * (size_t)h->n,
                             on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
  }
  LSA_CUDA(cudaStreamSynchronize(h->stream));
  return cnt;
  LSA_API_END(h)
  return LSA_ERR_INTERNAL;
}

int lsa_get_residuals(lsa_handle* h, double* out, int32_t capacity) {
  if (!h || !out) return LSA_ERR_ARG;
  if (int rc = need_device(h)) return rc;
  LSA_API_BEGIN
  const int cnt = std::min<int>(capacity, h->nconv);
  if (h->last_params.adjoint) ensure_transposes(*h);
  residual_norms(*h, out, cnt);
  return cnt;
  LSA_API_END(h)
  return LSA_ERR_INTERNAL;
}

int lsa_get_counters(const lsa_handle* h, lsa_counters* out) {
  if (!h || !out || !h->analyzed) return LSA_ERR_ARG;
  *out = h->counters;
  const Symbolic& s = h->sym;
  const double esz = h->scalar == LSA_C128 ? 16.0 : 8.0;
  double idx = 0;
  for (const Front& f : s.fronts) idx += 4.0 * (f.k + f.r);
  out->bytes_solve = (double)s.nnz_lu * esz + idx + 2.0 * s.n * 16.0;
  out->bytes_spmv_m = (double)h->nnz_m * ((h->m_complex ? 16.0 : 8.0) + 4.0) + 8.0 * (s.n + 1) + 2.0 * s.n * 16.0;
  out->bytes_spmv_a = (double)h->nnz_a * ((h->a_complex ? 16.0 : 8.0) + 4.0) + 8.0 * (s.n + 1) + 2.0 * s.n * 16.0;
  return LSA_OK;
}

int lsa_host_diag_is_zero(int32_t n, const void* indptr, int32_t indptr_is_64, const int32_t* colidx, const void* vals,
                          int32_t scalar, const int32_t* rows, int32_t nrows, uint8_t* out) {
  if (!indptr || !colidx || !vals || !out || n < 0 || nrows < 0) return LSA_ERR_ARG;
  const int32_t* p32 = (const int32_t*)indptr;
  const int64_t* p64 = (const int64_t*)indptr;
  const int cnt = rows ? nrows : n;
  // a few hundred thousand binary searches: a handful of threads (waking a whole many-core team costs more than the work)
  const int nth = std::max(1, std::min({omp_get_max_threads(), 16, cnt / 4096 + 1}));
#pragma omp parallel for schedule(static) num_threads(nth)
  for (int q = 0; q < cnt; ++q) {
    const int i = rows ? rows[q] : q;
    const long long b = indptr_is_64 ? p64[i] : p32[i], e = indptr_is_64 ? p64[i + 1] : p32[i + 1];
    const int32_t* lo = std::lower_bound(colidx + b, colidx + e, i);
    uint8_t zero = 1;
    if (lo != colidx + e && *lo == i) {
      const long long pos = lo - colidx;
      if (scalar == LSA_C128) zero = ((const double*)vals)[2 * pos] == 0.0 && ((const double*)vals)[2 * pos + 1] == 0.0;
      else zero = ((const double*)vals)[pos] == 0.0;
    }
    out[q] = zero;
  }
  return LSA_OK;
}

int lsa_host_equal(const void* a, const void* b, uint64_t bytes) {
  if (a == b || bytes == 0) return 1;
  if (!a || !b) return 0;
  const uint64_t chunk = 1u << 20;
  const long long nchunks = (long long)((bytes + chunk - 1) / chunk);
  const int nth = (int)std::max<long long>(1, std::min<long long>({(long long)omp_get_max_threads(), 12LL, nchunks / 8 + 1}));
  int differ = 0;
#pragma omp parallel for schedule(static) num_threads(nth) reduction(| : differ)
  for (long long c = 0; c < nchunks; ++c) {
    const uint64_t o = (uint64_t)c * chunk, len = std::min<uint64_t>(chunk, bytes - o);
    if (std::memcmp((const char*)a + o, (const char*)b + o, len) != 0) differ |= 1;
  }
  return differ ? 0 : 1;
}

int lsa_host_alloc(uint64_t bytes, void** ptr) {
  if (!ptr || bytes == 0) return LSA_ERR_ARG;
  *ptr = nullptr;
  return cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault) == cudaSuccess ? LSA_OK : LSA_ERR_CUDA;
}

int lsa_host_free(void* ptr) {
  if (!ptr) return LSA_OK;
  return cudaFreeHost(ptr) == cudaSuccess ? LSA_OK : LSA_ERR_CUDA;
}

int lsa_sync(lsa_handle* h) {
  if (!h) return LSA_ERR_ARG;
  if (int rc = need_device(h)) return rc;
  LSA_API_BEGIN
  LSA_CUDA(cudaStreamSynchronize(h->stream));
  LSA_API_END(h)
  return LSA_OK;
}

int lsa_dense_schur(lsa_handle* h, int32_t m, double* S_c128, int32_t ld, double* Q_c128, int32_t which,
                    int32_t transform, double sigma_re, double sigma_im) {
  if (!h || !S_c128 || !Q_c128 || m < 1 || ld < m) return LSA_ERR_ARG;
  if (int rc = need_device(h)) return rc;
  LSA_API_BEGIN
  z128* dS = dalloc<z128>((size_t)ld * m);
  z128* dQ = dalloc<z128>((size_t)m * m);
  LSA_CUDA(cudaMemcpyAsync(dS, S_c128, sizeof(z128) * (size_t)ld * m, cudaMemcpyHostToDevice, h->stream));
  try {
    dense_schur_device(*h, m, dS, ld, dQ, which, transform, mk(sigma_re, sigma_im));
  } catch (...) {
    cudaFree(dS);
    cudaFree(dQ);
    throw;
  }
  LSA_CUDA(cudaMemcpyAsync(S_c128, dS, sizeof(z128) * (size_t)ld * m, cudaMemcpyDeviceToHost, h->stream));
  LSA_CUDA(cudaMemcpyAsync(Q_c128, dQ, sizeof(z128) * (size_t)m * m, cudaMemcpyDeviceToHost, h->stream));
  LSA_CUDA(cudaStreamSynchronize(h->stream));
  cudaFree(dS);
  cudaFree(dQ);
  LSA_API_END(h)
  return LSA_OK;
}

int lsa_gemm_bench(lsa_handle* h, int32_t scalar, int32_t m, int32_t n, int32_t k, int32_t reps, double* ms,
                   double* max_abs_err) {
  if (!h || m < 1 || n < 1 || k < 1 || reps < 1) return LSA_ERR_ARG;
  if (int rc = need_device(h)) return rc;
  LSA_API_BEGIN
  const bool cplx = scalar == LSA_C128;
  const size_t S = cplx ? 2 : 1;
  cudaStream_t st = h->stream;
  double* A = dalloc<double>((size_t)m * k * S);
  double* B = dalloc<double>((size_t)k * n * S);
  double* C = dalloc<double>((size_t)m * n * S);
  double* C2 = dalloc<double>((size_t)m * n * S);
  double* red = dalloc<double>(256);
  k_fill_random<double><<<cdiv((long long)m * k * S, 256), 256, 0, st>>>(A, (long long)m * k * S, 1);
  k_fill_random<double><<<cdiv((long long)k * n * S, 256), 256, 0, st>>>(B, (long long)k * n * S, 2);
  LSA_CUDA(cudaMemsetAsync(C, 0, sizeof(double) * (size_t)m * n * S, st));
  LSA_CUDA(cudaMemsetAsync(C2, 0, sizeof(double) * (size_t)m * n * S, st));
  double err = -1.0;
  if ((double)m * n * k <= 4.0e9) {
    gemm_plain(st, cplx, A, m, B, k, C, m, m, n, k);
    if (cplx) k_gemm_naive<true><<<dim3(cdiv(m, 128), n), 128, 0, st>>>(A, B, C2, m, n, k);
    else k_gemm_naive<false><<<dim3(cdiv(m, 128), n), 128, 0, st>>>(A, B, C2, m, n, k);
    k_maxdiff<<<256, 256, 0, st>>>(C, C2, (long long)m * n * S, red);
    double hred[256];
    LSA_CUDA(cudaMemcpyAsync(hred, red, sizeof(hred), cudaMemcpyDeviceToHost, st));
    LSA_CUDA(cudaStreamSynchronize(st));
    err = 0;
    for (double v : hred) err = std::max(err, v);
  }
  for (int w = 0; w < 2; ++w) gemm_plain(st, cplx, A, m, B, k, C, m, m, n, k);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int r = 0; r < reps; ++r) gemm_plain(st, cplx, A, m, B, k, C, m, m, n, k);
  cudaEventRecord(e1, st);
  LSA_CUDA(cudaStreamSynchronize(st));
  float t = 0;
  cudaEventElapsedTime(&t, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (ms) *ms = t / reps;
  if (max_abs_err) *max_abs_err = err;
  cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(C2); cudaFree(red);
  LSA_API_END(h)
  return LSA_OK;
}

}  // extern "C"
