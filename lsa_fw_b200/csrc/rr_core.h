// Rayleigh-Ritz step of Krylov-Schur on the small projected matrix, written once for device and
// host: complex Schur form (Householder Hessenberg reduction + single-shift QR with Givens
// rotations), ordering of the Ritz values by `which`, residual estimates and restart bookkeeping.
//
// Stands in for SLEPc's DSSolve / DSSort / DSVectors on a DSNHEP object + EPSKrylovConvergence,
// which run inside `SLEPc.EPS.solve()` (reference call site Solver/utils.py:268-270).  Algorithms
// are the textbook ones (Golub & Van Loan ch. 7; Stewart 2001 for the Krylov-Schur restart).
//
// Execution model: a "team" of nt cooperating threads (one CUDA block, or nt = 1 on the host for
// the CPU unit tests of this file).  `TEAM_SYNC()` is __syncthreads() on the device and a no-op
// on the host; every loop is written `for (i = tid; i < n; i += nt)`.
#pragma once
#include "common.cuh"

namespace lsa {

struct RrParams {
  int m;          // current subspace dimension
  int ld;         // leading dimension of S (>= m + 1)
  int ldq;        // leading dimension of Q
  int nconv;      // locked leading block
  int nev;        // wanted pairs
  int which;      // lsa_which
  int transform;  // lsa_transform
  int last;       // 1: no further restart allowed (iteration limit or breakdown)
  double tol;
  z128 sigma;
  double beta_scale;  // 1.0 normally; 0.0 when the last Arnoldi step broke down (invariant subspace)
};

LSA_HD z128 rr_back(const RrParams& p, z128 theta) {
  if (p.transform == 1) {  // sinvert: lambda = sigma + 1/theta
    if (theta.x == 0.0 && theta.y == 0.0) return mk(INFINITY, 0.0);
    return p.sigma + recip(theta);
  }
  return theta + p.sigma;
}

// smaller key = preferred
LSA_HD double rr_key(const RrParams& p, z128 theta) {
  const z128 l = rr_back(p, theta);
  switch (p.which) {
    case 1: return -absz(l);
    case 2: return absz(l);
    case 3: return -l.x;
    case 4: return l.x;
    case 5: return -l.y;
    case 6: return l.y;
    case 7: return absz(l - p.sigma);
    case 8: return fabs(l.x - p.sigma.x);
    case 9: return fabs(l.y - p.sigma.y);
    default: return -absz(l);
  }
}

// Givens rotation [c s; -conj(s) c] [f; g] = [r; 0], c real.
LSA_HD void rr_givens(z128 f, z128 g, double& c, z128& s) {
  const double ag = absz(g);
  if (ag == 0.0) {
    c = 1.0;
    s = mk(0, 0);
    return;
  }
  const double af = absz(f);
  if (af == 0.0) {
    c = 0.0;
    s = conj_(g) * (1.0 / ag);
    return;
  }
  const double d = hypot(af, ag);
  c = af / d;
  s = (f * (1.0 / af)) * conj_(g) * (1.0 / d);
}

#ifdef __CUDA_ARCH__
#define TEAM_SYNC() __syncthreads()
#else
#define TEAM_SYNC() ((void)0)
#endif

#define S_(i, j) S[(i) + (long long)(j) * ld]
#define Q_(i, j) Q[(i) + (long long)(j) * ldq]

// Workspace (team-shared): rot_c[m], rot_s[m], vec[m] (z128), flags[4] (int), dwork[m] doubles.
struct RrWork {
  double* rot_c;
  z128* rot_s;
  z128* vec;
  double* key;
  int* iflag;
};

// Swap adjacent diagonal entries i, i+1 of the upper triangular S (rows/cols 0..m-1), update Q.
LSA_HD void rr_swap(z128* S, int ld, z128* Q, int ldq, int m, int i, int tid, int nt, RrWork& w) {
  if (tid == 0) {
    const z128 t11 = S_(i, i), t22 = S_(i + 1, i + 1);
    double c;
    z128 s;
    rr_givens(S_(i, i + 1), t22 - t11, c, s);
    w.rot_c[0] = c;
    w.rot_s[0] = s;
  }
  TEAM_SYNC();
  const double c = w.rot_c[0];
  const z128 s = w.rot_s[0];
  for (int j = i + 2 + tid; j < m; j += nt) {
    const z128 a = S_(i, j), b = S_(i + 1, j);
    S_(i, j) = c * a + s * b;
    S_(i + 1, j) = c * b - conj_(s) * a;
  }
  for (int r = tid; r < i; r += nt) {
    const z128 a = S_(r, i), b = S_(r, i + 1);
    S_(r, i) = c * a + conj_(s) * b;
    S_(r, i + 1) = c * b - s * a;
  }
  for (int r = tid; r < m; r += nt) {
    const z128 a = Q_(r, i), b = Q_(r, i + 1);
    Q_(r, i) = c * a + conj_(s) * b;
    Q_(r, i + 1) = c * b - s * a;
  }
  TEAM_SYNC();
  if (tid == 0) {
    const z128 t11 = S_(i, i), t22 = S_(i + 1, i + 1);
    S_(i, i) = t22;
    S_(i + 1, i + 1) = t11;
    S_(i + 1, i) = mk(0, 0);
  }
  TEAM_SYNC();
}

// Schur form of S[lo:m, lo:m] in place (S[0:lo, :] is carried along), Q = accumulated unitary
// (identity on the leading lo x lo block).  Returns 0, or 1 if the QR iteration did not converge.
LSA_HD int rr_schur(z128* S, int ld, z128* Q, int ldq, int m, int lo, int tid, int nt, RrWork& w) {
  // Q = I
  for (int e = tid; e < m * m; e += nt) Q_(e % m, e / m) = (e % m == e / m) ? mk(1, 0) : mk(0, 0);
  TEAM_SYNC();
  // ---- Householder reduction of the active block to upper Hessenberg form
  for (int j = lo; j + 2 < m; ++j) {
    // v = x - alpha e1 with x = S[j+1:m, j]
    if (tid == 0) {
      double tail2 = 0.0;
      for (int i = j + 2; i < m; ++i) tail2 += abs2(S_(i, j));
      const z128 x0 = S_(j + 1, j);
      const double nrm = sqrt(tail2 + abs2(x0));
      if (tail2 == 0.0 || nrm == 0.0) {
        w.iflag[0] = 0;  // nothing to annihilate
      } else {
        const double a0 = absz(x0);
        const z128 phase = a0 > 0.0 ? x0 * (1.0 / a0) : mk(1, 0);
        const z128 alpha = mk(0, 0) - phase * nrm;
        z128 v0 = x0 - alpha;
        double vn2 = abs2(v0) + tail2;
        const double inv = 1.0 / sqrt(vn2);
        w.vec[j + 1] = v0 * inv;
        for (int i = j + 2; i < m; ++i) w.vec[i] = S_(i, j) * inv;
        w.iflag[0] = 1;
        S_(j + 1, j) = alpha;
        for (int i = j + 2; i < m; ++i) S_(i, j) = mk(0, 0);
      }
    }
    TEAM_SYNC();
    if (w.iflag[0]) {
      // left: S[j+1:m, c] -= 2 v (v^H S[j+1:m, c]) for columns c > j
      for (int c = j + 1 + tid; c < m; c += nt) {
        z128 d = mk(0, 0);
        for (int i = j + 1; i < m; ++i) d += conj_(w.vec[i]) * S_(i, c);
        d = d * 2.0;
        for (int i = j + 1; i < m; ++i) S_(i, c) -= w.vec[i] * d;
      }
      TEAM_SYNC();
      // right: S[r, j+1:m] -= 2 (S[r, j+1:m] v) v^H for all rows, same for Q
      for (int r = tid; r < m; r += nt) {
        z128 d = mk(0, 0);
        for (int i = j + 1; i < m; ++i) d += S_(r, i) * w.vec[i];
        d = d * 2.0;
        for (int i = j + 1; i < m; ++i) S_(r, i) -= d * conj_(w.vec[i]);
        z128 q = mk(0, 0);
        for (int i = j + 1; i < m; ++i) q += Q_(r, i) * w.vec[i];
        q = q * 2.0;
        for (int i = j + 1; i < m; ++i) Q_(r, i) -= q * conj_(w.vec[i]);
      }
    }
    TEAM_SYNC();
  }
  // ---- shifted QR iteration with deflation on the window [l, ihi]
  const double eps = 2.220446049250313e-16;
  int ihi = m - 1;
  int its = 0, total = 0;
  const int max_total = 60 * (m - lo + 1);
  while (ihi > lo) {
    // deflation scan (all threads compute the same l from shared data)
    int l = lo;
    for (int i = ihi; i > lo; --i) {
      const double sub = abs1(S_(i, i - 1));
      double sc = abs1(S_(i - 1, i - 1)) + abs1(S_(i, i));
      if (sc == 0.0) sc = 1.0;
      if (sub <= eps * sc) {
        l = i;
        break;
      }
    }
    TEAM_SYNC();
    if (l > lo && tid == 0) S_(l, l - 1) = mk(0, 0);
    if (l == ihi) {
      ihi--;
      its = 0;
      TEAM_SYNC();
      continue;
    }
    if (total >= max_total) return 1;
    its++;
    total++;
    // shift
    z128 mu;
    if (its % 11 == 10) {
      mu = S_(ihi, ihi) + mk(0.75 * abs1(S_(ihi, ihi - 1)), 0.0);  // exceptional shift
    } else {
      // Wilkinson: eigenvalue of trailing 2x2 closest to S[ihi, ihi]
      const z128 a = S_(ihi - 1, ihi - 1), b = S_(ihi - 1, ihi), c = S_(ihi, ihi - 1), d = S_(ihi, ihi);
      const z128 tr2 = (a + d) * 0.5;
      const z128 det = (a - d) * 0.5;
      const z128 disc = det * det + b * c;
      // complex sqrt
      const double ad = absz(disc);
      z128 sq;
      if (ad == 0.0) sq = mk(0, 0);
      else {
        const double re = sqrt(0.5 * (ad + fabs(disc.x)));
        const double im = 0.5 * disc.y / re;
        sq = disc.x >= 0.0 ? mk(re, im) : mk(fabs(im), disc.y >= 0 ? re : -re);
      }
      const z128 e1 = tr2 + sq, e2 = tr2 - sq;
      mu = absz(e1 - d) <= absz(e2 - d) ? e1 : e2;
    }
    TEAM_SYNC();
    for (int i = l + tid; i <= ihi; i += nt) S_(i, i) -= mu;
    TEAM_SYNC();
    // left phase: thread per column, rotation i published by the owner of column i
    for (int i = l; i < ihi; ++i) {
      if ((i % nt) == tid) {
        double c;
        z128 s;
        rr_givens(S_(i, i), S_(i + 1, i), c, s);
        w.rot_c[i] = c;
        w.rot_s[i] = s;
      }
      TEAM_SYNC();
      const double c = w.rot_c[i];
      const z128 s = w.rot_s[i];
      for (int j = i + ((tid - i % nt + nt) % nt); j < m; j += nt) {
        const z128 a = S_(i, j), b = S_(i + 1, j);
        S_(i, j) = c * a + s * b;
        S_(i + 1, j) = (j == i) ? mk(0, 0) : c * b - conj_(s) * a;
      }
    }
    TEAM_SYNC();
    // right phase: thread per row applies all rotations in order (no barrier needed inside)
    for (int r = tid; r <= ihi; r += nt) {
      for (int i = (r - 1 > l ? r - 1 : l); i < ihi; ++i) {
        const double c = w.rot_c[i];
        const z128 s = w.rot_s[i];
        const z128 a = S_(r, i), b = S_(r, i + 1);
        S_(r, i) = c * a + conj_(s) * b;
        S_(r, i + 1) = c * b - s * a;
      }
    }
    for (int r = tid; r < m; r += nt) {
      for (int i = l; i < ihi; ++i) {
        const double c = w.rot_c[i];
        const z128 s = w.rot_s[i];
        const z128 a = Q_(r, i), b = Q_(r, i + 1);
        Q_(r, i) = c * a + conj_(s) * b;
        Q_(r, i + 1) = c * b - s * a;
      }
    }
    TEAM_SYNC();
    for (int i = l + tid; i <= ihi; i += nt) S_(i, i) += mu;
    TEAM_SYNC();
  }
  // clean the strictly lower part of the active block
  for (int e = tid; e < m * m; e += nt) {
    const int i = e % m, j = e / m;
    if (i > j && j >= lo) S_(i, j) = mk(0, 0);
  }
  TEAM_SYNC();
  return 0;
}

// Order the diagonal of the triangular S[lo:m, lo:m] by ascending key.
LSA_HD void rr_sort(z128* S, int ld, z128* Q, int ldq, int m, int lo, const RrParams& p, int tid, int nt, RrWork& w) {
  for (int pos = lo; pos + 1 < m; ++pos) {
    if (tid == 0) {
      int best = pos;
      double kb = rr_key(p, S_(pos, pos));
      for (int j = pos + 1; j < m; ++j) {
        const double kj = rr_key(p, S_(j, j));
        if (kj < kb) {
          kb = kj;
          best = j;
        }
      }
      w.iflag[1] = best;
    }
    TEAM_SYNC();
    const int best = w.iflag[1];
    for (int j = best - 1; j >= pos; --j) rr_swap(S, ld, Q, ldq, m, j, tid, nt, w);
    TEAM_SYNC();
  }
}

// Residual estimate of Ritz pair k: | brow[0:k+1] . y | with y the unit eigenvector of S[0:k+1, 0:k+1]
// for S[k, k].  y is written to `yv` (length >= k+1) when non-null.  Sequential (one thread).
LSA_HD double rr_resid(const z128* S, int ld, int k, const z128* brow, z128* ywork) {
  const z128 tkk = S_(k, k);
  double smax = 0.0;
  for (int i = 0; i <= k; ++i) smax = fmax(smax, abs1(S_(i, i)));
  const double smin = fmax(smax * 2.220446049250313e-16, 1e-300);
  ywork[k] = mk(1, 0);
  for (int i = k - 1; i >= 0; --i) {
    z128 acc = mk(0, 0);
    for (int j = i + 1; j <= k; ++j) acc += S_(i, j) * ywork[j];
    z128 d = S_(i, i) - tkk;
    if (abs1(d) < smin) d = mk(smin, 0);
    ywork[i] = mk(0, 0) - acc / d;
  }
  double nrm2 = 0.0;
  for (int i = 0; i <= k; ++i) nrm2 += abs2(ywork[i]);
  const double inv = 1.0 / sqrt(nrm2);
  z128 dot = mk(0, 0);
  for (int i = 0; i <= k; ++i) {
    ywork[i] = ywork[i] * inv;
    dot += brow[i] * ywork[i];
  }
  return absz(dot);
}

struct RrOut {
  int nconv, keep, status, pad;
};

// Full Rayleigh-Ritz + restart bookkeeping.  On entry S holds the projected matrix (m x m) and
// S[m, m-1] the last Arnoldi norm beta.  On exit S holds the restarted matrix
// (triangular keep x keep block + coupling row `keep`), Q the m x m transformation,
// theta[0:m] the Ritz values in the new order, resid[0:m] residual estimates (-1 where not
// evaluated), brow is scratch of length m, ywork_all scratch of m * m.
LSA_HD void rr_full(z128* S, z128* Q, const RrParams& p, z128* theta, double* resid, z128* brow, z128* ywork_all,
                    RrOut* out, int tid, int nt, RrWork& w) {
  const int m = p.m, ld = p.ld, ldq = p.ldq, lo = p.nconv;
  const z128 beta = S_(m, m - 1) * p.beta_scale;
  const int status = rr_schur(S, ld, Q, ldq, m, lo, tid, nt, w);
  rr_sort(S, ld, Q, ldq, m, lo, p, tid, nt, w);
  for (int j = tid; j < m; j += nt) {
    brow[j] = beta * Q_(m - 1, j);
    theta[j] = S_(j, j);
    resid[j] = -1.0;
  }
  TEAM_SYNC();
  // residual estimates: one thread per candidate (candidates beyond lo + nev + 8 are not needed)
  int ncand = m - lo;
  const int cap = p.nev + 8;
  if (ncand > cap && p.beta_scale != 0.0) ncand = cap;
  for (int q = tid; q < ncand; q += nt) {
    const int k = lo + q;
    resid[k] = rr_resid(S, ld, k, brow, ywork_all + (long long)q * m);
  }
  for (int j = tid; j < lo; j += nt) resid[j] = 0.0;
  TEAM_SYNC();
  if (tid == 0) {
    int k = lo;
    while (k < lo + ncand) {
      const double ref = absz(theta[k]);
      if (resid[k] <= p.tol * ref || p.beta_scale == 0.0) k++;
      else break;
    }
    const int done = (k >= p.nev) || p.last;
    int l = 0;
    if (!done) {
      l = (m - k) / 2;
      if (l < 1) l = 1;
      if (k + l > m) l = m - k;
    }
    out->nconv = k;
    out->keep = k + l;
    out->status = status;
    w.iflag[2] = k;
    w.iflag[3] = k + l;
  }
  TEAM_SYNC();
  const int k = w.iflag[2], keep = w.iflag[3];
  // restarted projected matrix: keep the leading triangle, coupling row at `keep`, zero elsewhere
  for (int e = tid; e < (m + 1) * m; e += nt) {
    const int i = e % (m + 1), j = e / (m + 1);
    if (j >= keep || i > keep) S_(i, j) = mk(0, 0);
    else if (i == keep) S_(i, j) = j < k ? mk(0, 0) : brow[j];
  }
  TEAM_SYNC();
}

#undef S_
#undef Q_

}  // namespace lsa
