"""NumPy emulation of the multifrontal numeric phase, driven by the C++ symbolic structures.

TEST INFRASTRUCTURE (CPU).  It follows the CUDA kernels of lsa_fw_b200/csrc/factor.cu and solve.cu
step by step (scatter -> extend-add -> partial LU with pivoting restricted to the pivot block ->
Schur complement; up/down sweeps for trans = N and trans = H) so that the host-side index maps and the
algebra of the sweeps can be verified against SciPy without a GPU.  Never imported by the package.
"""

from __future__ import annotations

import numpy as np
import scipy.linalg as sla


class Emulator:
    def __init__(self, handle, n: int):
        g = handle.symbolic_array
        self.n = n
        self.perm = g("perm")
        self.sn_ptr = g("sn_ptr")
        self.st_ptr = g("st_ptr")
        self.st_idx = g("st_idx")
        self.ea_map = g("ea_map")
        self.parent = g("parent")
        self.level = g("level")
        self.k = g("front_k")
        self.r = g("front_r")
        self.p_off = g("p_off")
        self.q_off = g("q_off")
        self.a_dst = g("a_dst")
        try:
            self.m_dst = g("m_dst")
        except Exception:
            self.m_dst = None
        info = handle.symbolic_info()
        self.fac_size = info.factor_entries
        self.n_iso = info.n_decoupled
        self.ns = info.n_fronts
        self.diag_off = self.fac_size - ((self.n_iso + 3) // 4) * 4 if self.n_iso else self.fac_size
        self.nlevels = info.n_levels

    # ---- factorisation of alpha A + beta M (values in the caller's CSR entry order)
    def factor(self, a_vals, m_vals, alpha, beta, dtype=np.complex128, pivot_block=128):
        fac = np.zeros(self.fac_size, dtype=dtype)
        fac[self.a_dst] = alpha * a_vals
        if m_vals is not None:
            np.add.at(fac, self.m_dst, beta * m_vals)
        self.fac = fac
        self.max_multiplier = 0.0
        ns = self.ns
        self.cb = [None] * ns
        self.piv = [None] * ns
        children = [[] for _ in range(ns)]
        for s in range(ns):
            if self.parent[s] >= 0:
                children[self.parent[s]].append(s)
        self.children = children
        order = np.argsort(-self.level, kind="stable")
        for s in order:
            k, r = int(self.k[s]), int(self.r[s])
            m = k + r
            P = fac[self.p_off[s]: self.p_off[s] + m * k].reshape((m, k), order="F")
            Q = fac[self.q_off[s]: self.q_off[s] + k * r].reshape((k, r), order="F")
            Cb = np.zeros((r, r), dtype=dtype)
            for c in children[s]:
                mp = self.ea_map[self.st_ptr[c]: self.st_ptr[c + 1]]
                cbc = self.cb[c]
                top = mp < k
                bot = ~top
                P[np.ix_(mp, mp[top])] += cbc[:, top]
                Q[np.ix_(mp[top], mp[bot] - k)] += cbc[np.ix_(top, bot)]
                Cb[np.ix_(mp[bot] - k, mp[bot] - k)] += cbc[np.ix_(bot, bot)]
                self.cb[c] = None
            # partial LU, pivot search restricted to the rows of the current 128-wide outer block of
            # the pivot block (as k_panel_lu does: rows below it still miss deferred updates)
            piv = np.arange(k)
            for j in range(k):
                hi = min(k, (j // pivot_block + 1) * pivot_block)
                p = j + int(np.argmax(np.abs(P[j:hi, j].real) + np.abs(P[j:hi, j].imag)))
                self.max_multiplier = max(self.max_multiplier, float(np.abs(P[j + 1:, j] / P[p, j]).max()) if j + 1 < m else 0.0)
                piv[j] = p
                if p != j:
                    P[[j, p], :] = P[[p, j], :]
                    Q[[j, p], :] = Q[[p, j], :]
                P[j + 1:, j] /= P[j, j]
                P[j + 1:, j + 1:] -= np.outer(P[j + 1:, j], P[j, j + 1:])
                Q[j + 1:, :] -= np.outer(P[j + 1: k, j], Q[j, :])
            Cb -= P[k:, :] @ Q
            self.cb[s] = Cb
            self.piv[s] = piv
        if self.n_iso:
            self.diag = fac[self.diag_off: self.diag_off + self.n_iso]

    def _front(self, s):
        k, r = int(self.k[s]), int(self.r[s])
        m = k + r
        P = self.fac[self.p_off[s]: self.p_off[s] + m * k].reshape((m, k), order="F")
        Q = self.fac[self.q_off[s]: self.q_off[s] + k * r].reshape((k, r), order="F")
        return k, r, P, Q

    def solve(self, b, trans="N"):
        """x = F^-1 b (trans='N') or F^-H b (trans='H'); b in the caller's ordering."""
        x = np.asarray(b, dtype=complex)[self.perm].copy()
        H = trans == "H"
        if self.n_iso:
            d = np.conj(self.diag) if H else self.diag
            x[: self.n_iso] = x[: self.n_iso] / d
        ns = self.ns
        up = np.argsort(-self.level, kind="stable")
        cbv = [None] * ns
        for s in up:
            k, r, P, Q = self._front(s)
            c0 = self.sn_ptr[s]
            bot = np.zeros(r, dtype=complex)
            for c in self.children[s]:
                mp = self.ea_map[self.st_ptr[c]: self.st_ptr[c + 1]]
                top = mp < k
                x[c0 + mp[top]] += cbv[c][top]
                bot[mp[~top] - k] += cbv[c][~top]
                cbv[c] = None
            xt = x[c0: c0 + k].copy()
            if not H:
                for j in range(k):
                    p = self.piv[s][j]
                    if p != j:
                        xt[[j, p]] = xt[[p, j]]
                xt = sla.solve_triangular(P[:k, :k], xt, lower=True, unit_diagonal=True)
                bot -= P[k:, :] @ xt
            else:
                xt = sla.solve_triangular(P[:k, :k].conj().T, xt, lower=True)
                bot -= Q.conj().T @ xt
            x[c0: c0 + k] = xt
            cbv[s] = bot
        for s in up[::-1]:
            k, r, P, Q = self._front(s)
            c0 = self.sn_ptr[s]
            anc = x[self.st_idx[self.st_ptr[s]: self.st_ptr[s + 1]]]
            xt = x[c0: c0 + k]
            if not H:
                xt = sla.solve_triangular(P[:k, :k], xt - Q @ anc, lower=False)
            else:
                xt = sla.solve_triangular(P[:k, :k].conj().T, xt - P[k:, :].conj().T @ anc, lower=False,
                                          unit_diagonal=True)
                for j in range(k - 1, -1, -1):
                    p = self.piv[s][j]
                    if p != j:
                        xt[[j, p]] = xt[[p, j]]
            x[c0: c0 + k] = xt
        out = np.empty_like(x)
        out[self.perm] = x
        return out


class PartEmulator(Emulator):
    """One rank of the PARTITIONED solve (lsa_fw_b200/csrc/partition.cpp), driven by the rank-local symbolic
    structures of a handle created with (rank, world): own sub-trees first, exchange, then the replicated top.
    `comm` provides `bcast(array, root)` and `allreduce(array)` (in place, NumPy) -- torch.distributed (gloo) in
    the CPU tests.  Mirrors factor.cu (cut pool, broadcast of the sub-tree roots' contribution blocks) and
    solve.cu (k_cut_scatter + all-reduce over the replicated rows, gather skipping flagged children)."""

    def __init__(self, handle, n: int, comm):
        super().__init__(handle, n)
        g = handle.symbolic_array
        self.comm = comm
        self.flags = g("front_flags")
        self.col0 = g("front_col0")
        self.owner, self.g2l, self.cut_roots = g("owner"), g("g2l"), g("cut_roots")
        pi = handle.partition_info()
        self.rank, self.world, self.n_top_levels = pi.rank, pi.world, pi.n_top_levels
        self.top_rows = np.concatenate([np.arange(a, b) for a, b in zip(g("top_lo"), g("top_hi"))] or [np.zeros(0, int)]).astype(int)
        self.own_rows = np.concatenate([np.arange(a, b) for a, b in zip(g("own_lo"), g("own_hi"))] or [np.zeros(0, int)]).astype(int)
        self.sn_ptr = self.col0          # col0 of every local front (the global sn_ptr is not contiguous here)
        self.children = [[] for _ in range(self.ns)]
        for s in range(self.ns):
            if self.parent[s] >= 0:
                self.children[self.parent[s]].append(s)

    def _order(self, top: bool):
        idx = [s for s in range(self.ns) if self.flags[s] != 1 and (self.level[s] < self.n_top_levels) == top]
        return sorted(idx, key=lambda s: -self.level[s])      # stable: deepest level first

    def factor(self, a_vals, m_vals, alpha, beta, dtype=np.complex128, pivot_block=128):
        fac = np.zeros(self.fac_size, dtype=dtype)
        ok = self.a_dst >= 0
        fac[self.a_dst[ok]] = alpha * a_vals[ok]
        if m_vals is not None:
            okm = self.m_dst >= 0
            np.add.at(fac, self.m_dst[okm], beta * m_vals[okm])
        self.fac = fac
        self.cb = [None] * self.ns
        self.piv = [None] * self.ns
        self.max_multiplier = 0.0
        for phase_top in (False, True):
            if phase_top:
                # contribution blocks of ALL sub-tree roots: broadcast from their owners (ascending global id)
                for gs in self.cut_roots:
                    l = self.g2l[gs]
                    r = int(self.r[l])
                    buf = self.cb[l] if self.owner[gs] == self.rank else np.zeros((r, r), dtype=dtype)
                    buf = np.ascontiguousarray(buf)
                    self.comm.bcast(buf, int(self.owner[gs]))
                    self.cb[l] = buf
            for s in self._order(phase_top):
                self._factor_front(s, dtype, pivot_block)
        if self.n_iso:
            self.diag = fac[self.diag_off: self.diag_off + self.n_iso]

    def _factor_front(self, s, dtype, pivot_block):
        k, r = int(self.k[s]), int(self.r[s])
        m = k + r
        P = self.fac[self.p_off[s]: self.p_off[s] + m * k].reshape((m, k), order="F")
        Q = self.fac[self.q_off[s]: self.q_off[s] + k * r].reshape((k, r), order="F")
        Cb = np.zeros((r, r), dtype=dtype)
        for c in self.children[s]:
            mp = self.ea_map[self.st_ptr[c]: self.st_ptr[c + 1]]
            cbc = self.cb[c]
            top = mp < k
            bot = ~top
            P[np.ix_(mp, mp[top])] += cbc[:, top]
            Q[np.ix_(mp[top], mp[bot] - k)] += cbc[np.ix_(top, bot)]
            Cb[np.ix_(mp[bot] - k, mp[bot] - k)] += cbc[np.ix_(bot, bot)]
        piv = np.arange(k)
        for j in range(k):
            hi = min(k, (j // pivot_block + 1) * pivot_block)
            p = j + int(np.argmax(np.abs(P[j:hi, j].real) + np.abs(P[j:hi, j].imag)))
            piv[j] = p
            if p != j:
                P[[j, p], :] = P[[p, j], :]
                Q[[j, p], :] = Q[[p, j], :]
            P[j + 1:, j] /= P[j, j]
            P[j + 1:, j + 1:] -= np.outer(P[j + 1:, j], P[j, j + 1:])
            Q[j + 1:, :] -= np.outer(P[j + 1: k, j], Q[j, :])
        Cb -= P[k:, :] @ Q
        self.cb[s] = Cb
        self.piv[s] = piv

    def solve(self, b, trans="N"):
        """x = F^-1 b / F^-H b, b complete on every rank; returns the complete solution (summed over ranks)."""
        x = np.asarray(b, dtype=complex)[self.perm].copy()
        H = trans == "H"
        if self.rank != 0:
            x[self.top_rows] = 0.0          # replicated rows must SUM to the right-hand side over the ranks
        cbv = [None] * self.ns
        for phase_top in (False, True):
            if phase_top:
                for gs in self.cut_roots:   # k_cut_scatter: own sub-tree roots straight into their final rows
                    if self.owner[gs] == self.rank:
                        l = self.g2l[gs]
                        x[self.st_idx[self.st_ptr[l]: self.st_ptr[l + 1]]] += cbv[l]
                buf = np.ascontiguousarray(x[self.top_rows])
                self.comm.allreduce(buf)
                x[self.top_rows] = buf
                if self.n_iso:
                    d = np.conj(self.diag) if H else self.diag
                    x[: self.n_iso] = x[: self.n_iso] / d
            for s in self._order(phase_top):
                k, r, P, Q = self._front(s)
                c0 = self.col0[s]
                bot = np.zeros(r, dtype=complex)
                for c in self.children[s]:
                    if self.flags[c] != 0:
                        continue
                    mp = self.ea_map[self.st_ptr[c]: self.st_ptr[c + 1]]
                    top = mp < k
                    x[c0 + mp[top]] += cbv[c][top]
                    bot[mp[~top] - k] += cbv[c][~top]
                xt = x[c0: c0 + k].copy()
                if not H:
                    for j in range(k):
                        p = self.piv[s][j]
                        if p != j:
                            xt[[j, p]] = xt[[p, j]]
                    xt = sla.solve_triangular(P[:k, :k], xt, lower=True, unit_diagonal=True)
                    bot -= P[k:, :] @ xt
                else:
                    xt = sla.solve_triangular(P[:k, :k].conj().T, xt, lower=True)
                    bot -= Q.conj().T @ xt
                x[c0: c0 + k] = xt
                cbv[s] = bot
        for phase_top in (True, False):
            for s in self._order(phase_top)[::-1]:
                k, r, P, Q = self._front(s)
                c0 = self.col0[s]
                anc = x[self.st_idx[self.st_ptr[s]: self.st_ptr[s + 1]]]
                xt = x[c0: c0 + k]
                if not H:
                    xt = sla.solve_triangular(P[:k, :k], xt - Q @ anc, lower=False)
                else:
                    xt = sla.solve_triangular(P[:k, :k].conj().T, xt - P[k:, :].conj().T @ anc, lower=False,
                                              unit_diagonal=True)
                    for j in range(k - 1, -1, -1):
                        p = self.piv[s][j]
                        if p != j:
                            xt[[j, p]] = xt[[p, j]]
                x[c0: c0 + k] = xt
        full = np.zeros_like(x)
        full[self.own_rows] = x[self.own_rows]
        if self.rank == 0:
            full[self.top_rows] = x[self.top_rows]
        self.comm.allreduce(full)
        out = np.empty_like(full)
        out[self.perm] = full
        return out


class SymEmulator(Emulator):
    """Symmetric factorisation F = L D L^T (option "symmetric": real FP64, diagonal pivots, no Q blocks), following the
    symmetric branches of factor.cu / solve.cu: the scatter maps skip the U12 entries, extend-add reads only the lower
    triangle of a contribution block (mirroring it inside the parent's pivot block) and skips the Q part, the Schur
    complement C -= L21 (D L21^T) is formed on / below the diagonal only (the strict upper triangle is poisoned here to
    prove nothing reads it), a solve is L-sweep, D^-1, L^T-sweep."""

    def factor(self, a_vals, m_vals, alpha, beta, dtype=np.float64, pivot_block=128):
        fac = np.zeros(self.fac_size, dtype=dtype)
        ok = self.a_dst >= 0
        fac[self.a_dst[ok]] = alpha * a_vals[ok]
        if m_vals is not None:
            okm = self.m_dst >= 0
            np.add.at(fac, self.m_dst[okm], beta * m_vals[okm])
        self.fac = fac
        ns = self.ns
        self.cb = [None] * ns
        self.children = [[] for _ in range(ns)]
        for s in range(ns):
            if self.parent[s] >= 0:
                self.children[self.parent[s]].append(s)
        for s in np.argsort(-self.level, kind="stable"):
            k, r = int(self.k[s]), int(self.r[s])
            m = k + r
            P = fac[self.p_off[s]: self.p_off[s] + m * k].reshape((m, k), order="F")
            Cb = np.zeros((r, r), dtype=dtype)
            for c in self.children[s]:
                mp = self.ea_map[self.st_ptr[c]: self.st_ptr[c + 1]]
                assert np.all(np.diff(mp) > 0)          # increasing map: lower triangle lands in the lower triangle
                cbc = np.tril(self.cb[c])               # entries on / below the diagonal only
                assert np.isfinite(cbc).all()
                top = mp < k
                bot = ~top
                P[np.ix_(mp, mp[top])] += cbc[:, top]
                nt = int(top.sum())
                P[np.ix_(mp[top], mp[top])] += np.tril(cbc[:nt, :nt], -1).T      # mirror inside the pivot block
                Cb[np.ix_(mp[bot] - k, mp[bot] - k)] += cbc[np.ix_(bot, bot)]
                self.cb[c] = None
            for j in range(k):                       # no pivoting: the diagonal entry is the pivot
                P[j + 1:, j] /= P[j, j]
                P[j + 1:, j + 1:] -= np.outer(P[j + 1:, j], P[j, j + 1:])
            d = np.diag(P[:k, :k]).copy()
            Cb -= np.tril(P[k:, :] @ (d[:, None] * P[k:, :].T))
            Cb[np.triu_indices(r, 1)] = np.nan
            self.cb[s] = Cb
        if self.n_iso:
            self.diag = fac[self.diag_off: self.diag_off + self.n_iso]

    def solve(self, b, trans="N"):
        x = np.asarray(b, dtype=complex)[self.perm].copy()
        if self.n_iso:
            x[: self.n_iso] = x[: self.n_iso] / self.diag
        up = np.argsort(-self.level, kind="stable")
        cbv = [None] * self.ns
        for s in up:
            k, r, P, _ = self._front_p(s)
            c0 = self.sn_ptr[s]
            bot = np.zeros(r, dtype=complex)
            for c in self.children[s]:
                mp = self.ea_map[self.st_ptr[c]: self.st_ptr[c + 1]]
                top = mp < k
                x[c0 + mp[top]] += cbv[c][top]
                bot[mp[~top] - k] += cbv[c][~top]
            xt = sla.solve_triangular(P[:k, :k], x[c0: c0 + k], lower=True, unit_diagonal=True)
            bot -= P[k:, :] @ xt
            x[c0: c0 + k] = xt / np.diag(P[:k, :k])
            cbv[s] = bot
        for s in up[::-1]:
            k, r, P, _ = self._front_p(s)
            c0 = self.sn_ptr[s]
            anc = x[self.st_idx[self.st_ptr[s]: self.st_ptr[s + 1]]]
            x[c0: c0 + k] = sla.solve_triangular(P[:k, :k].T, x[c0: c0 + k] - P[k:, :].T @ anc, lower=False, unit_diagonal=True)
        out = np.empty_like(x)
        out[self.perm] = x
        return out

    def _front_p(self, s):
        k, r = int(self.k[s]), int(self.r[s])
        m = k + r
        return k, r, self.fac[self.p_off[s]: self.p_off[s] + m * k].reshape((m, k), order="F"), None
