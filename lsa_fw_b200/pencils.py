"""Synthetic linearised Navier-Stokes pencils (A, M) on structured simplex meshes.

The reference assembles (A, M) with dolfinx (`FEM/operators.py:449-509`); dolfinx is not
available where this package runs, so the benchmark workloads of BASELINE.json are produced
by this small, self-contained finite-element assembler.  It reproduces the *structure* the
eigensolve has to digest:

* forms of `FEM/operators.py:461-481` and `:502`
    A = -((u.grad)U, v) - ((U.grad)u, v) - (1/Re)(grad u, grad v) + (p, div v) + (q, div u)
    M = (u, v)                                   (velocity block only -> M is singular)
* Taylor-Hood P2/P1, "SIMPLE" P1/P1 and MINI (P1+bubble)/P1 spaces (`FEM/spaces.py:103-179`)
* mixed DOF numbering with velocity and pressure interleaved node by node
* Dirichlet rows AND columns zeroed with a unit diagonal in both A and M
  (`FEM/operators.py:483-486,504-507`), i.e. a spurious eigenvalue 1 per Dirichlet DOF
* optional pinned pressure DOF (`FEM/utils.py:596-602`) for enclosed flows.

Meshes: `nx x ny` rectangles split into 2 triangles, or `nx x ny x nz` bricks split into
6 Kuhn tetrahedra.  On those meshes the P2 nodes are exactly the points of the twice-refined
grid, which gives the DOF formulas quoted in SURVEY.md section 8d.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from itertools import permutations
from typing import Callable

import numpy as np
import scipy.sparse as sp
from scipy.special import roots_jacobi

BaseFlow = Callable[[np.ndarray], tuple[np.ndarray, np.ndarray]]
"""x (npts, dim) -> (U (npts, dim), gradU (npts, dim, dim) with gradU[:, i, j] = dU_i/dx_j)."""


@dataclass
class Pencil:
    """Assembled generalized eigenproblem A x = lambda M x (CSR, float64, int32 indices)."""

    A: sp.csr_matrix
    M: sp.csr_matrix
    dofs_u: np.ndarray
    dofs_p: np.ndarray
    dirichlet: np.ndarray
    coords: np.ndarray
    """(n, dim) coordinates of every DOF (geometric hint for the ordering; optional to use)."""
    meta: dict = field(default_factory=dict)
    a_visc: np.ndarray | None = None
    """dA/d(1/Re) on A's pattern (same entry order as `A.data`), filled with `split_viscous=True`: the viscous
    term is the only Reynolds-dependent one for a fixed base flow, so a Reynolds sweep on one mesh
    (`.examples/eigenvalues.py:61-108`, BASELINE config 3) re-uses pattern AND assembly."""

    @property
    def n(self) -> int:
        return self.A.shape[0]

    def a_data_at(self, re: float) -> np.ndarray:
        """`A.data` of the same pencil at Reynolds number `re` (same base flow, same pattern)."""
        if self.a_visc is None:
            raise ValueError("assemble with split_viscous=True to sweep the Reynolds number")
        return self.A.data + (1.0 / re - 1.0 / self.meta["re"]) * self.a_visc


# --------------------------------------------------------------------------- quadrature


def simplex_quadrature(dim: int, q: int) -> tuple[np.ndarray, np.ndarray]:
    """Collapsed Gauss-Jacobi rule on the unit simplex: barycentric points and weights.

    Exact for total degree 2q-1.  Returns (lam (npts, dim+1), w (npts,)), sum(w) = 1/dim!.
    """
    pts = [np.zeros(0)] * dim
    wts = [np.zeros(0)] * dim
    for a in range(dim):
        x, w = roots_jacobi(q, dim - 1 - a, 0.0)
        pts[a] = 0.5 * (x + 1.0)
        wts[a] = w / 2.0 ** (dim - a)
    grids = np.meshgrid(*pts, indexing="ij")
    wgrid = np.meshgrid(*wts, indexing="ij")
    t = np.stack([g.ravel() for g in grids], axis=1)
    w = np.prod(np.stack([g.ravel() for g in wgrid], axis=1), axis=1)
    # Duffy map: x_0 = t_0, x_1 = t_1 (1 - t_0), x_2 = t_2 (1 - t_0)(1 - t_1)
    x = np.zeros_like(t)
    rem = np.ones(len(t))
    for a in range(dim):
        x[:, a] = t[:, a] * rem
        rem = rem * (1.0 - t[:, a])
    lam = np.concatenate([1.0 - x.sum(axis=1, keepdims=True), x], axis=1)
    return lam, w


# --------------------------------------------------------------------------- base flows


def wake_profile(a: float = 0.6, b: float = 2.0, y0: float = 0.0) -> BaseFlow:
    """Parallel wake U = (1 - a sech^2(b (y - y0)), 0[, 0]) -- cylinder-wake surrogate."""

    def flow(x: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        n, dim = x.shape
        s = 1.0 / np.cosh(b * (x[:, 1] - y0))
        U = np.zeros((n, dim))
        G = np.zeros((n, dim, dim))
        U[:, 0] = 1.0 - a * s * s
        G[:, 0, 1] = 2.0 * a * b * s * s * np.tanh(b * (x[:, 1] - y0))
        return U, G

    return flow


def step_profile(h: float = 0.5, umax: float = 1.0) -> BaseFlow:
    """Shear layer behind a backward-facing step: U_x = umax/2 (1 + tanh((y - h)/delta(x)))."""

    def flow(x: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        n, dim = x.shape
        delta = 0.05 + 0.02 * x[:, 0]
        arg = (x[:, 1] - h) / delta
        th = np.tanh(arg)
        sech2 = 1.0 - th * th
        U = np.zeros((n, dim))
        G = np.zeros((n, dim, dim))
        U[:, 0] = 0.5 * umax * (1.0 + th)
        G[:, 0, 1] = 0.5 * umax * sech2 / delta
        G[:, 0, 0] = 0.5 * umax * sech2 * (-arg / delta) * 0.02
        return U, G

    return flow


def cavity_vortex() -> BaseFlow:
    """Smooth divergence-free single vortex in the unit box (lid-driven-cavity surrogate)."""

    def flow(x: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        n, dim = x.shape
        X, Z = x[:, 0], x[:, dim - 1]
        sx, cx = np.sin(np.pi * X), np.cos(np.pi * X)
        sz, cz = np.sin(np.pi * Z), np.cos(np.pi * Z)
        U = np.zeros((n, dim))
        G = np.zeros((n, dim, dim))
        # stream function psi = sin^2(pi x) sin^2(pi z)/pi ; u = dpsi/dz, w = -dpsi/dx
        U[:, 0] = 2.0 * sx * sx * sz * cz
        U[:, dim - 1] = -2.0 * sx * cx * sz * sz
        G[:, 0, 0] = 4.0 * np.pi * sx * cx * sz * cz
        G[:, 0, dim - 1] = 2.0 * np.pi * sx * sx * (cz * cz - sz * sz)
        G[:, dim - 1, 0] = -2.0 * np.pi * (cx * cx - sx * sx) * sz * sz
        G[:, dim - 1, dim - 1] = -4.0 * np.pi * sx * cx * sz * cz
        return U, G

    return flow


def zero_flow() -> BaseFlow:
    def flow(x: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
        n, dim = x.shape
        return np.zeros((n, dim)), np.zeros((n, dim, dim))

    return flow


# --------------------------------------------------------------------------- mesh


def _simplices(shape: tuple[int, ...]) -> np.ndarray:
    """Vertices of all simplices in fine-grid integer coordinates: (ncell, dim+1, dim)."""
    dim = len(shape)
    idx = np.stack(
        np.meshgrid(*[np.arange(s) for s in shape], indexing="ij"), axis=-1
    ).reshape(-1, dim)
    out = []
    for perm in permutations(range(dim)):
        verts = np.zeros((dim + 1, dim), dtype=np.int64)
        for step, axis in enumerate(perm):
            verts[step + 1] = verts[step]
            verts[step + 1, axis] += 1
        out.append(2 * (idx[:, None, :] + verts[None, :, :]))
    return np.concatenate(out, axis=0)


def _grid_spacing(n: int, length: float, grading: float) -> np.ndarray:
    """Fine-grid node coordinates (2n+1 of them); grading > 1 clusters cells near the centre."""
    t = np.linspace(-1.0, 1.0, n + 1)
    if grading != 1.0:
        t = np.sign(t) * np.abs(t) ** grading
    v = 0.5 * (t + 1.0) * length
    fine = np.zeros(2 * n + 1)
    fine[0::2] = v
    fine[1::2] = 0.5 * (v[:-1] + v[1:])
    return fine


# --------------------------------------------------------------------------- assembly


def assemble_pencil(
    shape: tuple[int, ...],
    lengths: tuple[float, ...],
    *,
    re: float,
    baseflow: BaseFlow | None = None,
    space: str = "TH",
    dirichlet_faces: tuple[str, ...] | None = None,
    pin_pressure: bool = False,
    grading: tuple[float, ...] | None = None,
    quad_order: int = 4,
    chunk: int = 200_000,
    split_viscous: bool = False,
) -> Pencil:
    """Assemble (A, M) for the linearised Navier-Stokes operator around `baseflow`.

    `dirichlet_faces`: names out of {"x0","x1","y0","y1","z0","z1"} on which all velocity
    components are homogeneous-Dirichlet (default: everything except the outlet "x1").
    """
    dim = len(shape)
    if dim not in (2, 3):
        raise ValueError("only 2-D and 3-D meshes are supported")
    space = space.upper()
    if space not in ("TH", "SIMPLE", "MINI"):
        raise ValueError(f"unknown space {space!r}; choose TH, SIMPLE or MINI")
    baseflow = baseflow or zero_flow()
    grading = grading or (1.0,) * dim
    if dirichlet_faces is None:
        dirichlet_faces = tuple(
            f for f in ("x0", "y0", "y1", "z0", "z1")[: 2 * dim - 1] if f[0] in "xyz"[:dim]
        )

    fine_shape = tuple(2 * s + 1 for s in shape)
    axes = [_grid_spacing(s, l, g) for s, l, g in zip(shape, lengths, grading)]
    strides = np.array(
        [int(np.prod(fine_shape[a + 1 :])) for a in range(dim)], dtype=np.int64
    )
    nfine = int(np.prod(fine_shape))
    fine_idx = np.stack(
        np.meshgrid(*[np.arange(s) for s in fine_shape], indexing="ij"), axis=-1
    ).reshape(-1, dim)
    is_vertex = np.all(fine_idx % 2 == 0, axis=1)

    cells = _simplices(shape)  # (nc, dim+1, dim) even fine coords
    ncell = cells.shape[0]

    # ---- DOF numbering: node-interleaved [u_0..u_{d-1}, (p)] per fine node
    has_vel = np.ones(nfine, dtype=bool) if space == "TH" else is_vertex.copy()
    per_node = dim * has_vel.astype(np.int64) + is_vertex.astype(np.int64)
    node_off = np.concatenate([[0], np.cumsum(per_node)])
    n_nodal = int(node_off[-1])
    n_bubble = dim * ncell if space == "MINI" else 0
    n = n_nodal + n_bubble

    def vel_dof(node: np.ndarray, comp: int) -> np.ndarray:
        return node_off[node] + comp

    def prs_dof(node: np.ndarray) -> np.ndarray:
        return node_off[node] + dim * has_vel[node]

    # ---- local scalar bases on the reference simplex
    lam, wq = simplex_quadrature(dim, quad_order)
    nq = len(wq)
    nv = dim + 1
    pairs = [(i, i) for i in range(nv)]
    if space == "TH":
        pairs += [(i, j) for i in range(nv) for j in range(i + 1, nv)]
    # velocity scalar basis values phi (nq, nb) and barycentric derivatives dphi/dlam (nq, nb, nv)
    nb_nodal = len(pairs)
    nbv = nb_nodal + (1 if space == "MINI" else 0)
    phi = np.zeros((nq, nbv))
    dphi = np.zeros((nq, nbv, nv))
    for b, (i, j) in enumerate(pairs):
        if space == "TH":
            if i == j:
                phi[:, b] = lam[:, i] * (2.0 * lam[:, i] - 1.0)
                dphi[:, b, i] = 4.0 * lam[:, i] - 1.0
            else:
                phi[:, b] = 4.0 * lam[:, i] * lam[:, j]
                dphi[:, b, i] = 4.0 * lam[:, j]
                dphi[:, b, j] = 4.0 * lam[:, i]
        else:
            phi[:, b] = lam[:, i]
            dphi[:, b, i] = 1.0
    if space == "MINI":
        scale = float(nv) ** nv
        prod = np.prod(lam, axis=1)
        phi[:, -1] = scale * prod
        for i in range(nv):
            dphi[:, -1, i] = scale * np.prod(np.delete(lam, i, axis=1), axis=1)
    psi = lam  # pressure P1 basis (nq, nv)

    # reference-element tensors of the element-matrix GEMMs below
    ref_PP = np.einsum("qa,qb->qab", phi, phi).reshape(nq, nbv * nbv)
    ref_S = np.einsum("q,qai,qbj->ijab", wq, dphi, dphi).reshape(nv * nv, nbv * nbv)
    ref_T = np.einsum("qa,qbi->qiab", phi, dphi).reshape(nq * nv, nbv * nbv)
    ref_R = np.einsum("q,qc,qai->iac", wq, psi, dphi)

    rows_A: list[np.ndarray] = []
    cols_A: list[np.ndarray] = []
    vals_A: list[np.ndarray] = []
    rows_M: list[np.ndarray] = []
    cols_M: list[np.ndarray] = []
    vals_M: list[np.ndarray] = []
    rows_K: list[np.ndarray] = []
    cols_K: list[np.ndarray] = []
    vals_K: list[np.ndarray] = []

    for c0 in range(0, ncell, chunk):
        cv = cells[c0 : c0 + chunk]
        ne = cv.shape[0]
        # physical vertex coordinates (ne, nv, dim)
        X = np.stack([axes[a][cv[:, :, a]] for a in range(dim)], axis=-1)
        # affine map: x = X0 + J t, J[:, :, a] = X_{a+1} - X_0
        J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))
        detJ = np.abs(np.linalg.det(J))
        Jinv = np.linalg.inv(J)
        # grad lam_i: rows of Jinv for i>=1, minus their sum for i = 0  -> (ne, nv, dim)
        glam = np.concatenate([-Jinv.sum(axis=1, keepdims=True), Jinv], axis=1)
        xq = np.einsum("qi,eid->eqd", lam, X)
        U, GU = baseflow(xq.reshape(-1, dim))
        U = U.reshape(ne, nq, dim)
        GU = GU.reshape(ne, nq, dim, dim)
        w = wq[None, :] * detJ[:, None]  # (ne, nq)

        # Element matrices as GEMMs against reference-element tensors (the elements are affine, so the basis
        # gradients are dphi/dlam contracted with the constant grad lam of the element):
        #   mass  = sum_q w phi_a phi_b                         = w (ne x nq) @ PP (nq x nb^2)
        #   stiff = sum_q w grad phi_a . grad phi_b             = G (ne x nv^2) @ S (nv^2 x nb^2)
        #   conv  = sum_q w phi_a (U . grad phi_b)              = Cq (ne x nq nv) @ T (nq nv x nb^2)
        #   shear = sum_q w phi_a phi_b dU_i/dx_j               = WG (ne dim^2 x nq) @ PP
        #   grad_p= sum_q w psi_c d phi_a / dx_d                = detJ glam . R
        mass = (w @ ref_PP).reshape(ne, nbv, nbv)
        G = np.einsum("e,eid,ejd->eij", detJ, glam, glam).reshape(ne, nv * nv)
        stiff = (G @ ref_S).reshape(ne, nbv, nbv)
        Cq = np.einsum("eq,eqd,eid->eqi", w, U, glam, optimize=True).reshape(ne, nq * nv)
        conv = (Cq @ ref_T).reshape(ne, nbv, nbv)
        WG = np.ascontiguousarray((w[:, :, None, None] * GU).transpose(0, 2, 3, 1)).reshape(ne * dim * dim, nq)
        shear = np.ascontiguousarray((WG @ ref_PP).reshape(ne, dim, dim, nbv, nbv).transpose(0, 3, 1, 4, 2))
        grad_p = np.einsum("e,eid,iac->eadc", detJ, glam, ref_R, optimize=True)  # (ne, nbv, dim, nv)

        # global DOF ids
        vert_node = (cv * strides[None, None, :]).sum(axis=2)  # (ne, nv)
        node = np.zeros((ne, nbv), dtype=np.int64)
        for b, (i, j) in enumerate(pairs):
            node[:, b] = (((cv[:, i, :] + cv[:, j, :]) // 2) * strides[None, :]).sum(axis=1)
        vd = np.zeros((ne, nbv, dim), dtype=np.int64)
        for comp in range(dim):
            vd[:, :nb_nodal, comp] = vel_dof(node[:, :nb_nodal], comp)
            if space == "MINI":
                vd[:, -1, comp] = n_nodal + dim * (c0 + np.arange(ne)) + comp
        pd = prs_dof(vert_node)  # (ne, nv)

        # velocity-velocity block
        Kuu = -shear
        diag_block = -conv - stiff / re
        for comp in range(dim):
            Kuu[:, :, comp, :, comp] += diag_block
        r = np.broadcast_to(vd[:, :, :, None, None], Kuu.shape)
        c = np.broadcast_to(vd[:, None, None, :, :], Kuu.shape)
        rows_A.append(r.ravel()); cols_A.append(c.ravel()); vals_A.append(Kuu.ravel())
        if split_viscous:
            for comp in range(dim):
                rows_K.append(np.broadcast_to(vd[:, :, comp, None], stiff.shape).ravel())
                cols_K.append(np.broadcast_to(vd[:, None, :, comp], stiff.shape).ravel())
                vals_K.append(-stiff.ravel())
        # pressure gradient (p, div v) and divergence (q, div u)
        r = np.broadcast_to(vd[:, :, :, None], grad_p.shape)
        c = np.broadcast_to(pd[:, None, None, :], grad_p.shape)
        rows_A.append(r.ravel()); cols_A.append(c.ravel()); vals_A.append(grad_p.ravel())
        rows_A.append(c.ravel()); cols_A.append(r.ravel()); vals_A.append(grad_p.ravel())
        # mass
        Muu = np.zeros_like(Kuu)
        for comp in range(dim):
            Muu[:, :, comp, :, comp] = mass
        keep = np.zeros(Kuu.shape[1:], dtype=bool)
        for comp in range(dim):
            keep[:, comp, :, comp] = True
        r = np.broadcast_to(vd[:, :, :, None, None], Kuu.shape)[:, keep]
        c = np.broadcast_to(vd[:, None, None, :, :], Kuu.shape)[:, keep]
        rows_M.append(r.ravel()); cols_M.append(c.ravel()); vals_M.append(Muu[:, keep].ravel())

    A = sp.coo_matrix(
        (np.concatenate(vals_A), (np.concatenate(rows_A), np.concatenate(cols_A))), shape=(n, n)
    ).tocsr()
    M = sp.coo_matrix(
        (np.concatenate(vals_M), (np.concatenate(rows_M), np.concatenate(cols_M))), shape=(n, n)
    ).tocsr()
    Kv = None
    if split_viscous:
        Kv = sp.coo_matrix(
            (np.concatenate(vals_K), (np.concatenate(rows_K), np.concatenate(cols_K))), shape=(n, n)
        ).tocsr()
    del rows_A, cols_A, vals_A, rows_M, cols_M, vals_M, rows_K, cols_K, vals_K

    # ---- DOF bookkeeping
    vel_nodes = np.nonzero(has_vel)[0]
    dofs_u = (node_off[vel_nodes][:, None] + np.arange(dim)[None, :]).ravel()
    if n_bubble:
        dofs_u = np.concatenate([dofs_u, np.arange(n_nodal, n)])
    vtx_nodes = np.nonzero(is_vertex)[0]
    dofs_p = prs_dof(vtx_nodes)
    coords = np.zeros((n, dim))
    node_xyz = np.stack([axes[a][fine_idx[:, a]] for a in range(dim)], axis=1)
    for comp in range(dim):
        coords[vel_dof(vel_nodes, comp)] = node_xyz[vel_nodes]
    coords[dofs_p] = node_xyz[vtx_nodes]
    if n_bubble:
        cent = np.stack([axes[a][cells[:, :, a]].mean(axis=1) for a in range(dim)], axis=1)
        coords[n_nodal:] = np.repeat(cent, dim, axis=0)

    # ---- Dirichlet rows/columns: identity in both A and M
    on_face = np.zeros(nfine, dtype=bool)
    for f in dirichlet_faces:
        a = "xyz".index(f[0])
        on_face |= fine_idx[:, a] == (0 if f[1] == "0" else fine_shape[a] - 1)
    bc_nodes = np.nonzero(on_face & has_vel)[0]
    bc = (node_off[bc_nodes][:, None] + np.arange(dim)[None, :]).ravel()
    pinned = np.array([dofs_p[0]], dtype=np.int64) if pin_pressure else np.zeros(0, np.int64)
    A = _apply_identity_rows_cols(A, np.concatenate([bc, pinned]))
    M = _apply_identity_rows_cols(M, bc)
    if pin_pressure:
        # pin_dof is applied to A only in the reference (FEM/utils.py:596-602); M row stays zero
        pass

    A = _canonical(A)
    a_visc = None
    if Kv is not None:
        # the viscous part on A's pattern: Dirichlet rows / columns dropped, entries located by (row, col) key
        drop = np.zeros(n, dtype=bool)
        drop[np.concatenate([bc, pinned])] = True
        Kc = Kv.tocoo()
        keep = ~(drop[Kc.row] | drop[Kc.col])
        key_a = np.repeat(np.arange(n, dtype=np.int64), np.diff(A.indptr)) * n + A.indices
        key_k = Kc.row[keep].astype(np.int64) * n + Kc.col[keep]
        pos = np.searchsorted(key_a, key_k)
        if not np.array_equal(key_a[pos], key_k):
            raise RuntimeError("viscous entries outside the pattern of A")
        a_visc = np.zeros(A.nnz)
        a_visc[pos] = Kc.data[keep]
        del Kc, key_a, key_k, pos
    meta = dict(
        dim=dim, shape=tuple(shape), lengths=tuple(lengths), re=re, space=space,
        n_u=len(dofs_u), n_p=len(dofs_p), n_dirichlet=len(bc), pinned=pinned.tolist(),
    )
    return Pencil(
        A=A, M=_canonical(M), dofs_u=dofs_u.astype(np.int64),
        dofs_p=dofs_p.astype(np.int64), dirichlet=bc.astype(np.int64), coords=coords, meta=meta, a_visc=a_visc,
    )


def _apply_identity_rows_cols(mat: sp.csr_matrix, dofs: np.ndarray, diagonal: float = 1.0) -> sp.csr_matrix:
    """Rows and columns `dofs` removed from the pattern and replaced by a unit diagonal
    (`FEM/operators.py:483-486`).  Purely structural: no entry is dropped because of its VALUE, so the pattern
    is the same for every Reynolds number / base flow (one symbolic analysis per mesh)."""
    if len(dofs) == 0:
        return mat
    n = mat.shape[0]
    drop = np.zeros(n, dtype=bool)
    drop[dofs] = True
    coo = mat.tocoo()
    keep = ~(drop[coo.row] | drop[coo.col])
    rows = np.concatenate([coo.row[keep], dofs])
    cols = np.concatenate([coo.col[keep], dofs])
    vals = np.concatenate([coo.data[keep], np.full(len(dofs), float(diagonal))])
    return sp.csr_matrix((vals, (rows, cols)), shape=(n, n))


def _canonical(mat: sp.csr_matrix) -> sp.csr_matrix:
    mat = mat.tocsr()
    mat.sum_duplicates()
    mat.sort_indices()
    return sp.csr_matrix(
        (mat.data.astype(np.float64), mat.indices.astype(np.int32), mat.indptr.astype(np.int32)),
        shape=mat.shape,
    )


# --------------------------------------------------------------------------- named workloads


def th_dofs(shape: tuple[int, ...]) -> int:
    """DOF count of the Taylor-Hood pencil (SURVEY.md section 8d formulas)."""
    dim = len(shape)
    return dim * int(np.prod([2 * s + 1 for s in shape])) + int(np.prod([s + 1 for s in shape]))


def cylinder_wake_2d(nx: int = 110, ny: int = 50, re: float = 50.0, **kw) -> Pencil:
    """BASELINE config 1 surrogate: 2-D channel with wake profile (110x50 -> 50 303 DOFs)."""
    ly = 10.0
    return assemble_pencil(
        (nx, ny), (25.0, ly), re=re, baseflow=wake_profile(0.9, 1.2, ly / 2), **kw
    )


def backward_step_2d(nx: int = 667, ny: int = 167, re: float = 500.0, **kw) -> Pencil:
    """BASELINE config 2 surrogate: 2-D channel with a step shear layer (~1.0 M DOFs)."""
    return assemble_pencil((nx, ny), (20.0, 1.0), re=re, baseflow=step_profile(0.5), **kw)


def adapted_wake_2d(nx: int = 1155, ny: int = 289, re: float = 100.0, **kw) -> Pencil:
    """BASELINE config 3 surrogate: graded ("adapted") wake mesh (~3.0 M DOFs)."""
    ly = 10.0
    kw.setdefault("grading", (1.0, 1.6))
    return assemble_pencil(
        (nx, ny), (25.0, ly), re=re, baseflow=wake_profile(0.9, 1.2, ly / 2), **kw
    )


def cavity_3d(n: int = 54, re: float = 100.0, **kw) -> Pencil:
    """BASELINE config 4 surrogate: unit-cube lid-driven cavity, pressure pinned."""
    kw.setdefault("dirichlet_faces", ("x0", "x1", "y0", "y1", "z0", "z1"))
    kw.setdefault("pin_pressure", True)
    return assemble_pencil((n, n, n), (1.0, 1.0, 1.0), re=re, baseflow=cavity_vortex(), **kw)


def cylinder_wake_3d(n: int = 74, re: float = 300.0, **kw) -> Pencil:
    """BASELINE config 5 surrogate: 3-D box with wake profile."""
    return assemble_pencil(
        (n, n, n), (12.0, 6.0, 6.0), re=re, baseflow=wake_profile(0.9, 1.5, 3.0), **kw
    )


def membrane_pencil(nx: int, ny: int, a: float = 1.0, b: float = 1.0, m_bc_diagonal: float = 1.0) -> Pencil:
    """Vibrating membrane: P2 Laplace eigenproblem K x = lambda M x on an a x b rectangle with
    homogeneous Dirichlet boundary (reference `tests/benchmark/vibrating_membrane.py:130-173`):
    analytic spectrum pi^2 (m^2/a^2 + n^2/b^2); Dirichlet rows are identity in K and M, which adds
    the spurious eigenvalue 1 the reference filters out (`:169-173`); `m_bc_diagonal` (dolfinx's `diagonal=` of the
    mass assembly) moves it to 1 / m_bc_diagonal."""
    shape = (nx, ny)
    dim = 2
    fine_shape = (2 * nx + 1, 2 * ny + 1)
    axes = [_grid_spacing(nx, a, 1.0), _grid_spacing(ny, b, 1.0)]
    strides = np.array([fine_shape[1], 1], dtype=np.int64)
    n = fine_shape[0] * fine_shape[1]
    cells = _simplices(shape)
    lam, wq = simplex_quadrature(dim, 4)
    nv = 3
    pairs = [(i, i) for i in range(nv)] + [(i, j) for i in range(nv) for j in range(i + 1, nv)]
    nq = len(wq)
    phi = np.zeros((nq, 6))
    dphi = np.zeros((nq, 6, nv))
    for k, (i, j) in enumerate(pairs):
        if i == j:
            phi[:, k] = lam[:, i] * (2.0 * lam[:, i] - 1.0)
            dphi[:, k, i] = 4.0 * lam[:, i] - 1.0
        else:
            phi[:, k] = 4.0 * lam[:, i] * lam[:, j]
            dphi[:, k, i] = 4.0 * lam[:, j]
            dphi[:, k, j] = 4.0 * lam[:, i]
    X = np.stack([axes[d][cells[:, :, d]] for d in range(dim)], axis=-1)
    J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))
    detJ = np.abs(np.linalg.det(J))
    Jinv = np.linalg.inv(J)
    glam = np.concatenate([-Jinv.sum(axis=1, keepdims=True), Jinv], axis=1)
    gphi = np.einsum("qbi,eid->eqbd", dphi, glam)
    w = wq[None, :] * detJ[:, None]
    mass = np.einsum("eq,qa,qb->eab", w, phi, phi)
    stiff = np.einsum("eq,eqad,eqbd->eab", w, gphi, gphi)
    node = np.zeros((cells.shape[0], 6), dtype=np.int64)
    for k, (i, j) in enumerate(pairs):
        node[:, k] = (((cells[:, i, :] + cells[:, j, :]) // 2) * strides[None, :]).sum(axis=1)
    r = np.broadcast_to(node[:, :, None], mass.shape).ravel()
    c = np.broadcast_to(node[:, None, :], mass.shape).ravel()
    K = sp.coo_matrix((stiff.ravel(), (r, c)), shape=(n, n)).tocsr()
    M = sp.coo_matrix((mass.ravel(), (r, c)), shape=(n, n)).tocsr()
    idx = np.stack(np.meshgrid(np.arange(fine_shape[0]), np.arange(fine_shape[1]), indexing="ij"), -1).reshape(-1, 2)
    on_bnd = (idx[:, 0] == 0) | (idx[:, 0] == fine_shape[0] - 1) | (idx[:, 1] == 0) | (idx[:, 1] == fine_shape[1] - 1)
    bc = np.nonzero(on_bnd)[0]
    K = _apply_identity_rows_cols(K, bc)
    M = _apply_identity_rows_cols(M, bc, m_bc_diagonal)
    coords = np.stack([axes[0][idx[:, 0]], axes[1][idx[:, 1]]], axis=1)
    return Pencil(A=_canonical(K), M=_canonical(M), dofs_u=np.arange(n), dofs_p=np.zeros(0, np.int64),
                  dirichlet=bc.astype(np.int64), coords=coords, meta=dict(kind="membrane", a=a, b=b, shape=shape))


def membrane_analytic(count: int, a: float = 1.0, b: float = 1.0) -> np.ndarray:
    """First `count` analytic eigenvalues pi^2 (m^2/a^2 + n^2/b^2), m, n >= 1, ascending."""
    vals = sorted(np.pi**2 * (m * m / a**2 + n * n / b**2) for m in range(1, 40) for n in range(1, 40))
    return np.array(vals[:count])


def elasticity_pencil(nx: int, ny: int, lx: float = 10.0, ly: float = 2.0, young: float = 200e9, nu: float = 0.3,
                      rho: float = 8000.0, clamp: str = "left", m_bc_diagonal: float = 1.0) -> Pencil:
    """Free vibration of a plane-stress plate, K x = omega^2 M x (the modal problem of the reference's elasticity
    module, `Elasticity/operators.py:228-270`, `Elasticity/utils.py:139-155`; a deep cantilever in the spirit of the
    NAFEMS free-vibration benchmarks): P1 triangles, two displacement DOFs per node (node-major), consistent mass,
    clamped edge as identity rows and columns in K and M (dolfinx's `assemble_matrix(bcs=...)`), which adds the
    spurious eigenvalue 1 the reference drops with `skip_below_hz` (1 / m_bc_diagonal in general)."""
    shape = (nx, ny)
    axes = [np.linspace(0.0, lx, nx + 1), np.linspace(0.0, ly, ny + 1)]
    pts_shape = (nx + 1, ny + 1)
    cells2 = _simplices(shape)                      # vertices on the doubled (P2) lattice: even coordinates
    cells = cells2 // 2
    X = np.stack([axes[d][cells[:, :, d]] for d in range(2)], axis=-1)
    J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))
    detJ = np.abs(np.linalg.det(J))
    Jinv = np.linalg.inv(J)
    glam = np.concatenate([-Jinv.sum(axis=1, keepdims=True), Jinv], axis=1)     # (cells, 3 vertices, 2): grad of hat functions
    ne = cells.shape[0]
    B = np.zeros((ne, 3, 6))
    B[:, 0, 0::2] = glam[:, :, 0]
    B[:, 1, 1::2] = glam[:, :, 1]
    B[:, 2, 0::2] = glam[:, :, 1]
    B[:, 2, 1::2] = glam[:, :, 0]
    D = young / (1.0 - nu * nu) * np.array([[1.0, nu, 0.0], [nu, 1.0, 0.0], [0.0, 0.0, 0.5 * (1.0 - nu)]])
    area = 0.5 * detJ
    ke = np.einsum("e,eia,ij,ejb->eab", area, B, D, B)
    m_s = (np.ones((3, 3)) + np.eye(3)) / 12.0                                   # P1 consistent mass of a unit-area triangle
    me = np.zeros((ne, 6, 6))
    for c in range(2):
        me[:, c::2, c::2] = rho * area[:, None, None] * m_s[None]
    node = cells[:, :, 0] * pts_shape[1] + cells[:, :, 1]
    dof = np.stack([2 * node, 2 * node + 1], axis=-1).reshape(ne, 6)
    n = 2 * pts_shape[0] * pts_shape[1]
    r = np.broadcast_to(dof[:, :, None], ke.shape).ravel()
    c = np.broadcast_to(dof[:, None, :], ke.shape).ravel()
    K = sp.coo_matrix((ke.ravel(), (r, c)), shape=(n, n)).tocsr()
    M = sp.coo_matrix((me.ravel(), (r, c)), shape=(n, n)).tocsr()
    K = ((K + K.T) * 0.5).tocsr()
    M = ((M + M.T) * 0.5).tocsr()
    idx = np.stack(np.meshgrid(np.arange(pts_shape[0]), np.arange(pts_shape[1]), indexing="ij"), -1).reshape(-1, 2)
    on = {"left": idx[:, 0] == 0, "right": idx[:, 0] == nx, "none": np.zeros(len(idx), bool)}[clamp]
    bnodes = np.nonzero(on)[0]
    bc = np.sort(np.concatenate([2 * bnodes, 2 * bnodes + 1]))
    K = _apply_identity_rows_cols(K, bc)
    M = _apply_identity_rows_cols(M, bc, m_bc_diagonal)
    xy = np.stack([axes[0][idx[:, 0]], axes[1][idx[:, 1]]], axis=1)
    coords = np.repeat(xy, 2, axis=0)
    return Pencil(A=_canonical(K), M=_canonical(M), dofs_u=np.arange(n), dofs_p=np.zeros(0, np.int64),
                  dirichlet=bc.astype(np.int64), coords=coords,
                  meta=dict(kind="elasticity", shape=shape, lx=lx, ly=ly, young=young, nu=nu, rho=rho))
