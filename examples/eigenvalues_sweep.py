"""Reynolds sweep of shift-and-invert eigensolves on stored (A, M) pencils -- the workflow of the reference's
`.examples/eigenvalues.py:61-108` with the B200 backend: per Reynolds number load `A.mtx` / `M.mtx`, solve for the
eigenvalues next to a literature target, write the leading one to `sigma_eig0.txt`.

    python examples/eigenvalues_sweep.py CASES_DIR              # CASES_DIR/reynolds_<Re>/matrices/{A,M}.mtx
    python examples/eigenvalues_sweep.py CASES_DIR --synthetic  # first write small synthetic wake pencils there

Only the two import lines differ from the reference script.  Every pencil of the sweep has the
same sparsity pattern, so the host analysis runs once and each further Reynolds number costs one numeric
factorisation + one Krylov-Schur run (`es.solver.stats["symbolic_seconds"]` drops to milliseconds after the first).
"""

from __future__ import annotations

import argparse
import logging
import os
import sys
from pathlib import Path

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from lsa_fw_b200 import EigenSolver, EigensolverConfig, PreconditionerType, iPETScMatrix, iSTType  # noqa: E402

NUM_EIG, EIG_INDEX, ATOL = 5, 0, 1e-3
REYNOLDS = tuple(float(r) for r in range(40, 91, 5))
TARGETS = (-0.03 + 0.7197388769374216j, 0.7316769290210628j, 0.018 + 0.7379601143282424j, 0.03 + 0.742986662573986j,
           0.05 + 0.744243299635422j, 0.061 + 0.7461282552275759j, 0.072 + 0.7461282552275759j,
           0.085 + 0.744557458900781j, 0.09 + 0.742986662573986j, 0.1 + 0.7398450699203962j,
           0.115 + 0.7351326809400116j)          # shift targets of the reference's table (DOI:10.1115/1.4042737)

logger = logging.getLogger("eigenvalues_sweep")


def write_synthetic_cases(root: Path, nx: int = 60, ny: int = 24) -> None:
    """Small cylinder-wake surrogate pencils (one pattern, Reynolds-dependent values) in the directory layout of the
    reference's assembly script."""
    from lsa_fw_b200 import pencils

    pc = pencils.adapted_wake_2d(nx=nx, ny=ny, re=REYNOLDS[0], split_viscous=True)
    M = iPETScMatrix(pc.M)
    for re in REYNOLDS:
        mat_dir = root / f"reynolds_{re:.1f}" / "matrices"
        A = iPETScMatrix(pc.A.__class__((pc.a_data_at(re), pc.A.indices, pc.A.indptr), shape=pc.A.shape))
        A.export(mat_dir / "A.mtx")
        M.export(mat_dir / "M.mtx")
    logger.info("wrote %d synthetic cases (%d DOFs each) under '%s'", len(REYNOLDS), pc.n, root)


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("cases", type=Path)
    ap.add_argument("--synthetic", action="store_true", help="write synthetic pencils into CASES first")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--dry-run", action="store_true", help="load and report the matrices, do not solve")
    args = ap.parse_args()
    logging.basicConfig(level=logging.INFO, format="%(levelname)s %(name)s: %(message)s")
    if args.synthetic:
        write_synthetic_cases(args.cases)

    for re, target in zip(REYNOLDS, TARGETS):
        case_dir = args.cases / f"reynolds_{re:.1f}"
        a_path, m_path = case_dir / "matrices" / "A.mtx", case_dir / "matrices" / "M.mtx"
        if not a_path.exists() or not m_path.exists():
            logger.warning("Skipping Re = %.1f: missing matrices in '%s'", re, a_path.parent)
            continue
        A = iPETScMatrix.from_path(a_path)
        A.assemble()
        M = iPETScMatrix.from_path(m_path)
        M.assemble()
        logger.info("[Re=%.1f] A: shape=%s, nnz=%d, norm=%.3e", re, A.shape, A.nonzero_entries, A.norm)
        logger.info("[Re=%.1f] M: shape=%s, nnz=%d, norm=%.3e", re, M.shape, M.nonzero_entries, M.norm)
        if args.dry_run:
            continue

        cfg = EigensolverConfig(num_eig=NUM_EIG, atol=ATOL)
        es = EigenSolver(A, M, cfg=cfg, check_hermitian=False)
        es.solver.set_st_type(iSTType.SINVERT)
        es.solver.set_target(target)
        es.solver.set_st_pc_type(PreconditionerType.LU)
        es.solver.set_backend_options(device=args.device)
        es.solver.solve()

        sigma = es.solver.get_eigenvalue(EIG_INDEX)
        (case_dir / f"sigma_eig{EIG_INDEX}.txt").write_text(f"{sigma.real} {sigma.imag}\n", encoding="utf-8")
        st = es.solver.stats
        logger.info("[Re=%.1f] sigma = %.6f%+.6fj  (%d converged; analysis %.3f s, factor %.3f s, eigs %.3f s)", re, sigma.real,
                    sigma.imag, es.solver.get_num_converged(), st.get("symbolic_seconds", 0.0), st.get("factor_seconds", 0.0),
                    st.get("eigs_seconds", 0.0))
    logger.info("All cases processed.")


if __name__ == "__main__":
    main()
