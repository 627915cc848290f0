"""Digests of the host symbolic analysis (ordering, fronts, scatter maps) of three small pencils.

Written by the analysis as it stood BEFORE the round-2 host-side rework (parallel graph build, compact per-node
subgraphs in the nested dissection, branch-free sweeps, first-touch arrays): the rework must not move a single
entry, because every GPU measurement of the round was taken on these structures.  `tests/test_host_logic.py::
test_symbolic_analysis_is_pinned_and_thread_count_independent` recomputes them.

    python tests/golden/make_symbolic_digests.py          # rewrites tests/golden/symbolic_digests.json
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

NAMES = ["perm", "iperm", "sn_ptr", "st_ptr", "st_idx", "ea_map", "lvl_ptr", "lvl_front", "a_dst", "m_dst", "parent",
         "level", "front_k", "front_r", "p_off", "q_off", "c_off", "child_idx", "child0", "nchild"]


def digest(h) -> dict:
    return {nm: hashlib.sha1(h.symbolic_array(nm).tobytes()).hexdigest()[:12] for nm in NAMES}


def cases():
    from lsa_fw_b200 import pencils

    pc = pencils.assemble_pencil((24, 12), (8.0, 3.0), re=50.0, baseflow=pencils.wake_profile(0.9, 1.2, 1.5))
    ol = (pc.A.diagonal() == 0).astype(np.uint8)
    yield "small2d", pc, dict(leaf_size=24, order_last=ol), True, (0, 1)
    yield "small2d_nom", pc, dict(leaf_size=32), False, (0, 1)
    pc = pencils.cavity_3d(8)
    ol = (pc.A.diagonal() == 0).astype(np.uint8)
    yield "cav8", pc, dict(leaf_size=64, order_last=ol), True, (0, 1)
    for r in range(2):
        yield f"cav8_w2r{r}", pc, dict(leaf_size=64, order_last=ol), True, (r, 2)


def hub_pattern():
    """4000 unknowns, random sparse pattern with one dense row and one dense column (hub vertices)."""
    import scipy.sparse as sp

    n = 4000
    a = sp.random(n, n, density=6e-3, random_state=3, format="lil", dtype=np.float64)
    a.setdiag(4.0)
    a[17, ::2] = 1.0
    a[:, 2001] = 1.0
    a = a.tocsr()
    a.sort_indices()
    return a


def compute() -> dict:
    from lsa_fw_b200 import _lib

    out = {}
    for tag, pc, kw, with_m, (rank, world) in cases():
        h = _lib.Handle(pc.n, device=-1, rank=rank, world=world)
        if with_m:
            h.analyze(pc.A.indptr, pc.A.indices, pc.M.indptr, pc.M.indices, **kw)
        else:
            h.analyze(pc.A.indptr, pc.A.indices, **kw)
        out[tag] = digest(h)
        h.close()
    return out


if __name__ == "__main__":
    from lsa_fw_b200 import _lib

    out = compute()
    hub = hub_pattern()
    h = _lib.Handle(hub.shape[0], device=-1)
    h.analyze(hub.indptr, hub.indices, leaf_size=32)
    out["hub"] = digest(h)
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "symbolic_digests.json"), "w"), indent=1)
