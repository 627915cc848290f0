"""Writes tests/golden/kat_eigen.json: the known-answer cases the REFERENCE's own tests hold for
the eigensolver boundary.

The reference cannot be imported here (petsc4py/slepc4py/dolfinx are not installable in this
image), so the fixtures are transcribed from the literal matrices and expected values in

* /root/reference/tests/unit/Solver/test_eigen.py:36-47,107-304   (tiny dense matrices)
* /root/reference/tests/benchmark/vibrating_membrane.md:102-110   (P2 membrane, (a,b)=(2,4), 32x32:
  relative errors of modes 1-3 and the mean over 15 modes, as published by the reference)

Run:  python tests/golden/make_golden.py
"""

import json
from pathlib import Path

import numpy as np

rs = np.random.RandomState(42)  # test_eigen.py:73-78 (random_spd_matrix fixture)
X = rs.randn(5, 5)
spd = (X.T @ X + np.eye(5) * 1e-3).T

cases = {
    "diag3_standard": {  # test_eigen.py:107-118
        "ref": "tests/unit/Solver/test_eigen.py:107-118", "A": np.diag([1.0, 1.5, -42.0]).tolist(), "M": None,
        "problem_type": "GHEP", "num_eig": 3, "atol": 1e-3, "max_it": 100,
        "expected_sorted": [-42.0, 1.0, 1.5], "abs_tol": 1e-3},
    "diag3_generalized_identity": {  # test_eigen.py:121-129
        "ref": "tests/unit/Solver/test_eigen.py:121-129", "A": np.diag([1.0, 1.5, -42.0]).tolist(),
        "M": np.eye(3).tolist(), "problem_type": "GHEP", "num_eig": 3, "atol": 1e-3, "max_it": 100,
        "expected_sorted": [-42.0, 1.0, 1.5], "abs_tol": 1e-3},
    "jordan2": {  # test_eigen.py:132-139
        "ref": "tests/unit/Solver/test_eigen.py:132-139", "A": [[1, 1], [0, 1]], "M": None,
        "problem_type": "GNHEP", "num_eig": 2, "atol": 1e-6, "max_it": 100,
        "expected_sorted_real": [1.0, 1.0], "abs_tol": 1e-6},
    "complex_pair": {  # test_eigen.py:142-172
        "ref": "tests/unit/Solver/test_eigen.py:142-172", "A": [[5, -5], [1, 1]], "M": None,
        "problem_type": "GNHEP", "num_eig": 2, "atol": 1e-6, "max_it": 100,
        "expected_by_imag": [[3.0, -1.0], [3.0, 1.0]], "ratios_by_imag": [[2.0, -1.0], [2.0, 1.0]], "abs_tol": 1e-6},
    "smallest_magnitude_alias": {  # test_eigen.py:175-185 (passes through the LARGEST_REAL alias)
        "ref": "tests/unit/Solver/test_eigen.py:175-185", "A": np.diag([1.0, 1.5, -42.0]).tolist(), "M": None,
        "problem_type": "GHEP", "num_eig": 3, "atol": 1e-3, "max_it": 100, "which": "SMALLEST_MAGNITUDE",
        "first_two_sorted": [1.0, 1.5], "abs_tol": 1e-3},
    "random_spd5": {  # test_eigen.py:242-252
        "ref": "tests/unit/Solver/test_eigen.py:242-252", "A": spd.tolist(), "M": None, "problem_type": "HEP",
        "num_eig": 5, "atol": 1e-8, "max_it": 200,
        "expected_sorted": sorted(np.linalg.eigvalsh(X.T @ X + np.eye(5) * 1e-3).tolist()), "rel_tol": 1e-6},
    "shift_invert_epsilon": {  # test_eigen.py:255-269
        "ref": "tests/unit/Solver/test_eigen.py:255-269",
        "A": np.diag(np.array([1.0, 1.0 + 1e-8, 1.0 + 2e-8]) + 1e-9).tolist(), "M": None, "problem_type": "HEP",
        "num_eig": 3, "atol": 1e-12, "max_it": 500, "st": "SINVERT", "target": 1.0,
        "expected_sorted": [1.0, 1.0 + 1e-8, 1.0 + 2e-8], "rel_tol": 1e-6},
    "singular_m_raises": {  # test_eigen.py:272-281
        "ref": "tests/unit/Solver/test_eigen.py:272-281", "A": np.diag([1.0, 1.5, -42.0]).tolist(),
        "M": np.diag([1.0, 0.0, 0.0]).tolist(), "problem_type": "GHEP", "num_eig": 2, "atol": 1e-6, "max_it": 200,
        "raises": True},
    "repeated_223": {  # test_eigen.py:284-304
        "ref": "tests/unit/Solver/test_eigen.py:284-304", "A": np.diag([2.0, 2.0, 3.0]).tolist(), "M": None,
        "problem_type": "HEP", "num_eig": 3, "atol": 1e-8, "max_it": 200,
        "expected_sorted": [2.0, 2.0, 3.0], "abs_tol": 1e-8, "rank": 3},
}
membrane = {  # vibrating_membrane.md:102-110 (published by the reference)
    "ref": "tests/benchmark/vibrating_membrane.md:102-110", "a": 2.0, "b": 4.0, "mesh": [32, 32], "modes": 15,
    "lambda_num": [3.084254, 4.934827, 8.019193], "lambda_ana": [3.084251, 4.934802, 8.019054],
    "rel_err_first3": [9.01e-7, 5.04e-6, 1.73e-5], "rel_err_mean": 6.06e-5,
}
Path(__file__).with_name("kat_eigen.json").write_text(json.dumps({"cases": cases, "membrane": membrane}, indent=1))
print("wrote", Path(__file__).with_name("kat_eigen.json"))
