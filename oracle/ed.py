"""Exact diagonalisation for small N (TEST INFRASTRUCTURE).

Stands in for netket/exact.py:24-131 (``lanczos_ed`` / ``full_ed``): dense symmetric
eigensolve of the matrix assembled from connected elements.
"""

import numpy as np

from .operators import to_dense


def full_ed(conn_fn, N, total_sz=None, k=None, compute_eigenvectors=False):
    H = to_dense(conn_fn, N, total_sz)
    assert np.allclose(H, H.T, atol=1e-13), "operator is not hermitian"
    if compute_eigenvectors:
        w, v = np.linalg.eigh(H)
        return (w if k is None else w[:k]), (v if k is None else v[:, :k])
    w = np.linalg.eigvalsh(H)
    return w if k is None else w[:k]


def expectation(conn_fn, N, psi, total_sz=None):
    """<psi|H|psi>/<psi|psi> for a real amplitude vector in the reference basis ordering."""
    H = to_dense(conn_fn, N, total_sz)
    return float(psi @ H @ psi / (psi @ psi))
