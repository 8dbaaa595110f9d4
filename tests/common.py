"""Shared helpers for the test-suite: package / oracle loading and problem fixtures."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def package():
    libs = [os.path.join(entry.PKG_DIR, "libgomp_host.so"), os.path.join(entry.PKG_DIR, "libgomp_cuda.so")]
    if not all(os.path.exists(p) for p in libs):
        entry.build()
    return entry.load_package()


def oracle():
    return entry.load_oracle()


def random_block_qp(rng, bs, T, batch, rows_per_block=None, with_q=True):
    """A batch of random QPs whose reduced KKT matrix is block tridiagonal with T blocks of size bs under the
    identity order: P = banded SPD (half bandwidth < bs, never crossing more than one block boundary), every row
    of A lives in one block or two adjacent ones.  Bounds mix equality, loose, one-sided and box rows.
    Returns (P csc upper, A csc pattern, Ax [batch, nnz], l, u, q)."""
    import numpy as np
    import scipy.sparse as sp
    n = bs * T
    rows_per_block = rows_per_block or bs + 2
    # P: block tridiagonal SPD
    M = np.zeros((n, n))
    for t in range(T):
        blk = rng.standard_normal((bs, bs))
        M[t * bs:(t + 1) * bs, t * bs:(t + 1) * bs] = blk @ blk.T + 0.5 * np.eye(bs)
        if t + 1 < T:
            off = 0.1 * rng.standard_normal((bs, bs)) * (rng.random((bs, bs)) < 0.3)
            M[(t + 1) * bs:(t + 2) * bs, t * bs:(t + 1) * bs] = off
            M[t * bs:(t + 1) * bs, (t + 1) * bs:(t + 2) * bs] = off.T
    M += (np.abs(M).sum(axis=1).max() * 0.05) * np.eye(n)     # diagonally dominant enough to stay SPD
    P = sp.csc_matrix(np.triu(M))
    rows = []
    for t in range(T):
        for r in range(rows_per_block):
            row = np.zeros(n)
            span = 2 if (t + 1 < T and r % 3 == 0) else 1            # a third of the rows couple block t and t+1
            cols = rng.choice(np.arange(t * bs, (t + span) * bs), size=min(3 + r % 3, span * bs), replace=False)
            row[cols] = rng.standard_normal(len(cols))
            rows.append(row)
    rows.append(np.zeros(n))                                            # an empty row (0 in [l, u])
    A = sp.csc_matrix(np.array(rows))
    A.sort_indices()
    m = A.shape[0]
    Ax = np.stack([A.data * (1 + 0.2 * rng.standard_normal(A.nnz)) for _ in range(batch)])
    l = -rng.random((batch, m)) - 0.05
    u = rng.random((batch, m)) + 0.05
    l[:, 1::7] = u[:, 1::7]                 # equality rows
    l[:, 2::7], u[:, 2::7] = -1e30, 1e30    # loose rows
    u[:, 3::7] = 1e30                       # one-sided rows
    l[:, -1], u[:, -1] = -1.0, 1.0          # the empty row
    q = rng.standard_normal((batch, n)) if with_q else None
    return P, A, Ax, l, u, q
