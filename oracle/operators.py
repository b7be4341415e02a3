"""Connected elements of Ising / Heisenberg / 2-site LocalOperator (TEST INFRASTRUCTURE).

Restates, in NumPy:
  * Ising:   netket/operator/_ising/jax.py:125-175 (``_ising_mels_jax``,
             ``_ising_conn_states_jax``, ``_ising_kernel_jax``, ``_ising_n_conn_jax``),
             ``max_conn_size = N+1`` (netket/operator/_ising/base.py:158-161).
  * Heisenberg bond matrices: netket/operator/_heisenberg.py:77-122.
  * GraphOperator term list: netket/operator/_graph_operator.py:103-145.
  * LocalOperator canonicalisation: netket/operator/_local_operator/helpers.py:75-213
             (sort ``acting_on``, permute the matrix), same-support terms are summed
             (netket/operator/_local_operator/base.py:136-147).
  * table packing: netket/operator/_local_operator/compile_helpers.py:29-218,221-257,
             301-314,319-366.
  * connected elements + compaction: netket/operator/_local_operator/jax.py:37-201.

All functions work on sigma = +/-1 arrays (any integer/float dtype) and convert to the
reference's local indices (0 <-> +1, 1 <-> -1) internally, like ``get_conn_padded``
does (netket/operator/_ising/jax.py:82-88, _local_operator/jax.py:256-284).
"""

import numpy as np

from .hilbert import states_to_local_indices, local_indices_to_states, all_states, states_to_numbers

MEL_CUTOFF = 1.0e-10  # netket/operator/_local_operator/base.py:90


# --------------------------------------------------------------------------- Ising
def ising_conn_padded(x, edges, h, J=1.0, mel_dtype=np.float64):
    """x[..., N] -> (xp[..., K, N] dtype(x), mels[..., K]) with K = N+1 (K = 1 if h == 0)."""
    x = np.asarray(x)
    batch = x.shape[:-1]
    N = x.shape[-1]
    xi = states_to_local_indices(x.reshape(-1, N))
    edges = np.asarray(edges)
    same = xi[:, edges[:, 0]] == xi[:, edges[:, 1]]
    static_zero = h is None or h == 0
    K = 1 if static_zero else N + 1
    mels = np.zeros((xi.shape[0], K), dtype=mel_dtype)
    mels[:, 0] = np.asarray(J, dtype=mel_dtype) * (2 * same.astype(np.int64) - 1).sum(axis=-1)
    if static_zero:
        xpi = xi[:, None, :]
    else:
        mels[:, 1:] = -np.asarray(h, dtype=mel_dtype)
        flip = np.eye(K, N, k=-1, dtype=bool)
        was0 = xi[:, None, :] == 0
        xpi = np.where(flip[None] ^ was0, 0, 1)
    xp = local_indices_to_states(xpi, dtype=x.dtype)
    return xp.reshape(batch + (K, N)), mels.reshape(batch + (K,))


def ising_n_conn(x, edges, h, J=1.0):
    """_ising_n_conn_jax (jax.py:168-175)."""
    x = np.asarray(x)
    xi = states_to_local_indices(x)
    edges = np.asarray(edges)
    n_x = 0 if (h is None or h == 0) else x.shape[-1]
    same = xi[..., edges[:, 0]] == xi[..., edges[:, 1]]
    zz = J * (2 * same.astype(np.int64) - 1).sum(axis=-1)
    return n_x + (zz != 0).astype(np.int32)


# --------------------------------------------------------------------------- Heisenberg / GraphOperator
SZ_SZ = np.array([[1, 0, 0, 0], [0, -1, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]], dtype=np.float64)
EXCHANGE = np.array([[0, 0, 0, 0], [0, 0, 2, 0], [0, 2, 0, 0], [0, 0, 0, 0]], dtype=np.float64)


def heisenberg_bond_ops(J=1.0, sign_rule=False):
    """_heisenberg.py:97-122.  J scalar or sequence (one per edge colour)."""
    if isinstance(J, (list, tuple, np.ndarray)):
        if isinstance(sign_rule, bool):
            sign_rule = [sign_rule] * len(J)
        return [Jc * (SZ_SZ - EXCHANGE if s else SZ_SZ + EXCHANGE) for Jc, s in zip(J, sign_rule)], list(range(len(J)))
    return [J * (SZ_SZ - EXCHANGE if sign_rule else SZ_SZ + EXCHANGE)], []


def graph_operator_terms(edges, colors, bond_ops, bond_ops_colors):
    """(operators, acting_on) lists in the order GraphOperator emits them (:124-143)."""
    ops, aon = [], []
    if len(bond_ops_colors) > 0:
        for (u, v), color in zip(np.asarray(edges).tolist(), np.asarray(colors).tolist()):
            for c, bc in enumerate(bond_ops_colors):
                if bc == color:
                    ops.append(np.asarray(bond_ops[c]))
                    aon.append((u, v))
    else:
        assert len(bond_ops) == 1
        for u, v in np.asarray(edges).tolist():
            ops.append(np.asarray(bond_ops[0]))
            aon.append((u, v))
    return ops, aon


# --------------------------------------------------------------------------- LocalOperator
def _reorder_kronecker_product(mat, acting_on, d=2):
    """helpers.py:151-213 for a uniform local dimension d."""
    acting_on = tuple(int(a) for a in acting_on)
    srt = tuple(sorted(acting_on))
    if srt == acting_on:
        return np.asarray(mat), acting_on
    k = len(acting_on)
    unsorted_ids = [srt.index(s) for s in acting_on]
    n = d**k
    # digits (first site most significant) of every number in the sorted ordering
    v = np.stack([(np.arange(n) // d ** (k - 1 - p)) % d for p in range(k)], axis=1)
    v_unsorted = v[:, unsorted_ids]
    n_unsorted = (v_unsorted * (d ** np.arange(k - 1, -1, -1))[None, :]).sum(axis=1)
    mat = np.asarray(mat)
    return mat[n_unsorted, :][:, n_unsorted], srt


def canonical_operators_dict(operators, acting_on, dtype=np.float64):
    """{sorted acting_on tuple: matrix}, insertion ordered, same-support terms summed."""
    d = {}
    for op, aon in zip(operators, acting_on):
        op, aon = _reorder_kronecker_product(np.asarray(op, dtype=dtype), aon)
        if aon in d:
            d[aon] = d[aon] + op
        else:
            d[aon] = op.copy()
    return d


def _number_to_state(number, k, d=2):
    out = np.zeros(k, dtype=np.float64)
    ip = number
    p = k - 1
    while ip > 0:
        out[p] = ip % d
        ip //= d
        p -= 1
    return out


def pack_internals(operators_dict, constant=0.0, dtype=np.float64, mel_cutoff=MEL_CUTOFF, d=2):
    """compile_helpers.py:174-218 (grouped by number of sites) + :29-171 per group."""
    keys = list(operators_dict.keys())
    sizes = np.array([len(k) for k in keys], dtype=np.int64)
    data = {
        "acting_on": [], "basis": [], "diag_mels": [], "n_conns": [], "mels": [], "x_prime": [],
        "constant": np.asarray(constant, dtype=dtype),
    }
    nonzero_diagonal = bool(np.abs(constant) >= mel_cutoff)
    max_conn_size = 0
    for s in (np.unique(sizes) if len(sizes) > 0 else []):
        group = [k for k in keys if len(k) == s]
        n_ops = len(group)
        op_size = d ** int(s)
        # max_nonzero_per_row (:319-366): >= cutoff, diagonal excluded
        row_nnz_max = []
        for k in group:
            m = np.abs(operators_dict[k]) >= mel_cutoff
            np.fill_diagonal(m, False)
            row_nnz_max.append(int(np.count_nonzero(m, axis=1).max()))
        ncmax = max(row_nnz_max)
        acting_on = np.asarray(group, dtype=np.int64).reshape(n_ops, int(s))
        basis = np.tile((d ** np.arange(int(s), dtype=np.int64))[None, :], (n_ops, 1))
        diag_mels = np.full((n_ops, op_size), np.nan, dtype=dtype)
        mels = np.full((n_ops, op_size, ncmax), np.nan, dtype=dtype)
        x_prime = np.full((n_ops, op_size, ncmax, int(s)), -1, dtype=np.float64)
        n_conns = np.zeros((n_ops, op_size), dtype=np.int64)
        for o, k in enumerate(group):
            op = operators_dict[k]
            for i in range(op_size):  # _append_matrix (:221-257): strictly > epsilon
                diag_mels[o, i] = op[i, i]
                for j in range(op_size):
                    if i != j and np.abs(op[i, j]) > mel_cutoff:
                        c = n_conns[o, i]
                        mels[o, i, c] = op[i, j]
                        x_prime[o, i, c, :] = _number_to_state(j, int(s), d)
                        n_conns[o, i] += 1
        if np.any(np.abs(diag_mels) >= mel_cutoff):
            nonzero_diagonal = True
        max_conn_size += int(np.sum(row_nnz_max))
        for name, v in (("acting_on", acting_on), ("basis", basis), ("diag_mels", diag_mels),
                        ("n_conns", n_conns), ("mels", mels), ("x_prime", x_prime)):
            data[name].append(v)
    if nonzero_diagonal:
        max_conn_size += 1
    data["nonzero_diagonal"] = nonzero_diagonal
    data["max_conn_size"] = max_conn_size
    data["mel_cutoff"] = mel_cutoff
    return data


def local_operator_conn_padded(x, tables):
    """_local_operator_kernel_jax (jax.py:74-201) + _get_conn_padded (:256-284).

    Returns (xp[..., K, N] dtype(x), mels[..., K], n_conn[...]).
    """
    x = np.asarray(x)
    batch = x.shape[:-1]
    N = x.shape[-1]
    xi = states_to_local_indices(x.reshape(-1, N))
    B = xi.shape[0]
    K = tables["max_conn_size"]
    cutoff = tables["mel_cutoff"]
    dtype = tables["constant"].dtype

    cand_mels = []
    cand_xp = []
    rows = []
    for g in range(len(tables["acting_on"])):
        aon = tables["acting_on"][g]
        basis = tables["basis"][g]
        s = aon.shape[1]
        xs = xi[:, aon]  # (B, n_ops, s)
        # _state_to_number: sum_k basis[k] * x[s-k-1]
        rows.append((xs[:, :, ::-1] * basis[None, :, :]).sum(axis=-1))
    if tables["nonzero_diagonal"]:
        md = np.full((B,), tables["constant"], dtype=dtype)
        for g in range(len(tables["acting_on"])):
            dm = tables["diag_mels"][g]
            a = np.arange(dm.shape[0])
            md = md + dm[a[None, :], rows[g]].sum(axis=-1)
        cand_mels.append(md[:, None])
        cand_xp.append(xi[:, None, :])
    for g in range(len(tables["acting_on"])):
        aon = tables["acting_on"][g]
        n_ops, s = aon.shape
        ncmax = tables["mels"][g].shape[2]
        a = np.arange(n_ops)
        r = rows[g]
        nc = tables["n_conns"][g][a[None, :], r]  # (B, n_ops)
        maskall = np.arange(ncmax)[None, None, :] < nc[:, :, None]
        with np.errstate(invalid="ignore"):
            m = tables["mels"][g][a[None, :], r] * maskall  # NaN padding stays NaN, as in the reference
        new = tables["x_prime"][g][a[None, :], r].astype(np.int64)  # (B, n_ops, ncmax, s)
        old = np.broadcast_to(xi[:, aon][:, :, None, :], new.shape)
        new = np.where(maskall[..., None], new, old)
        xp = np.broadcast_to(xi[:, None, None, :], (B, n_ops, ncmax, N)).copy()
        for o in range(n_ops):
            xp[:, o, :, aon[o]] = np.moveaxis(new[:, o, :, :], -1, 0)
        cand_mels.append(m.reshape(B, n_ops * ncmax))
        cand_xp.append(xp.reshape(B, n_ops * ncmax, N))
    # trailing pad row (x, 0), selected by fill index -1
    cand_mels.append(np.zeros((B, 1), dtype=dtype))
    cand_xp.append(xi[:, None, :])
    mels_all = np.concatenate(cand_mels, axis=1)
    xp_all = np.concatenate(cand_xp, axis=1)
    with np.errstate(invalid="ignore"):
        mask = np.abs(mels_all) > cutoff
    n_conn = mask.sum(axis=-1)
    out_m = np.zeros((B, K), dtype=dtype)
    out_x = np.empty((B, K, N), dtype=np.int64)
    for bi in range(B):
        ind = np.flatnonzero(mask[bi])[:K]
        full = np.full(K, -1, dtype=np.int64)
        full[: len(ind)] = ind
        out_m[bi] = mels_all[bi][full]
        out_x[bi] = xp_all[bi][full]
    xp = local_indices_to_states(out_x, dtype=x.dtype)
    return xp.reshape(batch + (K, N)), out_m.reshape(batch + (K,)), n_conn.reshape(batch)


def heisenberg_tables(edges, colors=None, J=1.0, sign_rule=False, dtype=np.float64):
    bond_ops, bond_colors = heisenberg_bond_ops(J, sign_rule)
    if colors is None:
        colors = np.zeros(len(edges), dtype=np.int32)
    ops, aon = graph_operator_terms(edges, colors, bond_ops, bond_colors)
    return pack_internals(canonical_operators_dict(ops, aon, dtype=dtype), 0.0, dtype=dtype)


# --------------------------------------------------------------------------- dense matrices (small N)
def to_dense(conn_fn, N, total_sz=None):
    """Dense matrix in the reference's basis ordering from a get_conn_padded-style function.

    conn_fn(x[B,N] int8) -> (xp[B,K,N], mels[B,K]).  Duplicated x' accumulate, exactly as
    DiscreteOperator.to_sparse does.
    """
    states = all_states(N, total_sz)
    nums = states_to_numbers(states, N)
    lut = -np.ones(1 << N, dtype=np.int64)
    lut[nums] = np.arange(len(states))
    xp, mels = conn_fn(states)[:2]
    D = len(states)
    H = np.zeros((D, D), dtype=np.float64)
    cols = lut[states_to_numbers(xp.reshape(-1, N), N)].reshape(D, -1)
    for r in range(D):
        nz = mels[r] != 0
        if np.any(cols[r][nz] < 0):
            raise ValueError("connected state outside the (constrained) Hilbert space")
        np.add.at(H[r], cols[r][nz], mels[r][nz])
    return H


def kron_dense(N, one_site=(), two_site=()):
    """Independent construction by Kronecker products (site 0 most significant).

    one_site: iterable of (i, 2x2), two_site: iterable of (i, j, 4x4) with i<j and the 4x4
    indexed by 2*idx_i + idx_j.  Used to check ``to_dense``.
    """
    dim = 1 << N
    H = np.zeros((dim, dim))
    nums = np.arange(dim)
    bit = lambda n, s: (n >> (N - 1 - s)) & 1  # noqa: E731
    for i, m in one_site:
        m = np.asarray(m, dtype=np.float64)
        for bi in range(2):
            for bj in range(2):
                if m[bi, bj] == 0:
                    continue
                src = nums[bit(nums, i) == bi]
                dst = src ^ ((bi ^ bj) << (N - 1 - i))
                H[src, dst] += m[bi, bj]
    for i, j, m in two_site:
        m = np.asarray(m, dtype=np.float64)
        for r in range(4):
            for c in range(4):
                if m[r, c] == 0:
                    continue
                ri, rj = r >> 1, r & 1
                ci, cj = c >> 1, c & 1
                src = nums[(bit(nums, i) == ri) & (bit(nums, j) == rj)]
                dst = src ^ ((ri ^ ci) << (N - 1 - i)) ^ ((rj ^ cj) << (N - 1 - j))
                H[src, dst] += m[r, c]
    return H
