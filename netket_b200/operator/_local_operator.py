"""``LocalOperator`` (sum of 1- and 2-site terms), ``GraphOperator`` and ``Heisenberg``.

Host side (one-time, NumPy): canonicalise terms and pack the per-term lookup tables that the kernels index
with the local row number — the job of netket/operator/_local_operator/helpers.py:75-213 (sort
``acting_on``, permute the matrix), base.py:136-147 (terms with the same support are summed) and
compile_helpers.py:29-218 (``diag_mels``, ``n_conns``, ``mels``, ``x_prime`` grouped by number of sites).
Device side: ``nk_localop_conn`` (get_conn_padded with the reference's compaction semantics,
netket/operator/_local_operator/jax.py:74-201) and ``nk_eloc_localop_rbm`` / the fused sweep.
"""

import ctypes as C
import numbers

import numpy as np
import torch

from .. import _lib
from ._base import DiscreteJaxOperator


def _sort_support(mat, sites):
    """Permute a k-site matrix so that its support is ascending (first site = most significant digit)."""
    sites = tuple(int(s) for s in sites)
    order = tuple(sorted(sites))
    if order == sites:
        return mat, sites
    k = len(sites)
    where = [order.index(s) for s in sites]
    dim = 2 ** k
    digits = (np.arange(dim)[:, None] >> np.arange(k - 1, -1, -1)[None, :]) & 1  # sorted ordering
    src = (digits[:, where] << np.arange(k - 1, -1, -1)[None, :]).sum(axis=1)    # same state, original ordering
    return mat[np.ix_(src, src)], order


class LocalOperatorJax(DiscreteJaxOperator):
    def __init__(self, hilbert, operators=[], acting_on=[], constant=0, dtype=None, *, mel_cutoff=1.0e-10):
        if isinstance(acting_on, numbers.Number):
            acting_on = [acting_on]
        nested = any(hasattr(a, "__len__") for a in acting_on)
        if not nested:
            operators, acting_on = [operators], [acting_on]
        if all(len(a) == 0 for a in acting_on):
            operators, acting_on = [], []
        mats = [np.asarray(op.todense() if hasattr(op, "todense") else op) for op in operators]
        if dtype is None:
            dtype = np.result_type(float, *[m.dtype for m in mats], np.asarray(constant).dtype)
        if np.dtype(dtype) not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise NotImplementedError("netket_b200 LocalOperator supports real float32/float64 matrix elements")
        super().__init__(hilbert, dtype)
        if mel_cutoff < 0:
            raise ValueError("mel_cutoff must be non-negative")
        self._mel_cutoff = float(mel_cutoff)
        self._constant = float(np.real(constant))
        self._terms = {}
        for m, sites in zip(mats, acting_on):
            sites = tuple(int(s) for s in sites)
            if len(sites) != len(set(sites)):
                raise ValueError(f"The operator acts on duplicated sites {sites}")
            if min(sites) < 0 or max(sites) >= hilbert.size:
                raise ValueError("An operator acts on an invalid set of sites.")
            if m.shape != (2 ** len(sites),) * 2:
                raise ValueError(f"The matrix of the sub-operator acting on sites {sites} must have shape "
                                 f"{(2 ** len(sites),) * 2}, but it has shape {m.shape}.")
            if len(sites) > 2:
                raise NotImplementedError("netket_b200 LocalOperator supports terms acting on 1 or 2 sites")
            m, sites = _sort_support(m.astype(np.float64), sites)
            self._terms[sites] = self._terms[sites] + m if sites in self._terms else m.copy()
        self._tables = None
        self._dev = {}

    # ---------------------------------------------------------------- reference-facing properties
    @property
    def operators(self):
        return [m.astype(self.dtype) for m in self._terms.values()]

    @property
    def acting_on(self):
        return list(self._terms.keys())

    @property
    def n_operators(self):
        return len(self._terms)

    @property
    def constant(self):
        return self._constant

    @property
    def mel_cutoff(self):
        return self._mel_cutoff

    @property
    def is_hermitian(self):
        return all(np.allclose(m, m.T.conj()) for m in self._terms.values())

    @property
    def max_conn_size(self):
        return self._pack()["max_conn_size"]

    # ---------------------------------------------------------------- table packing (host)
    def _pack(self):
        if self._tables is not None:
            return self._tables
        cut = self._mel_cutoff
        groups = []
        nonzero_diag = abs(self._constant) >= cut
        K = 0
        for k in sorted({len(s) for s in self._terms}):
            sel = [(s, m) for s, m in self._terms.items() if len(s) == k]
            dim = 2 ** k
            row_nnz = []
            for _, m in sel:
                off = np.abs(m) >= cut          # max_nonzero_per_row uses >=  (compile_helpers.py:319-366)
                np.fill_diagonal(off, False)
                row_nnz.append(int(off.sum(axis=1).max()))
            ncmax = max(row_nnz)
            n_ops = len(sel)
            acting = np.array([s for s, _ in sel], dtype=np.int32).reshape(n_ops, k)
            diag = np.zeros((n_ops, dim), dtype=np.float64)
            nconn = np.zeros((n_ops, dim), dtype=np.int32)
            mels = np.full((n_ops, dim, max(ncmax, 1)), np.nan, dtype=np.float64)
            xprime = np.zeros((n_ops, dim, max(ncmax, 1), k), dtype=np.int8)
            for o, (_, m) in enumerate(sel):
                diag[o] = np.diag(m)
                for r in range(dim):
                    c = 0
                    for col in range(dim):
                        if col != r and abs(m[r, col]) > cut:   # _append_matrix uses >  (compile_helpers.py:221-257)
                            mels[o, r, c] = m[r, col]
                            xprime[o, r, c] = [(col >> (k - 1 - p)) & 1 for p in range(k)]
                            c += 1
                    nconn[o, r] = c
            if np.any(np.abs(diag) >= cut):
                nonzero_diag = True
            K += int(np.sum(row_nnz))
            groups.append(dict(n_sites=k, n_ops=n_ops, ncmax=ncmax, acting_on=acting, diag_mels=diag, n_conns=nconn,
                               mels=mels, x_prime=xprime))
        if nonzero_diag:
            K += 1
        self._tables = dict(groups=groups, nonzero_diagonal=bool(nonzero_diag), max_conn_size=int(K))
        return self._tables

    def _c_struct(self, device):
        """nk_localop_t for ``device`` (tables uploaded once and cached)."""
        key = str(device)
        if key not in self._dev:
            t = self._pack()
            op = _lib.nk_localop_t()
            keep = []
            for g, G in enumerate(t["groups"]):
                dev = {n: torch.from_numpy(np.ascontiguousarray(G[n])).to(device)
                       for n in ("acting_on", "diag_mels", "n_conns", "mels", "x_prime")}
                keep.append(dev)
                cg = op.groups[g]
                cg.n_ops, cg.n_sites, cg.ncmax = G["n_ops"], G["n_sites"], G["ncmax"]
                cg.acting_on = dev["acting_on"].data_ptr()
                cg.diag_mels = dev["diag_mels"].data_ptr()
                cg.n_conns = dev["n_conns"].data_ptr()
                cg.mels = dev["mels"].data_ptr()
                cg.x_prime = dev["x_prime"].data_ptr()
            op.n_groups = len(t["groups"])
            op.nonzero_diagonal = int(t["nonzero_diagonal"])
            op.max_conn_size = t["max_conn_size"]
            op.constant = self._constant
            op.mel_cutoff = self._mel_cutoff
            self._dev[key] = (op, keep)
        return self._dev[key][0]

    # ---------------------------------------------------------------- device calls
    def _conn(self, x8, want_nconn=True):
        B, N = x8.shape
        K = self.max_conn_size
        dev = x8.device
        xp = torch.empty((B, K, N), dtype=torch.int8, device=dev)
        mels = torch.empty((B, K), dtype=_lib.torch_dtype(self.dtype), device=dev)
        nconn = torch.zeros((B,), dtype=torch.int32, device=dev)
        op = self._c_struct(dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().nk_localop_conn(_lib.stream_ptr(dev), C.byref(op), _lib.ptr(x8), B, N, _lib.ptr(xp),
                                                  _lib.ptr(mels), _lib.dtype_code(self.dtype), _lib.ptr(nconn)))
        return xp, mels, nconn

    def _n_conn(self, x8):
        return self._conn(x8)[2]

    def __repr__(self):
        return (f"LocalOperatorJax(dim={self.hilbert.size}, #acting_on={self.n_operators} locations, "
                f"constant={self._constant}, dtype={self.dtype})")


LocalOperator = LocalOperatorJax


def GraphOperator(hilbert, graph, site_ops=[], bond_ops=[], bond_ops_colors=[], dtype=None, *, cls=LocalOperatorJax):
    """Sum of site / bond operators over a graph (netket/operator/_graph_operator.py:47-145): site terms first
    (node order), then one bond term per edge in ``graph.edges()`` order, matched by colour if colours are given."""
    if len(bond_ops) == 0 and len(site_ops) == 0:
        raise ValueError("Must input at least site_ops or bond_ops.")
    operators, acting_on = [], []
    for i in range(graph.n_nodes if len(site_ops) > 0 else 0):
        for op in site_ops:
            operators.append(np.asarray(op))
            acting_on.append([i])
    if len(bond_ops_colors) > 0:
        if len(bond_ops) != len(bond_ops_colors):
            raise ValueError("The GraphHamiltonian definition is inconsistent. "
                             "The sizes of bond_ops and bond_ops_colors do not match.")
        for u, v, color in graph.edges(return_color=True):
            for c, bc in enumerate(bond_ops_colors):
                if bc == color:
                    operators.append(np.asarray(bond_ops[c]))
                    acting_on.append([u, v])
    elif len(bond_ops) > 0:
        assert len(bond_ops) == 1
        for u, v in graph.edges():
            operators.append(np.asarray(bond_ops[0]))
            acting_on.append([u, v])
    return cls(hilbert, operators, acting_on, dtype=dtype)


def Heisenberg(hilbert, graph, J=1.0, sign_rule=None, dtype=None, *, cls=LocalOperatorJax):
    """Heisenberg Hamiltonian sum_b J_b (sx sx + sy sy + sz sz) in Pauli convention; with Marshall's sign rule the
    exchange part changes sign (netket/operator/_heisenberg.py:33-132)."""
    from ..graph import Graph

    sz_sz = np.diag([1.0, -1.0, -1.0, 1.0])
    exchange = np.zeros((4, 4))
    exchange[1, 2] = exchange[2, 1] = 2.0
    if isinstance(J, (list, tuple, np.ndarray)):
        assert len(J) == max(graph.edge_colors) + 1
        if sign_rule is None:
            sign_rule = [False] * len(J)
        else:
            assert len(sign_rule) == len(J)
            for i in range(len(J)):
                if sign_rule[i] and not Graph(graph.edges(filter_color=i), n_nodes=graph.n_nodes).is_bipartite():
                    raise ValueError("sign_rule=True specified for a non-bipartite lattice")
        bond_ops = [J[i] * (sz_sz - exchange if sign_rule[i] else sz_sz + exchange) for i in range(len(J))]
        colors = list(range(len(J)))
    else:
        if sign_rule is None:
            sign_rule = graph.is_bipartite()
        elif sign_rule and not graph.is_bipartite():
            raise ValueError("sign_rule=True specified for a non-bipartite lattice")
        bond_ops = [J * (sz_sz - exchange if sign_rule else sz_sz + exchange)]
        colors = []
    return GraphOperator(hilbert, graph, bond_ops=bond_ops, bond_ops_colors=colors, dtype=dtype, cls=cls)
