"""Minimal lattice/graph objects: just enough of netket.graph for the hot path's inputs.

The reference's graph package (igraph-backed, 3.1 kLoC) is out of scope; the kernels
consume ``edges[E,2]`` / ``clusters[C,2]`` arrays.  API mirrored: ``Graph(edges=, n_nodes=)``,
``Hypercube(length, n_dim, pbc, max_neighbor_order)``, ``Chain``, ``Square``, ``.edges(return_color=,
filter_color=)``, ``.edge_colors``, ``.n_nodes``, ``.n_edges``, ``.distances()``, ``.is_bipartite()``
(netket/graph/graph.py:135-180, netket/graph/common_lattices.py:143-217).

Edge order: the reference's order comes from Python ``set`` iteration
(netket/graph/_lattice_edge_logic.py:123,135-136) and is therefore unpinned; here edges are
``(min, max)`` pairs sorted lexicographically, colour by colour.
"""

import itertools

import numpy as np


class Graph:
    def __init__(self, edges, n_nodes=None, edge_colors=None):
        edges = [tuple(int(x) for x in e[:2]) for e in edges]
        if edge_colors is None:
            edge_colors = [0] * len(edges)
        self._edges = edges
        self._colors = [int(c) for c in edge_colors]
        if n_nodes is None:
            n_nodes = 1 + max((max(e) for e in edges), default=-1)
        self._n_nodes = int(n_nodes)
        self._dist = None

    @property
    def n_nodes(self):
        return self._n_nodes

    @property
    def n_edges(self):
        return len(self._edges)

    def nodes(self):
        return range(self._n_nodes)

    @property
    def edge_colors(self):
        return list(self._colors)

    def edges(self, *, return_color=False, filter_color=None):
        out = []
        for e, c in zip(self._edges, self._colors):
            if filter_color is not None and c != filter_color:
                continue
            out.append((*e, c) if return_color else e)
        return out

    def _adjacency(self):
        adj = [[] for _ in range(self._n_nodes)]
        for a, b in self._edges:
            adj[a].append(b)
            adj[b].append(a)
        return adj

    def distances(self):
        if self._dist is None:
            n = self._n_nodes
            adj = self._adjacency()
            D = np.full((n, n), np.iinfo(np.int64).max, dtype=np.int64)
            for s in range(n):
                D[s, s] = 0
                frontier, d = [s], 0
                while frontier:
                    d += 1
                    nxt = []
                    for u in frontier:
                        for v in adj[u]:
                            if D[s, v] > d:
                                D[s, v] = d
                                nxt.append(v)
                    frontier = nxt
            self._dist = D
        return self._dist

    def is_bipartite(self):
        adj = self._adjacency()
        color = [-1] * self._n_nodes
        for s in range(self._n_nodes):
            if color[s] >= 0:
                continue
            color[s] = 0
            stack = [s]
            while stack:
                u = stack.pop()
                for v in adj[u]:
                    if color[v] < 0:
                        color[v] = 1 - color[u]
                        stack.append(v)
                    elif color[v] == color[u]:
                        return False
        return True

    def __repr__(self):
        return f"{type(self).__name__}(n_nodes={self.n_nodes}, n_edges={self.n_edges})"


class Hypercube(Graph):
    """Periodic/open hypercubic lattice, sites numbered row-major (last coordinate fastest)."""

    def __init__(self, length, n_dim=1, *, pbc=True, max_neighbor_order=1):
        if length < 1 or n_dim < 1:
            raise ValueError("length and n_dim must be positive")
        if pbc and length <= 2 and length > 1:
            raise ValueError("periodic lattices need length > 2 (netket/graph/lattice.py)")
        L = int(length)
        coords = list(itertools.product(range(L), repeat=n_dim))
        index = {c: i for i, c in enumerate(coords)}
        offsets = [o for o in itertools.product(range(-2, 3), repeat=n_dim) if any(o)]
        shells = sorted({sum(x * x for x in o) for o in offsets})[:max_neighbor_order]
        all_e, all_c = [], []
        for color, d2 in enumerate(shells):
            es = set()
            for o in offsets:
                if sum(x * x for x in o) != d2:
                    continue
                for c in coords:
                    t = tuple(ci + oi for ci, oi in zip(c, o))
                    if pbc:
                        t = tuple(x % L for x in t)
                    elif any(x < 0 or x >= L for x in t):
                        continue
                    a, b = index[c], index[t]
                    if a != b:
                        es.add((min(a, b), max(a, b)))
            es = sorted(es)
            all_e += es
            all_c += [color] * len(es)
        super().__init__(all_e, n_nodes=len(coords), edge_colors=all_c)
        self.length, self.n_dim, self.pbc = L, n_dim, pbc


def Chain(length, *, pbc=True, max_neighbor_order=1):
    return Hypercube(length, 1, pbc=pbc, max_neighbor_order=max_neighbor_order)


def Square(length, *, pbc=True, max_neighbor_order=1):
    return Hypercube(length, 2, pbc=pbc, max_neighbor_order=max_neighbor_order)
