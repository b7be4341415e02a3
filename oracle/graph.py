"""Canonical periodic hypercube lattices (TEST INFRASTRUCTURE).

The reference builds lattices with igraph and its edge order comes from Python ``set``
iteration (netket/graph/_lattice_edge_logic.py:123,135-136), i.e. it is unpinned.  The
kernels take explicit ``edges[E,2]`` arrays; this module is the canonical generator used
for synthetic inputs: sites are numbered row-major (last coordinate fastest, as
netket/graph/common_lattices.py Hypercube/Square do), every edge is stored as
``(min, max)`` and the list is sorted lexicographically; with ``max_neighbor_order=k``
edges are coloured 0..k-1 by increasing Euclidean length (``get_nn_edges``,
_lattice_edge_logic.py:102-137) and listed colour by colour.
"""

import itertools

import numpy as np


def hypercube_edges(length, n_dim=1, pbc=True, max_neighbor_order=1):
    """Returns (edges[E,2] int32, colors[E] int32)."""
    L = length
    coords = list(itertools.product(range(L), repeat=n_dim))
    index = {c: i for i, c in enumerate(coords)}
    # squared distances of the first neighbour shells on a hypercubic lattice
    offsets = [o for o in itertools.product(range(-2, 3), repeat=n_dim) if any(o)]
    shells = sorted({sum(x * x for x in o) for o in offsets})[:max_neighbor_order]
    out_e, out_c = [], []
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
                if a == b:
                    continue
                es.add((min(a, b), max(a, b)))
        es = sorted(es)
        out_e += es
        out_c += [color] * len(es)
    return np.asarray(out_e, dtype=np.int32).reshape(-1, 2), np.asarray(out_c, dtype=np.int32)


def distances(n_nodes, edges):
    """All-pairs graph distances (``Graph.distances``, netket/graph/graph.py:179-180)."""
    adj = [[] for _ in range(n_nodes)]
    for a, b in np.asarray(edges):
        adj[int(a)].append(int(b))
        adj[int(b)].append(int(a))
    D = np.full((n_nodes, n_nodes), np.iinfo(np.int64).max, dtype=np.int64)
    for s in range(n_nodes):
        D[s, s] = 0
        frontier = [s]
        d = 0
        while frontier:
            d += 1
            nxt = []
            for u in frontier:
                for v in adj[u]:
                    if D[s, v] > d:
                        D[s, v] = d
                        nxt.append(v)
            frontier = nxt
    return D


def is_bipartite(n_nodes, edges):
    adj = [[] for _ in range(n_nodes)]
    for a, b in np.asarray(edges):
        adj[int(a)].append(int(b))
        adj[int(b)].append(int(a))
    color = [-1] * n_nodes
    for s in range(n_nodes):
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


def compute_clusters(n_nodes, edges, d_max=1):
    """Exchange clusters: all i<j with graph distance <= d_max, row-major sorted.

    netket/sampler/rules/exchange.py:190-205 (np.argwhere order).
    """
    D = distances(n_nodes, edges)
    cl = np.argwhere(D <= d_max)
    cl = cl[cl[:, 0] < cl[:, 1]]
    return cl.astype(np.int32)
