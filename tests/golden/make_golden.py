"""Generates tests/golden/reference_vectors.npz by EXECUTING THE REFERENCE'S OWN SOURCE for the hot-path functions.

Runs only in the build container (needs /root/reference; the GPU box does not have it — tests read the committed
.npz).  `import netket` is impossible here (jax/flax/plum are not installed), so the function definitions are pulled out
of the reference files with `ast`, unchanged, and executed with

  * real numpy / numba / scipy for the jax-free pieces (numba table packing of compile_helpers.py), and
  * tests/golden/jnp_shim.py standing in for jax.numpy / jax.vmap / jax.lax for the jax kernels.

    python tests/golden/make_golden.py
"""

import ast
import os
import sys
import types
import warnings
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import jnp_shim  # noqa: E402

from make_golden_lib import REF, base_ns, extract, jax, jnp  # noqa: E402

out = {}
rs = np.random.default_rng(20240901)

# ------------------------------------------------------------------ log_cosh  (netket/nn/activation.py:78-84)
ns = extract("nn/activation.py", ["log_cosh"], base_ns())
x = np.concatenate([np.linspace(-40, 40, 161), rs.normal(size=64) * 3])
out["log_cosh_x"] = x
out["log_cosh_y"] = np.asarray(ns["log_cosh"](x))
out["log_cosh_y32"] = np.asarray(ns["log_cosh"](x.astype(np.float32)))

# ------------------------------------------------------------------ StaticRange maps (netket/utils/static_range.py:148-192)
ns_sr = extract("utils/static_range.py", ["states_to_numbers", "numbers_to_states"], {**base_ns(), "DType": object,
                                                                                    "bottom_int_dtype": lambda n: np.uint8},
                class_name="StaticRange")
spin_range = types.SimpleNamespace(start=1, step=-2, length=2, dtype=np.int8)  # Spin(1/2): netket/hilbert/spin.py:165-171
to_idx = lambda x: np.asarray(ns_sr["states_to_numbers"](spin_range, np.asarray(x)))  # noqa: E731
to_state = lambda i, dtype=np.int8: np.asarray(ns_sr["numbers_to_states"](spin_range, np.asarray(i), dtype=dtype))  # noqa: E731
out["spin_states"] = np.array([1, -1], dtype=np.int8)
out["spin_indices"] = to_idx(np.array([1, -1], dtype=np.int8))


def pbc_edges(L, n_dim, order=1):
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import graph as ograph  # canonical synthetic edge lists (inputs, not reference logic)

    return ograph.hypercube_edges(L, n_dim, True, order)


# ------------------------------------------------------------------ Ising kernels (netket/operator/_ising/jax.py:125-175)
class StaticZero:  # netket/utils/numbers.py: marker type used when h == 0
    pass


ns_is = extract("operator/_ising/jax.py", ["_ising_mels_jax", "_ising_conn_states_jax", "_ising_kernel_jax", "_ising_n_conn_jax"],
                {**base_ns(), "StaticZero": StaticZero})
for tag, (L, nd, h, J) in {"ising1d": (10, 1, 1.321, 1.0), "ising2d": (4, 2, 3.0, 1.0), "ising_h0": (6, 1, 0.0, 2.0),
                           "ising_negJ": (5, 1, 0.7, -0.5)}.items():
    edges, _ = pbc_edges(L, nd)
    N = L ** nd
    sig = (1 - 2 * rs.integers(0, 2, size=(40, N))).astype(np.int8)
    sig[0] = 1
    sig[1] = -1
    hh = StaticZero() if h == 0 else jnp.array(h, dtype=np.float64)
    xp_ids, mels = ns_is["_ising_kernel_jax"](jnp.asarray(to_idx(sig)), jnp.asarray(edges), hh, jnp.array(J, dtype=np.float64))
    nconn = ns_is["_ising_n_conn_jax"](jnp.asarray(to_idx(sig)), jnp.asarray(edges), hh, jnp.array(J, dtype=np.float64))
    out[f"{tag}_cfg"] = np.array([L, nd, h, J])
    out[f"{tag}_edges"] = edges
    out[f"{tag}_sigma"] = sig
    out[f"{tag}_xp"] = to_state(np.asarray(xp_ids))
    out[f"{tag}_mels"] = np.asarray(mels)
    out[f"{tag}_nconn"] = np.asarray(nconn)


# ------------------------------------------------------------------ Heisenberg bond matrices (netket/operator/_heisenberg.py:97-122)
def heisenberg_literals():
    tree = ast.parse(open(os.path.join(REF, "operator/_heisenberg.py")).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "Heisenberg"][0]
    vals = {}
    for node in ast.walk(fn):
        if isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name) and node.targets[0].id in ("sz_sz", "exchange"):
            vals[node.targets[0].id] = eval(compile(ast.Expression(node.value), "<heis>", "eval"), {"np": np})
    return vals["sz_sz"], vals["exchange"]


sz_sz, exchange = heisenberg_literals()
out["heis_sz_sz"], out["heis_exchange"] = sz_sz, exchange

# ------------------------------------------------------------------ table packing (numba; compile_helpers.py:29-366)
import numba  # noqa: E402
from scipy import sparse  # noqa: E402

ns_ph = extract("operator/_local_operator/compile_helpers.py",
                ["pack_internals", "pack_internals_jax", "_append_matrix", "_append_matrix_sparse", "_number_to_state",
                 "max_nonzero_per_row"],
                {**base_ns(), "numba": numba, "sparse": sparse, "AbstractHilbert": object, "DType": object})
ns_lo = extract("operator/_local_operator/jax.py", ["_state_to_number", "_set_at", "_index_at", "_local_operator_kernel_jax"], base_ns())
ns_lv = extract("vqs/mc/kernels.py", ["local_value_kernel_jax"], {**base_ns(), "Callable": object, "PyTree": object, "Array": object,
                                                                  "DiscreteJaxOperator": object})


class FakeHilbert:
    def __init__(self, N):
        self.size = N
        self.shape = [2] * N

    def size_at_index(self, i):
        return 2


def localop_case(tag, N, operators_dict, constant=0.0, total_sz=None, n_samples=48):
    data = ns_ph["pack_internals_jax"](FakeHilbert(N), operators_dict, constant, np.float64, 1e-10)
    if total_sz is None:
        sig = (1 - 2 * rs.integers(0, 2, size=(n_samples, N))).astype(np.int8)
    else:
        base = np.array([1] * (N // 2 + total_sz) + [-1] * (N - N // 2 - total_sz), dtype=np.int8)
        sig = np.stack([rs.permutation(base) for _ in range(n_samples)])
    sig[0] = np.sort(sig[0])[::-1]
    op_args = (data["acting_on"], data["n_conns"], data["diag_mels"], data["x_prime"], data["mels"], data["basis"],
               np.asarray(constant, dtype=np.float64))
    xp_ids, mels, nconn = ns_lo["_local_operator_kernel_jax"](bool(data["nonzero_diagonal"]), int(data["max_conn_size"]), 1e-10, op_args,
                                                              jnp.asarray(to_idx(sig).astype(np.int64)))
    out[f"{tag}_sigma"] = sig
    out[f"{tag}_xp"] = to_state(np.asarray(xp_ids))
    out[f"{tag}_mels"] = np.asarray(mels)
    out[f"{tag}_nconn"] = np.asarray(nconn)
    out[f"{tag}_K"] = np.array(int(data["max_conn_size"]))
    out[f"{tag}_nonzero_diagonal"] = np.array(bool(data["nonzero_diagonal"]))
    for g in range(len(data["acting_on"])):
        for name in ("acting_on", "n_conns", "diag_mels", "x_prime", "mels", "basis"):
            out[f"{tag}_g{g}_{name}"] = np.asarray(data[name][g])
    return data, sig


# Heisenberg 1D L=10 total_sz=0 with sign rule (bipartite) — the reference's operator zoo (test/operator/test_operator.py:16-81)
e10, _ = pbc_edges(10, 1)
out["heis1d_edges"] = e10
localop_case("heis1d", 10, {tuple(map(int, e)): 1.0 * (sz_sz - exchange) for e in e10}, total_sz=0)
# J1-J2 on 4x4 (two colours, no sign rule): Examples/HeisenbergJ1J2
e44, c44 = pbc_edges(4, 2, order=2)
out["j1j2_edges"], out["j1j2_colors"] = e44, c44
Js = [1.0, 0.5]
localop_case("j1j2", 16, {tuple(map(int, e)): Js[c] * (sz_sz + exchange) for e, c in zip(e44, c44)}, total_sz=0)
# generic 1- and 2-site terms, with a constant
mats1 = [(lambda m: m + m.T)(rs.normal(size=(2, 2))) for _ in range(5)]
mats2 = []
for _ in range(4):
    m = rs.normal(size=(4, 4))
    m[np.abs(m) < 0.5] = 0.0
    mats2.append(m + m.T)
pairs = [(0, 1), (1, 3), (2, 4), (0, 4)]
od = {(i,): mats1[i] for i in range(5)}
od.update({p: m for p, m in zip(pairs, mats2)})
out["generic_mats1"] = np.stack(mats1)
out["generic_mats2"] = np.stack(mats2)
out["generic_pairs"] = np.array(pairs)
localop_case("generic", 5, od, constant=0.25)

# ------------------------------------------------------------------ local_value_kernel_jax (netket/vqs/mc/kernels.py:62-71)
Nl, Ml = 16, 32
Wl, bl, al = rs.normal(size=(Nl, Ml)) * 0.2, rs.normal(size=Ml) * 0.2, rs.normal(size=Nl) * 0.2
log_cosh = ns["log_cosh"]


def rbm_apply(pars, sigma):  # flax nn.Dense is x @ kernel + bias (netket/models/rbm.py:57-81 composes it with log_cosh and the visible bias)
    W_, b_, a_ = pars
    s = np.asarray(sigma, dtype=W_.dtype)
    return jnp.asarray(np.asarray(log_cosh(jnp.asarray(s @ W_ + b_))).sum(axis=-1) + s @ a_)


class IsingOp:
    def __init__(self, edges, h, J):
        self.edges, self.h, self.J = edges, h, J

    def get_conn_padded(self, x):
        xp_ids, mels = ns_is["_ising_kernel_jax"](jnp.asarray(to_idx(np.asarray(x))), jnp.asarray(self.edges), jnp.array(self.h),
                                                  jnp.array(self.J))
        return jnp.asarray(to_state(np.asarray(xp_ids))), mels


edges44, _ = pbc_edges(4, 2)
sig_l = (1 - 2 * rs.integers(0, 2, size=(32, Nl))).astype(np.int8)
el = ns_lv["local_value_kernel_jax"](rbm_apply, (Wl, bl, al), jnp.asarray(sig_l), IsingOp(edges44, 3.0, 1.0))
out["eloc_W"], out["eloc_b"], out["eloc_a"], out["eloc_sigma"], out["eloc_edges"] = Wl, bl, al, sig_l, edges44
out["eloc_ising_h3"] = np.asarray(el)
out["eloc_logpsi"] = np.asarray(rbm_apply((Wl, bl, al), sig_l))

# ------------------------------------------------------------------ statistics (netket/stats/mc_stats_old.py:28-196)
Stats = lambda *a: a  # noqa: E731
ns_st = extract("stats/mc_stats_old.py", ["_get_blocks", "_block_variance", "_batch_variance", "_statistics"],
                {**base_ns(), "Stats": Stats, "config": types.SimpleNamespace(netket_use_plain_rhat=False),
                 "nkjax": types.SimpleNamespace(dtype_real=lambda dt: np.dtype(np.float64))})
for tag, shape in {"stats_16x63": (16, 63), "stats_64x100": (64, 100), "stats_1x1000": (1, 1000), "stats_33x65": (33, 65),
                   "stats_40x1": (40, 1), "stats_5x7": (5, 7)}.items():
    data = rs.normal(size=shape).cumsum(axis=1) * 0.2 + rs.normal(size=shape) - 11.0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mean, err, var, tau, rhat = ns_st["_statistics"](jnp.asarray(data), 32)
    out[f"{tag}_data"] = data
    out[f"{tag}_result"] = np.array([float(mean), float(err), float(var), float(tau), float(rhat)])

# ------------------------------------------------------------------ exchange clusters (netket/sampler/rules/exchange.py:190-218)
ns_ex = extract("sampler/rules/exchange.py", ["compute_clusters", "_compute_different_clusters_mask"],
                {**base_ns(), "AbstractGraph": object})
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import graph as ograph  # noqa: E402

for tag, (L, nd, dmax) in {"clusters_chain8_d2": (8, 1, 2), "clusters_sq4_d1": (4, 2, 1)}.items():
    e, _ = pbc_edges(L, nd)
    D = ograph.distances(L ** nd, e)
    fake_graph = types.SimpleNamespace(distances=lambda D=D: D)
    cl, _ = ns_ex["compute_clusters"](fake_graph, dmax, None)
    out[f"{tag}"] = np.asarray(cl)
    out[f"{tag}_dist"] = D
sig_c = (1 - 2 * rs.integers(0, 2, size=(6, 8))).astype(np.int8)
out["clusters_mask_sigma"] = sig_c
out["clusters_mask"] = np.asarray(ns_ex["_compute_different_clusters_mask"](jnp.asarray(out["clusters_chain8_d2"]), jnp.asarray(sig_c)))

# ------------------------------------------------------------------ chain length rounding (netket/vqs/mc/mc_state/state.py:60-79)
jax.device_count = lambda: 1
ns_cl = extract("vqs/mc/mc_state/state.py", ["compute_chain_length"], base_ns())
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    out["chain_length_cases"] = np.array([[nc, ns_, ns_cl["compute_chain_length"](nc, ns_)] for nc, ns_ in
                                          [(16, 1000), (16, 1008), (32, 1), (7, 50), (65536, 2 ** 20)]])

path = os.path.join(HERE, "reference_vectors.npz")
np.savez_compressed(path, **out)
print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")
