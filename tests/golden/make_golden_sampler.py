"""Generates tests/golden/sampler_vectors.npz by EXECUTING THE REFERENCE'S OWN SOURCE of the Metropolis loop and of the
force estimator, with injected random draws.

Runs only in the build container (needs /root/reference).  What is executed, unchanged, pulled out of the files with `ast`:

  * `MetropolisSampler._sample_next` incl. its `loop_body`           netket/sampler/metropolis.py:416-464
  * `LocalRule.transition`                                            netket/sampler/rules/local.py:40-49
  * `flip_state`, `flip_state_batch`, `flip_state_scalar`             netket/hilbert/random/base.py:91-124, homogeneous.py:146-170
  * `ExchangeRule.transition`, `_compute_different_clusters_mask`     netket/sampler/rules/exchange.py:143-184,208-218
    (with and without `probabilities`)
  * `forces_expect_hermitian`                                         netket/vqs/mc/mc_state/expect_forces.py:67-112
  * `_statistics`, `log_cosh`, `StaticRange` maps                     (as in make_golden.py)

on tests/golden/jnp_shim.py plus a stand-in for `jax.random` below.  JAX's threefry bit stream cannot be reproduced without
JAX, so the stand-in does NOT generate randomness: every `randint` / `uniform` / `choice` call returns the draw of an
injected stream (the Philox proposal stream of oracle/rng.py).  `choice(key, a, p)` re-states JAX's algorithm for
`replace=True` (jax/_src/random.py: `p_cuml = cumsum(p); r = p_cuml[-1] * (1 - uniform(key)); ind = searchsorted(p_cuml, r)`)
and is fed the uniform `1 - (k + 1/2) / n_hop` for the unweighted rule, which makes it return the k-th hoppable cluster in
cluster order, k = floor(w0 n_hop / 2^32); for the weighted rule it is fed `1 - (w0 + 1/2) / 2^32`.
`nkjax.vjp` of the RBM is a central-difference (Richardson) derivative of the reference's own forward pass.

    python tests/golden/make_golden_sampler.py
"""

import ast
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from make_golden_lib import REF, base_ns, extract, jax, jnp  # noqa: E402

from oracle import graph as ograph  # noqa: E402  (canonical synthetic edge lists: inputs, not reference logic)
from oracle import rng as orng  # noqa: E402  (the injected draws)

out = {}
rs = np.random.default_rng(20241017)


# ------------------------------------------------------------------------------------------------ jax.random stand-in
class Key:
    """Carries which draw of the injected stream a jax.random call stands for."""

    def __init__(self, stream, role, t, chain=None):
        self.stream, self.role, self.t, self.chain = stream, role, t, chain


class Stream:
    def __init__(self, w0, u):
        self.w0, self.u = np.asarray(w0, dtype=np.uint32), np.asarray(u)  # [T, B]
        self.weighted = False


def r_split(key, num=2):
    if key.role == "root":  # metropolis.py:429: new_rng, key1 (rule), key2 (uniform)
        assert num == 3
        return [Key(key.stream, "root", key.t + 1), Key(key.stream, "rule", key.t), Key(key.stream, "accept", key.t)]
    if key.role == "rule" and num == 2:  # rules/local.py:41: key1 (site), key2 (flip)
        return [Key(key.stream, "site", key.t), Key(key.stream, "flip", key.t)]
    if key.role == "rule":  # rules/exchange.py:149: one key per chain
        ks = np.empty(num, dtype=object)
        for b in range(num):
            ks[b] = Key(key.stream, "cluster", key.t, b)
        return ks
    if key.role == "flip":  # hilbert/random/base.py:120: one key per chain; unused for 2 local states
        ks = np.empty(num, dtype=object)
        for b in range(num):
            ks[b] = Key(key.stream, "flip1", key.t, b)
        return ks
    raise AssertionError(key.role)


def r_randint(key, shape, minval, maxval):
    assert key.role == "site" and minval == 0
    return jnp.asarray(orng.index_from_word(key.stream.w0[key.t], maxval).astype(np.int64))


def r_uniform(key, shape=(), **kw):
    assert key.role == "accept", key.role
    return jnp.asarray(key.stream.u[key.t])


def r_choice(key, a, p=None, replace=True, **kw):
    assert key.role == "cluster" and replace
    p = np.asarray(p, dtype=np.float64)
    w0 = int(key.stream.w0[key.t, key.chain])
    if key.stream.weighted:
        v = 1.0 - (w0 + 0.5) / 2.0 ** 32
    else:
        n_hop = int(round(p.sum()))
        k = (w0 * max(n_hop, 1)) >> 32
        v = 1.0 - (k + 0.5) / max(n_hop, 1)
    p_cuml = np.cumsum(p)  # jax/_src/random.py, choice(..., replace=True)
    r = p_cuml[-1] * (1.0 - v)
    ind = int(np.searchsorted(p_cuml, r))
    return np.asarray(a)[min(ind, len(p) - 1)]


random = types.ModuleType("jax.random")
random.split, random.randint, random.uniform, random.choice = r_split, r_randint, r_uniform, r_choice
jax.random = random
jax.lax.fori_loop = lambda lo, hi, body, init: __import__("functools").reduce(lambda s, i: body(i, s), range(lo, hi), init)
jax.lax.collapse = lambda x, a, b: jnp.asarray(np.asarray(x).reshape(x.shape[:a] + (-1,) + x.shape[b:]))
jnp.isclose = lambda a, b: np.isclose(a, b)


def _asarray(x, **kw):
    if isinstance(x, np.ndarray) and x.dtype == object:
        return x
    return np.asarray(x, **kw).view(type(jnp.zeros(1)))


jnp.asarray = _asarray

# ------------------------------------------------------------------------------------------------ reference pieces
ns_act = extract("nn/activation.py", ["log_cosh"], base_ns())
log_cosh = ns_act["log_cosh"]
ns_sr = extract("utils/static_range.py", ["states_to_numbers", "numbers_to_states"], {**base_ns(), "DType": object,
                                                                                    "bottom_int_dtype": lambda n: np.uint8},
                class_name="StaticRange")
spin_range = types.SimpleNamespace(start=1, step=-2, length=2, dtype=np.int8)  # Spin(1/2): netket/hilbert/spin.py:165-171


class Hilbert:  # the two maps and the two attributes the rules use (netket/hilbert/homogeneous.py)
    def __init__(self, N):
        self.size = N
        self._local_states = spin_range

    def states_to_local_indices(self, x):
        return jnp.asarray(np.asarray(ns_sr["states_to_numbers"](spin_range, np.asarray(x))))

    def local_indices_to_states(self, i, dtype=np.int8):
        return jnp.asarray(np.asarray(ns_sr["numbers_to_states"](spin_range, np.asarray(i), dtype=dtype)))


class _Len2:
    start, step, length, dtype = 1, -2, 2, np.int8

    def __len__(self):
        return 2


def rbm_apply(pars, sigma):  # flax nn.Dense is x @ kernel + bias (netket/models/rbm.py:57-81 composes it with log_cosh and the visible bias)
    W_, b_, a_ = pars
    s = np.asarray(sigma, dtype=W_.dtype)
    return jnp.asarray(np.asarray(log_cosh(jnp.asarray(s @ W_ + b_))).sum(axis=-1) + s @ a_)


class State:  # MetropolisSamplerState: the fields loop_body touches, functional `replace`
    def __init__(self, **kw):
        self.__dict__.update(kw)

    def replace(self, **kw):
        return State(**{**self.__dict__, **kw})


def dispatch(f):
    return f


ns_flip = extract("hilbert/random/base.py", ["flip_state", "flip_state_batch"], {**base_ns(), "dispatch": dispatch, "HomogeneousHilbert": object})
extract("hilbert/random/homogeneous.py", ["flip_state_scalar"], ns_flip)  # same namespace: flip_state_batch vmaps over it
ns_local = extract("sampler/rules/local.py", ["transition"], {**base_ns(), "flip_state": ns_flip["flip_state"]}, class_name="LocalRule")
ns_ex = extract("sampler/rules/exchange.py", ["_compute_different_clusters_mask"], {**base_ns(), "AbstractGraph": object})
ns_ex = extract("sampler/rules/exchange.py", ["transition"], ns_ex, class_name="ExchangeRule")
ns_mh = extract("sampler/metropolis.py", ["_sample_next"],
                {**base_ns(), "apply_chunked": lambda f, in_axes, chunk_size: f, "_assert_good_sample_shape": lambda *a: None,
                 "_assert_good_log_prob_shape": lambda *a: None}, class_name="MetropolisSampler")


class Rule:
    def __init__(self, fn, **kw):
        self._fn = fn
        self.__dict__.update(kw)

    def transition(self, *a):
        return self._fn(self, *a)


def run_reference(rule, sig0, pars, stream, n_sweeps, sweep_size, machine_pow):
    B, N = sig0.shape
    hilb = Hilbert(N)
    hilb._local_states = _Len2()
    sampler = types.SimpleNamespace(rule=rule, hilbert=hilb, n_batches=B, dtype=np.int8, machine_pow=machine_pow, sweep_size=sweep_size,
                                    chunk_size=None)
    machine = types.SimpleNamespace(apply=rbm_apply)
    state = State(σ=jnp.asarray(sig0), log_prob=jnp.asarray(machine_pow * np.asarray(rbm_apply(pars, sig0))), rng=Key(stream, "root", 0),
                  n_accepted_proc=jnp.asarray(np.zeros(B, dtype=np.int64)), n_steps_proc=0)
    samples, logps = [], []
    for _ in range(n_sweeps):
        state, (s, lp) = ns_mh["_sample_next"](sampler, machine, pars, state)
        samples.append(np.asarray(s).copy())
        logps.append(np.asarray(lp).copy())
    return (np.stack(samples, axis=1).astype(np.int8), np.stack(logps, axis=1), np.asarray(state.n_accepted_proc).astype(np.int64),
            int(state.n_steps_proc))


def case(tag, rule_kind, L, n_dim, alpha, std, B, n_sweeps, *, total_sz=None, d_max=1, machine_pow=2.0, sweep_size=None, probs=None, seed=3):
    N = L ** n_dim
    M = alpha * N
    W, b, a = rs.normal(size=(N, M)) * std, rs.normal(size=M) * std, rs.normal(size=N) * std
    if total_sz is None:
        sig0 = (1 - 2 * rs.integers(0, 2, size=(B, N))).astype(np.int8)
    else:
        base = np.array([1] * (N // 2 + total_sz) + [-1] * (N - N // 2 - total_sz), dtype=np.int8)
        sig0 = np.stack([rs.permutation(base) for _ in range(B)])
    sweep_size = N if sweep_size is None else sweep_size
    T = n_sweeps * sweep_size
    words, u = orng.proposal_stream(seed, 0, T, np.arange(B), np.float64)
    stream = Stream(words[..., 0], u)
    out[f"{tag}_cfg"] = np.array([L, n_dim, alpha, B, n_sweeps, sweep_size, d_max, -99 if total_sz is None else total_sz], dtype=np.int64)
    out[f"{tag}_pow"] = np.array(machine_pow)
    if rule_kind == "local":
        rule = Rule(ns_local["transition"])
    else:
        e, _ = ograph.hypercube_edges(L, n_dim)
        clusters = ograph.compute_clusters(N, e, d_max)
        p = None
        if probs is not None:
            D = ograph.distances(N, e)
            p = np.asarray(probs, dtype=np.float64)[D[clusters[:, 0], clusters[:, 1]] - 1]  # exchange.py:201-203
            out[f"{tag}_probs"] = p
            stream.weighted = True
        rule = Rule(ns_ex["transition"], clusters=jnp.asarray(clusters), probabilities=None if p is None else jnp.asarray(p))
        out[f"{tag}_clusters"] = clusters
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        samples, logp, nacc, nsteps = run_reference(rule, sig0, (W, b, a), stream, n_sweeps, sweep_size, machine_pow)
    for k, v in dict(W=W, b=b, a=a, sigma0=sig0, w0=stream.w0, u=stream.u, samples=samples, logp=logp, nacc=nacc,
                     nsteps=np.array(nsteps)).items():
        out[f"{tag}_{k}"] = v
    print(f"{tag}: acceptance {nacc.sum() / nsteps:.3f}")


case("mh_local_1d", "local", 12, 1, 2, 0.3, 8, 3)
case("mh_local_2d", "local", 4, 2, 3, 0.2, 6, 2, machine_pow=1.0)
case("mh_local_sweep5", "local", 10, 1, 1, 0.5, 5, 4, sweep_size=5, machine_pow=3.0)
case("mh_exch_1d", "exchange", 12, 1, 2, 0.4, 8, 3, total_sz=0)
case("mh_exch_2d_d2", "exchange", 4, 2, 2, 0.3, 6, 2, total_sz=0, d_max=2)
case("mh_exch_sz1", "exchange", 10, 1, 1, 0.4, 5, 3, total_sz=1, d_max=2, machine_pow=1.0)
case("mh_exch_weighted", "exchange", 12, 1, 2, 0.4, 8, 3, total_sz=0, d_max=2, probs=[0.7, 0.3])
case("mh_exch_weighted3", "exchange", 4, 2, 2, 0.3, 6, 3, total_sz=0, d_max=3, probs=[0.5, 0.3, 0.2])

# ------------------------------------------------------------------------------------------------ forces (expect_forces.py:67-112)
Stats = lambda mean, err, var, tau, rhat: types.SimpleNamespace(mean=mean, error_of_mean=err, variance=var, tau_corr=tau, R_hat=rhat)  # noqa: E731
ns_st = extract("stats/mc_stats_old.py", ["_get_blocks", "_block_variance", "_batch_variance", "_statistics"],
                {**base_ns(), "Stats": Stats, "config": types.SimpleNamespace(netket_use_plain_rhat=False),
                 "nkjax": types.SimpleNamespace(dtype_real=lambda dt: np.dtype(np.float64))})
ns_is = extract("operator/_ising/jax.py", ["_ising_mels_jax", "_ising_conn_states_jax", "_ising_kernel_jax"],
                {**base_ns(), "StaticZero": type("StaticZero", (), {})})
ns_lv = extract("vqs/mc/kernels.py", ["local_value_kernel_jax"], {**base_ns(), "Callable": object, "PyTree": object, "Array": object,
                                                                  "DiscreteJaxOperator": object})
hil = Hilbert(1)


class IsingOp:
    def __init__(self, edges, h, J):
        self.edges, self.h, self.J = edges, h, J

    def get_conn_padded(self, x):
        xp_ids, mels = ns_is["_ising_kernel_jax"](hil.states_to_local_indices(np.asarray(x)), jnp.asarray(self.edges), jnp.array(self.h),
                                                  jnp.array(self.J))
        return hil.local_indices_to_states(np.asarray(xp_ids)), mels


def fd_vjp(fun, params, conjugate=True, has_aux=False):
    """nkjax.vjp for a real function of real parameters: cotangent -> sum_s w_s d f_s / d p, by Richardson-extrapolated
    central differences of the reference's own forward pass (error ~1e-10 at these sizes)."""
    flat, shapes = [], []
    for k in ("kernel", "bias", "visible_bias"):
        flat.append(np.asarray(params[k], dtype=np.float64).ravel())
        shapes.append((k, np.shape(params[k])))
    x0 = np.concatenate(flat)

    def unflat(x):
        d, o = {}, 0
        for k, shp in shapes:
            n = int(np.prod(shp))
            d[k] = x[o:o + n].reshape(shp)
            o += n
        return d

    def vjp_fun(w):
        w = np.asarray(w, dtype=np.float64)
        g = np.zeros_like(x0)
        for i in range(x0.size):
            def d(h):
                xp, xm = x0.copy(), x0.copy()
                xp[i] += h
                xm[i] -= h
                return (np.asarray(fun(unflat(xp))) - np.asarray(fun(unflat(xm)))) / (2 * h)
            h = 1e-3
            g[i] = np.dot(w, (4.0 * d(h / 2) - d(h)) / 3.0)
        return (unflat(g),)

    return fun(unflat(x0)), vjp_fun


def model_apply(variables, sigma, mutable=False):
    p = variables["params"]
    return rbm_apply((np.asarray(p["kernel"]), np.asarray(p["bias"]), np.asarray(p["visible_bias"])), sigma)


jnp.conjugate = lambda x: jnp.asarray(np.conjugate(np.asarray(x)))
ns_f = extract("vqs/mc/mc_state/expect_forces.py", ["forces_expect_hermitian"],
               {**base_ns(), "Callable": object, "PyTree": object, "CollectionFilter": object, "statistics": lambda x: ns_st["_statistics"](x, 32),
                "nkjax": types.SimpleNamespace(vjp=fd_vjp)})
for tag, (L, nd, alpha, std, nch, cl) in {"forces_1d": (6, 1, 2, 0.3, 8, 6), "forces_2d": (3, 2, 1, 0.2, 4, 9)}.items():
    N = L ** nd
    M = alpha * N
    Wf, bf, af = rs.normal(size=(N, M)) * std, rs.normal(size=M) * std, rs.normal(size=N) * std
    edges, _ = ograph.hypercube_edges(L, nd)
    sig = (1 - 2 * rs.integers(0, 2, size=(nch, cl, N))).astype(np.int8)
    op = IsingOp(edges, 1.3, 1.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Obar, grad, _ = ns_f["forces_expect_hermitian"](ns_lv["local_value_kernel_jax"], model_apply, False,
                                                         {"kernel": Wf, "bias": bf, "visible_bias": af}, {}, jnp.asarray(sig), op)
    out[f"{tag}_cfg"] = np.array([L, nd, alpha, nch, cl])
    out[f"{tag}_h"] = np.array(1.3)
    out[f"{tag}_W"], out[f"{tag}_b"], out[f"{tag}_a"], out[f"{tag}_sigma"], out[f"{tag}_edges"] = Wf, bf, af, sig, edges
    out[f"{tag}_mean"] = np.array(float(Obar.mean))
    out[f"{tag}_F_kernel"], out[f"{tag}_F_bias"], out[f"{tag}_F_visible"] = (np.asarray(grad["kernel"]), np.asarray(grad["bias"]),
                                                                         np.asarray(grad["visible_bias"]))
    print(f"{tag}: <E> = {float(Obar.mean):.6f}, |F| = {np.linalg.norm(np.asarray(grad['kernel'])):.4f}")

path = os.path.join(HERE, "sampler_vectors.npz")
np.savez_compressed(path, **out)
print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")
