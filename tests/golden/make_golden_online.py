"""Generates tests/golden/online_stats_vectors.npz by EXECUTING THE REFERENCE'S OWN SOURCE of the streaming statistics.

Same method as make_golden.py (build container only; jax is not installed, so the definitions are pulled out of the
reference files with `ast`, unchanged, and run on tests/golden/jnp_shim.py):

  * `_acf_core`, `_update_arrays`                    netket/_src/stats/online_stats/kernels.py:25-190
  * the properties of `OnlineStats`                  netket/_src/stats/online_stats/accumulator.py:226-447
  * `expand_max_lag`, `thin_acf_by_2`                netket/_src/stats/online_stats/operations.py:132-261
  * `acf_window_saturated`, `tau_corr_reliable`      netket/_src/vqs/check_mc_convergence.py:243-272

    python tests/golden/make_golden_online.py
"""

import copy
import math
import os
import sys
from math import isnan, sqrt

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import jnp_shim  # noqa: E402
from make_golden_lib import base_ns, extract, jax, jnp  # noqa: E402

lax = jax.lax
ns = extract("_src/stats/online_stats/kernels.py", ["_acf_core", "_update_arrays"], {**base_ns(), "lax": lax})
_update_arrays = ns["_update_arrays"]

PROPS = ["n_chains", "mean", "variance", "tau_corr", "tau_corr_batch", "tau_corr_acf", "acf", "R_hat", "_compute_error_of_mean"]
ns_p = extract("_src/stats/online_stats/accumulator.py", PROPS, {**base_ns(), "isnan": isnan, "sqrt": sqrt, "_NaN": np.nan},
               class_name="OnlineStats")


class Est:
    """Plain-attribute stand-in for the struct.Pytree container: the reference's own property bodies run on it."""

    FIELDS = ("_chain_count", "_chain_mean", "_chain_M2", "_cross_sum", "_m1_sum", "_m2_sum", "_pair_count", "_chain_buf")

    def __init__(self, n_chains, dtype, decay=None, max_lag=64):  # shapes of accumulator.py:96-131
        acf_len = max_lag + 1 if max_lag > 0 else 0
        self.max_lag, self._decay = max_lag, decay
        self._chain_count = jnp.zeros(n_chains, dtype=np.float64)
        self._chain_mean = jnp.zeros(n_chains, dtype=dtype)
        self._chain_M2 = jnp.zeros(n_chains, dtype=np.float64)
        for f in ("_cross_sum", "_m1_sum", "_m2_sum", "_pair_count"):
            setattr(self, f, jnp.zeros((n_chains, acf_len), dtype=np.float64))
        self._chain_buf = jnp.zeros((n_chains, max_lag), dtype=np.float64)
        self._buf_len = jnp.array(0, dtype=np.int32)
        self._n_samples_total = 0

    def replace(self, **kw):
        new = copy.copy(self)
        for k, v in kw.items():
            setattr(new, k, v)
        return new

    def update(self, data):  # the call of accumulator.py:181-221, source of _update_arrays unchanged
        data = jnp.asarray(data)
        res = _update_arrays(self._chain_count, self._chain_mean, self._chain_M2, self._cross_sum, self._m1_sum, self._m2_sum,
                             self._pair_count, self._chain_buf, self._decay, self.max_lag, self._buf_len, data)
        new = self.replace(**dict(zip(self.FIELDS, res)))
        new._buf_len = jnp.minimum(self._buf_len + data.shape[1], self.max_lag)
        new._n_samples_total = self._n_samples_total + data.shape[0] * data.shape[1]
        return new


for name in PROPS[:-1]:
    setattr(Est, name, property(ns_p[name]))
Est._compute_error_of_mean = ns_p["_compute_error_of_mean"]

ns_o = extract("_src/stats/online_stats/operations.py", ["expand_max_lag", "thin_acf_by_2"], {**base_ns(), "OnlineStats": Est})
ns_c = extract("_src/vqs/check_mc_convergence.py", ["acf_window_saturated", "tau_corr_reliable"], {**base_ns(), "math": math})
ns_c["acf_window_saturated"].__globals__["acf_window_saturated"] = ns_c["acf_window_saturated"]


def summary(e):
    acf = e.acf
    var = e.variance
    vals = [float(e.mean), float(var), float(e.tau_corr), float(e.tau_corr_batch), float(e.tau_corr_acf), float(e.R_hat),
            float(e._compute_error_of_mean(var)), float(e._n_samples_total), float(ns_c["acf_window_saturated"](e)),
            float(ns_c["tau_corr_reliable"](e))]
    return np.array(vals), (np.full(0, np.nan) if acf is None else np.asarray(acf))


def dump(out, tag, e, state=True):
    for f in Est.FIELDS if state else ():
        out[f"{tag}{f}"] = np.asarray(getattr(e, f))
    out[f"{tag}_buf_len"] = np.asarray(int(e._buf_len))
    out[f"{tag}_summary"], out[f"{tag}_acf"] = summary(e)


def ar1(rs, n_chains, n, phi, offset=-3.0, scale=1.0):
    x = np.zeros((n_chains, n))
    x[:, 0] = rs.normal(size=n_chains)
    for t in range(1, n):
        x[:, t] = phi * x[:, t - 1] + math.sqrt(1 - phi * phi) * rs.normal(size=n_chains)
    return offset + scale * x + 0.3 * rs.normal(size=(n_chains, 1))


out = {}
rs = np.random.default_rng(20240917)
# (tag, n_chains, max_lag, decay, batch lengths, phi, dtype): batches shorter and longer than max_lag, a single chain,
# max_lag = 0, EMA decay (thermalise_mcmc: decay 0.9, max_lag 0), float32 data.
CASES = [
    ("a", 16, 64, None, [16, 16, 16, 16, 100, 3, 70], 0.6, np.float64),
    ("b", 5, 8, None, [3, 1, 2, 30, 8, 9], 0.8, np.float64),
    ("c", 1, 32, None, [200, 17, 300], 0.5, np.float64),
    ("d", 12, 0, 0.9, [8, 8, 8, 8, 8], 0.3, np.float64),
    ("e", 7, 6, 0.8, [4, 9, 5], 0.7, np.float64),
    ("f", 33, 64, None, [16, 16, 16], 0.4, np.float32),
    ("g", 2, 5, None, [1, 1, 1, 1, 1, 1, 1, 1], 0.9, np.float64),
]
cases = []
for tag, nc, L, decay, lens, phi, dt in CASES:
    data = ar1(rs, nc, sum(lens), phi).astype(dt)
    e = Est(nc, dt, decay=decay, max_lag=L)
    pos = 0
    for i, n in enumerate(lens):
        e = e.update(data[:, pos:pos + n])
        pos += n
        dump(out, f"{tag}_s{i}", e, state=(i == len(lens) - 1 or i == 1))
    out[f"{tag}_data"] = data
    out[f"{tag}_lens"] = np.array(lens)
    out[f"{tag}_cfg"] = np.array([nc, L, np.nan if decay is None else decay])
    cases.append(tag)
    if tag in ("a", "b"):  # the coarsening step of check_mc_convergence.py:196-207, then one more batch at the new cadence
        old = e.max_lag
        t = ns_o["thin_acf_by_2"](e)
        dump(out, f"{tag}_thin", t)
        x = ns_o["expand_max_lag"](t, old)
        dump(out, f"{tag}_expand", x)
        more = ar1(rs, nc, 20, phi)
        out[f"{tag}_more"] = more
        dump(out, f"{tag}_after", x.update(more))
out["cases"] = np.array(cases)

# ------------------------------------------------------------------ FFT statistics (netket/stats/mc_stats.py:303-331, _autocorr.py:32-86)
import types  # noqa: E402

ns_a = extract("stats/_autocorr.py", ["next_pow_two", "autocorr_1d", "auto_window", "integrated_time"], {**base_ns(), "lax": lax})
FftStats = lambda *a: a  # noqa: E731
ns_f = extract("stats/mc_stats.py", ["_get_blocks", "_block_variance", "_batch_variance", "_split_R_hat", "_statistics"],
               {**base_ns(), "Stats": FftStats, "config": types.SimpleNamespace(netket_use_plain_rhat=False), "BLOCK_SIZE": 32,
                "integrated_time": ns_a["integrated_time"]})
for tag, shape, phi in [("fft_16x200", (16, 200), 0.7), ("fft_1x500", (1, 500), 0.5), ("fft_8x33", (8, 33), 0.2), ("fft_3x7", (3, 7), 0.9)]:
    data = ar1(rs, shape[0], shape[1], phi)
    res = ns_f["_statistics"](jnp.asarray(data))
    out[f"{tag}_data"] = data
    out[f"{tag}_result"] = np.array([float(np.asarray(v)) for v in res])     # mean, error, variance, tau_avg, R_hat, tau_max
    out[f"{tag}_acf0"] = np.asarray(ns_a["autocorr_1d"](jnp.asarray(data[0])))
out["fft_cases"] = np.array(["fft_16x200", "fft_1x500", "fft_8x33", "fft_3x7"])

path = os.path.join(HERE, "online_stats_vectors.npz")
np.savez_compressed(path, **out)
print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1024:.1f} KiB")
