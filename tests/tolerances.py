"""Parity bars of the -m gpu tests (BASELINE.json north_star): logpsi and E_loc on identical sample batches agree with the
oracle within 1e-12 relative in fp64 and 1e-5 in fp32.

`assert_rel` is element-wise: |x - ref| <= tol * |ref|, except that elements smaller than 1 % of the batch's largest
magnitude are held to the absolute floor tol * 0.01 * max|ref| (a local energy can pass through zero by cancellation
between its diagonal and off-diagonal parts; the reference's own rounding error is relative to the terms, not to the sum).
Every call also records the observed error (max and median relative error, in units of `tol`) in
gpurun_out/parity_errors.jsonl, from which DESIGN.md's table of measured errors is taken.
"""

import json
import os

import numpy as np

F64_TOL = 1e-12
F32_TOL = 1e-5
# fp32 product-form kernels on weights far beyond the benchmark's (std >= 0.2, i.e. >= 20 x the default init, max|W| up to 3):
# the running (A, B) pairs take one rounding per accepted move and the sensitivity of a flip ratio to them grows with |W|;
# measured 1.2e-5 .. 2.6e-5, asserted at:
F32_TOL_LARGE_W = 3e-5

_LOG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_errors.jsonl")


def assert_rel(x, ref, tol, what=""):
    x, ref = np.asarray(x, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    scale = np.abs(ref).max() if ref.size else 1.0
    floor = 0.01 * scale
    err = np.abs(x - ref) / np.maximum(np.abs(ref), floor) if ref.size else np.zeros(0)
    try:
        os.makedirs(os.path.dirname(_LOG), exist_ok=True)
        with open(_LOG, "a") as fh:
            fh.write(json.dumps({"what": what or os.environ.get("PYTEST_CURRENT_TEST", ""), "tol": tol, "n": int(ref.size),
                                 "max_rel": float(err.max()) if err.size else 0.0,
                                 "median_rel": float(np.median(err)) if err.size else 0.0}) + "\n")
    except OSError:
        pass
    np.testing.assert_allclose(x, ref, rtol=tol, atol=tol * floor, err_msg=what)
