"""CPU oracle for the VMC inner loop (TEST INFRASTRUCTURE — not the product).

This package is a NumPy restatement of the reference (NetKet, mounted read-only at
/root/reference) for the one hot path this repository accelerates: Metropolis sampling
of an RBM wave-function + local-energy estimation + MC statistics.

Rules (enforced by tests/test_no_oracle_in_product.py):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import anything under ``oracle/``;
  * the product package ``netket_b200`` never imports it and has no CPU fallback.

Pinning status
--------------
The reference cannot be imported here (jax/flax/plum are not installed, no network), and
its own tests hold **no golden vectors** for this path — they are
cross-implementation and statistical checks.  What pins this oracle:

  * ``tests/golden/`` fixtures produced by executing the reference's *own source* for
    the jax-free pieces (numba table packing ``_append_matrix``/``pack_internals``, the
    numba Ising / LocalOperator connected-element kernels) and, under a NumPy stand-in
    for ``jax.numpy`` (``tests/golden/jnp_shim.py``), the jax kernels
    ``_ising_kernel_jax``, ``_local_operator_kernel_jax``, ``log_cosh``,
    ``local_value_kernel_jax`` and the block ``statistics``.  Generating script:
    ``tests/golden/make_golden.py`` (runs only in the build container, where
    /root/reference exists).
  * the reference's known answers: ``lanczos_ed`` doctest eigenvalues
    (netket/exact.py:61-63), the invariants of test/operator/test_operator.py
    (hermiticity, dense equality, n_conn, trailing-zero padding) and the chi-square /
    5-sigma sampler and expect checks of test/sampler, test/variational.

What stays **parity unpinned**: the RNG bit-stream (the reference uses JAX threefry, which
cannot be reproduced without JAX).  The oracle and the CUDA kernels share an explicit
Philox4x32-10 proposal stream instead (see ``oracle/rng.py``), and the sampler is
validated statistically exactly as the reference's own tests do.
"""

from . import rng, hilbert, graph, rbm, operators, sampler, estimators, stats, ed  # noqa: F401
