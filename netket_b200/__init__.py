"""netket_b200 — the VMC inner loop of NetKet (Metropolis sampling of an RBM + local energies + MC statistics)
as hand-written sm_100a CUDA behind NetKet's own Python surface.

    import netket_b200 as nk
    g  = nk.graph.Hypercube(length=10, n_dim=2, pbc=True)
    hi = nk.hilbert.Spin(s=1/2, N=g.n_nodes)
    ha = nk.operator.Ising(hi, g, h=3.0)
    vs = nk.vqs.MCState(nk.sampler.MetropolisLocal(hi, n_chains=2**16), nk.models.RBM(alpha=4, param_dtype="float32"))
    vs.expect(ha)

Scope: SURVEY.md §8 / DESIGN.md.  There is no CPU path: every numerical entry point calls libnkb200.so
(include/nkb200.h) and raises if it is missing.
"""

from . import convergence, driver, graph, hilbert, models, operator, optimizer, sampler, serialization, stats, vqs  # noqa: F401
from ._lib import NkError, LIB_PATH  # noqa: F401
from .config import config  # noqa: F401

# netket/stats/__init__.py:20 exports the estimator containers from `nk.stats` (they are defined next to the multimethods that
# return them)
stats.LocalEstimators = vqs.LocalEstimators
stats.LocalEstimatorsBatch = vqs.LocalEstimatorsBatch

__version__ = "0.1.0"
