"""Operators on the hot path: Ising, Heisenberg, GraphOperator, LocalOperator (1- and 2-site terms).

Everything else in netket.operator (operator algebra, Pauli strings, bosons, fermions, continuous,
Liouvillians) is out of scope.
"""

from ._base import DiscreteJaxOperator  # noqa: F401
from ._ising import Ising, IsingJax  # noqa: F401
from ._local_operator import LocalOperator, LocalOperatorJax, GraphOperator, Heisenberg  # noqa: F401
