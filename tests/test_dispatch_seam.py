"""Seam S4 (netket/vqs/mc/common.py:31-99): the local-estimator multimethods dispatch on (state, operator, chunk_size) types as
plum does in the reference - most specific wins, `precedence` orders catch-alls, None vs int chunk sizes are distinct."""

import pytest

from netket_b200._dispatch import Dispatcher


class A:
    pass


class B(A):
    pass


def test_most_specific_method_wins_and_precedence_orders_catch_alls():
    f = Dispatcher("f")

    @f.dispatch
    def _(x: A, c: None):
        return "A-None"

    @f.dispatch
    def _(x: A, c: int):
        return "A-int"

    @f.dispatch
    def _(x: B, c: None):
        return "B-None"

    @f.dispatch(precedence=-100)
    def _(x, c):
        return "any"

    assert f(A(), None) == "A-None" and f(B(), None) == "B-None" and f(B(), 3) == "A-int" and f(3, "x") == "any"
    with pytest.raises(NotImplementedError):
        f(1, 2, 3)


def test_ambiguity_raises_and_reregistration_overrides():
    f = Dispatcher("f")

    @f.dispatch
    def _(x: A, y):
        return 1

    @f.dispatch
    def _(x, y: A):
        return 2

    with pytest.raises(TypeError, match="ambiguous"):
        f(A(), A())

    @f.dispatch
    def _(x: A, y: A):
        return 3

    assert f(A(), A()) == 3

    @f.dispatch
    def _(x: A, y: A):  # same signature registered again: replaces
        return 4

    assert f(A(), A()) == 4


def test_module_level_multimethods_exist_and_fall_through_with_the_reference_s_message():
    import netket_b200 as nk

    for name in ("get_local_kernel_arguments", "get_local_kernel", "local_estimators", "expect"):
        assert isinstance(getattr(nk.vqs, name), Dispatcher)

    class MyOp:
        pass

    with pytest.raises(NotImplementedError, match="local_estimators is not implemented for the combination"):
        nk.vqs.local_estimators(object(), MyOp(), None)
    with pytest.raises(NotImplementedError):
        nk.vqs.get_local_kernel_arguments(object(), MyOp())
