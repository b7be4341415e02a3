"""Multiple dispatch on argument types for the local-estimator seam S4 (netket/vqs/mc/common.py:31-99 uses plum).

    @local_estimators.dispatch
    def _(vstate: MCState, op: MyOperator, chunk_size: None): ...

Resolution: every registered signature whose annotations accept the call's arguments is a candidate; the most specific one
wins (argument-wise subclass order), `precedence` breaks ties / orders catch-alls as in plum, an ambiguous call raises.
Annotations: a class, a tuple / union of classes, `None` (the value None), `typing.Any` or nothing (anything).
"""

import inspect
import types
import typing


def _norm(ann):
    if ann is inspect.Parameter.empty or ann is typing.Any:
        return (object,)
    if ann is None:
        return (type(None),)
    origin = typing.get_origin(ann)
    if origin is typing.Union or (hasattr(types, "UnionType") and isinstance(ann, types.UnionType)):
        out = ()
        for a in typing.get_args(ann):
            out += _norm(a)
        return out
    if isinstance(ann, tuple):
        out = ()
        for a in ann:
            out += _norm(a)
        return out
    if isinstance(ann, type):
        return (ann,)
    raise TypeError(f"unsupported annotation {ann!r} in a dispatch signature")


class Dispatcher:
    def __init__(self, name, doc=None):
        self.__name__ = name
        self.__doc__ = doc
        self._methods = []  # (signature: tuple of tuples of types, precedence, fn)
        self._cache = {}    # argument types -> resolved method (as plum: resolution runs once per type tuple)

    def dispatch(self, fn=None, *, precedence=0):
        if fn is None:
            return lambda f: self.dispatch(f, precedence=precedence)
        params = [p for p in inspect.signature(fn).parameters.values()
                  if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
        sig = tuple(_norm(p.annotation) for p in params)
        self._methods = [m for m in self._methods if m[0] != sig]  # re-registration overrides
        self._methods.append((sig, precedence, fn))
        self._cache.clear()
        return self

    @staticmethod
    def _accepts(sig, args):
        return len(sig) == len(args) and all(isinstance(a, t) for a, t in zip(args, sig))

    @staticmethod
    def _more_specific(a, b):
        """signature a <= b argument-wise (every type a accepts, b accepts too)"""
        return all(all(any(issubclass(x, y) for y in tb) for x in ta) for ta, tb in zip(a, b))

    def resolve(self, *args):
        key = tuple(type(a) for a in args)
        fn = self._cache.get(key)
        if fn is None:
            fn = self._cache[key] = self._resolve(args)
        return fn

    def _resolve(self, args):
        cands = [m for m in self._methods if self._accepts(m[0], args)]
        if not cands:
            raise NotImplementedError(f"{self.__name__}: no method registered for ({', '.join(type(a).__name__ for a in args)})")
        best = [m for m in cands if not any(self._strictly(o, m) for o in cands)]
        if len(best) > 1:
            hi = max(p for _, p, _ in best)
            best = [m for m in best if m[1] == hi]
        if len(best) != 1:
            raise TypeError(f"{self.__name__}: ambiguous call for ({', '.join(type(a).__name__ for a in args)}); "
                            "register a more specific method or give one a higher precedence")
        return best[0][2]

    def _strictly(self, o, m):
        """o beats m: strictly more specific, or equally specific with a higher precedence"""
        o_le_m, m_le_o = self._more_specific(o[0], m[0]), self._more_specific(m[0], o[0])
        if o_le_m and not m_le_o:
            return True
        if o_le_m and m_le_o:
            return o[1] > m[1]
        return False

    def __call__(self, *args):
        return self.resolve(*args)(*args)
