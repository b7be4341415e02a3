"""Shared helpers of the golden-vector generators: pull function definitions out of the reference's files with `ast`,
unchanged, and execute them on the NumPy stand-in for jax (tests/golden/jnp_shim.py)."""

import ast
import os
import sys
import warnings
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import jnp_shim  # noqa: E402

REF = "/root/reference/netket"
jax, jnp = jnp_shim.make_jax()


def extract(path, names, ns, class_name=None):
    """exec the named top-level functions (or methods of `class_name`) of a reference file, source unchanged."""
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    body = tree.body
    if class_name is not None:
        body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name][0].body
    found = []
    for node in body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            if class_name is not None:
                node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, os.path.join(REF, path), "exec"), ns)
            found.append(node.name)
    missing = set(names) - set(found)
    assert not missing, f"{path}: {missing} not found"
    return ns


def base_ns():
    return {"np": np, "jnp": jnp, "jax": jax, "partial": partial, "warnings": warnings, "__builtins__": __builtins__}


