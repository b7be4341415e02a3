"""The oracle is test infrastructure: the product package must never import it, and only tests/, smoke() and
bench.py's cpu_baseline / --impl reference legs may."""

import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _imports(path):
    tree = ast.parse(open(path).read())
    mods = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            mods += [(a.name, node.lineno) for a in node.names]
        elif isinstance(node, ast.ImportFrom) and node.module:
            mods.append((node.module, node.lineno))
    return mods


def test_product_never_imports_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "netket_b200")):
        for f in files:
            if f.endswith(".py"):
                for mod, line in _imports(os.path.join(d, f)):
                    assert not (mod == "oracle" or mod.startswith("oracle.")), f"{f}:{line} imports {mod}"
            if f.endswith((".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(d, f)).read().replace("oracle/rng.py", "").replace("oracle/hilbert.py", "")


def test_bench_uses_oracle_only_for_the_cpu_baseline():
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    allowed = {"run_cpu", "edges_np"}  # the cpu_baseline / --impl reference legs
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("oracle"):
                assert fn.name in allowed, f"bench.py:{node.lineno}: {fn.name} imports {node.module}"
    # edges_np is only reachable from run_cpu
    assert src.count("edges_np()") == 2  # its definition's docstring-free body is used once, in run_cpu


def test_graft_entry_uses_oracle_only_in_smoke():
    src = open(os.path.join(ROOT, "__graft_entry__.py")).read()
    tree = ast.parse(src)
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("oracle"):
                assert fn.name == "smoke"
