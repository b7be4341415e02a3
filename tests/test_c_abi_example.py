"""The boundary is plain C: include/nkb200.h compiles as pedantic C99, and a C program with host buffers only
(examples/c_abi_step.c) links against libnkb200.so without Python, torch or CUDA headers and - on a GPU - runs VMC steps."""

import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "netket_b200", "lib")


def _build(tmp_path):
    exe = str(tmp_path / "c_abi_step")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_step.c"), "-L", LIBDIR, "-lnkb200", f"-Wl,-rpath,{LIBDIR}", "-lm", "-o", exe])
    return exe


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_header_is_c99_and_the_c_example_links(lib_built, tmp_path):
    src = tmp_path / "hdr.c"
    src.write_text('#include "nkb200.h"\nint main(void) { nk_sweep_t s; nk_ctx_desc_t d; (void)s; (void)d; return 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])
    exe = _build(tmp_path)
    needed = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libnkb200.so" in needed and "torch" not in needed and "python" not in needed.lower()


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_c_example_runs_vmc_steps(cuda, tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "6", "2", "2048", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("step ")]
    assert len(lines) == 3
    # TFIM 6x6, h = 3, near-zero weights: E_loc is about -h N + J <sum zz> ~ -108 with an acceptance close to one
    for ln in lines:
        e = float(ln.split("E = ")[1].split()[0])
        acc = float(ln.split("acceptance = ")[1].split()[0])
        assert -130.0 < e < -90.0 and 0.5 < acc <= 1.0, ln
