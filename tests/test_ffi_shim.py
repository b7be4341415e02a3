"""netket_b200/csrc/ffi/nkb200_jax_ffi.cc against a stub of the XLA typed-FFI API (tests/ffi_stub): jaxlib's headers are not in
this image, so the shim cannot be built for real here; this checks what can be checked without them - the file is valid C++17,
every C ABI call in it matches include/nkb200.h, and every handler's parameter list matches its Bind().Ctx/Arg/Ret/Attr chain."""

import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "netket_b200", "csrc", "ffi", "nkb200_jax_ffi.cc")
CUDA_INC = "/usr/local/cuda/include"


def _compile(src, extra=()):
    return subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "ffi_stub"), "-I",
                           os.path.join(ROOT, "include"), "-I", CUDA_INC, *extra, src], capture_output=True, text=True)


@pytest.mark.skipif(shutil.which("g++") is None or not os.path.isdir(CUDA_INC), reason="g++ or the CUDA headers are not available")
def test_ffi_shim_type_checks_against_the_stub_api(tmp_path):
    r = _compile(SHIM)
    assert r.returncode == 0, r.stderr
    # the check has teeth: a handler whose parameters disagree with its binding is rejected
    bad = tmp_path / "bad.cc"
    src = open(SHIM).read()
    assert ".Attr<double>(\"mean\")" in src
    bad.write_text(src.replace(".Attr<double>(\"mean\")", ".Attr<int32_t>(\"mean\").Attr<int32_t>(\"extra\")"))
    r2 = _compile(str(bad))
    assert r2.returncode != 0 and "do not match the binding" in r2.stderr
