"""Developer tool: build libnkb200 variants with extra nvcc defines for A/B timing on the GPU box.

    python tools/build_variant.py NAME -DNK_FAST_WARPS=32 ...   ->  netket_b200/lib/variants/libnkb200_NAME.so
    NKB200_LIB=netket_b200/lib/variants/libnkb200_NAME.so python tools/fast_probe.py
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from netket_b200 import build as B  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "netket_b200", "lib", "variants")
obj_dir = os.path.join("/tmp", "nkb200_variant_" + name)  # outside the tree: the snapshot sent to the GPU box stays small
os.makedirs(out_dir, exist_ok=True)
os.makedirs(obj_dir, exist_ok=True)
objs = []
procs = []
for src in B._sources():
    obj = os.path.join(obj_dir, src[:-3] + ".o")
    objs.append(obj)
    procs.append(subprocess.Popen([B._nvcc()] + B.NVCC_FLAGS + flags + ["-c", os.path.join(B.CSRC, src), "-o", obj]))
for p in procs:
    if p.wait() != 0:
        raise SystemExit("nvcc failed")
lib = os.path.join(out_dir, f"libnkb200_{name}.so")
subprocess.check_call([B._nvcc(), "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-lrt", "-lpthread", "-ldl"])
print(lib)
