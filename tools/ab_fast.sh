#!/bin/bash
# A/B timing of sweep_fast variants built by tools/build_variant.py (developer tool; run under gpurun)
for v in "" pf0 pf2w28 pf2w24 pf1w32 pf0w32; do
  if [ -z "$v" ]; then lib=netket_b200/lib/libnkb200.so; else lib=netket_b200/lib/variants/libnkb200_$v.so; fi
  [ -f "$lib" ] || continue
  echo "== variant ${v:-default}"
  NKB200_LIB=$lib python tools/fast_probe.py --reps 5 --check 16 2>&1 | head -1
  NKB200_LIB=$lib python tools/fast_probe.py --reps 5 --no-eloc --cl 21 --check 16 2>&1 | head -1
done
