#!/bin/bash
# A/B of builds of the grid family's blocked kernel (bwgr_b200/lib/variants/lib_<markers per block>_<accumulator copies>.so), same probe each
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/grid_variants.log
: > $OUT
for lib in bwgr_b200/lib/variants/lib_*.so; do
  for shape in "100000 4000" "200000 2000"; do
    echo "== $lib $shape" >> $OUT
    GRID_ONLY=1 BWGR_LIB=$PWD/$lib timeout 120 python tools/grid_probe.py $shape emRR 3 >> $OUT 2>&1
  done
done
cat $OUT
