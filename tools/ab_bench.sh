#!/bin/bash
# usage (GPU box): tools/ab_bench.sh <alt .so> [rounds]  -- alternate the built library (A) and an alternative build (B) on the same box
alt=$1; rounds=${2:-3}
mkdir -p gpurun_out
cp conformer_ocr_b200/libcocr_b200.so /tmp/lib_A.so
for r in $(seq $rounds); do
  for v in A B; do
    if [ $v = A ]; then cp /tmp/lib_A.so conformer_ocr_b200/libcocr_b200.so; else cp $alt conformer_ocr_b200/libcocr_b200.so; fi
    timeout 200 python bench.py --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/ab_${v}_$r.json
    echo -n "$v round $r: "; python tools/summarize_bench.py gpurun_out/ab_${v}_$r.json 2>/dev/null | grep "lines/s\|attention_tc" | tr '\n' ' '; echo
  done
done
cp /tmp/lib_A.so conformer_ocr_b200/libcocr_b200.so
