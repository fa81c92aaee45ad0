#!/bin/bash
# Everything profiles/README.md quotes, collected on one B200 box (run through gpurun from the repo root):
#   bash profiles/tools/collect_evidence.sh            -> gpurun_out/r2_*.{json,log,csv,ncu-rep}
# Numbers printed under ncu are never bench values: the bench lines below run without a profiler.
set -u
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2_tests.log 2>&1; tail -2 $O/r2_tests.log
python bench.py > $O/r2_bench_1gpu.json 2> $O/r2_bench_1gpu.err
python bench.py --impl reference > $O/r2_bench_reference_arm.json 2> /dev/null
python bench.py --scale-long 4000 --pages 64 --no-e2e --no-cpu-baseline --no-skew > $O/r2_bench_sl4000_cli.json 2> /dev/null
python bench.py --preset gui --no-e2e --no-cpu-baseline --no-skew > $O/r2_bench_sl1600_gui.json 2> /dev/null
python bench_ops.py > $O/r2_ops.jsonl 2> $O/r2_ops.err
B="python bench.py --pages 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r2_launches.csv $B > $O/r2_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"warp_|tc_blur|lut|morph|otsu|adaptive|mask_blend|affine|close3" -c 18 -o $O/prof_r2 -f $B --no-skew > $O/r2_ncu2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"canny|hough|skew_finish" -c 5 -o $O/prof_r2_skew -f $B > $O/r2_ncu3.log 2>&1
ls -la $O/*.ncu-rep
