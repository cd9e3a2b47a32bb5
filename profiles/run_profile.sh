#!/bin/bash
# Run on the GPU box (gpurun).  Launch list of the whole short bench run (post-processed to the last step by
# profiles/summarize_launches.py) + optional full captures.  Numbers printed under ncu are never bench values.
R=${1:-r1}
FULL=${2:-0}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_$R.log 2>&1
if [ "$FULL" = "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:"gram_pair|colsum_partial" -c 6 -o gpurun_out/prof_gram_$R -f \
    python tests/prof_bilinear.py > gpurun_out/prof_gram_$R.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv3x3_igemm|conv3x3_wgrad" -s 105 -c 35 -o gpurun_out/prof_conv_$R -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof_conv_$R.log 2>&1
fi
ls -la gpurun_out | tail -8
