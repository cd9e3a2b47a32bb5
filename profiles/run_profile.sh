#!/bin/bash
# Run on the GPU box (gpurun).  usage: run_profile.sh <tag> [gram] [conv] [mpn]
#   always : ncu launch list (gpu__time_duration.sum per launch) of a short bench.py run -> gpurun_out/launches_<tag>.csv
#   gram   : --set full captures of the bilinear-pool forward (B=32 and B=256) and backward (B=32), raw metrics exported to CSV
#   conv   : --set full capture of one launch of each conv kernel shape class, raw metrics exported to CSV
#   mpn    : --set full capture of the covariance / Newton-Schulz GEMMs of one MPN-COV head forward+backward
# Numbers printed under ncu are never bench values.  .ncu-rep files are kept only while gpurun_out stays small.
R=${1:-r2}
shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-micro --no-eager > gpurun_out/bench_under_ncu_$R.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_$R.csv > gpurun_out/launches_$R.summary.txt 2>&1
for what in "$@"; do
  if [ "$what" = "gram" ]; then
    for B in 32 256; do
      ncu --set full --clock-control none --import-source on -k regex:"bcnn_gram_fwd|bcnn_cluster_fwd|bcnn_super_fwd" -c 2 -o gpurun_out/prof_gram_${R}_B$B -f \
          python tests/prof_bilinear.py $B > gpurun_out/prof_gram_${R}_B$B.log 2>&1
      ncu -i gpurun_out/prof_gram_${R}_B$B.ncu-rep --page raw --csv > gpurun_out/prof_gram_${R}_B$B.raw.csv 2>/dev/null
    done
    ncu --set full --clock-control none -k regex:"gram_pair|umma_gemm|colsum" -c 6 -o gpurun_out/prof_gramb_${R}_B32 -f \
        python tests/prof_bilinear.py 32 bwd > gpurun_out/prof_gramb_${R}_B32.log 2>&1
    ncu -i gpurun_out/prof_gramb_${R}_B32.ncu-rep --page raw --csv > gpurun_out/prof_gramb_${R}_B32.raw.csv 2>/dev/null
  fi
  if [ "$what" = "conv" ]; then
    ncu --set full --clock-control none -k regex:"conv3x3_igemm|conv3x3_wgrad" -s 105 -c 35 -o gpurun_out/prof_conv_$R -f \
        python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-micro --no-eager > gpurun_out/prof_conv_$R.log 2>&1
    ncu -i gpurun_out/prof_conv_$R.ncu-rep --page raw --csv > gpurun_out/prof_conv_$R.raw.csv 2>/dev/null
  fi
  if [ "$what" = "mpn" ]; then
    ncu --set full --clock-control none -k regex:"umma_gemm" -s 40 -c 24 -o gpurun_out/prof_mpncov_$R -f \
        python tests/prof_mpncov.py > gpurun_out/prof_mpncov_$R.log 2>&1
    ncu -i gpurun_out/prof_mpncov_$R.ncu-rep --page raw --csv > gpurun_out/prof_mpncov_$R.raw.csv 2>/dev/null
  fi
done
python profiles/extract_metrics.py gpurun_out $R > gpurun_out/metrics_$R.txt 2>&1
# gpurun copies back at most 64 MiB: drop the biggest reports first
while [ $(du -sm gpurun_out | cut -f1) -gt 48 ]; do
  f=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1)
  [ -z "$f" ] && break
  echo "dropping $f ($(du -m $f | cut -f1) MB)"; rm -f "$f"
done
du -sh gpurun_out; ls -la gpurun_out | tail -12
