# One gpurun call: ncu captures of the headline kernels (full set, source, NO cache flush between replays so that
# L2-resident inputs stay L2-resident as in the real tick) + the launch list of a short bench run.
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --cache-control none --import-source on -f"
$NCU -k regex:sfw_score -s 4 -c 2 -o gpurun_out/r2_c1 python scripts/prof_workload.py C1 > gpurun_out/r2_c1.log 2>&1
$NCU -k regex:sfw_score_crowd -s 1 -c 1 -o gpurun_out/r2_c2 python scripts/prof_workload.py C2 40 48 > gpurun_out/r2_c2.log 2>&1
$NCU -k regex:sfw_score -s 6 -c 3 -o gpurun_out/r2_c4 python scripts/prof_workload.py C4 > gpurun_out/r2_c4.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_b_ncu.log 2>&1
ls -la gpurun_out | tail -12
