# One gpurun call: ncu captures (full set, source, no cache flush) of the final build's headline kernels + sanitizer
# runs of the dense-crowd ring walk and the prefetching cross-pair loop.
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --cache-control none --import-source on -f"
$NCU -k regex:sfw_score -s 4 -c 2 -o gpurun_out/r2n_c1 python scripts/prof_workload.py C1 > gpurun_out/r2n_c1.log 2>&1
$NCU -k regex:sfw_score_crowd -s 1 -c 1 -o gpurun_out/r2n_c2 python scripts/prof_workload.py C2 40 48 > gpurun_out/r2n_c2.log 2>&1
$NCU -k regex:sfw_score -s 6 -c 3 -o gpurun_out/r2n_c4 python scripts/prof_workload.py C4 > gpurun_out/r2n_c4.log 2>&1
$NCU -k regex:sfw_score_crowd -s 2 -c 1 -o gpurun_out/r2n_c0 python scripts/prof_workload.py C0 > gpurun_out/r2n_c0.log 2>&1
ls -la gpurun_out | grep r2n_ | tail -8
