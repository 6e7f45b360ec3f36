# One gpurun call: ncu captures (full set, source, no cache flush) of the r2m build's headline kernels + sanitizer
# runs of the dense-crowd ring walk and the prefetching cross-pair loop.
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --cache-control none --import-source on -f"
$NCU -k regex:sfw_score -s 4 -c 2 -o gpurun_out/r2m_c1 python scripts/prof_workload.py C1 > gpurun_out/r2m_c1.log 2>&1
$NCU -k regex:sfw_score_crowd -s 1 -c 1 -o gpurun_out/r2m_c2 python scripts/prof_workload.py C2 40 48 > gpurun_out/r2m_c2.log 2>&1
$NCU -k regex:sfw_score -s 6 -c 3 -o gpurun_out/r2m_c4 python scripts/prof_workload.py C4 > gpurun_out/r2m_c4.log 2>&1
OUT=gpurun_out/r2m_sanitizer.txt
: > $OUT
run() { # title, tool, pytest selection
  echo "--- $1 ($2)" >> $OUT
  timeout 900 compute-sanitizer --tool $2 --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$3" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -20 >> $OUT
  echo "$2 rc=${PIPESTATUS[0]}" >> $OUT
}
run "dense-crowd ring walk (crowd layouts, odd / even rings, tails of the last warp)" memcheck "crowd_kernel_parity or crowd_kernel_with_groups or layout_boundar or crowd_batch"
run "dense-crowd ring walk" racecheck "crowd_kernel_parity"
run "thread-per-trajectory cross-pair loop with the one-trip-ahead fetch (golden cases)" memcheck "golden"
cat $OUT
ls -la gpurun_out | tail -8
