# compute-sanitizer on the code paths added in round 2 (one gpurun call)
set -x
mkdir -p gpurun_out
OUT=gpurun_out/r2_sanitizer.txt
: > $OUT
run() { # title, tool, pytest selection
  echo "--- $1 ($2)" >> $OUT
  timeout 600 compute-sanitizer --tool $2 --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$3" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" | head -20 >> $OUT
  echo "$2 rc=${PIPESTATUS[0]}" >> $OUT
}
run "winner exchange: rings of 2/3/8 contexts, ragged counts, device-side slab merge, empty slabs, lost peer" memcheck "test_gpu_exchange"
run "prefix sharing on row slabs, pipelined batch (copy stream + launch groups), parallel packing" memcheck "row_slabs_is_bit or pipelined_batch or parallel_packing"
run "tie-break scenes through both kernel families" memcheck "test_gpu_tiebreak and (no_zero_w or all_costs_equal or all_invalid)"
run "device-side slab merge + fused exchange epilogue" racecheck "row_slab_merge_on_the_device and no_zero_w"
run "prefix sharing on row slabs (PDL launches)" racecheck "row_slabs_is_bit and 64-64"
cat $OUT
