# One gpurun call: GPU parity tests, smoke, bench (own + reference arm), ncu launch list, one full capture.
set -x
mkdir -p gpurun_out
nvidia-smi -L
nproc
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py 2>gpurun_out/bench.err | tee gpurun_out/bench.json
python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
# a tick = the path launch (warp-per-path writer) + the sample launch: capture one of each
ncu --set full --clock-control none --import-source on -k regex:sfw_score -s 2 -c 2 -f -o gpurun_out/prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu_full.log 2>&1
ls -la gpurun_out
