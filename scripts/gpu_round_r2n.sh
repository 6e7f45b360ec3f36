# One gpurun call at the end of a change: GPU suite, smoke, bench (own + reference arm), launch list.
set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_r2n.log
python __graft_entry__.py --smoke 2>&1 | tail -2 | cut -c1-400 | tee gpurun_out/smoke_r2n.log
python bench.py 2>gpurun_out/bench_r2n.err | tail -1 > gpurun_out/bench_r2n.json; cut -c1-300 gpurun_out/bench_r2n.json
python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_ref_r2n.err | tail -1 > gpurun_out/bench_ref_r2n.json; cut -c1-300 gpurun_out/bench_ref_r2n.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2n_b_ncu.log 2>&1
tail -3 gpurun_out/bench_r2n.err
