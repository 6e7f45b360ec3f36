set -x
nvidia-smi -L
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s 2>&1 | tail -40
