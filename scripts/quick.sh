# quick GPU check: parity tests + device-time probe
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/perf_probe.py C1 2>&1 | grep -E "kernel ms|valid"
python scripts/perf_probe.py C4 2>&1 | grep -E "kernel ms"
python scripts/perf_probe.py C3 512 2>&1 | grep -E "kernel ms"
python scripts/perf_probe.py C0 2>&1 | grep -E "kernel ms"
python scripts/perf_probe.py C2 2>&1 | grep -E "kernel ms"
