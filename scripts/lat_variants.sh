for v in t256 t512 t128; do
  echo "== $v"
  SFW_B200_LIB=build/variants/libsfw_$v.so python scripts/latency_probe.py 2>&1 | tail -4 | cut -c1-100
  SFW_B200_LIB=build/variants/libsfw_$v.so SFW_PROBE_WL=C0 SFW_PROBE_CHILD=1 python scripts/variant_probe.py 2>&1 | tail -1
  SFW_B200_LIB=build/variants/libsfw_$v.so python -m pytest tests -m gpu -x -q -k "crowd or tiebreak" 2>&1 | tail -1
done
