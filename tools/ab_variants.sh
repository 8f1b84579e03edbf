for v in "" park6_128x7 park6_256x4 park6_256x3 park0_128x6; do
  L=${v:+$PWD/plonky2_aes_b200/variants/libp2gpu_$v.so}
  echo "== ${v:-default}"
  P2G_LIB_PATH=$L python tools/poseidon_peak.py
  P2G_LIB_PATH=$L python bench.py --no-cpu-baseline --config5-proofs 0 --steps 6 2>/dev/null | python -c "import json,sys; p=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('value', round(p['value'],1), 'merkle_ms', round(p['roofline']['merkle_ms'],3), 'perm/s', round(p['roofline']['int_pipe']['perms_per_s']/1e9,3))"
done
