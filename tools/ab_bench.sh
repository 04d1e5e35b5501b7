# A/B of library variants (tools/build_variant.py) through the contract benchmark, short run: tools/ab_bench.sh product v_x ...
for v in "$@"; do echo "== $v"
if [ "$v" = product ]; then unset MOCAP_B200_LIB; else export MOCAP_B200_LIB=mocapv2_b200/_variants/$v.so; fi
python bench.py --steps 100 --no-e2e --no-geometry --no-extra --cpu-sample 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['phase_ms']
print(round(d['value']), round(d['ms_per_step'],4), 'detect_in_flight', round(p['detect_in_flight_ms'],4), 'geom_latency', round(p['geometry_ms'],4), d['output_checksum']['sha1_16'])"
done
