for v in "" tools/_variants/lib_add3all.so; do env ${v:+LM_B200_LIB=$v} timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-exact --min-seconds 0 > gpurun_out/bench_v.json 2>/dev/null; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v.json').read().strip().splitlines()[-1])
print(round(d['value']), d['roofline']['kernel_ms'], {k:round(v['kernel_ms_best'],3) for k,v in d['inputs'].items()}, d['configs']['config2_many_motifs']['kernel_ms'], d['configs']['config3_protein']['kernel_ms'])
PY
done
