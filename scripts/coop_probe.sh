run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2v_$name.json 2> gpurun_out/r2v_$name.err; python - <<PY
import json
j=json.loads(open('gpurun_out/r2v_$name.json').read().strip().splitlines()[-1])
print('$name', 'value %.0f'%j['value'], 'ms/stream %.4f'%j['pipeline']['ms_per_stream'], 'single %.4f'%j['single_stream']['ms'], 'launches', j['pipeline']['launches_per_stream'])
PY
}
run default A=1
run nocoop CSB_NO_COOP_SCAN=1
run default2 A=1
run nocoop2 CSB_NO_COOP_SCAN=1
