mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/r2g_$name.json 2> gpurun_out/r2g_$name.err; python - <<PY
import json
j=json.loads(open('gpurun_out/r2g_$name.json').read().strip().splitlines()[-1])
print('$name', 'value %.0f'%j['value'], 'ms/stream %.4f'%j['pipeline']['ms_per_stream'], 'single %.4f'%j['single_stream']['ms'], 'solve %.4f'%j['stages_ms']['kmeans_solve'])
PY
}
run base A=1
run nodraws CSB_PAR_DRAWS=0
run nodraws_fpw2 CSB_PAR_DRAWS=0 CSB_SOLVER_FPW=2
run nodraws_noearly CSB_PAR_DRAWS=0 CSB_NOEARLY=1
