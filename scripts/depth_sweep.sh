run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --e2e-steps 1 $EXTRA > gpurun_out/r2q_$name.json 2> gpurun_out/r2q_$name.err; python - <<PY
import json
j=json.loads(open('gpurun_out/r2q_$name.json').read().strip().splitlines()[-1])
print('$name', 'value %.0f'%j['value'], 'ms/stream %.4f'%j['pipeline']['ms_per_stream'], 'single %.4f'%j['single_stream']['ms'])
PY
}
EXTRA="--depth 6" run d6 A=1
EXTRA="--depth 4" run d4 A=1
EXTRA="--depth 8" run d8 A=1
EXTRA="--depth 12" run d12 A=1
EXTRA="--depth 6" run d6_pix48 CSB_EARLY_PIX_CTAS=48
EXTRA="--depth 6" run d6_pix96 CSB_EARLY_PIX_CTAS=96
EXTRA="--depth 8" run d8_pix96 CSB_EARLY_PIX_CTAS=96
EXTRA="--depth 6 --early-pixel 0" run d6_noearly A=1
EXTRA="--depth 6" run d6_again A=1
