# Multi-GPU measurements of one round (run on an N-GPU box: gpurun --gpus N -- bash scripts/multi_gpu_suite.sh N)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/mg_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29500
run() { # name nproc args...
  name=$1; np=$2; shift 2; port=$((port+1))
  if [ "$np" = "1" ]; then timeout 600 python "$@" > gpurun_out/mg_$name.json 2> gpurun_out/mg_$name.err
  else timeout 600 $TR --nproc-per-node $np --master-port $port "$@" > gpurun_out/mg_$name.json 2> gpurun_out/mg_$name.err; fi
  echo "$name rc=$?"; tail -c 300 gpurun_out/mg_$name.json | head -c 300; echo
}
# copy-only probe: H2D / D2H per rank at 1, 2, 4, N ranks, default vs NUMA-local pinning
for np in 1 4 $N; do
  for mode in default numa; do
    port=$((port+1))
    if [ "$np" = "1" ]; then PROBE_MODE=$mode timeout 120 python scripts/h2d_probe.py >> gpurun_out/mg_h2d_probe.jsonl 2>> gpurun_out/mg_h2d_probe.err
    else PROBE_MODE=$mode timeout 120 $TR --nproc-per-node $np --master-port $port scripts/h2d_probe.py >> gpurun_out/mg_h2d_probe.jsonl 2>> gpurun_out/mg_h2d_probe.err; fi
  done
done
cat gpurun_out/mg_h2d_probe.jsonl
run config2_n$N $N bench.py --gpus $N --no-cpu-baseline
run config3_n$N $N bench.py --gpus $N --workload config3 --steps 5 --no-cpu-baseline
run config5_n$N $N bench.py --gpus $N --workload config5 --steps 10 --no-cpu-baseline
run config4_n$N $N bench.py --gpus $N --workload config4 --steps 5 --no-cpu-baseline
for np in 2 $N; do
  port=$((port+1))
  timeout 600 $TR --nproc-per-node $np --master-port $port scripts/sharded_config4.py >> gpurun_out/mg_sharded_config4.jsonl 2>> gpurun_out/mg_sharded_config4.err
done
timeout 300 python scripts/sharded_config4.py >> gpurun_out/mg_sharded_config4.jsonl 2>> gpurun_out/mg_sharded_config4.err
cat gpurun_out/mg_sharded_config4.jsonl
