# End-of-round multi-GPU check (gpurun --gpus N -- bash scripts/multi_gpu_final.sh N): the default bench line at N ranks, the CPU
# reference arm under the same launcher (rank 0 works, the others exit 0), and the frame-sharded config-4 k-means at N ranks.
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --no-cpu-baseline > gpurun_out/mgf_config2_n$N.json 2> gpurun_out/mgf_config2_n$N.err; echo "config2 rc=$?"; cut -c1-330 gpurun_out/mgf_config2_n$N.json
timeout 600 $TR --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/mgf_ref_n$N.json 2> gpurun_out/mgf_ref_n$N.err; echo "reference rc=$?"; cut -c1-200 gpurun_out/mgf_ref_n$N.json
timeout 600 $TR --master-port 29513 scripts/sharded_config4.py > gpurun_out/mgf_sharded_n$N.jsonl 2> gpurun_out/mgf_sharded_n$N.err; echo "sharded rc=$?"; cat gpurun_out/mgf_sharded_n$N.jsonl | cut -c1-400
