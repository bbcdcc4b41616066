set -x
N=$1
export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=false
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps $2 --warmup 3 --workload $3 --no-cpu-baseline > gpurun_out/r22_bench_$3_n${N}.json 2> gpurun_out/r22_bench_$3_n${N}.err; echo "rc=$?" >> gpurun_out/r22_bench_$3_n${N}.err; }
run 29511 10 c3
run 29513 5 c5
run 29515 20 c1
tail -n 2 gpurun_out/r22_bench_c3_n${N}.err gpurun_out/r22_bench_c5_n${N}.err
