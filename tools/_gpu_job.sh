set -x
N=$1
export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=false FGL_BENCH_DEBUG=1
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps $2 --warmup 3 --workload $3 --no-cpu-baseline --handoff $4 > gpurun_out/r19_bench_$3_n${N}_$4.json 2> gpurun_out/r19_bench_$3_n${N}_$4.err; echo "rc=$?" >> gpurun_out/r19_bench_$3_n${N}_$4.err; }
run 29513 5 c5 peer
run 29514 5 c5 host
grep "e2e step" gpurun_out/r19_bench_c5_n${N}_peer.err | head -30
grep "e2e step" gpurun_out/r19_bench_c5_n${N}_host.err | head -30
