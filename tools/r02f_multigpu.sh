#!/bin/bash
# Multi-GPU train-step scaling: one process per GPU over NCCL, DDP gradient all-reduce (G 45 MB + D 150 MB fp32).
#   /usr/local/graft/bin/gpurun --gpus N --timeout 900 -- 'bash tools/r02f_multigpu.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py > $O/r02f_build_n$N.log 2>&1
run() {   # config, extra flags
  local c=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + c)) \
    tools/bench_train_step.py --config $c --cips-backend fused "$@" --profile $O/r02f_prof_c${c}_n$N.txt > $O/r02f_train_c${c}_n$N.json 2> $O/r02f_train_c${c}_n$N.err
  echo "config $c on $N GPUs: exit $?"; cut -c1-260 $O/r02f_train_c${c}_n$N.json; grep -o '"comm": {[^}]*}' $O/r02f_train_c${c}_n$N.json | cut -c1-200
}
run 5
run 4
run 3 --film-backend fused --integrate-backend fused
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N --steps 10 --warmup 3 > $O/r02f_bench_n$N.json 2> $O/r02f_bench_n$N.err; echo "bench N=$N: $?"; cut -c1-400 $O/r02f_bench_n$N.json
grep -m3 -i "NVLS\|via P2P\|NET/" $O/r02f_bench_n$N.err | cut -c1-200
