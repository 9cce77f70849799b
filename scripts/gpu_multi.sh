#!/bin/bash
# N-GPU session: GPU test of the communicator is single-GPU; here bench.py at N ranks (library communicator) and the H2D ceiling.
N=${1:-2}
TAG=${2:-multi$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
PORT=29511
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT "$@"; PORT=$((PORT+1)); }
( run scripts/h2d_ceiling.py ) > $OUT/h2d_numa.json 2> $OUT/h2d_numa.err
cat $OUT/h2d_numa.json; tail -2 $OUT/h2d_numa.err
( DFDB_NO_NUMA=1 run scripts/h2d_ceiling.py ) > $OUT/h2d_nonuma.json 2> $OUT/h2d_nonuma.err
cat $OUT/h2d_nonuma.json
( run bench.py --gpus $N --steps ${3:-10} --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
tail -4 $OUT/bench.err; cat $OUT/bench.json
nvidia-smi topo -m > $OUT/topo.txt 2>&1; lscpu | grep -i "numa\|model name\|socket" > $OUT/cpu.txt
