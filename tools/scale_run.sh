#!/bin/bash
# Runs ON THE GPU BOX: tools/scale_run.sh N  ->  gpurun_out/final/bench_{c2,c5}_nN.json (the lines under profiles/r02_bench_c*_n*.json)
N=$1; O=gpurun_out/final; mkdir -p $O
if [ "$N" == "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"; fi
[ "$N" != "1" ] && { $L bench.py --gpus $N > $O/bench_c2_n$N.json 2> $O/bench_c2_n$N.err; python tools/show_bench.py $O/bench_c2_n$N.json; tail -2 $O/bench_c2_n$N.err; }
$L bench.py --gpus $N --config c5 --steps 1 --warmup 3 > $O/bench_c5_n$N.json 2> $O/bench_c5_n$N.err; python tools/show_bench.py $O/bench_c5_n$N.json; tail -2 $O/bench_c5_n$N.err
[ "$N" != "1" ] && { $L tools/dist_check.py 2>&1 | tail -4; }
