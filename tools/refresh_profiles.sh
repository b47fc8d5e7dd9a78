#!/bin/bash
# Runs ON THE GPU BOX (gpurun -- tools/refresh_profiles.sh): every capture that profiles/r02_* is built from, each ncu pass
# after the identical command has exited 0 without ncu.  Outputs land in gpurun_out/final/; tools/make_profiles.py turns
# them into the files under profiles/ (run in the build container).
set -u
O=gpurun_out/final; mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
B="python bench.py --no-extra --no-rmse --no-cpu-baseline --steps 2 --warmup 1"
B3="python bench.py --config c3 --no-extra --no-rmse --no-cpu-baseline --steps 1 --warmup 1"
$B > $O/plain_c2.json 2> $O/plain_c2.err && \
  ncu --metrics $M --clock-control none -c 60 --csv --log-file $O/r02_launches_bench_c2.csv $B > /dev/null 2>&1
$B3 > $O/plain_c3.json 2> $O/plain_c3.err && \
  ncu --metrics $M,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:walk_megakernel -c 2 --csv --log-file $O/r02_traffic_bench_c3.csv $B3 > /dev/null 2>&1
full() {  # name, command...
  local name=$1; shift
  "$@" > $O/$name.plain.log 2>&1 && \
    ncu --set full --clock-control none --import-source on -k regex:walk_megakernel -c 1 -s 1 -f -o $O/$name "$@" > $O/$name.ncu.log 2>&1
}
full c2walk32 python tools/run_once.py megakernel compat f32 512 64
full c2walk64 python tools/run_once.py megakernel compat f64 512 64
full c3lbvh2 python tools/run_mesh.py --spp 4 --reps 2 --trav lbvh2
ls -la $O
