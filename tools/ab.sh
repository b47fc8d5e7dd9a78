#!/bin/bash
# Dev: run a bench line (no CPU arm, no RMSE) for each variant library:  tools/ab.sh [bench args --] name...
args=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do args+=("$1"); shift; done
if [ "$1" == "--" ]; then shift; else set -- "${args[@]}"; args=(); fi
for v in "$@"; do
  lib=variants/liblt_$v.so; [ "$v" == "base" ] && lib=light_transport_b200/liblt_b200.so
  python bench.py --lib $PWD/$lib --no-cpu-baseline --no-rmse --no-extra --steps 3 --warmup 3 "${args[@]}" > gpurun_out/ab_$v.json 2>gpurun_out/ab_$v.err
  echo -n "$v: "; python tools/show_bench.py gpurun_out/ab_$v.json
done
