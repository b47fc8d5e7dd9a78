#!/bin/bash
# Dev: C3 mesh probe (tools/run_mesh.py) for each variant library:  tools/ab_mesh.sh [run_mesh args --] name...
args=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do args+=("$1"); shift; done
if [ "$1" == "--" ]; then shift; else set -- "${args[@]}"; args=(); fi
for v in "$@"; do
  lib=variants/liblt_$v.so; [ "$v" == "base" ] && lib=light_transport_b200/liblt_b200.so
  echo "== $v"; python tools/run_mesh.py --lib $PWD/$lib "${args[@]}" 2>&1 | tail -n +1
done
