#!/bin/bash
# Dev: C3 mesh probe (tools/run_mesh.py) for each variant library:  tools/ab_mesh.sh name...
for v in "$@"; do
  lib=variants/liblt_$v.so; [ "$v" == "base" ] && lib=light_transport_b200/liblt_b200.so
  echo -n "$v: "; LT_B200_LIB=$PWD/$lib python tools/run_mesh.py megakernel 8 2>&1 | tail -1
done
