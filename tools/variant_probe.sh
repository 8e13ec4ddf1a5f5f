#!/bin/bash
# ms/frame of bench workloads for the default library and the tuning variants named on the command line
WL="${PROBE_WL:-bunny4k mirror1080 spheres1080}"
echo "== default"; python tools/tiny_probe_ms.py $WL
for v in "$@"; do
  echo "== $v"; CUTRACE_B200_LIB=$PWD/cutrace_b200/lib/variants/libcutrace_b200_$v.so python tools/tiny_probe_ms.py $WL
done
