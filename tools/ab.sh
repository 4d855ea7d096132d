#!/bin/bash
# development aid: time library variants (variants/*.so) back to back on one box
# usage: tools/ab.sh "<probe args>" variant1.so[:ENV=VAL,...] variant2.so ...
ARGS="$1"; shift
LIB=pyfem_gpu_testflight_b200/libpyfem_b200.so
cp $LIB /tmp/lib_orig.so
for spec in "$@"; do
  so="${spec%%:*}"; envs=""
  if [[ "$spec" == *:* ]]; then envs="${spec#*:}"; envs="${envs//,/ }"; fi
  cp "variants/$so" $LIB
  echo "=== $spec"
  env $envs python tools/probe.py $ARGS 2>&1 | grep -v "^mesh:"
done
cp /tmp/lib_orig.so $LIB
