#!/bin/bash
# development aid: run one command under several library variants back to back on one box
# usage: tools/ab_libs.sh "<command>" variant1.so variant2.so ...
CMD="$1"; shift
LIB=pyfem_gpu_testflight_b200/libpyfem_b200.so
cp $LIB /tmp/lib_orig.so
for so in "$@"; do
  cp "variants/$so" $LIB
  echo "=== $so: $CMD"
  bash -c "$CMD" 2>&1 | grep -v "^mesh:"
done
cp /tmp/lib_orig.so $LIB
