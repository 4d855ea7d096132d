#!/bin/bash
# development aid: build variants/<name>.so from the sources under <srcdir> (default pyfem_gpu_testflight_b200/csrc)
# usage: tools/build_variant.sh name [srcdir] [extra nvcc flags...]
set -e
cd "$(dirname "$0")/.."
name=$1; src=${2:-pyfem_gpu_testflight_b200/csrc}; shift; shift || true
mkdir -p variants /tmp/var_$name
FLAGS="-std=c++20 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -Wno-deprecated-declarations -I include"
pids=()
for f in pfg_setup pfg_assemble pfg_solve pfg_probe; do
  nvcc $FLAGS "$@" -c $src/$f.cu -o /tmp/var_$name/$f.o & pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o variants/$name.so /tmp/var_$name/*.o -gencode arch=compute_100a,code=sm_100a
echo built variants/$name.so
