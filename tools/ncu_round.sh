#!/bin/bash
# One gpurun call: ncu captures of the dominant kernel of every workload (after a plain run of the same command has
# exited 0), plus the launch list of the default workload.  Reports land in gpurun_out/; summaries are made here with
# tools/ncu_summary.py and committed under profiles/.
R=${1:-r02}
cd "$(dirname "$0")/.."
cap() {  # name, kernel regex, skip, count, bench args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  python bench.py --quick --steps 3 --warmup 3 "$@" > gpurun_out/${R}_plain_${name}.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:${rx} -s ${skip} -c ${cnt} -f -o gpurun_out/${R}_prof_${name} \
      python bench.py --quick --steps 3 --warmup 3 "$@" > gpurun_out/${R}_ncu_${name}.log 2>&1
  echo "${name}: rc=$?"
}
# (kernel base names: "^k_tile$" keeps the plan-builder kernels k_tile_dir, k_tile_fill, ... out)
cap c2 '^k_tile$' 3 1 --workload c2
cap c3 '^k_tile$' 3 1 --workload c3
cap c4 '^k_tile$' 3 1 --workload c4
if [ -z "$SKIP_C5" ]; then cap c5 k_hex8 6 2 --workload c5 --n 128; fi
python bench.py --quick --steps 3 --warmup 3 > gpurun_out/${R}_plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_c2.csv \
    python bench.py --quick --steps 3 --warmup 3 > gpurun_out/${R}_ncu_launches.log 2>&1
echo "launch list: rc=$?"
