#!/bin/bash
# ncu --set full of the v2 estimation kernel: the two-group (opposite-phase) build, then the one-group build
set -u
mkdir -p gpurun_out
O=gpurun_out
prof() {
  ncu --set full --clock-control none --import-source on -k regex:'v2_fused_kernel' --launch-skip 3 -c 1 \
      -o $O/u_$1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/u_ncu_$1.log 2>&1
  ncu -i $O/u_$1.ncu-rep --page details > $O/u_$1_details.txt 2>/dev/null
  ncu -i $O/u_$1.ncu-rep --page raw --csv > $O/u_$1_raw.csv 2>/dev/null
  ncu -i $O/u_$1.ncu-rep --page source --csv > $O/u_$1_source.csv 2>/dev/null
  rm -f $O/u_$1.ncu-rep
  grep -E "Duration|Issue Slots Busy|Executed Ipc Active" $O/u_$1_details.txt | head -5
}
prof pp
cp sfft_b200/libsfft_k512.so sfft_b200/libsfft.so
prof k512
timeout 300 python bench.py --no-extras --no-cpu-baseline > $O/u_bench_k512.json 2>/dev/null
cut -c1-200 $O/u_bench_k512.json
