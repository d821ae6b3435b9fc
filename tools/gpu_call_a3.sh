#!/bin/bash
# round 2, call A3 (1 GPU): parity suite on the new gather / vote / v3 team kernels, A/B timings, ncu of C2
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/a3_pytest.log 2>&1; echo "pytest rc $?" >> $O/a3_pytest.log
: > $O/a3_ab.jsonl
for wl in C2 C4 C5 C1; do
  for un in 4 8; do
    SFFTB_GATHER_UNROLL=$un timeout 300 python tools/gather_ab.py $wl >> $O/a3_ab.jsonl 2>> $O/a3_ab.err
  done
done
: > $O/a3_v3_peel_profile.txt
for tm in 1 4 8 16; do
  SFFTB_V3_TEAM=$tm timeout 300 python tools/gather_ab.py C3 >> $O/a3_ab.jsonl 2>> $O/a3_ab.err
  echo "team $tm" >> $O/a3_v3_peel_profile.txt
  SFFTB_V3_TEAM=$tm timeout 300 python tools/v3_peel_profile.py >> $O/a3_v3_peel_profile.txt 2>> $O/a3_ab.err
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'v2_fused_kernel|gather_kernel' -s 4 -c 2 \
  -o $O/a3_prof_C2 python tools/gather_ab.py C2 2 > /dev/null 2>> $O/a3_ab.err
ncu -i $O/a3_prof_C2.ncu-rep --page raw --csv > $O/a3_prof_C2_raw.csv 2>/dev/null
ls -la $O | tail -8
