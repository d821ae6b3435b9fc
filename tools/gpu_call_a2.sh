#!/bin/bash
# round 2, call A2 (1 GPU): persistent gather + interleaved v2 estimation: parity suite, A/B timings, ncu
set -u
mkdir -p gpurun_out
O=gpurun_out
SFFTB_V2_INTERLEAVE=1 timeout 900 python -m pytest tests -m gpu -x -q > $O/a2_pytest.log 2>&1; echo "pytest rc $?" >> $O/a2_pytest.log
: > $O/a2_ab.jsonl
for il in 0 1; do
  SFFTB_V2_INTERLEAVE=$il timeout 300 python tools/gather_ab.py C2 20 >> $O/a2_ab.jsonl 2>> $O/a2_ab.err
done
for wl in C2 C4 C5 C1; do
  for un in 4 8 12; do
    SFFTB_GATHER_UNROLL=$un timeout 300 python tools/gather_ab.py $wl >> $O/a2_ab.jsonl 2>> $O/a2_ab.err
  done
done
for tm in 1 4 8 16; do
  SFFTB_V3_TEAM=$tm timeout 300 python tools/gather_ab.py C3 >> $O/a2_ab.jsonl 2>> $O/a2_ab.err
  SFFTB_V3_TEAM=$tm timeout 300 python tools/v3_peel_profile.py >> $O/a2_v3_peel_profile.txt 2>> $O/a2_ab.err
done
SFFTB_V2_INTERLEAVE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'v2_fused_kernel|gather_kernel' -s 4 -c 2 \
  -o $O/a2_prof_C2 python tools/gather_ab.py C2 2 > /dev/null 2>> $O/a2_ab.err
ncu -i $O/a2_prof_C2.ncu-rep --page raw --csv > $O/a2_prof_C2_raw.csv 2>/dev/null
ls -la $O | tail -8
