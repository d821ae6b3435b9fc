#!/bin/bash
# round 2, call A (1 GPU): full GPU test-suite, bench with extras at N=1, gather A/B + ncu bytes per sample
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/a_pytest.log 2>&1; echo "pytest rc $?" >> $O/a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/a_bench_C2.json 2> $O/a_bench_C2.err
: > $O/a_gather_ab.jsonl
for wl in C2 C5 C4; do
  for fill in 64 128; do
    for l2 in default 32; do
      if [ $l2 = default ]; then
        SFFTB_GATHER_FILL=$fill timeout 300 python tools/gather_ab.py $wl >> $O/a_gather_ab.jsonl 2>> $O/a_gather_ab.err
      else
        SFFTB_GATHER_FILL=$fill SFFTB_L2_FETCH=$l2 timeout 300 python tools/gather_ab.py $wl >> $O/a_gather_ab.jsonl 2>> $O/a_gather_ab.err
      fi
    done
  done
done
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,dram__sectors_read.sum,smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct
for wl in C2 C5 C4; do
  for fill in 64 128; do
    SFFTB_GATHER_FILL=$fill timeout 300 ncu --metrics $M --clock-control none -k regex:gather_kernel -s 2 -c 2 --csv \
      --log-file $O/a_ncu_gather_${wl}_fill${fill}.csv python tools/gather_ab.py $wl 3 > /dev/null 2>> $O/a_gather_ab.err
  done
done
SFFTB_GATHER_FILL=64 SFFTB_L2_FETCH=32 timeout 300 ncu --metrics $M --clock-control none -k regex:gather_kernel -s 2 -c 2 --csv \
  --log-file $O/a_ncu_gather_C2_fill64_l2f32.csv python tools/gather_ab.py C2 3 > /dev/null 2>> $O/a_gather_ab.err
for lg in 24 27; do
  timeout 120 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/a_ld_variants_${lg}.csv tools/microbench/ld_variants $lg > $O/a_ld_variants_${lg}.out 2>&1
done
ls -la $O | tail -30
