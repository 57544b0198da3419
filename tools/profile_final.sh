#!/bin/bash
# Final-state profiling pass of round 2 (kernels renamed / added late in the round): launch list, GEMM DRAM traffic, `--set full` captures of
# the specialised-epilogue GEMM, the tensor-core rel-pos dense attention, the fused RVSA backward tail and the sampling heads' forward,
# compute-sanitizer over the tests that exercise them.     usage: tools/profile_final.sh <tag>
TAG=${1:-r2final}
OUT=gpurun_out/${TAG}prof
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gemm-share"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file $OUT/launches.csv $B > $OUT/launches.out 2>&1
ncu -k regex:gemm_bf16 --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 203 -c 203 --csv \
    --log-file $OUT/gemm_dram.csv $B > /dev/null 2>&1
cap() {   # name regex skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $OUT/$1 $B > $OUT/$1.out 2>&1
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  if [ "$4" = "src" ]; then ncu -i $OUT/$1.ncu-rep --page source --csv > $OUT/$1_src.csv 2>/dev/null; fi
  rm -f $OUT/$1.ncu-rep
}
cap gemm_fc1_fwd gemm_bf16 210 src
cap gemm_fc2_dual gemm_bf16 300
cap dense_attn_bwd full_attn_bwd_tc 5
cap dense_attn_fwd full_attn_fwd_tc 9
cap rvsa_bwd_tail rvsa_bwd_tail 22
cap sampling_fwd rvsa_sampling_fused_fwd 22
S=$OUT/sanitizer.log
: > $S
run_san() {   # tool, pytest args
  echo "==== compute-sanitizer --tool $1 :: ${@:2}" >> $S
  timeout 900 compute-sanitizer --tool $1 --error-exitcode 9 --print-limit 20 python -m pytest "${@:2}" -m gpu -q -x -p no:cacheprovider 2>&1 | tail -12 >> $S
  echo "exit code: ${PIPESTATUS[0]}" >> $S
}
run_san memcheck tests/test_gemm_gpu.py -k "specialised and (64 or 1256)"
run_san memcheck tests/test_attention_gpu.py -k "14-2-2 or 10-2-2"
run_san racecheck tests/test_attention_gpu.py -k "14-2-2"
run_san synccheck tests/test_attention_gpu.py -k "14-2-2"
ls -la $OUT | head -40
