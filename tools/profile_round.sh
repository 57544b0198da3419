#!/bin/bash
# Profiling pass of one round (run under gpurun on ONE B200): launch list, GEMM DRAM traffic, one `--set full` capture per kernel family
# (exported to raw / source CSV on the box -- the .ncu-rep files would exceed gpurun's 64 MiB return limit), compute-sanitizer over the
# smallest test of every kernel family.   usage: tools/profile_round.sh <tag>
TAG=${1:-r2}
OUT=gpurun_out/${TAG}prof
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gemm-share"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv $B > $OUT/launches.out 2>&1
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
cap rvsa_attn_bwd rvsa_attn_bwd_tc 22 src
cap rvsa_attn_fwd rvsa_attn_fwd_tc 22 src
cap dense_attn_bwd full_attn_bwd_tc 5
cap dense_attn_fwd full_attn_fwd_tc 9
cap ln_bwd ln_bwd_kernel 60
cap ln_fwd ln_fwd_kernel 60
cap adamw adamw_kernel 1
cap kv_finalize rvsa_kv_finalize 22
cap tok_to_nchw tok_to_nchw 5
cap sampling_fwd rvsa_sampling_fused_fwd 22
# ---- compute-sanitizer (memcheck / synccheck / racecheck) over small tests of each family
S=$OUT/sanitizer.log
: > $S
run_san() {   # tool, pytest args
  echo "==== compute-sanitizer --tool $1 :: ${@:2}" >> $S
  timeout 900 compute-sanitizer --tool $1 --error-exitcode 9 --print-limit 20 python -m pytest "${@:2}" -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 >> $S
  echo "exit code: ${PIPESTATUS[0]}" >> $S
}
run_san memcheck tests/test_rowops_gpu.py tests/test_layout_gpu.py
run_san memcheck tests/test_gemm_gpu.py -k "epilogue or layouts_tiles or grouped"
run_san memcheck tests/test_attention_gpu.py -k "10-3-2 or 10-2-2 or 10-1-3 or 13-1-2"
run_san memcheck tests/test_precise_gpu.py -k "hilo or 160"
run_san synccheck tests/test_rowops_gpu.py tests/test_attention_gpu.py -k "10-3-2 or 10-2-2 or 10-1-3 or layernorm"
run_san racecheck tests/test_rowops_gpu.py tests/test_layout_gpu.py
run_san racecheck tests/test_attention_gpu.py -k "10-3-2 or 10-2-2 or 10-1-3"
ls -la $OUT | head -60
