#!/bin/bash
# compute-sanitizer over the smallest tests of every kernel family (run under gpurun); full logs under gpurun_out/<tag>san/
TAG=${1:-r2}
OUT=gpurun_out/${TAG}san
mkdir -p $OUT
run_san() {   # name tool pytest-args...
  timeout 900 compute-sanitizer --tool $2 --error-exitcode 9 --print-limit 5 python -m pytest "${@:3}" -m gpu -q -x -p no:cacheprovider > $OUT/$1.log 2>&1
  echo "$1: exit code $? ; $(grep -E 'ERROR SUMMARY|passed|failed' $OUT/$1.log | tr '\n' ' ')"
}
run_san memcheck_rowops_layout memcheck tests/test_rowops_gpu.py tests/test_layout_gpu.py
run_san memcheck_gemm memcheck tests/test_gemm_gpu.py -k "epilogue or layouts_tiles or grouped"
run_san memcheck_attention memcheck tests/test_attention_gpu.py -k "10-3-2 or 10-2-2 or 10-1-3 or 13-1-2 or 19-1-2"
run_san memcheck_precise memcheck tests/test_precise_gpu.py -k "hilo or 160"
run_san memcheck_trainer memcheck tests/test_trainer_gpu.py -k "fused_step or stand_in or three_task"
run_san synccheck_rowops_attention synccheck tests/test_rowops_gpu.py tests/test_attention_gpu.py -k "10-3-2 or 10-2-2 or 10-1-3 or layernorm"
run_san racecheck_rowops_layout racecheck tests/test_rowops_gpu.py tests/test_layout_gpu.py
run_san racecheck_attention racecheck tests/test_attention_gpu.py -k "10-3-2 or 10-2-2 or 10-1-3"
