#!/bin/bash
# Second profiling pass: `--set full` captures of the remaining kernels of the headline step and of the long-sequence kernels (config c4).
TAG=${1:-r2}
OUT=gpurun_out/${TAG}prof2
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gemm-share"
cap() {   # name regex skip [extra bench args]
  ncu --set full --clock-control none -k regex:$2 -s $3 -c 1 -f -o $OUT/$1 $B $4 $5 > $OUT/$1.out 2>&1
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  rm -f $OUT/$1.ncu-rep
}
cap gemm_pair_dual "gemm_bf16_kernel<256" 60
cap gemm_bn128_resid "gemm_bf16_kernel<128" 60
cap nchw_to_tok nchw_to_tok 5
cap patchify_u8 patchify_u8 2
cap sqloss sqloss 14
cap sampling_bwd rvsa_sampling_bwd_kernel 22
cap sampling_wgrad rvsa_sampling_wgrad 22
cap partials_reduce rvsa_partials_reduce 22
cap colsum colsum_bf16 8
cap scale_cast scale_cast 8
cap maxpool maxpool2_tok_fwd 2
cap ln_fwd_gelu "ln_fwd_kernel<__nv_bfloat16" 2
cap dense_stream_fwd full_attn_fwd_stream_tc 5 --config c4
cap dense_stream_bwd full_attn_bwd_stream_tc 5 --config c4
cap dense_bwd_prep dense_bwd_prep 5 --config c4
cap dense_bwd_finish dense_bwd_finish 5 --config c4
ls $OUT | head -50
