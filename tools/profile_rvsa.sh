OUT=gpurun_out/r3iprof; mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gemm-share"
cap() { ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $OUT/$1 $B > $OUT/$1.out 2>&1
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  ncu -i $OUT/$1.ncu-rep --page source --csv > $OUT/$1_src.csv 2>/dev/null
  rm -f $OUT/$1.ncu-rep; }
cap rvsa_attn_bwd rvsa_attn_bwd_tc 22
cap rvsa_attn_fwd rvsa_attn_fwd_tc 22
