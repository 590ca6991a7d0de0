# end-of-round captures of the external product (pair kernel) and the tcgen05 key switch
P="--set full --clock-control none --import-source on"
cap() { name=$1; rx=$2; skip=$3; shift 3
  "$@" > gpurun_out/plain_$name.log 2>&1 && ncu $P -k regex:$rx -s $skip -c 1 -o gpurun_out/$name "$@" > gpurun_out/ncu_$name.log 2>&1; echo "$name rc=$?"
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep gpurun_out/$name.csv > /dev/null 2>&1 && rm -f gpurun_out/$name.ncu-rep; }
cap f37_extprod extprod_fused 2 python tools/prof.py extprod 1024 1 1184
cap f37_kstc ks_tc_kernel 2 python tools/prof.py bootstrap 8192
cap f37_bfv bfv_mul_kernel 2 python tools/prof.py bfv 1048576
