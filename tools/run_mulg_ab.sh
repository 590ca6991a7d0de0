# A/B: output-row-park polymul at N=8192 with three resident 256-thread CTAs (80 registers, no spills) against four (64, spills)
for v in "" "FHE_B200_LIB=fhe_study_b200/variants/lib_mulg3.so" "FHE_NTT_LOGE=4" ""; do
  echo "== $v" >> gpurun_out/f34_ab.log
  env $v timeout 200 python tools/ntt_ab.py 13 14 2>&1 | grep -v 4611686 >> gpurun_out/f34_ab.log
done
