for v in base base_d4 base_d5 r2s1 r2s2_3 r2s2_4 r2s2_5; do
  echo "== $v steady"; VBOC_LIB=$PWD/vboc_b200/variants/$v.so python tools/prof_run.py 37888 3 | tail -1
  echo "== $v lone"; VBOC_LIB=$PWD/vboc_b200/variants/$v.so python tools/prof_run.py 1 20 | tail -2
done > gpurun_out/r2_ab2.log 2>&1
cat gpurun_out/r2_ab2.log
