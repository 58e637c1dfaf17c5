set -x
for v in base r2s1; do
  echo "== $v steady"; VBOC_LIB=$PWD/vboc_b200/variants/$v.so python tools/prof_run.py 37888 3 | tail -2
  echo "== $v bounded8"; VBOC_LIB=$PWD/vboc_b200/variants/$v.so python tools/prof_run.py 8192 8 | tail -2
  echo "== $v lone"; VBOC_LIB=$PWD/vboc_b200/variants/$v.so python tools/prof_run.py 1 20 | tail -2
done > gpurun_out/r2_ab1.log 2>&1
cat gpurun_out/r2_ab1.log
python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_gputest2.log 2>&1
tail -8 gpurun_out/r2_gputest2.log
