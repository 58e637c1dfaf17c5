for v in w2 w1c64 w1c72 w1c100 w2c72; do
  echo "== $v"
  VBOC_LIB=$PWD/vboc_b200/variants/$v.so timeout 600 python tools/prof_run.py 37888 3 2>&1 | tail -1
done
