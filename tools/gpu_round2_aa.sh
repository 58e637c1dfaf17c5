for v in base scs scg swt base scs; do
  echo "== $v"
  VBOC_LIB=$PWD/vboc_b200/variants/$v.so timeout 600 python tools/prof_run.py 37888 3 2>&1 | tail -1
done
