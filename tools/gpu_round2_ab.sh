for v in base o16 o16u2 o20u2 base o16u2; do
  echo "== $v"
  VBOC_LIB=$PWD/vboc_b200/variants/$v.so timeout 600 python tools/prof_run.py 37888 3 2>&1 | tail -1
done
