for v in "" _nowide _nowres _legacy; do
  for rep in 1 2; do
    echo "== variant '${v}' rep $rep"
    GG_LIB=$PWD/gnnome_assembly_b200/libgnnome_b200${v}.so timeout 100 python tools/sweep.py 1000000 128,256 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    j=json.loads(l); print(j['d'], round(j['fwd_ms'],3), round(j['fwd_bwd_ms'],3))
"
  done
done
