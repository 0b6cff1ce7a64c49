O=gpurun_out
timeout 200 python scripts/gelu_table.py 2>&1 | grep -v Warn > $O/r2_gelu_table.log; cat $O/r2_gelu_table.log
for g in half t32; do
  echo "== CFRB_X2_GELU=$g"
  CFRB_X2_GELU=$g MODES=tcx2 timeout 300 python scripts/p5_debug.py 1 6 64 2>&1 | grep -v Warn | grep "tcx2\|mean target"
  CFRB_X2_GELU=$g MODES=tcx2 timeout 300 python scripts/p5_debug.py 1 4 64 2>&1 | grep -v Warn | grep "tcx2"
  CFRB_X2_GELU=$g timeout 200 python scripts/datagen_probe.py --waves 4
  CFRB_X2_GELU=$g timeout 200 python scripts/tc3_check.py 2>&1 | grep "NET_TC_F16X2" | tail -3
done > $O/r2_gelu_variants.log 2>&1
cat $O/r2_gelu_variants.log
