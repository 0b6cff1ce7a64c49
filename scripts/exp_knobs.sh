O=gpurun_out
timeout 800 python -m pytest tests/test_rela_module.py -q -m gpu -k "datagen_distribution or config5_with" 2>&1 | grep -v Warning | tail -40 > $O/r2_newtests.log; tail -4 $O/r2_newtests.log
cp $O/parity_notes.log $O/r2_parity_notes_new.log 2>/dev/null
for v in 8 7 6 5; do echo "CTAS_PER_SM=$v"; CFRB_D2_CTAS_PER_SM=$v timeout 200 python scripts/datagen_probe.py --waves 4; done > $O/r2_knobs.log 2>&1
for v in 2 1; do echo "D2_GROUPS=$v"; CFRB_D2_GROUPS=$v timeout 200 python scripts/datagen_probe.py --waves 4; done >> $O/r2_knobs.log 2>&1
cat $O/r2_knobs.log
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum -k regex:'cfr_iter_d2|leaf_mlp_tc3' -s 600 -c 12 --csv --log-file $O/r2_warm_traffic.csv python scripts/datagen_probe.py --iters 64 --waves 1 --warm 6 > $O/r2_warm_traffic.log 2>&1
tail -30 $O/r2_warm_traffic.csv | cut -c1-200
