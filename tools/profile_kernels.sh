out=gpurun_out; tag=r01r
ncu --set full --clock-control none --import-source on -k regex:"hash_leaves_colmajor" -s 1 -c 1 -o $out/${tag}_leaf python tools/quick_bench.py 22 128 > $out/${tag}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pass_strided|pass_last" -s 6 -c 6 -o $out/${tag}_ntt python tools/quick_bench.py 22 128 >> $out/${tag}_ncu.log 2>&1
ncu -i $out/${tag}_leaf.ncu-rep --page raw --csv > $out/${tag}_leaf_raw.csv 2>/dev/null
ncu -i $out/${tag}_ntt.ncu-rep --page raw --csv > $out/${tag}_ntt_raw.csv 2>/dev/null
