#!/bin/bash
# One `ncu --set full` capture per hot kernel of the commit at the bench size (run under gpurun; never a bench number).
#   leaf hashing: the 2nd launch; NTT: the six passes of the 2nd commit (3 iFFT + 3 coset-LDE); tree levels: the widest one.
out=gpurun_out; tag=${1:-r01v4}
ncu --set full --clock-control none --import-source on -k regex:"hash_leaves_colmajor" -s 1 -c 1 -o $out/${tag}_leaf python tools/quick_bench.py 22 128 > $out/${tag}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pass_strided|pass_last" -s 6 -c 6 -o $out/${tag}_ntt python tools/quick_bench.py 22 128 >> $out/${tag}_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hash_level" -s 19 -c 1 -o $out/${tag}_level python tools/quick_bench.py 22 128 >> $out/${tag}_ncu.log 2>&1
for k in leaf ntt level; do ncu -i $out/${tag}_$k.ncu-rep --page raw --csv > $out/${tag}_${k}_raw.csv 2>/dev/null; done
ncu -i $out/${tag}_leaf.ncu-rep --page source --csv > $out/${tag}_leaf_source.csv 2>/dev/null
# the reports themselves exceed the 64 MiB that gpurun copies back: keep the exported pages only
rm -f $out/${tag}_leaf.ncu-rep $out/${tag}_ntt.ncu-rep $out/${tag}_level.ncu-rep
ls -la $out | tail -12
