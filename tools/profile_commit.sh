#!/bin/bash
# Round-end measurement pass on one B200 (run under gpurun): bench line, reference arm, ncu launch list of the
# commit at the bench size and one `--set full` capture of the dominant kernels (never a bench number under ncu).
tag=${1:-r01q}
out=gpurun_out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
python bench.py --impl reference > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_2p22.csv \
    python tools/quick_bench.py 22 128 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hash_leaves_colmajor|pass_strided|pass_last|hash_level" \
    -s 14 -c 8 -o $out/${tag}_prof python tools/quick_bench.py 22 128 > $out/${tag}_ncu.log 2>&1
ncu -i $out/${tag}_prof.ncu-rep --page raw --csv > $out/${tag}_prof_raw.csv 2>/dev/null
ls -la $out | tail -8
