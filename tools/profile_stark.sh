#!/bin/bash
# `ncu --set full` of the openings / FRI / quotient kernels inside memory-table proofs at 2^22 rows (run under gpurun; never a
# bench number): two proofs, the captures come from both (the first one is the warm-up).  Exports the raw page and a
# compact summary; the report itself stays on the box (size).
out=gpurun_out; tag=${1:-r02_stark}; logn=${2:-22}
ncu --set full --clock-control none -k regex:"quotient_kernel|combine_values|combine_norms|fri_fold16|eval_polys_at_two_points|batch_inverse|aux_|lookup_|scan_" \
    -c 120 -o $out/${tag} python tools/prove_once.py $logn > $out/${tag}_ncu.log 2>&1
ncu -i $out/${tag}.ncu-rep --page raw --csv > $out/${tag}_raw.csv 2>/dev/null
python tools/summarize_ncu.py $out/${tag}_raw.csv $out/${tag}_summary.csv
rm -f $out/${tag}.ncu-rep $out/${tag}_raw.csv
tail -3 $out/${tag}_ncu.log
