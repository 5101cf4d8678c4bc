#!/bin/bash
# NVLink traffic of the fused "all-gather of row tiles + leaf hashing" kernel (needs >= 2 GPUs: gpurun --gpus 2).
out=gpurun_out; tag=${1:-r02_colsplit_nvlink}; logn=${2:-22}
python tools/colsplit_inprocess.py $logn 128 2 2>&1 | tail -3 | tee $out/${tag}.log
ncu --metrics gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_aperture_peer.sum,lts__t_sectors_srcunit_tex_aperture_peer_op_read.sum \
    --clock-control none -k regex:hash_leaves_colmajor --csv --log-file $out/${tag}_ncu.csv python tools/colsplit_inprocess.py $logn 128 2 > $out/${tag}_ncu.log 2>&1
tail -4 $out/${tag}_ncu.log; grep -c hash_leaves $out/${tag}_ncu.csv
