# usage: bash tools/ab_variants.sh "" _n4 _n2 ...   (library variants built with make VARIANT=..)
for v in "$@"; do echo "variant=$v"; ETP_B200_LIB=$PWD/eth_tx_proof_b200/libetp_b200$v.so python tools/quick_bench.py 22 128 2>&1 | grep -v Warn; done
