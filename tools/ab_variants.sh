for v in "" _b4 _b5 _b6 _b8; do echo "variant=$v"; ETP_B200_LIB=$PWD/eth_tx_proof_b200/libetp_b200$v.so python tools/quick_bench.py 20 128 2>&1 | grep -v Warn; done
