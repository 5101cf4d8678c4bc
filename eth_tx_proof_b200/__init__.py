"""eth_tx_proof_b200 — B200-native (sm_100a) STARK proving hot path behind plonky2's operator surface.

Host-side mirror (Python, for tests and benches) of the C ABI in ``include/etp_b200.h``; the names
follow plonky2 0.2.2 / starky 0.4.0 (``PolynomialBatch.from_values``, ``MerkleTree.new`` …), the crates
the reference worker reaches through ``generate_txn_proof`` (/root/reference/ops/src/lib.rs:52).

There is NO CPU fallback: importing is cheap, but any compute call raises ``EtpError`` unless
``libetp_b200.so`` is built (``python -c "import __graft_entry__ as g; g.build()"``) and a CUDA device
is present.  Nothing in this package imports ``oracle/``.
"""
from .api import (  # noqa: F401
    BatchShard,
    Challenger,
    Context,
    FriParams,
    FriState,
    EtpError,
    MerkleTree,
    PolynomialBatch,
    TABLE_FIBONACCI,
    TABLE_MEMORY,
    lib_path,
    load_library,
)
