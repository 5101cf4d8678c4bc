#!/usr/bin/env python
"""bench.py — headline measurement of the B200 STARK hot path (driver contract in the task statement).

Workload at any N: BASELINE.json configs[1], the `PolynomialBatch::from_values` commit microbench —
synthetic Goldilocks 2^22 rows x 128 columns, rate_bits=1, cap_height=4, Poseidon-12 — one "step" =
one commit (iFFT -> coset LDE -> leaf hashing -> Merkle levels -> cap) of one batch.  Each rank
(one process per GPU) commits its own batch: independent tables shard with no collective
(SURVEY.md 8(e)), scaling = "weak".  metric = "commit HBM GB/s": algorithmic bytes of SURVEY.md 8(d)
  B(N,C,r,h) = 8CN + 8CN + 8CN*2^r + 32*(2*(N*2^r - 2^h) + 2^h)
divided by the device time (CUDA events on the library's stream, max over ranks).

  value     inputs resident in HBM (etp_batch_recommit_values_dev)
  e2e       the same commit through the host-buffer C-ABI call etp_batch_from_values_host: pinned host
            columns -> H2D -> commit -> cap D2H, all inside the timed region
  roofline  dominant kernel (leaf hashing), live CUDA-event time; plus per-kernel lines in `kernels`
  cpu_baseline  the oracle (C restatement, OpenMP, all host cores) on a bounded sample of the workload
  stark     BASELINE.json configs[2]: single-table STARK prove (memory-shaped table 2^22 rows): ms and
            proofs/min (whole job, all ranks; two prover contexts per GPU)
  tx        synthetic transaction: seven table proofs of the evm_arithmetization shapes (BASELINE configs[3]/[4] shape
            only: no CTLs, no recursion), ms per transaction and transactions/min (whole job, all ranks)

`--impl reference`: the CPU implementation of the same path on the host cores.  The reference's own
prover is Rust in un-vendored crates and cannot be built here, so this arm runs the oracle
port (`oracle/liboracle.so`) with all host threads on a bounded sample, as the task statement allows.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N, COLS, RATE_BITS, CAP_HEIGHT = 22, 128, 1, 4
STARK_LOG_N = 22
STARK_CONTEXTS_PER_GPU = 2  # parallel.ProverPool: the tail of one proof overlaps the commits of the next (tools/prove_concurrent.py)
CPU_SAMPLE_LOG_N = 18  # bounded CPU sample: 2^18 x 128 (1/16 of the rows; ~10-30 s of CPU work)


def commit_bytes(log_n, cols, r=RATE_BITS, h=CAP_HEIGHT):
    n = 1 << log_n
    return 8 * cols * n * (2 + (1 << r)) + 32 * (2 * ((n << r) - (1 << h)) + (1 << h))


def commit_perms(log_n, cols, r=RATE_BITS, h=CAP_HEIGHT):
    n = 1 << log_n
    return (n << r) * ((cols + 7) // 8) + ((n << r) - (1 << h))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # median over samples taken under load (above 60% of max, else all)
        load = [x for x, m in zip(sm, mx) if x > 0.6 * m] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_commit_gbs(log_n, cols, repeats=1):
    """Oracle (port) commit on the host cores; returns (GB/s algorithmic, seconds, threads)."""
    import numpy as np

    import oracle
    from eth_tx_proof_b200 import synthetic as syn

    vals = syn.random_columns(cols, log_n, seed=0xB200)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        b = oracle.Batch.from_values(vals, RATE_BITS, CAP_HEIGHT)
        dt = time.perf_counter() - t0
        del b
        best = dt if best is None else min(best, dt)
    return commit_bytes(log_n, cols) / best / 1e9, best, oracle.num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm uses all host cores (set before liboracle loads)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    steps = max(1, args.steps)
    for _ in range(min(args.warmup, 1)):
        cpu_commit_gbs(14, COLS)
    times = []
    threads = 1
    for _ in range(min(steps, 3)):
        gbs, dt, threads = cpu_commit_gbs(CPU_SAMPLE_LOG_N, COLS)
        times.append(dt)
    ms = 1e3 * statistics.mean(times)
    value = commit_bytes(CPU_SAMPLE_LOG_N, COLS) / (ms / 1e3) / 1e9
    sample = f"from_values 2^{CPU_SAMPLE_LOG_N} x {COLS} (rows/16 of the 2^{LOG_N} workload), {len(times)} timed commits"
    print(json.dumps({
        "impl": "reference", "metric": "commit_hbm_gbs", "value": value, "unit": "GB/s", "n_gpus": world, "steps": len(times),
        "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"PolynomialBatch::from_values commit, 2^{LOG_N} x {COLS}, rate_bits=1, cap_height=4, Poseidon-12",
                   "note": "CPU arm = oracle port (C + OpenMP restatement of plonky2), NOT plonky2 itself: the Rust reference cannot be built here"},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def run_column_split(etp, ctx, torch, dist, rank, world, log_n, cols, n, nbytes, barrier, max_over_ranks):
    """One table column-split across all ranks: strong scaling of a single commit.  Rank g transforms
    columns [g*C/G, (g+1)*C/G) and hashes leaf rows [g*L/G, (g+1)*L/G), reading the peers' LDE columns over NVLink
    inside the hashing kernel; the cap parts are all-gathered over NCCL."""
    from eth_tx_proof_b200 import parallel

    g2 = torch.Generator(device="cuda").manual_seed(0xC0)
    c0, c1 = parallel.column_split_plan(cols, 2 * n, CAP_HEIGHT, rank, world)["cols"]
    xs = torch.randint(0, 2**62, (cols, n), dtype=torch.int64, device="cuda", generator=g2)[c0:c1].contiguous()
    torch.cuda.synchronize()  # the library works on its own stream
    shard = etp.BatchShard(ctx, cols, log_n, RATE_BITS, CAP_HEIGHT, rank, world)
    cap0 = parallel.commit_column_split(shard, values_dev=(xs.data_ptr(), n))
    for _ in range(2):
        parallel.recommit_column_split(shard, (xs.data_ptr(), n))
    barrier()
    t0 = time.perf_counter()
    reps = 4
    for _ in range(reps):
        cap1 = parallel.recommit_column_split(shard, (xs.data_ptr(), n))
    torch.cuda.synchronize()
    dt = max_over_ranks((time.perf_counter() - t0) / reps)
    assert (cap0 == cap1).all()
    parallel.finish_column_split(shard)
    del shard, xs
    return {"workload": f"ONE 2^{log_n} x {cols} table column-split over {world} GPUs (CUDA IPC + NVLink peer loads fused into the "
                        "leaf-hash kernel; cap parts all-gathered over NCCL)", "ms_per_commit": dt * 1e3,
            "value": nbytes / dt / 1e9, "unit": "GB/s", "scaling": "strong",
            "timed": "local columns resident in HBM -> whole cap on every rank (wall clock incl. 2 barriers, max over ranks)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=LOG_N)
    ap.add_argument("--skip-stark", action="store_true")
    ap.add_argument("--skip-tx", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-split", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import eth_tx_proof_b200 as etp

    warmup = max(args.warmup, 3)
    steps = max(args.steps, 1)
    log_n, cols = args.log_n, COLS
    n = 1 << log_n
    torch.cuda.set_device(local_rank)
    # stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner there) write to file
    # descriptor 1 directly, so everything but the final line is sent to stderr at the descriptor level
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = etp.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream)
    # synthetic trace, uniform in [0, p): hi == 0xFFFFFFFF and lo != 0 would be >= p -> fold back
    g = torch.Generator(device="cuda").manual_seed(0xB200 + rank)
    lo = torch.randint(0, 2**32, (cols, n), dtype=torch.int64, device="cuda", generator=g)
    hi = torch.randint(0, 2**32, (cols, n), dtype=torch.int64, device="cuda", generator=g)
    over = (hi == 0xFFFFFFFF) & (lo != 0)
    hi = torch.where(over, torch.zeros_like(hi), hi)
    lo = torch.where(over, lo - 1, lo)
    x = (hi << 32) | lo  # int64 storage of the u64 bit pattern
    del lo, hi, over
    torch.cuda.synchronize()

    batch = etp.PolynomialBatch.from_values_dev(ctx, x.data_ptr(), n, cols, log_n, RATE_BITS, False, CAP_HEIGHT)
    for _ in range(warmup):
        batch.recommit_values_dev(x.data_ptr(), n)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count
    barrier()
    phase = {"IFFT": 0.0, "FFT + blinding": 0.0, "build Merkle tree (leaves)": 0.0, "build Merkle tree (levels)": 0.0}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
    for _ in range(steps):
        batch.recommit_values_dev(x.data_ptr(), n)  # synchronous: returns once the cap is on the host
        for k, v in batch.last_commit_timings().items():
            phase[k] += v
    with torch.cuda.stream(stream):
        e1.record()
    barrier()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count - launches0
    clocks = sampler.stop()
    ms_per_step = total_ms / steps
    nbytes = commit_bytes(log_n, cols)
    value = world * nbytes / (ms_per_step / 1e3) / 1e9
    phase = {k: v / steps for k, v in phase.items()}

    # ---- e2e: host buffers through the C ABI, H2D + D2H inside the timed region
    e2e = None
    if not args.skip_e2e:
        host = torch.empty((cols, n), dtype=torch.int64).pin_memory()
        host.copy_(x)
        torch.cuda.synchronize()
        harr = host.numpy().view(np.uint64)
        del batch  # free ~17 GiB before the second resident copy
        e2e_steps = max(2, min(steps, 4))
        b2 = etp.PolynomialBatch.from_values(ctx, harr, RATE_BITS, False, CAP_HEIGHT)  # warm-up (allocations, pools)
        cap_ref = b2.cap.copy()
        del b2
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            b2 = etp.PolynomialBatch.from_values(ctx, harr, RATE_BITS, False, CAP_HEIGHT)
            cap = b2.cap
            del b2
        ctx.synchronize()
        dt = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        assert (cap == cap_ref).all()
        e2e = {"value": world * nbytes / dt / 1e9, "unit": "GB/s", "ms_per_step": dt * 1e3, "h2d_bytes_per_step": 8 * cols * n,
               "d2h_bytes_per_step": 32 << CAP_HEIGHT, "call": "etp_batch_from_values_host (pinned host columns) + etp_batch_cap"}
        del host, harr
    else:
        del batch

    # ---- BASELINE configs[2] / [4] shape: single-table STARK proofs (memory-shaped table), 8 independent
    # "segment" jobs sharded over the ranks with no collective (eth_tx_proof_b200/parallel.py)
    stark = None
    if not args.skip_stark:
        from eth_tx_proof_b200 import parallel, synthetic as syn

        sl = min(STARK_LOG_N, log_n)
        n_jobs = 8
        my_jobs = parallel.shard_jobs(n_jobs, rank, world)
        trace_pinned = torch.from_numpy(syn.memory_trace(sl, seed=7 + rank).view(np.int64)).pin_memory()
        trace = trace_pinned.cuda()
        ctx.stark_prove_dev(etp.TABLE_MEMORY, sl, trace.data_ptr(), 1 << sl)  # warm-up
        # latency of one proof (one context), then throughput with two prover contexts per GPU (parallel.ProverPool)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            proof = ctx.stark_prove_dev(etp.TABLE_MEMORY, sl, trace.data_ptr(), 1 << sl)
        prove_ms = (time.perf_counter() - t0) * 1e3 / 3
        phases = ctx.last_prove_timings()
        # the same proof through the host-buffer entry point: pinned host trace -> H2D (streamed under the trace commit) -> proof
        trace_host = trace_pinned.numpy().view(np.uint64)
        ctx.stark_prove(etp.TABLE_MEMORY, trace_host)
        t0 = time.perf_counter()
        for _ in range(3):
            proof_h = ctx.stark_prove(etp.TABLE_MEMORY, trace_host)
        prove_host_ms = (time.perf_counter() - t0) * 1e3 / 3
        assert (proof_h == proof).all(), "host-trace proof differs from the resident-trace proof"
        pool = parallel.ProverPool(local_rank, STARK_CONTEXTS_PER_GPU)
        jobs = [(trace.data_ptr(), 1 << sl)] * len(my_jobs)
        pool.stark_prove_dev(etp.TABLE_MEMORY, sl, jobs[:STARK_CONTEXTS_PER_GPU])  # warm-up of every context
        barrier()
        t0 = time.perf_counter()
        proofs = pool.stark_prove_dev(etp.TABLE_MEMORY, sl, jobs)
        local = time.perf_counter() - t0
        dt = max_over_ranks(local)
        assert all((p == proof).all() for p in proofs), "pooled proofs differ from the single-context proof"
        pool.close()
        stark = {"workload": f"starky prove, memory-shaped table 2^{sl} x 21 (+4 aux, 4 quotient), standard_fast_config; "
                             f"{n_jobs} independent segment jobs sharded over {world} GPU(s), {STARK_CONTEXTS_PER_GPU} prover contexts per GPU",
                 "prove_ms": prove_ms, "prove_host_ms": prove_host_ms, "h2d_bytes_per_proof": int(trace_host.nbytes),
                 "proofs_per_min": n_jobs * 60.0 / dt, "jobs": n_jobs, "contexts_per_gpu": STARK_CONTEXTS_PER_GPU,
                 "proof_bytes": int(proof.size * 8), "phases_ms": phases,
                 "timed": "prove_ms: one proof, one context, trace resident in HBM -> complete proof bytes on the host; "
                          "prove_host_ms: the same through etp_stark_prove_host from a pinned host trace (H2D inside); "
                          "proofs_per_min: all jobs through the pool (wall clock, max over ranks)"}
        del trace

    # ---- BASELINE configs[3] / [4] shape: a synthetic TRANSACTION = seven table proofs of the evm_arithmetization shapes
    # (arithmetic, byte packing, cpu, keccak 2400 columns, keccak sponge, logic, memory; cprog.EVM_TABLE_SHAPES, degree bits
    # at the low end of the reference's circuit ranges).  Shape only: no cross-table lookups, no recursion layers.  Tables are
    # registered (NVRTC) once per context, outside the timed region, like the reference builds its circuits at start-up.
    tx = None
    if not args.skip_stark and not args.skip_tx:
        from eth_tx_proof_b200 import parallel, synthetic as syn

        tables = syn.tx_job_tables()
        dev = [torch.from_numpy(t.view(np.int64)).cuda() for _, _, _, t in tables]
        torch.cuda.synchronize()
        pool = parallel.ProverPool(local_rank, STARK_CONTEXTS_PER_GPU)
        ids = [[(c.register_table(p, p.lookups) if p is not None else etp.TABLE_MEMORY) for _, p, _, _ in tables] for c in pool.contexts]

        def prove_table(c, k):
            w = pool.contexts.index(c)
            return c.stark_prove_dev(ids[w][k], tables[k][2], dev[k].data_ptr(), 1 << tables[k][2])

        n_tx = 8
        my_tx = parallel.shard_jobs(n_tx, rank, world)
        flat = [k for _ in my_tx for k in range(len(tables))]
        pool.map(prove_table, list(range(len(tables))) * STARK_CONTEXTS_PER_GPU)  # warm-up: every table on every context
        c0 = pool.contexts[0]
        per_table = {}
        t0 = time.perf_counter()
        for k, (name, _, bits, tr) in enumerate(tables):
            t1 = time.perf_counter()
            pr = prove_table(c0, k)
            per_table[f"{name} 2^{bits} x {tr.shape[0]}"] = {"ms": (time.perf_counter() - t1) * 1e3, "proof_bytes": int(pr.size * 8)}
        tx_ms = (time.perf_counter() - t0) * 1e3
        barrier()
        t0 = time.perf_counter()
        pool.map(prove_table, flat)
        dt = max_over_ranks(time.perf_counter() - t0)
        pool.close()
        del dev
        tx = {"workload": "synthetic transaction: 7 table STARK proofs of the evm_arithmetization shapes (SURVEY.md App. B), "
                          f"shape-only constraint programs, no CTLs, no recursion; {n_tx} transactions sharded over {world} GPU(s), "
                          f"{STARK_CONTEXTS_PER_GPU} prover contexts per GPU",
              "tx_ms": tx_ms, "tx_per_min": n_tx * 60.0 / dt, "transactions": n_tx, "tables": per_table,
              "timed": "tx_ms: the seven proofs in sequence on one context, traces resident in HBM -> proof bytes on the host; "
                       "tx_per_min: all table proofs of all transactions through the pool (wall clock, max over ranks)"}

    # ---- one table column-split across all ranks: strong scaling of a single commit.
    # Rank g transforms columns [g*C/G, (g+1)*C/G) and hashes leaf rows [g*L/G, (g+1)*L/G), reading the peers'
    # LDE columns over NVLink inside the hashing kernel; the cap parts are all-gathered over NCCL.
    split = None
    if world > 1 and not args.skip_split and (world & (world - 1)) == 0 and world <= 8:
        try:
            split = run_column_split(etp, ctx, torch, dist, rank, world, log_n, cols, n, nbytes, barrier, max_over_ranks)
        except Exception as e:  # the weak-scaling line above must survive a box without CUDA IPC between its GPUs
            split = {"error": f"{type(e).__name__}: {e}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines (denominators: MEASURED_PEAKS.json, else the B200_PROFILING.md fallback)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            hbm_peak = json.load(f)["hbm_gbs"]
        peak_src = "MEASURED_PEAKS.json (measured copy bandwidth)"
    else:
        hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    leaf_ms = phase["build Merkle tree (leaves)"]
    leaf_bytes = 8 * cols * (n << RATE_BITS) + 32 * (n << RATE_BITS)  # read the LDE once, write the leaf digests
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = (json.load(f).get(f"hash_leaves_colmajor@2^{log_n}x{cols}") or {}).get("total")
    achieved = leaf_bytes / (leaf_ms / 1e3) / 1e9
    perms_leaf = (n << RATE_BITS) * ((cols + 7) // 8)
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    # integer-pipe roofline of the Poseidon kernels (profiles/README.md, tools/microbench/pipes*.cu): on B200 the integer ALU
    # and the FP64 unit of a sub-partition share one issue port that accepts a warp instruction every 2 clk (16 lanes/clk), and
    # that port is what saturates (ncu: alu % + fp64 % of the leaf kernel).  Static SASS counts per permutation from
    # tools/sass_count.py: 4840 ALU + 4280 FP64 port instructions (the IMAD.WIDE multiplier chains run on the FMA pipe beside it).
    alu_per_perm, fp64_per_perm = 4840, 4280
    port_clk_per_perm = 2.0 * (alu_per_perm + fp64_per_perm)
    perm_peak = 148 * 4 * sm_mhz * 1e6 / port_clk_per_perm * 32
    roofline = {"kernel": "merkle::hash_leaves_colmajor", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "note": "dominant kernel is integer/FP64-issue bound (16 Poseidon permutations per 1 KiB row), not HBM bound; see int_pipe"}
    int_pipe = {"kernel": "merkle::hash_leaves_colmajor", "achieved": perms_leaf / (leaf_ms / 1e3), "peak": perm_peak, "unit": "perm/s",
                "frac": perms_leaf / (leaf_ms / 1e3) / perm_peak,
                "model": "shared ALU/FP64 issue port: 148 SM x 4 SMSP x f_sm x 32 lanes / (2 clk x (4840 ALU + 4280 FP64) port instructions per "
                         "permutation, tools/sass_count.py); f_sm = sampled clock; frac is that port's utilisation (cf. ncu alu % + fp64 %)"}
    ntt_bytes_ifft = 16 * cols * n
    ntt_bytes_lde = 8 * cols * n + 8 * cols * (n << RATE_BITS)
    kernels = [
        {"scope": "IFFT", "ms": phase["IFFT"], "bound": "hbm", "achieved_gbs": ntt_bytes_ifft / (phase["IFFT"] / 1e3) / 1e9,
         "frac": ntt_bytes_ifft / (phase["IFFT"] / 1e3) / 1e9 / hbm_peak},
        {"scope": "FFT + blinding (coset LDE)", "ms": phase["FFT + blinding"], "bound": "hbm",
         "achieved_gbs": ntt_bytes_lde / (phase["FFT + blinding"] / 1e3) / 1e9,
         "frac": ntt_bytes_lde / (phase["FFT + blinding"] / 1e3) / 1e9 / hbm_peak},
        {"scope": "build Merkle tree (leaf hashing)", "ms": leaf_ms, "bound": "int_pipe", "perm_per_s": perms_leaf / (leaf_ms / 1e3)},
        {"scope": "build Merkle tree (levels + cap)", "ms": phase["build Merkle tree (levels)"], "bound": "int_pipe"},
    ]

    cpu = None
    if not args.skip_cpu and world == 1:  # reported at N=1 only (rank 0's cores are shared with the other ranks otherwise)
        gbs, dt, threads = cpu_commit_gbs(CPU_SAMPLE_LOG_N, cols)
        cpu = {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port", "seconds": dt,
               "sample": f"oracle from_values 2^{CPU_SAMPLE_LOG_N} x {cols} (rows/16 of the workload), 1 commit, all host threads; "
                         "C+OpenMP restatement, NOT plonky2 (the Rust reference cannot be built here)"}

    out = {
        "metric": "commit_hbm_gbs", "value": value, "unit": "GB/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic",
        "config": {"workload": f"PolynomialBatch::from_values commit, 2^{log_n} x {cols} Goldilocks, rate_bits=1, cap_height=4, "
                               "Poseidon-12 (BASELINE.json configs[1]); one batch per GPU",
                   "l2": f"inputs ({8 * cols * n >> 20} MiB per batch) are larger than the 126 MB L2; no flush needed",
                   "algorithmic_bytes_per_step": nbytes, "poseidon_permutations_per_step": commit_perms(log_n, cols)},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "int_pipe": int_pipe,
        "kernels": kernels, "phases_ms": phase, "cpu_baseline": cpu, "stark": stark, "tx": tx, "column_split": split,
    }
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(out) + "\n").encode())
    os.close(json_fd)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
