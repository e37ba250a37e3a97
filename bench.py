#!/usr/bin/env python
"""bench.py -- KKW prover AND-gates/sec (GF(2), 8 players x 256 repetitions, 40 opened) on B200, next to the CPU port.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line on rank 0.
For N > 1 the driver launches it under torchrun, one rank per GPU; the 32 packed instances (src/proof/mod.rs:127-157) are
sharded 32/N per rank and the only exchange is the all-gather of the 256 x 32-byte repetition hashes
(src/proof/mod.rs:160-171) -> strong scaling (the proof is the same for every N).

A step = one batch of B independent `Proof::new` calls on the workload circuit (default: SHA-256 compression, SURVEY.md
8(d) config 2; B = --batch, default 32) issued together, the way a proving service sees its queue; B = 1 gives the
single-proof latency, which is also reported.  Small GF(2) circuits are held by multi-proof sessions (--per-session proofs
side by side, every kernel launch covering all of them); big / Z64 circuits one proof per session (--batch 1).
  value  device-resident: witnesses + seeds already in HBM; the step is one CUDA graph launch per phase for all sessions
         (rv_batch) on a leader stream that carries the CUDA-event pair of the step
  e2e    host buffers in, proof bytes out, through the public batched call (Proof.new_batch -> rv_prove_batch), all
         host<->device copies inside the timed region; N > 1: upload, sharded step, NCCL assembly on rank 0
  verify Proof.verify of the same proofs (host bytes in), B verifications in flight
The CPU arm (--impl reference) proves the same B proofs per step with the C restatement of the reference's dataflow.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "KKW prover AND-gates/sec (GF(2), 128-bit sec)"
UNIT = "AND-gates/s"


def make_workload(name: str):
    from reverie_b200 import circuits as C

    nz = np.zeros(0, dtype=np.uint64)
    zw = np.array([0x0123456789ABCDEF, 0xFEDCBA9876543210], dtype=np.uint64)
    if name.startswith("z64mul"):
        n = int(name[6:])
        ops, nw = C.z64_mul_circuit(n)
        return ops, np.zeros(0, dtype=np.uint8), zw, (nw, 0), f"Z64 synthetic arithmetic circuit: 2 inputs + {n} x Mul over a 1024-cell register file (SURVEY.md 8(d) config 3), 256 reps x 8 players"
    if name.startswith("z64flat"):
        n = int(name[7:])
        ops, wc = C.flat_mul_circuit(n, domain=C.Z64)
        return ops, np.zeros(0, dtype=np.uint8), zw, wc, f"Z64 flat circuit: 2 inputs + {n} x Mul(2,0,1) (src/proof/mod.rs:322-329 over Z64)"
    if name == "sha256":
        ops, wit, wc = C.sha256_abc_case()
        return ops, wit, nz, wc, "GF(2) SHA-256 compression circuit (generated Bristol-style: 22573 AND / 93666 XOR / 2147 INV, 768 inputs, 256 output asserts), 256 reps x 8 players"
    if name.startswith("flat"):
        n = int(name[4:])
        ops, wc = C.flat_mul_circuit(n)
        return ops, np.array([1, 1], dtype=np.uint8), nz, wc, f"GF(2) flat circuit: 2 inputs + {n} x Mul(2,0,1) (src/proof/mod.rs:322-329 scaled)"
    if name.startswith("layered"):
        n = int(name[7:])
        width = min(1 << 20, max(1024, n // 16))
        ops, nw = C.layered_and_circuit(width, n)
        wit = np.random.default_rng(0).integers(0, 2, size=width).astype(np.uint8)
        return ops, wit, nz, (0, nw), f"GF(2) layered circuit: {width} inputs + {n} ANDs in layers of {width}"
    raise SystemExit(f"unknown workload {name}")


def default_seeds() -> bytes:
    """256 x 16 bytes of repetition seeds, fixed so that runs are comparable (the reference draws them from OsRng,
    src/proof/mod.rs:131-134).  Both arms use the same ones."""
    return np.random.default_rng(20261017).integers(0, 256, size=256 * 16, dtype=np.uint8).tobytes()


# bench kernel label -> kernel(s) in the committed ncu capture (profiles/r1c_traffic.json; DRAM bytes per launch)
NCU_NAMES = {"linear": ["k_mask_vm"], "mask_gen": ["k_mask_gen_tt<0>", "k_mask_gen_tt<1>", "k_mask_gen_tt"], "items": ["k_items", "k_items_tile<0>", "k_items_tile<1>"], "chunk_cv": ["k_chunk_cv"],
             "values": ["k_values<1>"], "rep_hash": ["k_rep_hash"], "extract": ["k_extract"], "challenge": ["k_challenge"], "key_setup": ["k_key_setup"],
             "z.mask_gen": ["k_zmask_gen_tt"], "z.items": ["k_zitems_online"], "z.values": ["k_zvalues"], "z.extract": ["k_zextract"]}


def ncu_traffic(workload: str, label: str):
    """DRAM bytes per launch of the kernel behind `label`, from the committed `ncu --set full` capture of the same workload
    (None when that workload / kernel was not captured)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1c_traffic.json")) as f:
            w = json.load(f)["workloads"].get(workload)
        vals = [w[k]["dram_bytes_per_launch"] for k in NCU_NAMES.get(label, []) if k in w]
        return sum(vals) / len(vals) if vals else None
    except Exception:
        return None


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_run(ops, wit, wz, wc, seeds, min_seconds: float, min_proofs: int):
    """The CPU restatement of the reference's dataflow (oracle/c, `kind: port`), all host threads (capped at 32 like the
    reference's rayon fan-out over packed instances)."""
    import orc

    cores = min(os.cpu_count() or 1, 32)
    orc.prove(ops, wit, wz, wc, seeds, n_threads=cores)  # warm
    n, t0 = 0, time.perf_counter()
    while n < min_proofs or time.perf_counter() - t0 < min_seconds:
        rc, _ = orc.prove(ops, wit, wz, wc, seeds, n_threads=cores)
        assert rc == 0
        n += 1
    return n, time.perf_counter() - t0, cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    ops, wit, wz, wc, desc = make_workload(args.workload)
    n_and = int((ops["opcode"] == 6).sum())
    seeds = default_seeds()
    import orc

    cores = min(os.cpu_count() or 1, 32)
    B = args.batch
    for _ in range(max(args.warmup, 1)):
        orc.prove(ops, wit, wz, wc, seeds, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _b in range(B):
            rc, _ = orc.prove(ops, wit, wz, wc, seeds, n_threads=cores)
            assert rc == 0
    dt = time.perf_counter() - t0
    v = n_and * B * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": {"workload": desc, "batch": B, "parallelism": f"{cores} host threads over 32 packed instances, proofs of a batch one after the other"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps * B} whole proofs; C restatement of the reference's dataflow (oracle/c): the Rust reference cannot be built here (no cargo/rustc)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def set_metric(workload: str):
    """The headline metric is BASELINE.json's (GF(2) AND gates); a Z64 workload reports its multiplication gates instead."""
    global METRIC, UNIT
    if workload.startswith("z64"):
        METRIC, UNIT = "KKW prover MUL-gates/sec (Z64, 128-bit sec)", "MUL-gates/s"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sha256")
    ap.add_argument("--batch", type=int, default=32, help="independent proofs per step")
    ap.add_argument("--per-session", type=int, default=8,
                    help="proofs held side by side by one multi-proof session (small GF(2) circuits); the batch is batch/per-session such sessions")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shard", default="reps", choices=["reps", "proofs"],
                    help="N > 1: 'reps' shards the 32 packed instances of every proof over the ranks with the NCCL all-gather of repetition "
                         "hashes (BASELINE config 4, strong scaling); 'proofs' gives every rank its own whole proofs (no collective, weak scaling)")
    args = ap.parse_args()
    set_metric(args.workload)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist

    import reverie_b200 as rb
    from reverie_b200 import _native, sharding

    if not torch.cuda.is_available() or _native.lib().rv_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: reverie_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    _native.check(_native.lib().rv_set_device(local_rank))
    if world > 1:
        import datetime

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=120))
    if 32 % world:
        raise SystemExit("world size must divide the 32 packed instances")

    ops, wit, wz, wc, desc = make_workload(args.workload)
    seeds = default_seeds()
    circ = rb.Circuit(ops, wc)
    st = circ.stats()
    n_and = st["n_and"] + st["z64_mul"]  # multiplication gates of either domain
    st_proof_len = 0
    by_proofs = world > 1 and args.shard == "proofs"
    per = 32 if by_proofs else 32 // world
    B = max(1, args.batch)
    first = 0 if by_proofs else rank * per
    # Small GF(2) circuits: the batch is held by multi-proof sessions (P proofs side by side per session: every kernel launch
    # covers P proofs); other circuits: one proof per session.
    P = max(1, min(args.per_session, B))
    try:
        sessions = [rb.Session(circ, first, per, n_proofs=P)] if P > 1 else []
    except rb.ReverieError:
        P, sessions = 1, []
    B = (B + P - 1) // P * P
    sessions += [rb.Session(circ, first, per, n_proofs=P) for _ in range(B // P - len(sessions))]
    sess1 = sessions[0] if B == 1 else rb.Session(circ, first, per)  # one proof alone: latency and per-kernel times
    streams = [torch.cuda.ExternalStream(x.stream) for x in sessions]

    def upload_all(xs):
        for x in xs:
            for slot in range(x.n_proofs):
                x.upload(wit, wz, seeds, slot=slot)
    timing_stream = torch.cuda.Stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    recv_bufs = {}  # session -> (receive tensor over the session's own all-gather buffer, send tensor over its hashes)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The B sessions of a step are driven as one rv_batch: each phase (commit / open / prove) of all of them is ONE CUDA
    # graph launch on the leader's stream; the sessions' own streams fork from it and join back inside the graph.
    batches = {}

    def batch_of(sess_list):
        key = tuple(id(x) for x in sess_list)
        if key not in batches:
            bt = rb.Batch(sess_list)
            batches[key] = (bt, torch.cuda.ExternalStream(bt.stream))
        return batches[key]

    def step_device():
        """B proofs: commit + open with inputs resident in HBM; the all-gather of repetition hashes when sharded."""
        bt, lead = batch_of(sessions)
        if world == 1 or by_proofs:
            bt.prove()
            return
        bt.commit()
        # the one exchange of the protocol (src/proof/mod.rs:160-171): NCCL all-gather of the repetition hashes, device to
        # device from each session's hash buffer into its own receive buffer, ONE NCCL group launch for the B proofs in
        # flight, enqueued on the leader's stream between the two graphs -- no host round trip, no synchronisation
        for x in sessions:
            if x not in recv_bufs:
                recv_bufs[x] = (torch.as_tensor(x.all_hashes_device(), device="cuda"), torch.as_tensor(x.hashes_device(), device="cuda"))
        with torch.cuda.stream(lead):
            sharding.all_gather_hashes_batched([recv_bufs[x][0] for x in sessions], [recv_bufs[x][1] for x in sessions])
        bt.open()

    def timed_device(k: int):
        tot = 0.0
        _, lead = batch_of(sessions)
        for _ in range(k):
            with torch.cuda.stream(timing_stream):
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a.record(timing_stream)
            lead.wait_event(a)
            step_device()
            e = torch.cuda.Event()
            e.record(lead)
            timing_stream.wait_event(e)
            b.record(timing_stream)
            barrier()
            tot += a.elapsed_time(b)
        return tot  # ms

    upload_all(sessions + ([sess1] if sess1 is not sessions[0] else []))
    st_proof_len = len(torch.as_tensor(sess1.proof_device(), device="cuda"))
    for _ in range(args.warmup):
        step_device()
    for x in sessions:
        x.sync()
    sess = sess1
    launches0 = sum(x.launch_count for x in sessions)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total = timed_device(args.steps)
    launches = sum(x.launch_count for x in sessions) - launches0
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    n_jobs = B * (world if by_proofs else 1)  # proofs completed per step by the whole job
    value = n_and * n_jobs * args.steps / (ms_total * 1e-3)

    # single-proof device latency (B = 1), same rules
    lat_ms = None
    if world == 1:
        keep_s = sessions
        sessions = [sess1]
        for _ in range(3):  # eager run, graph capture, first replay
            step_device()
        lat_ms = timed_device(max(5, min(args.steps, 20))) / max(5, min(args.steps, 20))
        sessions = keep_s

    # ---- per-kernel device times -> roofline of the dominant kernel (rank 0) ----
    roofline, kernels = None, None
    peak, peak_src = load_peaks()
    reps = max(5, min(args.steps, 20))
    if rank == 0:
        sess.timing(True)
    keep_s = sessions
    sessions = [sess1]
    for _ in range(reps):  # every rank runs the steps (they contain the all-gather); only rank 0 brackets its kernels with events
        step_device()
    sessions = keep_s
    sess.sync()
    if rank == 0:
        kt = sess.kernel_times()
        sess.timing(False)
        comm, part = sess.fetch()
        kernels = {k["name"]: {"us_per_step": k["ms"] * 1e3 / reps, "launches_per_step": k["launches"] // reps,
                               "algorithmic_bytes_per_step": k["algorithmic_bytes"] // reps} for k in kt}
        # the dominant kernel of the critical stream; the plaintext value planes run on a side stream, overlapped with mask generation
        main_stream = [k for k in kt if k["name"] not in ("values", "z.values")] or kt
        top = max(main_stream, key=lambda k: k["ms"])
        per_launch_s = top["ms"] * 1e-3 / max(top["launches"], 1)
        bytes_per_launch = top["algorithmic_bytes"] / max(top["launches"], 1)
        ach = bytes_per_launch / per_launch_s / 1e9 if per_launch_s > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": top["name"], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": ncu_traffic(args.workload, top["name"]),
                    "peak_source": peak_src, "us_per_launch": per_launch_s * 1e6,
                    "per_kernel_frac": {k["name"]: (k["algorithmic_bytes"] / max(k["ms"], 1e-9) / 1e6) / peak for k in kt},
                    "path": {"algorithmic_bytes_per_step": st["algorithmic_bytes"], "achieved": B * st["algorithmic_bytes"] / (ms_total / args.steps * 1e-3) / 1e9,
                             "frac": B * st["algorithmic_bytes"] / (ms_total / args.steps * 1e-3) / 1e9 / peak,
                             "note": "SURVEY.md 8(d) bytes of the B proofs of a step / device time per step"}}

    big = st["n_masks"] * 256 + st["z64_masks"] * 16384 > (8 << 30)  # a session of this circuit holds tens of GB: one at a time
    if big and world == 1:
        del sess, sess1
        batches.clear()
        recv_bufs.clear()
        sessions.clear()
        streams.clear()
        import gc

        gc.collect()

    # ---- end to end through the public API (host buffers, copies inside the timed region) ----
    e2e = None
    verify = None
    single_latency_ms = None
    if world == 1:
        from concurrent.futures import ThreadPoolExecutor

        pool = ThreadPoolExecutor(max_workers=B)

        def one(_):
            return rb.Proof.new(circ, wit, wz, seeds=seeds)

        def step_api():  # the B queued requests of a step through the public batched call (host witnesses in, proof bytes out)
            return rb.Proof.new_batch(circ, [wit] * B, [wz] * B, seeds=[seeds] * B)

        for _ in range(args.warmup):
            proofs = step_api()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            proofs = step_api()
        dt = time.perf_counter() - t0
        e2e_v = n_and * B * args.steps / dt
        proof = proofs[0]
        for _ in range(3):  # session creation, eager run, graph capture
            one(0)
        t0 = time.perf_counter()
        for _ in range(10):
            one(0)
        single_latency_ms = (time.perf_counter() - t0) / 10 * 1e3
        d2h = B * (len(proof) + 36)
        # Proof::verify through the same API ("also reported"); bounded to circuits whose verifier tables
        # (kappa leaves + u-plane of 40 repetitions) fit next to a prover session
        if st["n_ops"] <= (32 << 20):
            def vfy(_):
                return proof.verify(circ)

            assert all(pool.map(vfy, range(B)))
            nv = max(2, args.steps // 4)
            t0 = time.perf_counter()
            for _ in range(nv):
                oks = list(pool.map(vfy, range(B)))
            dtv = time.perf_counter() - t0
            t0 = time.perf_counter()
            for _ in range(5):
                vfy(0)
            verify = {"value": n_and * B * nv / dtv, "unit": UNIT, "accepted": bool(all(oks)), "single_proof_ms": (time.perf_counter() - t0) / 5 * 1e3,
                      "note": "Proof.verify end to end (proof bytes in host memory), B verifications in flight"}
    else:
        def step_e2e():
            upload_all(sessions)
            step_device()
            if by_proofs:
                outs = [x.fetch(b) for x in sessions for b in range(x.n_proofs)]
                return outs, [p for _, p in outs]  # every rank holds its own whole proofs
            proofs = sharding.reduce_proofs(sessions)  # one NCCL reduce: rank 0 ends up with the B assembled proofs in host memory
            return [(None, proofs[0] if proofs else b"")], proofs
        for _ in range(args.warmup):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            outs, proofs = step_e2e()
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_v = n_and * n_jobs * args.steps / float(t.item())
        d2h = B * (st_proof_len + 36)
    e2e = {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": B * (st["n_inputs"] + 8 * st["z64_inputs"] + per * 8 * 16 + (256 * 32 if world > 1 else 0)), "d2h_bytes_per_step": d2h}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        c_ops, c_wit, c_wz, c_wc, c_n, c_note = ops, wit, wz, wc, n_and, "whole proofs of the same workload"
        if n_and > 4 * 10**6:  # bounded sample: the same circuit family at a size the CPU finishes in seconds
            small = args.workload.rstrip("0123456789") + str(2 * 10**6 if args.workload.startswith("z64") is False else 2 * 10**5)
            c_ops, c_wit, c_wz, c_wc, _ = make_workload(small)
            c_n = int((c_ops["opcode"] == 6).sum())
            c_note = f"whole proofs of the same circuit family at {c_n} multiplication gates ({small})"
        n, secs, cores = cpu_port_run(c_ops, c_wit, c_wz, c_wc, seeds, 10.0, 1)
        cpu = {"value": c_n * n / secs, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} {c_note} in {secs:.1f} s; C restatement of the reference's dataflow (oracle/c), threads over the 32 packed instances"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak" if by_proofs else "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": desc, "batch": B, "parallelism": (f"whole proofs per GPU, {B} proofs in flight per GPU per step, no collective" if by_proofs else
                                                                      f"{per} packed instances (= {per * 8} repetitions) per GPU, {B} proofs in flight per step, NCCL all-gather of the repetition hashes")
                                      + f"; {B // P} sessions x {P} proofs side by side",
                       "l2": "256 MiB memset between timed steps (outside the per-step CUDA-event pair)",
                       "timing": "one CUDA-event pair per step on a timing stream that forks to / joins the batch leader's stream (the B session streams fork from / join it inside the CUDA graph), summed over K steps"},
            "clocks": clocks, "e2e": e2e, "verify": verify, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "kernels": kernels, "single_proof_latency_ms": {"device": lat_ms, "e2e": single_latency_ms}, "circuit": {k: st[k] for k in ("n_and", "n_ops", "value_depth", "linear_depth", "z64_mul", "z64_value_depth")},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
