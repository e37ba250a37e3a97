#!/usr/bin/env python
"""bench.py -- KKW prover AND-gates/sec (GF(2), 8 players x 256 repetitions, 40 opened) on B200, next to the CPU port.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line on rank 0.
For N > 1 the driver launches it under torchrun, one rank per GPU; the 32 packed instances (src/proof/mod.rs:127-157) are
sharded 32/N per rank and the only exchange is the all-gather of the 256 x 32-byte repetition hashes
(src/proof/mod.rs:160-171) -> strong scaling (the proof is the same for every N).  By default that exchange and the
assembly of the proof run over NVLink peer memory inside the kernels (rv_session_peer_link: one CUDA graph launch per rank
and step, no collective library on the data path); `--exchange nccl` keeps the NCCL all-gather + reduce for comparison.

A step = one batch of B independent `Proof::new` calls on the workload circuit (default: SHA-256 compression, SURVEY.md
8(d) config 2; B = --batch, default 32) issued together, the way a proving service sees its queue; B = 1 gives the
single-proof latency, which is also reported.  Small GF(2) circuits are held by multi-proof sessions (--per-session proofs
side by side, every kernel launch covering all of them); big / Z64 circuits one proof per session (--batch 1).
  value    device-resident: witnesses + seeds already in HBM; the step is one CUDA graph launch for all sessions
           (rv_batch) on a leader stream that carries the CUDA-event pair of the step
  e2e      host buffers in, proof bytes out: N = 1 through the public batched call (Proof.new_batch -> rv_prove_batch);
           N > 1: every rank uploads witnesses + its seeds, runs the sharded step, rank 0 fetches the assembled proofs
  parity   after the timed region the proofs of the step are hashed (SHA-256 of the bincode bytes) and compared with the digest
           committed in tests/golden/bench_digests.json (computed by the CPU oracle: tests/golden/make_bench_digests.py)
  roofline the kernels of the timed (batched) configuration, CUDA events around every launch: see `roofline` in the line
  extra_workloads (default run only): BASELINE.json configs 3 and 5 -- z64mul1000000 (N = 1), flat100000000 and
           layered100000000 sharded over the N ranks -- each with value, e2e, path fraction and parity
The CPU arm (--impl reference) proves the same B proofs per step with the C restatement of the reference's dataflow.
"""
from __future__ import annotations

import argparse
import gc
import hashlib
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
EXTRAS = ("z64mul1000000", "flat100000000", "layered100000000")
STREAM_EXTRA = "flat300000000"  # 3 x 10^8 ANDs: resident proving would need ~330 GB of device memory; proved in streaming mode on one GPU


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


def golden_digest(workload: str):
    """SHA-256 of the oracle's proof bytes for `workload` with default_seeds() (tests/golden/bench_digests.json), or None."""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "bench_digests.json")) as f:
            return json.load(f)["digests"].get(workload)
    except Exception:
        return None


# bench kernel label -> kernel(s) in the committed ncu capture (profiles/r2_traffic.json; DRAM bytes per launch, pipe utilisation)
NCU_NAMES = {"linear": ["k_mask_vm"], "mask_gen": ["k_mask_gen_tt<0>", "k_mask_gen_tt<1>", "k_mask_gen_tt"], "items": ["k_items"], "chunk_cv": ["k_chunk_cv"],
             "values": ["k_values<1>"], "rep_hash": ["k_rep_hash"], "extract": ["k_extract"], "challenge": ["k_challenge"], "key_setup": ["k_key_setup"],
             "z.mask_gen": ["k_zmask_gen_tt"], "z.items": ["k_zitems_online"], "z.values": ["k_zvalues"], "z.extract": ["k_zextract"]}


def ncu_record(workload: str, label: str):
    """The committed `ncu --set full` figures of the kernel behind `label` for this workload's timed configuration
    (profiles/r2_traffic.json, falling back to round 1's capture), or None."""
    for fn in ("r2_traffic.json", "r1c_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                w = json.load(f)["workloads"].get(workload)
            for k in NCU_NAMES.get(label, []):
                if w and k in w:
                    return dict(w[k], capture=fn)
        except Exception:
            pass
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


def unit_of(workload: str):
    """The headline metric is BASELINE.json's (GF(2) AND gates); a Z64 workload reports its multiplication gates instead."""
    if workload.startswith("z64"):
        return "KKW prover MUL-gates/sec (Z64, 128-bit sec)", "MUL-gates/s"
    return METRIC, UNIT


class Env:
    """What every workload of one bench process shares: ranks, torch handles, peaks."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        from reverie_b200 import _native

        self.args, self.torch, self.dist = args, torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available() or _native.lib().rv_device_count() < 1:
            raise SystemExit("bench.py needs a CUDA device: reverie_b200 has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        _native.check(_native.lib().rv_set_device(self.local_rank))
        if self.world > 1:
            import datetime

            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank), timeout=datetime.timedelta(seconds=900))
        if 32 % self.world:
            raise SystemExit("world size must divide the 32 packed instances")
        self.peak, self.peak_src = load_peaks()
        self.timing_stream = torch.cuda.Stream()
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, ok: bool) -> bool:
        if self.world == 1:
            return ok
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())


def run_workload(env: Env, name: str, batch: int, per_session: int, steps: int, warmup: int, full: bool) -> dict:
    """One workload on the env's ranks.  full = the headline workload (adds latency, per-kernel roofline, verify, one-shot and
    the CPU baseline); otherwise the compact record of an extra workload."""
    import reverie_b200 as rb
    from reverie_b200 import sharding

    torch, dist, args = env.torch, env.dist, env.args
    rank, world = env.rank, env.world
    metric, unit = unit_of(name)
    t0 = time.perf_counter()
    ops, wit, wz, wc, desc = make_workload(name)
    gen_s = time.perf_counter() - t0
    seeds = default_seeds()
    t0 = time.perf_counter()
    n_ops = int(ops.size)
    # proving needs no verifier tables; the headline workload also times Proof.verify, big extras only prove
    circ = rb.Circuit(ops, wc, prove_only=not full)
    compile_s = time.perf_counter() - t0
    st = circ.stats()
    n_and = st["n_and"] + st["z64_mul"]  # multiplication gates of either domain
    by_proofs = world > 1 and args.shard == "proofs"
    linked = world > 1 and not by_proofs and args.exchange == "p2p"
    per = 32 if by_proofs else 32 // world
    first = 0 if by_proofs else rank * per
    B = max(1, batch)
    # Small GF(2) circuits: the batch is held by multi-proof sessions (P proofs side by side per session: every kernel launch
    # covers P proofs); other circuits: one proof per session.
    P = max(1, min(per_session, B))
    grp = None
    if linked:
        # the rank's part of a multi-GPU prover behind one handle (rv_group): B / P linked sessions of P proofs, launched as one graph
        try:
            grp = rb.Group.rank(circ, rank, world, n_sessions=(B + P - 1) // P, slots=P)
        except rb.ReverieError:
            P = 1
            grp = rb.Group.rank(circ, rank, world, n_sessions=B, slots=1)
        B = (B + P - 1) // P * P
        grp.link_distributed()
        sessions = list(grp.sessions)
    else:
        try:
            sessions = [rb.Session(circ, first, per, n_proofs=P)] if P > 1 else []
        except rb.ReverieError:
            P, sessions = 1, []
        B = (B + P - 1) // P * P
        sessions += [rb.Session(circ, first, per, n_proofs=P) for _ in range(B // P - len(sessions))]
    recv_bufs = {}  # NCCL exchange: session -> (receive tensor over the session's own all-gather buffer, send tensor over its hashes)
    batches = {}

    def upload_all(xs):
        for x in xs:
            for slot in range(x.n_proofs):
                x.upload(wit, wz, seeds, slot=slot)

    # N = 1 (and the NCCL / whole-proof modes): the sessions of a step are driven as one rv_batch -- each phase of all of them is
    # ONE CUDA graph launch on the leader's stream; the sessions' own streams fork from it and join back inside the graph.
    # Linked groups (N > 1): every session is its own CUDA graph launch on its own stream (rv_group: uploads of the next session
    # overlap the work of the previous one), so a step is B / P launches per rank.
    ext_streams = {}

    def batch_of(sess_list):
        key = tuple(id(x) for x in sess_list)
        if key not in batches:
            bt = rb.Batch(sess_list)
            batches[key] = (bt, torch.cuda.ExternalStream(bt.stream))
        return batches[key]

    def streams_of(sess_list):
        """The streams a step of sess_list runs on (what the step's CUDA-event pair has to fork to and join from)."""
        if grp is None:
            return [batch_of(sess_list)[1]]
        for x in sess_list:
            if x not in ext_streams:
                ext_streams[x] = torch.cuda.ExternalStream(x.stream)
        return [ext_streams[x] for x in sess_list]

    def step_device(sess_list):
        """The proofs held by sess_list: commit + exchange + open with inputs resident in HBM."""
        if grp is not None:
            for x in sess_list:
                x.prove()  # linked shards: the exchange of repetition hashes happens inside the challenge kernel, over peer memory
            return
        bt, lead = batch_of(sess_list)
        if world == 1 or by_proofs:
            bt.prove()
            return
        bt.commit()
        # --exchange nccl: all-gather of the repetition hashes device to device, ONE NCCL group launch for the proofs in flight,
        # enqueued on the leader's stream between the two graphs
        for x in sess_list:
            if x not in recv_bufs:
                recv_bufs[x] = (torch.as_tensor(x.all_hashes_device(), device="cuda"), torch.as_tensor(x.hashes_device(), device="cuda"))
        with torch.cuda.stream(lead):
            sharding.all_gather_hashes_batched([recv_bufs[x][0] for x in sess_list], [recv_bufs[x][1] for x in sess_list])
        bt.open()

    done_ms = []  # diagnostic: when each stream of a step finished, relative to the step's start (mean over the timed steps)

    def timed_device(sess_list, k: int) -> float:
        tot = 0.0
        sts = streams_of(sess_list)
        acc = [0.0] * len(sts)
        for _ in range(k):
            with torch.cuda.stream(env.timing_stream):
                env.flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            env.barrier()
            a.record(env.timing_stream)
            for st_ in sts:
                st_.wait_event(a)
            step_device(sess_list)
            ends = []
            for st_ in sts:
                e = torch.cuda.Event(enable_timing=True)
                e.record(st_)
                env.timing_stream.wait_event(e)
                ends.append(e)
            b.record(env.timing_stream)
            env.barrier()
            tot += a.elapsed_time(b)
            for i, e in enumerate(ends):
                acc[i] += a.elapsed_time(e)
        done_ms[:] = [x / max(k, 1) for x in acc]
        return tot  # ms

    def collect(sess_list):
        """Proof bytes of every slot on the assembling rank (rank 0; every rank with --shard proofs)."""
        if world == 1 or by_proofs or linked:
            return [x.fetch(b)[1] for x in sess_list for b in range(x.n_proofs)]
        out = sharding.reduce_proofs(sess_list)  # --exchange nccl: one NCCL sum-reduce of the shards' buffers
        return out if out is not None else [b""] * sum(x.n_proofs for x in sess_list)

    upload_all(sessions)
    for _ in range(warmup):
        step_device(sessions)
    for x in sessions:
        x.sync()
    launches0 = sum(x.launch_count for x in sessions)
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    ms_total = timed_device(sessions, steps)
    stream_done_ms = list(done_ms)
    launches = sum(x.launch_count for x in sessions) - launches0
    clocks = sampler.stop()
    ms_total = env.max_over_ranks(ms_total)
    n_jobs = B * (world if by_proofs else 1)  # proofs completed per step by the whole job
    value = n_and * n_jobs * steps / (ms_total * 1e-3)
    ms_per_step = ms_total / steps

    # ---- parity: the proofs the timed steps left behind vs the oracle's digest (committed; no oracle code runs here) ----
    proofs = collect(sessions)
    want = golden_digest(name)
    proof_len = len(proofs[0]) if proofs else 0
    digest, parity = None, None
    if rank == 0 or by_proofs:
        digests = {hashlib.sha256(memoryview(p)).hexdigest() for p in proofs}
        digest = sorted(digests)[0]
        parity = (len(digests) == 1 and digest == want) if want else None
    if want:
        parity_all = env.all_true(parity is not False)
        if not parity_all:
            raise SystemExit(f"PARITY FAILURE: workload {name}: proof digest {digest} != oracle digest {want} (rank {rank}, {world} GPUs)")
    del proofs

    # ---- per-kernel device times of the TIMED configuration: each session of the step run alone with CUDA events around every
    #      launch (a launch covers the P proofs of its session) -> which kernel dominates the step, its 8(d) bytes / its time ----
    peak = env.peak
    path_bytes = B * st["algorithmic_bytes"]
    path = {"algorithmic_bytes_per_step": path_bytes, "achieved": path_bytes / (ms_per_step * 1e-3) / 1e9, "peak": peak * world,
            "frac": path_bytes / (ms_per_step * 1e-3) / 1e9 / (peak * world),
            "note": f"SURVEY.md 8(d) bytes of the {B} proofs of a step / device time per step / ({world} x HBM peak)"}
    roofline, kernels = None, None
    if (full or world == 1) and (world == 1 or by_proofs or linked):
        reps = 3 if not full else max(5, min(steps, 10))
        if rank == 0:
            for x in sessions:
                x.timing(True)
        for _ in range(reps):  # timing keeps the phases eager; every rank runs them (linked sessions meet on the device)
            for x in sessions:
                x.prove()
                torch.cuda.synchronize()
        kt = {}
        if rank == 0:
            for x in sessions:
                for k in x.kernel_times():
                    a = kt.setdefault(k["name"], {"ms": 0.0, "launches": 0, "bytes": 0})
                    a["ms"] += k["ms"]
                    a["launches"] += k["launches"]
                    a["bytes"] += k["algorithmic_bytes"]
                x.timing(False)
            kernels = {n: {"us_per_step": a["ms"] * 1e3 / reps, "launches_per_step": a["launches"] // reps, "us_per_launch": a["ms"] * 1e3 / max(a["launches"], 1),
                           "algorithmic_bytes_per_launch": a["bytes"] // max(a["launches"], 1),
                           "hbm_frac": (a["bytes"] / max(a["ms"], 1e-9) / 1e6) / peak} for n, a in kt.items()}
            side = ("values", "z.values")  # plaintext value planes: a side stream, overlapped with mask generation
            main = {n: a for n, a in kt.items() if n not in side} or kt
            top = max(main, key=lambda n: main[n]["ms"])
            a = kt[top]
            us = a["ms"] * 1e3 / max(a["launches"], 1)
            bpl = a["bytes"] / max(a["launches"], 1)
            ach = bpl / (us * 1e-6) / 1e9 if us > 0 else 0.0
            rec = ncu_record(name, top) or {}
            main_sum_us = sum(v["ms"] for v in main.values()) * 1e3 / reps
            roofline = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": rec.get("dram_bytes_per_launch"),
                        "peak_source": env.peak_src, "us_per_launch": us, "algorithmic_bytes_per_launch": bpl, "proofs_per_launch": P,
                        "launches_per_step": a["launches"] // reps, "share_of_step": a["ms"] * 1e3 / reps / max(main_sum_us, 1e-9),
                        "main_stream_kernel_us_per_step": main_sum_us,
                        "how": "CUDA events around every launch of the step's sessions, each session run alone (eager) after the timed region; "
                               "achieved = SURVEY.md 8(d) bytes this kernel moves per launch / its time per launch",
                        "secondary": {"bound": "alu/lds issue (AES T-table + BLAKE3 are integer work, DESIGN.md section 5)",
                                      "ncu": {k: rec.get(k) for k in ("alu_pipe_pct", "lsu_pipe_pct", "issue_active_pct", "sm_throughput_pct", "dram_throughput_pct", "capture") if k in rec}},
                        "path": path}
    if roofline is None:
        roofline = {"bound": "hbm", "path": path, "peak": peak, "unit": "GB/s", "frac": path["frac"], "achieved": path["achieved"]}

    # ---- single-proof latency on the device (one proof alone in flight), headline workload at N = 1 ----
    lat_ms = None
    if full and world == 1:
        sess1 = sessions[0] if (B == 1) else rb.Session(circ, first, per)
        if sess1 is not sessions[0]:
            sess1.upload(wit, wz, seeds)
        for _ in range(3):  # eager run, graph capture, first replay
            step_device([sess1])
        nl = max(5, min(steps, 20))
        lat_ms = timed_device([sess1], nl) / nl
        del sess1

    big = st["n_masks"] * 256 + st["z64_masks"] * 16384 > (8 << 30)  # a session of this circuit holds tens of GB: one at a time
    x = None  # (loop variable: would keep the last session, tens of GB, alive)
    if big and world == 1:
        batches.clear()
        recv_bufs.clear()
        ext_streams.clear()
        sessions.clear()
        gc.collect()

    # ---- end to end (host buffers in, proof bytes out, copies inside the timed region) ----
    verify, single_latency_ms, oneshot = None, None, None
    if world == 1:
        def step_api():  # the B queued requests of a step through the public call (host witnesses in, proof bytes out)
            if B == 1:
                return [rb.Proof.new(circ, wit, wz, seeds=seeds)]
            return rb.Proof.new_batch(circ, [wit] * B, [wz] * B, seeds=[seeds] * B)

        for _ in range(warmup):
            out = step_api()
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = step_api()
        dt = time.perf_counter() - t0
        e2e_v = n_and * B * steps / dt
        proof = out[0]
        if want and hashlib.sha256(memoryview(proof._buf if isinstance(proof._buf, np.ndarray) else proof.data)).hexdigest() != want:
            raise SystemExit(f"PARITY FAILURE: workload {name}: the public API's proof differs from the oracle digest")
        d2h = B * (len(proof) + 36)
        if full:
            for _ in range(3):  # session creation, eager run, graph capture
                rb.Proof.new(circ, wit, wz, seeds=seeds)
            t0 = time.perf_counter()
            for _ in range(10):
                rb.Proof.new(circ, wit, wz, seeds=seeds)
            single_latency_ms = (time.perf_counter() - t0) / 10 * 1e3
            # the reference's own call shape: Proof::new(circuit, ...) with the op list every time (src/proof/mod.rs:119-124) ->
            # rv_proof_new: first call compiles (prove-only tables), later calls find the circuit in the content-addressed cache
            from reverie_b200 import _native

            _native.lib().rv_circuit_cache_clear()
        if full and not big:  # (a second resident copy of a 10^8-gate circuit does not fit next to the first)
            t0 = time.perf_counter()
            p1 = rb.Proof.new(ops, wit, wz, wc, seeds=seeds)
            first_ms = (time.perf_counter() - t0) * 1e3
            for _ in range(2):
                rb.Proof.new(ops, wit, wz, wc, seeds=seeds)
            t0 = time.perf_counter()
            for _ in range(10):
                p1 = rb.Proof.new(ops, wit, wz, wc, seeds=seeds)
            cached_ms = (time.perf_counter() - t0) / 10 * 1e3
            oneshot = {"first_call_ms": first_ms, "cached_call_ms": cached_ms, "value_first_call": n_and / (first_ms * 1e-3), "value_cached": n_and / (cached_ms * 1e-3),
                       "unit": unit, "parity_checked": (hashlib.sha256(p1.data).hexdigest() == want) if want else None,
                       "note": "rv_proof_new(ops, ...): one proof, op list passed with every call like the reference's Proof::new; first call = circuit compile + session + proof"}
            # Proof::verify through the same API ("also reported")
            if st["n_ops"] <= (32 << 20):
                from concurrent.futures import ThreadPoolExecutor

                pool = ThreadPoolExecutor(max_workers=B)

                def vfy(_):
                    return proof.verify(circ)

                for _ in range(3):  # every pooled session: eager run, graph capture, first replay
                    assert all(pool.map(vfy, range(B)))
                nv = max(2, steps // 4)
                t0 = time.perf_counter()
                for _ in range(nv):
                    oks = list(pool.map(vfy, range(B)))
                dtv = time.perf_counter() - t0
                t0 = time.perf_counter()
                for _ in range(5):
                    vfy(0)
                verify = {"value": n_and * B * nv / dtv, "unit": unit, "accepted": bool(all(oks)), "single_proof_ms": (time.perf_counter() - t0) / 5 * 1e3,
                          "note": "Proof.verify end to end (proof bytes in host memory), B verifications in flight"}
        del out, proof
    else:
        def step_e2e():
            if grp is not None:  # one C call per rank: uploads, the step's graph launch, the fetch of the assembled proofs on rank 0
                return grp.prove_batch([wit] * B, [wz] * B, [seeds] * B)  # Proof objects over the library's buffers on rank 0, None elsewhere
            upload_all(sessions)
            step_device(sessions)
            return collect(sessions)  # rank 0 (every rank with --shard proofs) ends up with the proof bytes in host memory

        for _ in range(warmup):
            out = step_e2e()  # (held like in the timed loop: big proofs come back in pooled pinned buffers, two are in rotation)
        env.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = step_e2e()
        env.barrier()
        dt = env.max_over_ranks(time.perf_counter() - t0)
        e2e_v = n_and * n_jobs * steps / dt
        if want and (rank == 0 or by_proofs) and hashlib.sha256(memoryview(out[0]._buf if isinstance(out[0], rb.Proof) else out[0])).hexdigest() != want:
            raise SystemExit(f"PARITY FAILURE: workload {name}: the end-to-end proof differs from the oracle digest")
        if full and grp is not None and st["n_ops"] <= (32 << 20):
            # Proof::verify of a queue of world x B proofs spread over the GPUs (rv_group_verify_batch: whole proofs per GPU -- verification
            # has no exchange step); every rank verifies its share
            box = [out[0].data if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            queue = [rb.Proof(box[0])] * (B * world)
            for _ in range(3):
                verdicts = grp.verify_batch(queue)
            nv = max(2, steps // 4)
            env.barrier()
            t0 = time.perf_counter()
            for _ in range(nv):
                verdicts = grp.verify_batch(queue)
            env.barrier()
            dtv = env.max_over_ranks(time.perf_counter() - t0)
            verify = {"value": n_and * B * world * nv / dtv, "unit": unit, "accepted": env.all_true(all(v for v in verdicts if v is not None)),
                      "note": f"Proof.verify of {B * world} proofs per round, whole proofs per GPU ({B} per rank, 8 in flight), proof bytes in host memory"}
        del out
        d2h = B * (proof_len + 36)
    e2e = {"value": e2e_v, "unit": unit, "h2d_bytes_per_step": B * (st["n_inputs"] + 8 * st["z64_inputs"] + per * 8 * 16),
           "d2h_bytes_per_step": d2h, "ms_per_step": dt / steps * 1e3}

    cpu = None
    if full and rank == 0 and world == 1 and not args.no_cpu_baseline:
        c_ops, c_wit, c_wz, c_wc, c_n, c_note = ops, wit, wz, wc, n_and, "whole proofs of the same workload"
        if n_and > 4 * 10**6:  # bounded sample: the same circuit family at a size the CPU finishes in seconds
            small = name.rstrip("0123456789") + str(2 * 10**6 if name.startswith("z64") is False else 2 * 10**5)
            c_ops, c_wit, c_wz, c_wc, _ = make_workload(small)
            c_n = int((c_ops["opcode"] == 6).sum())
            c_note = f"whole proofs of the same circuit family at {c_n} multiplication gates ({small})"
        n, secs, cores = cpu_port_run(c_ops, c_wit, c_wz, c_wc, seeds, 10.0, 1)
        cpu = {"value": c_n * n / secs, "unit": unit, "cores": cores, "kind": "port",
               "sample": f"{n} {c_note} in {secs:.1f} s; C restatement of the reference's dataflow (oracle/c), threads over the 32 packed instances"}

    exchange = ("none (one GPU)" if world == 1 else "none (whole proofs per GPU)" if by_proofs else
                "device-side: repetition hashes pushed into every rank's receive buffer over NVLink peer memory inside the challenge kernel, openings written "
                "straight into rank 0's proof buffer (rv_session_peer_link, CUDA IPC between the ranks)" if linked else "NCCL all-gather of the repetition hashes + NCCL sum-reduce of the shard buffers")
    res = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
        "scaling": "weak" if by_proofs else "strong",
        "config": {"workload": desc, "batch": B,
                   "parallelism": (f"whole proofs per GPU, {B} proofs in flight per GPU per step, no collective" if by_proofs else
                                   f"{per} packed instances (= {per * 8} repetitions) per GPU, {B} proofs in flight per step") + f"; {B // P} sessions x {P} proofs side by side",
                   "exchange": exchange,
                   "l2": "256 MiB memset between timed steps (outside the per-step CUDA-event pair)",
                   "timing": "one CUDA-event pair per step on a timing stream that forks to / joins the batch leader's stream (the session streams fork from / join it inside the CUDA graph), summed over K steps (linked groups: the pair forks to / joins every session's stream)"},
        "clocks": clocks, "stream_done_ms": stream_done_ms, "e2e": e2e, "parity_checked": parity, "proof_sha256": digest, "proof_bytes": proof_len,
        "compile_s": compile_s, "circuit_gen_s": gen_s, "gpu_launches": int(launches), "roofline": roofline,
        "circuit": {k: st[k] for k in ("n_and", "n_ops", "value_depth", "linear_depth", "z64_mul", "z64_value_depth")},
    }
    if full:
        res.update({"verify": verify, "oneshot": oneshot, "cpu_baseline": cpu, "kernels": kernels,
                    "single_proof_latency_ms": {"device": lat_ms, "e2e": single_latency_ms}})
    batches.clear()
    recv_bufs.clear()
    ext_streams.clear()
    sessions.clear()
    grp = None
    del circ
    gc.collect()
    return res


def run_streaming(env: Env, name: str, window: int) -> dict:
    """rv_prove_streaming on one GPU: a circuit whose share tensor and transcripts exceed HBM, proved in segments (two passes).
    One timed call, end to end: segmentation, compilation of the segments, both passes, the proof bytes in host memory."""
    import reverie_b200 as rb

    metric, unit = unit_of(name)
    t0 = time.perf_counter()
    ops, wit, wz, wc, desc = make_workload(name)
    gen_s = time.perf_counter() - t0
    n_and = int((ops["opcode"] == 6).sum())
    seeds = default_seeds()
    small = make_workload(name.rstrip("0123456789") + "1000000")
    rb.Proof.new_streaming(small[0], small[1], small[3], seeds=seeds, window_ops=1 << 18)  # warm: kernels loaded, pools primed
    t0 = time.perf_counter()
    proof = rb.Proof.new_streaming(ops, wit, wc, seeds=seeds, window_ops=window)
    dt = time.perf_counter() - t0
    want = golden_digest(name)
    digest = hashlib.sha256(memoryview(proof._buf)).hexdigest()
    if want and digest != want:
        raise SystemExit(f"PARITY FAILURE: streaming proof of {name}: digest {digest} != oracle digest {want}")
    n_bytes, ends = len(proof), (bytes(proof._buf[:4096]), bytes(proof._buf[-4096:]))
    proof = None
    gc.collect()  # (the 3 GB pinned block goes back to the library's pool)
    t0 = time.perf_counter()
    proof = rb.Proof.new_streaming(ops, wit, wc, seeds=seeds, window_ops=window)  # the same call again: pinned pool and page cache warm
    dt2 = time.perf_counter() - t0
    if len(proof) != n_bytes or (bytes(proof._buf[:4096]), bytes(proof._buf[-4096:])) != ends:
        raise SystemExit(f"PARITY FAILURE: streaming proof of {name}: the second call returned different bytes")
    return {"metric": metric, "value": n_and / dt, "unit": unit, "seconds": dt, "seconds_second_call": dt2, "window_ops": window, "parity_checked": (digest == want) if want else None,
            "proof_sha256": digest, "proof_bytes": len(proof), "circuit_gen_s": gen_s, "resident_device_bytes_needed": int(n_and) * 1100,
            "config": {"workload": desc, "mode": "streaming (rv_prove_streaming): segments of window_ops ops, wires carried on the device, two passes (hashes, then openings); "
                                                 "the timed call includes segmentation and compilation of the segments"}}


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
    ap.add_argument("--extras", default="auto", help="comma list of extra workloads reported under extra_workloads ('none' = skip; "
                                                     "auto = BASELINE configs 3 and 5 when the headline workload is the default one)")
    ap.add_argument("--shard", default="reps", choices=["reps", "proofs"],
                    help="N > 1: 'reps' shards the 32 packed instances of every proof over the ranks (BASELINE config 4, strong scaling); "
                         "'proofs' gives every rank its own whole proofs (no exchange, weak scaling)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1, --shard reps: 'p2p' = device-side exchange over peer memory inside the kernels; 'nccl' = NCCL all-gather + reduce from the host")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args, int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")))

    env = Env(args)
    line = run_workload(env, args.workload, args.batch, args.per_session, args.steps, args.warmup, full=True)
    gc.collect()
    extras = {}
    names = EXTRAS if (args.extras == "auto" and args.workload == "sha256" and args.shard == "reps") else () if args.extras in ("auto", "none") else tuple(args.extras.split(","))
    for name in names:
        if name.startswith("z64") and env.world > 1:
            continue  # BASELINE config 3 is a 1-GPU configuration
        try:
            r = run_workload(env, name, 1, 1, 3, 3, full=False)
            extras[name] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "warmup", "e2e", "parity_checked", "proof_sha256", "proof_bytes", "compile_s",
                                              "circuit_gen_s", "gpu_launches", "roofline", "circuit", "config", "clocks")}
        except SystemExit:
            raise
        except Exception as ex:  # an extra must not take the headline line down with it
            extras[name] = {"error": f"{type(ex).__name__}: {ex}"}
            ex = None
        gc.collect()
    if names and env.world == 1 and args.extras == "auto":
        try:
            extras["streaming:" + STREAM_EXTRA] = run_streaming(env, STREAM_EXTRA, 1 << 22)
        except SystemExit:
            raise
        except Exception as ex:
            extras["streaming:" + STREAM_EXTRA] = {"error": f"{type(ex).__name__}: {ex}"}
        gc.collect()
    if env.rank == 0:
        out = {"metric": line["metric"], "value": line["value"], "unit": line["unit"], "n_gpus": env.world, "steps": line["steps"], "warmup": line["warmup"],
               "ms_per_step": line["ms_per_step"], "higher_is_better": True, "scaling": line["scaling"], "vs_baseline": None, "dtype": "u64", "data": "synthetic"}
        out.update({k: v for k, v in line.items() if k not in out})
        if extras:
            out["extra_workloads"] = extras
        print(json.dumps(out), flush=True)
    if env.world > 1:
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
