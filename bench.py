#!/usr/bin/env python
"""bench.py -- KKW prover AND-gates/sec (GF(2), 8 players x 256 repetitions, 40 opened) on B200, next to the CPU port.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line on rank 0.
For N > 1 the driver launches it under torchrun, one rank per GPU; the 32 packed instances (src/proof/mod.rs:127-157) are
sharded 32/N per rank and the only exchange is the all-gather of the 256 x 32-byte repetition hashes
(src/proof/mod.rs:160-171) -> strong scaling (the proof is the same for every N).

A step = one `Proof::new` of the workload circuit (default: SHA-256 compression).
  value  device-resident: witness + seeds already in HBM, commit + open on the session stream, CUDA events per step
  e2e    host buffers in, proof bytes out, through the public API (Proof.new -> rv_prove), copies inside the timed region
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

    if name == "sha256":
        ops, wit, wc = C.sha256_abc_case()
        return ops, wit, wc, "GF(2) SHA-256 compression circuit (generated Bristol-style: 22573 AND / 93666 XOR / 2147 INV, 768 inputs, 256 output asserts), 256 reps x 8 players, 1 proof per step"
    if name.startswith("flat"):
        n = int(name[4:])
        ops, wc = C.flat_mul_circuit(n)
        return ops, np.array([1, 1], dtype=np.uint8), wc, f"GF(2) flat circuit: 2 inputs + {n} x Mul(2,0,1) (src/proof/mod.rs:322-329 scaled), 1 proof per step"
    if name.startswith("layered"):
        n = int(name[7:])
        width = min(1 << 20, max(1024, n // 16))
        ops, nw = C.layered_and_circuit(width, n)
        wit = np.random.default_rng(0).integers(0, 2, size=width).astype(np.uint8)
        return ops, wit, (0, nw), f"GF(2) layered circuit: {width} inputs + {n} ANDs in layers of {width}, 1 proof per step"
    raise SystemExit(f"unknown workload {name}")


def default_seeds() -> bytes:
    import reverie_oracle as R

    return b"".join(R.default_seeds())


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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
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


def cpu_port_run(ops, wit, wc, seeds, min_seconds: float, min_proofs: int):
    """The CPU restatement of the reference's dataflow (oracle/c, `kind: port`), all host threads (capped at 32 like the
    reference's rayon fan-out over packed instances)."""
    import orc

    cores = min(os.cpu_count() or 1, 32)
    orc.prove(ops, wit, [], wc, seeds, n_threads=cores)  # warm
    n, t0 = 0, time.perf_counter()
    while n < min_proofs or time.perf_counter() - t0 < min_seconds:
        rc, _ = orc.prove(ops, wit, [], wc, seeds, n_threads=cores)
        assert rc == 0
        n += 1
    return n, time.perf_counter() - t0, cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    ops, wit, wc, desc = make_workload(args.workload)
    n_and = int((ops["opcode"] == 6).sum())
    seeds = default_seeds()
    import orc

    cores = min(os.cpu_count() or 1, 32)
    for _ in range(max(args.warmup, 1)):
        orc.prove(ops, wit, [], wc, seeds, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc, _ = orc.prove(ops, wit, [], wc, seeds, n_threads=cores)
        assert rc == 0
    dt = time.perf_counter() - t0
    v = n_and * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": {"workload": desc, "parallelism": f"{cores} host threads over 32 packed instances"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} whole proofs; C restatement of the reference's dataflow (oracle/c): the Rust reference cannot be built here (no cargo/rustc)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sha256")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist

    import reverie_b200 as rb
    from reverie_b200 import _native

    if not torch.cuda.is_available() or _native.lib().rv_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: reverie_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    _native.check(_native.lib().rv_set_device(local_rank))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if 32 % world:
        raise SystemExit("world size must divide the 32 packed instances")

    ops, wit, wc, desc = make_workload(args.workload)
    seeds = default_seeds()
    circ = rb.Circuit(ops, wc)
    st = circ.stats()
    n_and = st["n_and"]
    per = 32 // world
    sess = rb.Session(circ, rank * per, per)
    stream = torch.cuda.ExternalStream(sess.stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    gathered = torch.empty(256 * 32, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        """commit + open with inputs resident in HBM; the all-gather of repetition hashes when sharded."""
        sess.commit()
        if world > 1:
            mine = torch.frombuffer(bytearray(sess.hashes()), dtype=torch.uint8).cuda()
            dist.all_gather_into_tensor(gathered, mine)
            torch.cuda.synchronize()
            sess.open(gathered.data_ptr())
        else:
            sess.open()

    def timed_device(k: int):
        tot = 0.0
        for _ in range(k):
            with torch.cuda.stream(stream):
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            a.record(stream)
            step_device()
            b.record(stream)
            barrier()
            tot += a.elapsed_time(b)
        return tot  # ms

    sess.upload(wit, (), seeds)
    for _ in range(args.warmup):
        step_device()
    sess.sync()
    launches0 = sess.launch_count
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms_total = timed_device(args.steps)
    launches = sess.launch_count - launches0
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    comm, part = sess.fetch()
    value = n_and * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public API (host buffers, copies inside the timed region) ----
    e2e = None
    single_latency_ms = None
    if world == 1:
        for _ in range(args.warmup):
            proof = rb.Proof.new(circ, wit, (), seeds=seeds)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            proof = rb.Proof.new(circ, wit, (), seeds=seeds)
        dt = time.perf_counter() - t0
        e2e_v = n_and * args.steps / dt
        single_latency_ms = dt / args.steps * 1e3
        d2h = len(proof) + 36
    else:
        def step_e2e():
            sess.upload(wit, (), seeds)
            step_device()
            return sess.fetch()
        for _ in range(args.warmup):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            c_, p_ = step_e2e()
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_v = n_and * args.steps / float(t.item())
        d2h = len(p_) + 36 + per * 8 * 32
    e2e = {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": st["n_inputs"] + per * 8 * 16 + (256 * 32 if world > 1 else 0), "d2h_bytes_per_step": d2h}

    # ---- per-kernel device times -> roofline of the dominant kernel (rank 0) ----
    roofline, kernels = None, None
    peak, peak_src = load_peaks()
    if rank == 0:
        sess.timing(True)
        reps = max(5, min(args.steps, 20))
        for _ in range(reps):
            step_device()
        kt = sess.kernel_times()
        sess.timing(False)
        kernels = {k["name"]: {"us_per_step": k["ms"] * 1e3 / reps, "launches_per_step": k["launches"] // reps,
                               "algorithmic_bytes_per_step": k["algorithmic_bytes"] // reps} for k in kt}
        main_stream = [k for k in kt if k["name"] != "values"]
        top = max(kt, key=lambda k: k["ms"])
        per_launch_s = top["ms"] * 1e-3 / max(top["launches"], 1)
        bytes_per_launch = top["algorithmic_bytes"] / max(top["launches"], 1)
        ach = bytes_per_launch / per_launch_s / 1e9 if per_launch_s > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": top["name"], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "peak_source": peak_src, "us_per_launch": per_launch_s * 1e6,
                    "path": {"algorithmic_bytes_per_step": st["algorithmic_bytes"], "achieved": st["algorithmic_bytes"] / (ms_total / args.steps * 1e-3) / 1e9,
                             "frac": st["algorithmic_bytes"] / (ms_total / args.steps * 1e-3) / 1e9 / peak,
                             "note": "SURVEY.md 8(d) bytes of the whole proof / device time per step"}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n, secs, cores = cpu_port_run(ops, wit, wc, seeds, 10.0, 3)
        cpu = {"value": n_and * n / secs, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} whole proofs of the same workload in {secs:.1f} s; C restatement of the reference's dataflow (oracle/c), threads over the 32 packed instances"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": desc, "parallelism": f"{per} packed instances (= {per * 8} repetitions) per GPU",
                       "l2": "256 MiB memset between timed steps (outside the per-step CUDA-event pair)",
                       "timing": "CUDA events on the library's stream, one pair per step, summed over K steps"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "kernels": kernels, "single_proof_latency_ms": single_latency_ms, "circuit": {k: st[k] for k in ("n_and", "n_ops", "value_depth", "linear_depth")},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
