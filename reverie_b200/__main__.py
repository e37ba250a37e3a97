"""Command line mirror of the reference's binary (src/main.rs:167-273): same operations and flag names.

    python -m reverie_b200 --operation prove      --program-path P --witness-path W --proof-path F
    python -m reverie_b200 --operation verify     --program-path P --proof-path F
    python -m reverie_b200 --operation oneshot    --program-path P --witness-path W     (cleartext evaluation)
    python -m reverie_b200 --operation oneshot-zk --program-path P --witness-path W     (prove, then verify)
    python -m reverie_b200 --operation version_info

Program files: the reference reads a bincode `Vec<mcircuit::CombineOperation>` (src/main.rs:66) whose serde variant order
cannot be derived from the reference tree, so this front end reads (a) Bristol-Fashion text
(`ngates nwires` header; XOR/AND/INV/EQW/EQ gates; `--assert-outputs BITS` pins the outputs with AddConst + AssertZero) and
(b) `.rvops` files: the packed 24-byte `rv_op` records of include/reverie_b200.h, written by `circuits.save_ops`.
Witness files are ASCII: every '0' / '1' is one GF(2) witness bit, all other bytes are skipped (src/witness.rs:17-34).
Proof files are `bincode::serialize(&Proof)` (src/main.rs:84,103).  Proving and verifying run on the GPU through the C ABI."""
from __future__ import annotations

import argparse
import sys

import numpy as np

from . import circuits as C


def load_program(path: str, assert_outputs: str | None):
    with open(path, "rb") as f:
        head = f.read(8)
    if head == C.RVOPS_MAGIC:
        return C.load_ops(path)
    with open(path, "r") as f:
        text = f.read()
    expected = [int(ch) for ch in assert_outputs if ch in "01"] if assert_outputs is not None else None
    ops, n_wires, _ = C.parse_bristol_fashion(text, expected)
    return ops, (0, n_wires)


def load_witness(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        raw = np.frombuffer(f.read(), dtype=np.uint8)
    return (raw[(raw == ord("0")) | (raw == ord("1"))] - ord("0")).astype(np.uint8)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="reverie_b200", description="B200-native KKW prover/verifier (reverie-compatible CLI)")
    ap.add_argument("--operation", required=True, choices=["prove", "verify", "oneshot", "oneshot-zk", "version_info"])
    ap.add_argument("--witness-path")
    ap.add_argument("--program-path")
    ap.add_argument("--proof-path")
    ap.add_argument("--reference-verify", action="store_true",
                    help="verify: the reference's exact verdict (commitment equality only); default also requires the opened repetitions' AssertZero checks to hold")
    ap.add_argument("--assert-outputs", help="Bristol programs: expected output bits (output-wire order); appends AddConst + AssertZero")
    a = ap.parse_args(argv)
    if a.operation == "version_info":
        from . import _native

        print("reverie_version:", _native.lib().rv_version().decode())
        return 0
    need = {"prove": ("program_path", "witness_path", "proof_path"), "verify": ("program_path", "proof_path"),
            "oneshot": ("program_path", "witness_path"), "oneshot-zk": ("program_path", "witness_path")}[a.operation]
    for n in need:
        if getattr(a, n) is None:
            ap.error(f"--{n.replace('_', '-')} is required for --operation {a.operation}")
    ops, wc = load_program(a.program_path, a.assert_outputs)
    if a.operation == "oneshot":
        print("Evaluating program in cleartext")
        _, ok = C.evaluate_gf2(ops, load_witness(a.witness_path), wc[1])
        if not ok:
            print("Invalid proof: an AssertZero failed", file=sys.stderr)
            return 255
        print("()")
        return 0
    from . import Circuit, Proof, ReverieError

    circ = Circuit(ops, wc)
    try:
        if a.operation in ("prove", "oneshot-zk"):
            print("Evaluating program in ~zero knowledge~")
            proof = Proof.new(circ, load_witness(a.witness_path), ())
            if a.operation == "prove":
                with open(a.proof_path, "wb") as f:
                    f.write(proof.serialize())
                print("Ok(())")
                return 0
        else:
            with open(a.proof_path, "rb") as f:
                proof = Proof.deserialize(f.read())
            print("Verifying Proof")
        if proof.verify(circ, strict=not a.reference_verify):
            print("Ok(())")
            return 0
        print('Err("Unverifiable Proof")')
        return 255
    except ReverieError as e:  # the reference panics here (invalid witness / malformed proof)
        print(f"Invalid proof: {e}", file=sys.stderr)
        return 255


if __name__ == "__main__":
    sys.exit(main())
