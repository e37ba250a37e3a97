"""reverie_b200 -- B200-native (sm_100a CUDA) KKW MPC-in-the-head prover/verifier core, results-compatible with
trailofbits/reverie 0.3.2.  See DESIGN.md.  The compute path is libreverie_b200.so (reverie_b200/csrc); this package
is the thin host-side mirror of the reference's Proof API plus the circuit front end."""
from .circuits import OP_DTYPE  # noqa: F401
from .proof import Batch, Circuit, Group, Proof, Session, assemble  # noqa: F401
from ._native import FormatError, ReverieError, WitnessError  # noqa: F401

__all__ = ["Batch", "Circuit", "Group", "Proof", "Session", "assemble", "OP_DTYPE", "ReverieError", "WitnessError", "FormatError"]
