"""Builds tests/cpp/test_proof_api.cpp (the C++ host mirror's tests) against the in-tree libraries.  Test infrastructure."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_cpp_api_test() -> str:
    import orc
    from reverie_b200 import _native

    _native.lib()  # builds libreverie_b200.so if needed
    orc.build()
    out = os.path.join(ROOT, "tests", "cpp", "test_proof_api")
    src = os.path.join(ROOT, "tests", "cpp", "test_proof_api.cpp")
    libdir, orcdir = os.path.join(ROOT, "reverie_b200", "_lib"), os.path.join(ROOT, "oracle", "c")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", out, src, "-L" + libdir, "-lreverie_b200", "-L" + orcdir, "-lorc",
                           "-Wl,-rpath," + libdir, "-Wl,-rpath," + orcdir])
    return out
