"""The C-ABI library loads and exports every entry point include/opensc2_b200.h
declares (no compute calls: this runs without a GPU)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "opensc2_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(osc2_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from opensc2_b200 import _lib

    L = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.EXPORTED)


def test_ctypes_structs_match_header_layout():
    """sizeof of the ctypes mirrors equals what the C compiler lays out."""
    import subprocess
    import tempfile

    from opensc2_b200._abi import Osc2StepInputs, Osc2Topology

    src = ('#include <stdio.h>\n#include "opensc2_b200.h"\n'
           'int main(void){printf("%zu %zu\\n", sizeof(osc2_topology), sizeof(osc2_step_inputs));return 0;}\n')
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        a, b = map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split())
    assert a == C.sizeof(Osc2Topology)
    assert b == C.sizeof(Osc2StepInputs)


def test_version_and_error_text():
    from opensc2_b200 import _lib

    L = _lib.load()
    assert L.osc2_version() >= 100
    assert isinstance(L.osc2_last_error(), bytes)
