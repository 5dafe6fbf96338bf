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
    """sizeof AND the offset of every member of the four ctypes mirrors equal what the C compiler lays out
    (osc2_topology, osc2_step_inputs, osc2_model, osc2_sweep)."""
    import subprocess
    import tempfile

    from opensc2_b200._abi import Osc2Model, Osc2StepInputs, Osc2Sweep, Osc2Topology

    mirrors = (("osc2_topology", Osc2Topology), ("osc2_step_inputs", Osc2StepInputs), ("osc2_model", Osc2Model),
               ("osc2_sweep", Osc2Sweep))
    lines = []
    for cname, cls in mirrors:
        lines.append(f'printf("%zu\\n", sizeof({cname}));')
        for fname, _t in cls._fields_:
            lines.append(f'printf("%zu\\n", offsetof({cname}, {fname}));')
    src = ('#include <stdio.h>\n#include <stddef.h>\n#include "opensc2_b200.h"\nint main(void){' + "".join(lines) + "return 0;}\n")
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        got = list(map(int, subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()))
    want = []
    for _cname, cls in mirrors:
        want.append(C.sizeof(cls))
        want += [getattr(cls, fname).offset for fname, _t in cls._fields_]
    assert got == want


def test_version_and_error_text():
    from opensc2_b200 import _lib

    L = _lib.load()
    assert L.osc2_version() >= 100
    assert isinstance(L.osc2_last_error(), bytes)
