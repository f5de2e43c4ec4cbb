"""CPU checks of the drop-in boundary: the library builds, loads, exports every symbol the header declares,
and the ctypes mirror of ``pdm_stats_args`` matches the C layout.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pdm_b200.h")


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build()
    from pdm_b200 import _cabi
    return _cabi.load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pdm_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from pdm_b200 import _cabi
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pdm_b200.h but not exported"
        assert n in _cabi.SIGNATURES, f"{n} has no ctypes signature in pdm_b200/_cabi.py"
    assert lib.pdm_abi_version() == 4


def test_stats_args_layout_matches_c(lib):
    from pdm_b200._cabi import StatsArgs
    fields = [f[0] for f in StatsArgs._fields_]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "pdm_b200.h"\nint main(){\n'
    prog += 'printf("%zu\\n", sizeof(pdm_stats_args));\n'
    for f in fields:
        prog += f'printf("%zu\\n", offsetof(pdm_stats_args, {f}));\n'
    prog += "return 0;}\n"
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "l.c"), os.path.join(td, "l")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert out[0] == C.sizeof(StatsArgs)
    for f, off in zip(fields, out[1:]):
        assert getattr(StatsArgs, f).offset == off, f


def test_no_cpu_fallback(lib):
    """Without a GPU the compute path must fail loudly, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pdm_b200 import PdmError
    from pdm_b200.engine import default_backend
    with pytest.raises(PdmError):
        default_backend()
    sm = C.c_int()
    assert lib.pdm_device_info(0, C.byref(sm), None, None) != 0
    assert lib.pdm_last_error()
    import utils
    with pytest.raises(PdmError):
        utils.compute_pw_dist_sqr(torch.zeros(2, 3))
