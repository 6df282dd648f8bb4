"""CPU suite: the C-ABI library loads and exports every symbol include/bourse_b200.h declares.
No compute calls — there is no GPU here and the library has no CPU fallback."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from bourse_b200 import abi, build

    build.build_lib()
    return abi.load()


def test_header_symbols_exported(lib):
    from bourse_b200 import abi

    hdr = open(os.path.join(ROOT, "include", "bourse_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(bb_\w+)\s*\(", hdr, flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(abi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bb_abi_version() == 1


def test_struct_layouts_match_header():
    from bourse_b200 import abi

    assert C.sizeof(abi.Config) == 104
    assert abi.INSTR_DTYPE.itemsize == 32 and abi.GROUP_DTYPE.itemsize == 80
    assert C.sizeof(abi.Stats) == 64
    assert abi.INSTR_DTYPE.fields["op_flags"][1] == 8 and abi.INSTR_DTYPE.fields["price"][1] == 16
    assert abi.GROUP_DTYPE.fields["decay"][1] == 32


def test_fails_loudly_without_gpu(lib):
    """bb_create must refuse to run without a CUDA device instead of falling back to the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from bourse_b200 import core

    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        core.OrderBook(0, 1)


def test_oracle_is_not_imported_by_the_product():
    import subprocess
    import sys

    code = "import sys, bourse_b200, bourse_b200.core, bourse_b200.workloads; print(any(m.startswith('oracle') for m in sys.modules))"
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.stdout.strip() == "False", out.stdout + out.stderr
    for dirpath, _, files in os.walk(os.path.join(ROOT, "bourse_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src.replace("oracle/_ref", ""), f
