"""CPU: the C-ABI shared library builds, loads, and exports exactly the symbols include/moldiff_b200.h declares;
the ctypes mirrors of its structs have the C layout.  No compute call is made (there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "moldiff_b200.h")


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as g
    return g.build()


def declared_functions():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(mdb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = declared_functions()
    assert {"mdb_net_forward", "mdb_moldiff_forward", "mdb_bondpred_forward", "mdb_bondpred_backward",
            "mdb_workspace_bytes", "mdb_tc_selftest", "mdb_profile_begin", "mdb_profile_end"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_exported_symbols_are_plain_c(lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    mdb = {s for s in exported if s.startswith("mdb_")}
    assert mdb == set(declared_functions())          # nothing undeclared leaks out, nothing is missing


def test_struct_layouts_match_the_header(tmp_path, lib_path):
    from moldiff_b200 import engine, packing
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "moldiff_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %d %d %d %d\\n",'
                   'sizeof(mdb_net_desc), offsetof(mdb_net_desc, head_off), offsetof(mdb_net_desc, block_off),'
                   'offsetof(mdb_net_desc, tc_blob), sizeof(mdb_plan), MDB_NUM_BLOCK_SLOTS, MDB_NUM_HEAD_SLOTS,'
                   'MDB_NUM_TC_SLOTS, MDB_NUM_KERNEL_CLASSES);return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    vals = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    assert vals[0] == ctypes.sizeof(engine.NetDesc)
    assert vals[1] == engine.NetDesc.head_off.offset and vals[2] == engine.NetDesc.block_off.offset
    assert vals[3] == engine.NetDesc.tc_blob.offset
    assert vals[4] == ctypes.sizeof(engine.Plan)
    assert vals[5] == len(packing.BLOCK_SLOTS) and vals[6] == len(packing.HEAD_SLOTS) and vals[7] == len(packing.TC_SLOTS)
    lib = engine.load_library()
    assert lib.mdb_num_kernel_classes() == vals[8]
    assert lib.mdb_version() >= 1


def test_workspace_query_without_gpu(lib_path):
    from moldiff_b200 import engine
    lib = engine.load_library()
    fwd = lib.mdb_workspace_bytes(6286, 157102, 0, 6)
    bwd = lib.mdb_workspace_bytes(6286, 157102, 1, 8)
    assert 0 < fwd < bwd < 2 ** 32


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of silently computing somewhere else."""
    import torch
    from moldiff_b200 import MolDiff
    from moldiff_b200.config import builtin_config
    from moldiff_b200.engine import MoldiffB200Error
    from tests.helpers import batch_inputs, doubled
    torch.manual_seed(0)
    m = MolDiff(builtin_config("train/train_MolDiff.yml").model, 8, 6).eval()
    inp = batch_inputs(B=1)
    ei, be, he = doubled(inp)
    with pytest.raises(MoldiffB200Error):
        m(inp["h_node"], inp["pos"], inp["batch_node"], he, ei, be, inp["t"])


def test_product_never_imports_the_oracle():
    for base in ("moldiff_b200", "models"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith(".py"):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dirpath, f)
