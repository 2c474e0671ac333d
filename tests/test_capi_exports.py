"""CPU: the C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol that
include/lgbm_b200.h declares.  No compute calls here (no GPU in this container); on a box without a
usable CUDA driver the compute entry points must fail loudly, never fall back to a CPU path."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_exports():
    src = open(os.path.join(ROOT, "include", "lgbm_b200.h")).read()
    return sorted(set(re.findall(r"LGBMB200_EXPORT\s+[\w\s\*]+?\b(LGBMB200_\w+)\s*\(", src)))


def test_header_and_loader_agree(built_lib):
    from lightgbm_b200 import _lib
    assert declared_exports() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in declared_exports():
        assert hasattr(lib, name), name


def test_library_is_sm100a_native(built_lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_cpu_fallback_without_gpu(built_lib):
    import lightgbm_b200 as lgb
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    L = lgb.B200TreeLearner(lgb.Config(num_leaves=4))
    with pytest.raises(RuntimeError, match="CUDA"):
        L.init(lgb.Layout.identity(np.zeros((64, 4), np.uint8)))


def test_product_never_imports_the_oracle():
    """The oracle is the checker: nothing shipped under lightgbm_b200/, include/ or integration/ may import,
    include, link or execute anything under oracle/ (integration/Makefile only reuses the compiled REFERENCE
    objects that oracle/Makefile.ref produced — the reference itself, not the oracle restatement)."""
    bad = re.compile(r"(^\s*(import|from)\s+oracle\b)|(lgbm_oracle)|(oracle_py)|(liblgbm_oracle)", re.M)
    for top in ("lightgbm_b200", "include", "integration"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            if "_build" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    txt = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert not bad.search(txt), f"{os.path.join(dirpath, f)} references the oracle"


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """The Python mirror structs (tree_learner.py) must have the size and field offsets of include/lgbm_b200.h as a C
    compiler lays them out — a drift here corrupts every call silently."""
    import shutil
    import subprocess
    from lightgbm_b200 import tree_learner as tl
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    structs = {"LGBMB200_Config": tl._CConfig, "LGBMB200_Tree": tl._CTree, "LGBMB200_Split": tl._CSplit, "LGBMB200_Layout": tl._CLayout}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "lgbm_b200.h")}"', "int main(void) {"]
    for cname, py in structs.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in py._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call([gcc, "-std=c11", "-o", str(exe), str(src)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    for line in out:
        parts = line.split()
        py = structs[parts[0]]
        assert ctypes.sizeof(py) == int(parts[1]), parts[0]
        assert [getattr(py, f).offset for f, _ in py._fields_] == [int(x) for x in parts[2:]], parts[0]
