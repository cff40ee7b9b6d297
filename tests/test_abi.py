"""The drop-in boundary without a GPU: the library builds for sm_100a, loads, exports every function include/eskf.h
declares, the ctypes structures have the sizes of the C structs, and the product refuses to run without CUDA
(no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "eskf.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(eskf_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from dvi_ekf_b200 import build

    return C.CDLL(build.build_cuda())


def test_every_declared_entry_point_is_exported(lib):
    names = _declared_functions()
    assert len(names) >= 18 and "eskf_run" in names and "eskf_prepass" in names and "eskf_noise_dump" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    from dvi_ekf_b200 import _lib

    assert set(_lib.EXPORTS) == set(names)  # the Python binding covers the whole ABI


def test_ctypes_structures_match_the_c_structs(tmp_path):
    """sizeof of every struct of the header, compiled with gcc, against the ctypes mirrors"""
    from dvi_ekf_b200 import _lib

    prog = tmp_path / "sizes.c"
    prog.write_text('#include <stdio.h>\n#include "eskf.h"\nint main(void){printf("%zu %zu %zu %zu\\n", sizeof(eskf_model_t), '
                    "sizeof(eskf_streams_t), sizeof(eskf_prepass_in_t), sizeof(eskf_prepass_out_t));return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(_lib.EskfModel), C.sizeof(_lib.EskfStreams), C.sizeof(_lib.EskfPrepassIn), C.sizeof(_lib.EskfPrepassOut)]


def test_no_cpu_fallback():
    """without a CUDA device the engine fails loudly instead of computing on the host"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dvi_ekf_b200 import BatchFilter
    from dvi_ekf_b200._lib import EskfError

    with pytest.raises(EskfError):
        BatchFilter(4, scope_length=50.0, cam_angle_rad=0.5, frozen_dofs=(1, 1, 1, 1, 1, 1))


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under dvi_ekf_b200/ may import it"""
    pkg = os.path.join(ROOT, "dvi_ekf_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
