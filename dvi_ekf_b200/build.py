"""Builds the in-tree native libraries.

``build_cuda()`` cross-compiles libeskf_b200.so for sm_100a with nvcc (no GPU
needed); the per-CTA-shape instantiations of the persistent kernel are separate
translation units compiled in parallel.  ``build_hostcheck()`` compiles the CPU
harness around the host/device math header used by the not-gpu tests.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libeskf_b200.so")
SHAPES = (4, 28)  # filters per CTA of eskf_kernel (v1, test / A-B variant only); keep in sync with eskf_api.cu
SHAPES3 = (4, 8, 16, 28)  # filters per CTA of eskf_kernel3 (v3)
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd, log):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        raise RuntimeError(f"build step failed: {' '.join(cmd)}\n{p.stdout}")
    return p.stdout


def build_cuda(force: bool = False, verbose: bool = False, defines=(), suffix: str = "") -> str:
    """``defines`` / ``suffix`` build an experiment variant (e.g. -DESKF_EXP_NO_COV) next to the product
    library as libeskf_b200<suffix>.so; select it at run time with ESKF_B200_LIB=<path>."""
    global OBJ, LIB
    if suffix:
        OBJ = os.path.join(HERE, "build" + suffix)
        LIB = os.path.join(HERE, f"libeskf_b200{suffix}.so")
    extra = [f"-D{d}" for d in defines] + os.environ.get("ESKF_B200_NVCC_EXTRA", "").split()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, h) for h in ("eskf_math.cuh", "eskf_rng.cuh", "eskf_kernel.cuh", "eskf_cov3.cuh", "eskf_kernel3.cuh")]
    headers.append(os.path.join(ROOT, "include", "eskf.h"))
    jobs = []
    for f in SHAPES:
        obj = os.path.join(OBJ, f"eskf_launch_f{f}.o")
        src = os.path.join(CSRC, "eskf_launch.cu")
        jobs.append((obj, [src] + headers, [_nvcc(), *NVCC_FLAGS, *extra, f"-DESKF_F={f}", "-c", src, "-o", obj]))
    for f in SHAPES3:
        obj = os.path.join(OBJ, f"eskf_launch3_f{f}.o")
        src = os.path.join(CSRC, "eskf_launch3.cu")
        jobs.append((obj, [src] + headers, [_nvcc(), *NVCC_FLAGS, *extra, f"-DESKF_F={f}", "-c", src, "-o", obj]))
    pp_obj = os.path.join(OBJ, "eskf_prepass.o")
    pp_src = os.path.join(CSRC, "eskf_prepass.cu")
    jobs.append((pp_obj, [pp_src] + headers, [_nvcc(), *NVCC_FLAGS, *extra, "-c", pp_src, "-o", pp_obj]))
    api_obj = os.path.join(OBJ, "eskf_api.o")
    api_src = os.path.join(CSRC, "eskf_api.cu")
    jobs.append((api_obj, [api_src] + headers, [_nvcc(), *NVCC_FLAGS, *extra, "-c", api_src, "-o", api_obj]))
    todo = [(o, c) for (o, srcs, c) in jobs if force or not _newer(o, srcs)]
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        outs = list(ex.map(lambda oc: _run(oc[1], oc[0] + ".log"), todo))
    if verbose:
        for o in outs:
            sys.stdout.write(o)
    objs = [o for (o, _, _) in jobs]
    if force or todo or not _newer(LIB, objs):
        _run([_nvcc(), "-Wno-deprecated-gpu-targets", "-shared", "-o", LIB, *objs, "-lcudart", "-ldl"], os.path.join(OBJ, "link.log"))
    return LIB


def build_hostcheck(force: bool = False) -> str:
    src = os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp")
    lib = os.path.join(ROOT, "tests", "hostcheck", "libhostcheck.so")
    deps = [src, os.path.join(CSRC, "eskf_math.cuh"), os.path.join(CSRC, "eskf_rng.cuh"), os.path.join(CSRC, "eskf_cov3.cuh")]
    if force or not _newer(lib, deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", lib, src], check=True)
    return lib


def ptxas_summary() -> str:
    """registers / spills / shared memory per kernel, from the saved nvcc logs."""
    lines = []
    for tag, shapes in (("", SHAPES), ("3", SHAPES3)):
        for f in shapes:
            log = os.path.join(OBJ, f"eskf_launch{tag}_f{f}.o.log")
            if os.path.exists(log):
                txt = open(log).read().splitlines()
                for i, l in enumerate(txt):
                    if "eskf_kernel" in l and "Compiling entry" in l:
                        lines.append(f"v{tag or 1} F={f}: " + " | ".join(x.strip() for x in txt[i + 2 : i + 4]))
    return "\n".join(lines)


if __name__ == "__main__":
    for a in sys.argv[1:]:
        if a.startswith("--exp="):  # e.g. --exp=ESKF_EXP_NO_COV
            d = a.split("=", 1)[1].split(",")  # several switches: --exp=A,B -> suffix _a_b
            print(build_cuda(defines=tuple(d), suffix="_" + "_".join(x.lower().replace("eskf_exp_", "") for x in d)))
            sys.exit(0)
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_hostcheck())
    print(ptxas_summary())
