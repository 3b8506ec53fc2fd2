"""In-tree build of libmoldyn_b200.so (sm_100a only). `python -m moldyn_b200.build` or __graft_entry__.build()."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "moldyn_b200.cu")
DEPS = [SRC, os.path.join(ROOT, "include", "moldyn_b200.h")] + [
    os.path.join(HERE, "csrc", f) for f in ("md_kernels.cuh", "md_common.cuh", "md_cells.cuh", "md_lists.cuh", "md_reduce.cuh",
                                            "md_force.cuh", "md_integrate.cuh", "md_loop.cuh", "md_tile.cuh", "md_dist_kernels.cuh", "md_dist.inc")]
LIB = os.path.join(HERE, "lib", "libmoldyn_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: moldyn_b200 needs the CUDA toolkit to build its sm_100a kernels")


def up_to_date() -> bool:
    return os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in DEPS)


def build_library(force: bool = False, verbose: bool = False, out: str | None = None, extra: tuple = ()) -> str:
    """`out` / `extra`: a variant build of the same sources (measurement aids such as -DMD_LOOP_TRACE) next to the library."""
    if out is None and not force and up_to_date():
        return LIB
    out = out or LIB
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [nvcc(), *NVCC_FLAGS, *os.environ.get("MD_NVCC_EXTRA", "").split(), *extra, "-o", out, SRC, "-ldl"]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
        print(" ".join(cmd), file=sys.stderr)
    env = dict(os.environ)
    env.pop("CC", None)  # the image's $CC wrapper is not a usable host compiler for nvcc
    env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
    return out


HOST_SRC = [os.path.join(HERE, "host", "moldyn.cpp"), os.path.join(HERE, "host", "moldyn_cli.cpp")]
HOST_DEPS = HOST_SRC + [os.path.join(HERE, "host", "moldyn.hpp"), os.path.join(ROOT, "include", "moldyn_b200.h")]
CLI = os.path.join(HERE, "lib", "moldyn_cli")


def build_cli(force: bool = False) -> str:
    """C++ host mirror + the moldyn_cli-compatible driver, linked against libmoldyn_b200.so (rpath $ORIGIN)."""
    build_library()
    if not force and os.path.exists(CLI) and all(os.path.getmtime(CLI) >= os.path.getmtime(d) for d in HOST_DEPS + [LIB]):
        return CLI
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [gxx, "-O2", "-std=c++17", "-Wall", "-Wextra", "-o", CLI, *HOST_SRC, "-L", os.path.dirname(LIB),
           "-lmoldyn_b200", "-Wl,-rpath,$ORIGIN", "-pthread"]
    subprocess.check_call(cmd)
    return CLI


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_cli(force="--force" in sys.argv))
