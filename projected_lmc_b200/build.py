"""In-tree build of libplmc_b200.so (sm_100a only).

``python -m projected_lmc_b200.build`` compiles every ``csrc/*.cu`` with
``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` and links one shared
library next to this file, so that it travels with the repo snapshot to the GPU
box.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
INCLUDE = PKG_DIR.parent / "include"
LIB_PATH = PKG_DIR / "libplmc_b200.so"
OBJ_DIR = PKG_DIR / "build"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-I", str(CSRC), "-I", str(INCLUDE),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _stamp() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp_file = OBJ_DIR / "stamp"
    return LIB_PATH.exists() and stamp_file.exists() and stamp_file.read_text() == _stamp()


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and is_current():
        return LIB_PATH
    OBJ_DIR.mkdir(exist_ok=True)
    nvcc = _nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: Path) -> Path:
        obj = OBJ_DIR / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            sys.stderr.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    # --cudart shared: the library uses the CUDA runtime torch already has in the process (and carries no copy
    # of the runtime's symbol table)
    cmd = [nvcc, "-shared", "--cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-o", str(LIB_PATH), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    (OBJ_DIR / "stamp").write_text(_stamp())
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
