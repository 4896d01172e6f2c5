"""Builds the C-ABI library `librs_b200.so` (include/rs_b200.h) with nvcc for sm_100a.

The library is built IN-TREE (recommendsystem_b200/_lib/librs_b200.so) so that it
travels to the GPU box with the repo snapshot.  There is exactly one target
architecture (sm_100a) and no fallback: if nvcc is missing the build fails.

Objects are cached under build/obj keyed by a hash of (source, headers, flags),
so an unchanged translation unit is not recompiled.
"""
from __future__ import annotations

import hashlib
import os
import re
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "_lib"
LIB_PATH = LIB_DIR / "librs_b200.so"
OBJ_DIR = ROOT / "build" / "obj"

# (D, U, H) instantiations of the fused InteractingLayer kernels.  Keep in sync
# with RS_INTERACT_SHAPES in csrc/interacting.cu.
INTERACT_SHAPES = [(16, 16, 1), (16, 16, 2), (16, 16, 4), (8, 8, 1), (8, 8, 2), (16, 8, 2)]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
] + os.environ.get("RS_NVCC_DEFS", "").split()      # developer switches, e.g. -DRS_ITB_PROFILE


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the sm_100a library cannot be built (no fallback exists)")


def _units():
    """(object name, source file, extra -D flags)"""
    units = []
    for src in sorted(CSRC.glob("*.cu")):
        if src.name == "interacting_inst.cu":
            continue
        units.append((src.stem, src, []))
    inst = CSRC / "interacting_inst.cu"
    for d, u, h in INTERACT_SHAPES:
        units.append((f"interacting_inst_{d}_{u}_{h}", inst, [f"-DRS_D={d}", f"-DRS_U={u}", f"-DRS_H={h}"]))
    return units


_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _deps_digest(src: Path) -> bytes:
    """Hash of the source plus every quoted header it (transitively) includes."""
    seen, todo = {}, [src]
    while todo:
        p = todo.pop().resolve()
        if p in seen or not p.exists():
            continue
        text = p.read_text()
        seen[p] = text
        for inc in _INC.findall(text):
            todo.append(p.parent / inc)
    h = hashlib.sha256()
    for p in sorted(seen):
        h.update(p.name.encode())
        h.update(seen[p].encode())
    return h.digest()


def _compile(nvcc: str, name: str, src: Path, defs, verbose: bool) -> Path:
    h = hashlib.sha256()
    h.update(_deps_digest(src))
    h.update(" ".join(NVCC_FLAGS + list(defs)).encode())
    obj = OBJ_DIR / f"{name}.{h.hexdigest()[:16]}.o"
    if obj.exists():
        return obj
    for stale in OBJ_DIR.glob(f"{name}.*.o"):
        stale.unlink()
    cmd = [nvcc, *NVCC_FLAGS, *defs, "-I", str(ROOT / "include"), "-c", str(src), "-o", str(obj)]
    if verbose:
        print("[build]", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name} {defs}:\n{r.stdout}\n{r.stderr}")
    return obj


def build_library(force: bool = False, verbose: bool = False, jobs: int | None = None) -> Path:
    """Compile every CUDA translation unit for sm_100a and link librs_b200.so."""
    nvcc = _nvcc()
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    if force:
        for o in OBJ_DIR.glob("*.o"):
            o.unlink()
    units = _units()
    jobs = jobs or min(len(units), os.cpu_count() or 4)
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        objs = list(ex.map(lambda u: _compile(nvcc, u[0], u[1], u[2], verbose), units))
    stamp = hashlib.sha256(" ".join(sorted(o.name for o in objs)).encode()).hexdigest()
    stamp_file = LIB_DIR / "librs_b200.stamp"
    if LIB_PATH.exists() and stamp_file.exists() and stamp_file.read_text() == stamp and not force:
        return LIB_PATH
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH),
           *map(str, objs)]
    if verbose:
        print("[build]", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp_file.write_text(stamp)
    return LIB_PATH


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose=True)
    print(p)
