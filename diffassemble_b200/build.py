"""In-tree nvcc build of the C-ABI library (sm_100a only).

``python -m diffassemble_b200.build`` compiles ``csrc/*.cu`` into
``diffassemble_b200/lib/libdiffassemble_b200.so``.  The ``.so`` is git-ignored but
travels with the working tree (it must sit in-tree so the GPU box loads it).
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "build"
LIBNAME = "libdiffassemble_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; the CUDA library cannot be built")


def _stamp(src: Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(src.read_bytes())
    for hdr in sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "diffassemble_b200.h"]:
        h.update(hdr.read_bytes())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    nvcc = _nvcc()
    LIBDIR.mkdir(exist_ok=True)
    OBJDIR.mkdir(exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    objs, jobs = [], []
    for src in sources:
        obj = OBJDIR / (src.stem + ".o")
        stamp_file = OBJDIR / (src.stem + ".stamp")
        stamp = _stamp(src)
        objs.append(obj)
        if not force and obj.exists() and stamp_file.exists() and stamp_file.read_text() == stamp:
            continue
        jobs.append((src, obj, stamp_file, stamp))

    def compile_one(job):
        src, obj, stamp_file, stamp = job
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = OBJDIR / (src.stem + ".log")
        log.write_text(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{res.stdout}\n{res.stderr}")
        stamp_file.write_text(stamp)
        if verbose:
            print(res.stderr, file=sys.stderr)
        return src.name

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    lib = LIBDIR / LIBNAME
    if jobs or not lib.exists():
        cmd = [nvcc, "-shared", "-o", str(lib), *map(str, objs), "-lcudart", "-lcuda"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
