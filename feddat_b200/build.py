"""In-tree build of ``libfeddat_sm100.so`` (the C-ABI library behind every kernel in this repo).

``nvcc`` cross-compiles sm_100a without a GPU, so this runs in the CPU container as the "does it
build" check and the resulting ``.so`` travels to the GPU box with the repo snapshot.  Objects are
rebuilt only when a source or header is newer than the object.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
OBJ_DIR = PKG_DIR / "build"
LIB_PATH = LIB_DIR / "libfeddat_sm100.so"
# the debug twin: the same sources with -DFEDDAT_DEBUG (trace hooks, probes, debug switches); tests / scripts only
DBG_LIB_PATH = LIB_DIR / "libfeddat_sm100_dbg.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the FedDAT kernels are CUDA-only and have no fallback")


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, debug: bool = False) -> Path:
    """Builds (if stale) the product library, or with ``debug`` its -DFEDDAT_DEBUG twin.  Concurrent callers
    (torchrun ranks that all find the library missing) serialise on a lock file, and the link step writes
    to a temporary name that is renamed into place, so no process can dlopen a half-written library."""
    import fcntl
    LIB_DIR.mkdir(exist_ok=True)
    with open(LIB_DIR / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose, debug)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool, debug: bool) -> Path:
    nvcc = _nvcc()
    obj_dir = OBJ_DIR / "dbg" if debug else OBJ_DIR
    lib_path = DBG_LIB_PATH if debug else LIB_PATH
    flags = NVCC_FLAGS + (["-DFEDDAT_DEBUG"] if debug else [])
    obj_dir.mkdir(parents=True, exist_ok=True)
    sources = sorted(CSRC.glob("*.cu"))
    headers = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + \
        sorted((PKG_DIR.parent / "include").glob("*.h"))
    include = ["-I", str(CSRC), "-I", str(PKG_DIR.parent / "include")]

    def compile_one(src: Path):
        obj = obj_dir / (src.stem + ".o")
        if not force and not _stale(obj, [src] + headers):
            return obj, ""
        cmd = [nvcc, *flags, *include, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        (obj_dir / (src.stem + ".ptxas.txt")).write_text(r.stderr)
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        results = list(ex.map(compile_one, sources))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log, file=sys.stderr)
    if force or _stale(lib_path, objs):
        tmp = lib_path.with_suffix(f".tmp{os.getpid()}.so")
        cmd = [nvcc, "-shared", "-o", str(tmp), *map(str, objs), "-lcudart_static", "-ldl",
               "-lpthread", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, lib_path)
    return lib_path


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
    print(build_library(force="--force" in sys.argv, verbose=False, debug=True))
