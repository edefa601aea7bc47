"""In-tree build of libwaveflow_b200.so (nvcc, sm_100a only).  `python -m waveflow_b200.build [--force] [-v]`.

Objects are compiled in parallel (the fused live-path kernels are fully unrolled and take minutes each) and only
re-compiled when their source or any header changed.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = CSRC / "_obj"
LIB = HERE / "libwaveflow_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--expt-relaxed-constexpr",
         "--extended-lambda", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-Xcudafe", "--diag_suppress=177"]


def sources():
    return sorted(CSRC.glob("*.cu"))


def _headers_mtime() -> float:
    hs = list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "waveflow_b200.h"]
    return max(p.stat().st_mtime for p in hs)


def _stale(src: Path, obj: Path, hm: float) -> bool:
    return (not obj.exists()) or obj.stat().st_mtime < max(src.stat().st_mtime, hm)


def needs_build() -> bool:
    if not LIB.exists():
        return True
    hm = _headers_mtime()
    return any(_stale(s, OBJ / (s.stem + ".o"), hm) for s in sources()) or any(
        (OBJ / (s.stem + ".o")).stat().st_mtime > LIB.stat().st_mtime for s in sources())


def _compile(src: Path, obj: Path):
    r = subprocess.run([NVCC, *FLAGS, "-c", str(src), "-o", str(obj)], capture_output=True, text=True)
    return src, r


def build(force: bool = False, verbose: bool = False, jobs: int | None = None) -> Path:
    """Compile + link under an exclusive file lock: when N ranks import the package at the same time (torchrun) exactly one
    of them builds, the others wait on the lock and then find everything up to date.  The library is linked to a temporary
    name and moved into place with os.replace, so no process can ever load a half-written file."""
    import fcntl
    OBJ.mkdir(exist_ok=True)
    with open(OBJ / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose, jobs)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool, jobs: int | None) -> Path:
    hm = _headers_mtime()
    todo = [(s, OBJ / (s.stem + ".o")) for s in sources() if force or _stale(s, OBJ / (s.stem + ".o"), hm)]
    if not todo and LIB.exists() and not needs_build():
        return LIB
    jobs = jobs or min(len(todo) or 1, os.cpu_count() or 4)
    logs = []
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        for src, r in ex.map(lambda a: _compile(*a), todo):
            logs.append(f"== {src.name}\n{r.stderr}")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stderr}\n{r.stdout}")
    objs = [str(OBJ / (s.stem + ".o")) for s in sources()]
    tmp = LIB.with_name(LIB.name + f".tmp{os.getpid()}")
    r = subprocess.run([NVCC, "-shared", "-o", str(tmp), *objs, "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        tmp.unlink(missing_ok=True)
        raise RuntimeError(f"link failed:\n{r.stderr}")
    os.replace(tmp, LIB)
    with open(CSRC / "ptxas.log", "w") as f:           # the log of the LAST build only (it used to grow without bound)
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
