"""Builds libpanoswin_b200.so (sm_100a only) in-tree with nvcc.  No torch involvement: the library is a
plain C-ABI shared object (include/panoswin_b200.h) that links the static CUDA runtime.

`build()` is the product library.  `build(diagnostics=True)` compiles the same sources with -DPSW_DIAGNOSTICS into
libpanoswin_b200_diag.so (profiling switches of include/panoswin_b200_debug.h; used by tools/microbench.py only)."""
from __future__ import annotations

import concurrent.futures as cf
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(CSRC, "build")
LIB_PATH = os.path.join(PKG_DIR, "libpanoswin_b200.so")
DIAG_LIB_PATH = os.path.join(PKG_DIR, "libpanoswin_b200_diag.so")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
              "--expt-relaxed-constexpr"]


def _nvcc(required: bool = True):
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    if required:
        raise RuntimeError("nvcc not found: libpanoswin_b200.so cannot be built (there is no CPU fallback)")
    return None


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint(extra=()) -> str:
    h = hashlib.sha256()
    files = _sources() + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")]
    inc = os.path.join(os.path.dirname(PKG_DIR), "include")
    files += [os.path.join(inc, "panoswin_b200.h"), os.path.join(inc, "panoswin_b200_debug.h")]
    for f in files:
        h.update(os.path.relpath(f, PKG_DIR).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(ARCH_FLAGS + NVCC_FLAGS + list(extra)).encode())
    return h.hexdigest()


def is_current(diagnostics: bool = False) -> bool:
    """True when the built library matches the sources (fingerprint stamp written by build())."""
    lib = DIAG_LIB_PATH if diagnostics else LIB_PATH
    stamp = os.path.join(OBJ_DIR, "fingerprint_diag.txt" if diagnostics else "fingerprint.txt")
    extra = ("-DPSW_DIAGNOSTICS",) if diagnostics else ()
    try:
        return os.path.isfile(lib) and open(stamp).read() == _fingerprint(extra)
    except OSError:
        return False


def build(force: bool = False, verbose: bool = False, diagnostics: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link the library next to this file.  Safe to call from several
    processes at once (torchrun ranks): an exclusive file lock serialises them and the losers find the work done."""
    lib = DIAG_LIB_PATH if diagnostics else LIB_PATH
    extra = ["-DPSW_DIAGNOSTICS"] if diagnostics else []
    if not force and is_current(diagnostics):
        return lib
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, "fingerprint_diag.txt" if diagnostics else "fingerprint.txt")
    with open(os.path.join(OBJ_DIR, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current(diagnostics):          # another process built it while we waited
                return lib
            fp = _fingerprint(extra)
            suffix = ".diag.o" if diagnostics else ".o"

            def compile_one(src):
                obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + suffix)
                cmd = [nvcc] + ARCH_FLAGS + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
                r = subprocess.run(cmd, capture_output=True, text=True)
                if r.returncode != 0:
                    raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
                if verbose and (r.stdout or r.stderr):
                    print(r.stdout, r.stderr, file=sys.stderr)
                return obj

            with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
                objs = list(ex.map(compile_one, _sources()))
            tmp = lib + ".tmp"
            cmd = [nvcc] + ARCH_FLAGS + ["-shared", "-o", tmp] + objs + ["-cudart", "static", "-Xlinker", "--exclude-libs,ALL"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
            os.replace(tmp, lib)
            with open(stamp, "w") as fh:
                fh.write(fp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, diagnostics="--diag" in sys.argv))
