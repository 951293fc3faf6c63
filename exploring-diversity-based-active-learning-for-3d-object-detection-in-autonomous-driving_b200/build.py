"""In-tree build of the CUDA library (sm_100a only) -- ``libadsel_b200.so`` next to the sources.

The built ``.so`` is git-ignored but travels with the tree to the GPU box.  No JIT cache, no torch
extension machinery: the library has a plain C ABI (include/adsel_b200.h) and is bound with ctypes.
"""
import os
import subprocess

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libadsel_b200.so")
SOURCES = ["api.cu", "pool_device.cu", "ego_check.cu", "knn_graph.cu", "geodesic.cu", "greedy.cu", "greedy_warp.cu", "lazy.cu", "fused.cu", "comm.cu", "bev.cu", "pairdist.cu", "featdist.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = _sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(CSRC), "..", "include", "adsel_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


class _BuildLock:
    """Inter-process lock for in-tree builds: under torchrun every rank imports the package at the same time, and
    two compilers writing the same output (or a rank loading a half-written library) must never happen."""

    def __enter__(self):
        import fcntl
        self._f = open(os.path.join(CSRC, ".build.lock"), "w")
        fcntl.flock(self._f, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        fcntl.flock(self._f, fcntl.LOCK_UN)
        self._f.close()


def _headers():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(CSRC), "..", "include", "adsel_b200.h"))
    return [d for d in deps if os.path.exists(d)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a (one nvcc per source, in parallel, objects cached under csrc/build/ and
    redone only when the source or a header is newer) and link them into one shared library; returns its path.
    The library is written to a temporary name and renamed into place, under a lock, so concurrent ranks never
    see a partial file."""
    if not force and not needs_build():
        return LIB
    with _BuildLock():
        if not force and not needs_build():          # another process built it while this one waited
            return LIB
        from concurrent.futures import ThreadPoolExecutor
        nvcc = os.environ.get("NVCC", "nvcc")
        objdir = os.path.join(CSRC, "build")
        os.makedirs(objdir, exist_ok=True)
        hdr_time = max(os.path.getmtime(d) for d in _headers())
        compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])

        def compile_one(src):
            obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
            if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                    and os.path.getmtime(obj) > hdr_time):
                return obj
            subprocess.check_call([nvcc] + compile_flags + ["-c", "-o", obj, src], cwd=CSRC)
            return obj

        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
            objs = list(ex.map(compile_one, _sources()))
        tmp = f"{LIB}.tmp{os.getpid()}"
        try:
            # (the link step gets the arch too: without one nvcc assumes its default sm_52 and warns about it)
            subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp] + objs + ["-lcuda"], cwd=CSRC)
            os.replace(tmp, LIB)
        finally:
            if os.path.exists(tmp):
                os.remove(tmp)
    return LIB


HOSTEXT = os.path.join(CSRC, "_adsel_hostext.so")


def build_hostext(force: bool = False):
    """Compile the optional CPython helper for the host-side extraction (csrc/hostext.c); returns its path or
    None when Python.h / gcc are not available (pool.py then keeps its pure-python implementation)."""
    import shutil
    import sysconfig
    src = os.path.join(CSRC, "hostext.c")
    if not force and os.path.exists(HOSTEXT) and os.path.getmtime(HOSTEXT) >= os.path.getmtime(src):
        return HOSTEXT
    inc = sysconfig.get_paths()["include"]
    if not os.path.exists(os.path.join(inc, "Python.h")) or shutil.which("gcc") is None:
        return HOSTEXT if os.path.exists(HOSTEXT) else None
    with _BuildLock():
        if not force and os.path.exists(HOSTEXT) and os.path.getmtime(HOSTEXT) >= os.path.getmtime(src):
            return HOSTEXT
        tmp = f"{HOSTEXT}.tmp{os.getpid()}"
        flags = []
        try:                                  # numpy's headers allow reading ndarray fields without the buffer protocol
            import numpy
            npinc = numpy.get_include()
            if os.path.exists(os.path.join(npinc, "numpy", "ndarraytypes.h")):
                flags = [f"-I{npinc}", "-DADSEL_HAVE_NUMPY_HEADERS"]
        except Exception:
            pass
        try:
            subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", f"-I{inc}"] + flags + [src, "-o", tmp, "-lm"])
            os.replace(tmp, HOSTEXT)
        finally:
            if os.path.exists(tmp):
                os.remove(tmp)
    return HOSTEXT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
