"""Compile the CUDA library in-tree for sm_100a:  python -m vjf_b200.build

nvcc cross-compiles without a GPU; the resulting vjf_b200/lib/libvjf_b200.so is git-ignored but
travels to the GPU box with the repository snapshot.

Every translation unit is compiled to its own object file (vjf_b200/lib/obj/*.o, in parallel) and only
re-compiled when it or a header it includes changed, then everything is linked into one shared library.
"""
import os
import re
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libvjf_b200.so")
HEADER = os.path.join(HERE, "..", "include", "vjf_b200.h")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]
LINK_FLAGS = ARCH + ["--shared", "-cudart", "static", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: vjf_b200 needs the CUDA toolkit to build its sm_100a library")


_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _deps(path, seen=None):
    """The file and every local header it includes (transitively)."""
    seen = set() if seen is None else seen
    path = os.path.normpath(path)
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    with open(path) as f:
        for inc in _INC.findall(f.read()):
            _deps(os.path.join(os.path.dirname(path), inc), seen)
    return seen


def _obj(src):
    return os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")


def _stale(src):
    o = _obj(src)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    return any(os.path.getmtime(d) > t for d in _deps(os.path.join(CSRC, src)))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


DEBUG_SRCS = ("k_tile.cu", "step.cu")  # translation units that carry development stamps (-DVJF_DEBUG_STAMPS)
LIB_DEBUG = os.path.join(LIBDIR, "libvjf_b200_dbg.so")


def build_debug():
    """Development library with globaltimer stamps in the tile pipeline: the release objects plus DEBUG_SRCS recompiled with
    -DVJF_DEBUG_STAMPS, linked as lib/libvjf_b200_dbg.so (select it with VJF_B200_LIB=<path>)."""
    build()
    nvcc = _nvcc()
    objs = []
    for src in sources():
        if src in DEBUG_SRCS:
            o = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".dbg.o")
            r = subprocess.run([nvcc] + NVCC_FLAGS + ["-DVJF_DEBUG_STAMPS", "-c", os.path.join(CSRC, src), "-o", o], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
            objs.append(o)
        else:
            objs.append(_obj(src))
    r = subprocess.run([nvcc] + LINK_FLAGS + objs + ["-o", LIB_DEBUG], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB_DEBUG


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    todo = [s for s in sources() if force or _stale(s)]

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", _obj(src) + ".tmp"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        os.replace(_obj(src) + ".tmp", _obj(src))
        return src, r.stderr

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        for src, log in ex.map(compile_one, todo):
            if verbose:
                print(f"==== {src}\n{log}")
    cmd = [nvcc] + LINK_FLAGS + [_obj(s) for s in sources()] + ["-o", LIB + ".tmp"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    if "--debug" in sys.argv:
        print(build_debug())
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
