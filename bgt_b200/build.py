"""Build libbgt_b200.so (CUDA kernels + C ABI) in-tree for sm_100a.  Used by __graft_entry__.build().

Every source is compiled to its own object (in parallel, only when it or a header changed) and the objects are linked
into bgt_b200/lib/libbgt_b200.so."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libbgt_b200.so")
SOURCES = ["api.cu", "pbwt_kernels.cu", "pairwalk.cu", "plane1.cu", "marginal.cu", "margpiece.cu", "compose.cu", "index.cu", "encode.cu", "inflate.cu", "sites.cu", "synth.cu", "flt.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function"]
LINK_FLAGS = ["--shared", "-cudart", "static", "-Xcompiler", "-fPIC", "-ldl"]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    return hs + [os.path.join(HERE, "..", "include", "bgt_b200.h"), __file__]


def _obj(src):
    return os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")


def _stale(src):
    o = _obj(src)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    return any(os.path.getmtime(d) > t for d in [os.path.join(CSRC, src)] + _headers())


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + _headers()
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    todo = [s for s in SOURCES if force or _stale(s)]

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", _obj(src)]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return src, r.returncode, r.stdout

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, todo))
    bad = [r for r in results if r[1] != 0]
    for src, rc, out in results:
        if rc != 0 or verbose:
            sys.stderr.write("---- %s\n%s" % (src, out))
    if bad:
        raise RuntimeError("nvcc failed on " + ", ".join(b[0] for b in bad))
    r = subprocess.run([nvcc] + LINK_FLAGS + [_obj(s) for s in SOURCES] + ["-o", LIB], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
