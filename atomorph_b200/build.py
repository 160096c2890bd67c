"""Builds the native libraries in-tree with nvcc / g++ (no JIT cache, no pip install).

    python -m atomorph_b200.build            # libatomorph_b200.so (+ oracle libs when possible)

* atomorph_b200/libatomorph_b200.so -- CUDA kernels + C-ABI (include/amx.h) + the am::morph host
  facade (include/atomorph/*.h) + its flat C wrappers (include/amx_morph.h), sm_100a only.
* oracle/libamoracle.so, oracle/_ref/libamref.so -- test infrastructure (see oracle/README.md);
  the latter only where /root/reference exists.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "atomorph_b200", "csrc")
# AMX_BUILD_TAG=x (tuning experiments): objects in build/obj_x, library libatomorph_b200_x.so (picked up through AMX_LIB)
_TAG = os.environ.get("AMX_BUILD_TAG", "")
OUT = os.path.join(ROOT, "atomorph_b200", "libatomorph_b200%s.so" % ("_" + _TAG if _TAG else ""))
OBJDIR = os.path.join(ROOT, "build", "obj" + ("_" + _TAG if _TAG else ""))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                      # keep the reference's double operation order (no FMA contraction)
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-Wall,-Wno-unused-function",
    "-I", os.path.join(ROOT, "include"),
]


def _newer(src, dst, deps=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(p) > t for p in (src,) + tuple(deps))


def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError("build failed: " + " ".join(cmd[:3]))
    return r.stdout


def build_cuda(verbose=False, force=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    incdir = os.path.join(ROOT, "include")
    for d, _, fs in os.walk(incdir):
        headers += [os.path.join(d, f) for f in fs if f.endswith(".h")]
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))
    objs = []
    procs = []
    for s in sources:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJDIR, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or _newer(src, obj, headers):
            extra = os.environ.get("AMX_NVCC_EXTRA", "").split()      # tuning experiments: extra -D switches
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", src, "-o", obj]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
            raise RuntimeError("nvcc failed on " + cmd[-3])
        if verbose:
            print(out)
    if procs or not os.path.exists(OUT):
        _run([nvcc, "-shared", "-o", OUT] + objs + ["-lcudart", "-lpthread"])
    return OUT


def build_oracle():
    odir = os.path.join(ROOT, "oracle")
    if os.path.exists(os.path.join(odir, "am_oracle.cpp")):
        _run(["make", "-C", odir, "port"])
    if os.path.isdir("/root/reference"):
        _run(["make", "-C", odir, "ref"])
        if os.path.exists(OUT):
            _run(["make", "-C", odir, "demo"])       # the reference's CLI against both libraries (tests/test_demo_gpu.py)


def build_all(verbose=False, force=False):
    build_cuda(verbose=verbose, force=force)
    build_oracle()


if __name__ == "__main__":
    build_all(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print("built", OUT)
