"""Build libbci_b200.so (sm_100a only) in-tree with nvcc; no torch headers involved.

    python -m lstm_ode_bci_b200.build [--force]

The shared library lands in lstm-ode-bci_b200/lib/ (git-ignored, but it travels to the GPU box
with the gpurun snapshot).  Objects are cached under build/ keyed by source mtime and a hash of the nvcc flags.
"""
import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libbci_b200.so")
OBJ_DIR = os.path.join(ROOT, "build", "obj")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--threads", "2"]
# extra -D switches for experiments (e.g. BCI_NVCC_DEFINES="-DBCI_REC_ACCURATE_ACT -DBCI_DEBUG_SWITCHES")
NVCC_FLAGS += os.environ.get("BCI_NVCC_DEFINES", "").split()


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers_mtime():
    hs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    return max(os.path.getmtime(h) for h in hs)


def _flags_tag():
    """Objects are cached per flag set: changing BCI_NVCC_DEFINES (or NVCC_FLAGS) never reuses objects built with other flags."""
    import hashlib
    return hashlib.sha1(" ".join(NVCC_FLAGS).encode()).hexdigest()[:10]


def _compile(src, verbose):
    obj = os.path.join(OBJ_DIR, "%s.%s.o" % (os.path.basename(src)[:-3], _flags_tag()))
    newest = max(os.path.getmtime(src), _headers_mtime())
    if os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj, ""
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    if force:
        for f in glob.glob(os.path.join(OBJ_DIR, "*.o")):
            os.remove(f)
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    tag_file = os.path.join(OBJ_DIR, "linked.tag")
    linked_tag = open(tag_file).read() if os.path.exists(tag_file) else ""
    if (not os.path.exists(LIB)) or linked_tag != _flags_tag() or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        open(tag_file, "w").write(_flags_tag())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
