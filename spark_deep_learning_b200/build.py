"""In-tree native builds.

* ``libsparkdl_core.so`` - C++17 engine / shm control plane (g++, no CUDA needed);
* ``libsparkdl_cuda.so`` - sm_100a kernels + CUDA host runtime (nvcc cross-compiles without a GPU).

Both are plain shared objects with a C ABI loaded through ctypes, so a rebuild takes seconds and
the artefacts travel to the GPU box with the source tree (they are git-ignored, not
gpurun-ignored).
"""
import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CORE_SRC = os.path.join(PKG, "csrc", "core")
CUDA_SRC = os.path.join(PKG, "csrc", "cuda")

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("native build failed: %s" % " ".join(cmd[:6]))
    if verbose and r.stdout.strip():
        print(r.stdout)
    return r.stdout


def build_core(force=False, verbose=False):
    out = os.path.join(PKG, "libsparkdl_core.so")
    srcs = sorted(glob.glob(os.path.join(CORE_SRC, "*.cc")))
    deps = srcs + glob.glob(os.path.join(CORE_SRC, "*.h"))
    if not force and not _newer(out, deps):
        return out
    tmp = out + ".tmp%d" % os.getpid()
    _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-o", tmp] + srcs
         + ["-lrt"], verbose)
    os.replace(tmp, out)
    return out


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def build_cuda(force=False, verbose=False, ptxas_info=False):
    out = os.path.join(PKG, "libsparkdl_cuda.so")
    srcs = sorted(glob.glob(os.path.join(CUDA_SRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CUDA_SRC, "*.cuh")) + glob.glob(os.path.join(CUDA_SRC, "*.h")) \
        + glob.glob(os.path.join(CORE_SRC, "*.h"))
    if not force and not _newer(out, deps):
        return out
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libsparkdl_cuda.so")
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = NVCC_ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                         "-I", CUDA_SRC, "-I", CORE_SRC, "--expt-relaxed-constexpr",
                         "-Xcudafe", "--diag_suppress=177"]
    if ptxas_info:
        flags += ["-Xptxas", "-v"]
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer(o, [s] + [d for d in deps if not d.endswith(".cu")]):
            cmd = [nvcc] + flags + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                              text=True)))
    log = []
    for s, p in procs:
        txt, _ = p.communicate()
        log.append(txt)
        if p.returncode != 0:
            sys.stderr.write(txt)
            raise RuntimeError("nvcc failed on %s" % s)
        if (verbose or ptxas_info) and txt.strip():
            print(txt)
    tmp = out + ".tmp%d" % os.getpid()
    _run([nvcc] + NVCC_ARCH + ["-shared", "-o", tmp] + objs + ["-lrt", "-ldl"], verbose)
    os.replace(tmp, out)
    return out


def build_all(force=False, verbose=False):
    a = build_core(force, verbose)
    b = build_cuda(force, verbose)
    return a, b


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
