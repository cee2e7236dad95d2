"""Build recipes for the native pieces (in-tree, so the .so files travel with a gpurun snapshot).

  libc2rt.so        CUDA kernels + the C ABI of include/c2rt.h      (nvcc, sm_100a only)
  libc2rt_host.so   C++ host mirror of the reference's scene API     (g++)
  chess2rt_headless headless executable                               (g++)
  oracle/liborc*.so CPU restatement of the reference — TEST INFRASTRUCTURE (make -C oracle)
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "chess2rt_b200")
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-Wextra", "-Wno-unused-parameter", "-pthread"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd, cwd=None):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd, cwd=cwd)


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build_variant(name, defines):
    """Tuning aid: a copy of both libraries under build_variants/<name>/ compiled with extra -D flags;
    select it with C2RT_LIB_DIR=build_variants/<name> (see api.py)."""
    vdir = os.path.join(ROOT, "build_variants", name)
    os.makedirs(vdir, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in ("render_kernel.cu", "c2rt_api.cu")]
    _run([_nvcc()] + list(NVCC_FLAGS) + ["-D" + d for d in defines] + ["-o", os.path.join(vdir, "libc2rt.so")] + srcs)
    shutil.copy(os.path.join(PKG, "libc2rt_host.so"), os.path.join(vdir, "libc2rt_host.so"))
    return vdir


def build_cuda(force=False, verbose_ptxas=False):
    out = os.path.join(PKG, "libc2rt.so")
    srcs = [os.path.join(CSRC, f) for f in ("render_kernel.cu", "c2rt_api.cu")]
    deps = srcs + [os.path.join(CSRC, "scene_dev.h"), os.path.join(ROOT, "include", "c2rt.h")]
    if not force and _newer(out, deps):
        return out
    flags = list(NVCC_FLAGS) + (["-Xptxas", "-v"] if verbose_ptxas else [])
    _run([_nvcc()] + flags + ["-o", out] + srcs)
    return out


def build_host(force=False):
    out = os.path.join(PKG, "libc2rt_host.so")
    srcs = [os.path.join(HOST, f) for f in ("rt.cpp", "scene_loader.cpp", "flatten.cpp", "renderer.cpp", "host_capi.cpp")]
    deps = srcs + [os.path.join(HOST, f) for f in ("rt.hpp", "flatten.hpp", "renderer.hpp", "scene_text.hpp")] + [
        os.path.join(ROOT, "include", "c2rt.h"), os.path.join(PKG, "libc2rt.so")]
    cxx = os.environ.get("CXX", "g++")
    if force or not _newer(out, deps):
        _run([cxx] + CXX_FLAGS + ["-shared", "-o", out] + srcs + ["-L" + PKG, "-lc2rt", "-Wl,-rpath,$ORIGIN"])
    exe = os.path.join(PKG, "chess2rt_headless")
    main = os.path.join(HOST, "main.cpp")
    if force or not _newer(exe, [main, out]):
        _run([cxx] + CXX_FLAGS + ["-o", exe, main, "-L" + PKG, "-lc2rt_host", "-lc2rt", "-Wl,-rpath,$ORIGIN"])
    return out


def build_oracle(force=False):
    """TEST INFRASTRUCTURE: the product never loads these."""
    odir = os.path.join(ROOT, "oracle")
    if force:
        _run(["make", "-C", odir, "clean"])
    _run(["make", "-C", odir, "-j2"])


def build_all(force=False):
    build_cuda(force)
    build_host(force)
    build_oracle(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
