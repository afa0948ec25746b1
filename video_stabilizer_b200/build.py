"""In-tree build of the native libraries.

  libvstab.so       CUDA kernels + C ABI (include/vstab.h), nvcc, sm_100a only
  libvstab_host.so  C++ host layer with the reference's public interfaces (imgproc.hpp,
                    alignment.hpp, stabilizer.hpp) on top of the C ABI, g++
  oracle/*.so       CPU oracle (test infrastructure; building it is not using it)

nvcc cross-compiles without a GPU, so this runs in the CPU-only container; the built
`.so` files travel to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(REPO, "include")
LIB_CUDA = os.path.join(PKG_DIR, "libvstab.so")
LIB_HOST = os.path.join(PKG_DIR, "libvstab_host.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # canonical rounding: no FMA contraction
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
    "-shared",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found; libvstab.so cannot be built")


def build_cuda(force: bool = False) -> str:
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(INCLUDE, "*.h")) + [os.path.abspath(__file__)]
    if not force and _newer(LIB_CUDA, deps):
        return LIB_CUDA
    _run([_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC, "-o", LIB_CUDA] + srcs)
    return LIB_CUDA


def build_host(force: bool = False) -> str | None:
    host_dir = os.path.join(CSRC, "host")
    srcs = sorted(glob.glob(os.path.join(host_dir, "*.cpp")))
    if not srcs:
        return None
    deps = srcs + glob.glob(os.path.join(host_dir, "*.hpp")) + glob.glob(os.path.join(host_dir, "*.h")) + \
        glob.glob(os.path.join(INCLUDE, "compat", "*.h")) + glob.glob(os.path.join(INCLUDE, "compat", "*", "*")) + \
        glob.glob(os.path.join(INCLUDE, "*.h")) + \
        [LIB_CUDA, os.path.abspath(__file__)]
    if not force and _newer(LIB_HOST, deps):
        return LIB_HOST
    cxx = shutil.which("g++") or "g++"
    _run([cxx, "-std=c++17", "-O2", "-mavx2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-pthread",
          "-I", INCLUDE, "-I", host_dir, "-idirafter", os.path.join(INCLUDE, "compat"),
          "-o", LIB_HOST] + srcs + ["-L", PKG_DIR, "-lvstab", "-lrt", "-Wl,-rpath,$ORIGIN"])
    return LIB_HOST


def build_tools(force: bool = False) -> list[str]:
    """align_test / video_test: the reference's driver programs rebuilt against the drop-in
    host library, from the public headers only."""
    host_dir = os.path.join(CSRC, "host")
    tools_dir = os.path.join(host_dir, "tools")
    bin_dir = os.path.join(PKG_DIR, "bin")
    os.makedirs(bin_dir, exist_ok=True)
    outs = []
    cxx = shutil.which("g++") or "g++"
    for name in ("align_test", "video_test", "stream_bench", "grid_search_align"):
        src = os.path.join(tools_dir, name + ".cpp")
        out = os.path.join(bin_dir, name)
        deps = [src, LIB_HOST, LIB_CUDA] + glob.glob(os.path.join(tools_dir, "*.hpp")) + glob.glob(os.path.join(host_dir, "*.hpp"))
        if force or not _newer(out, deps):
            _run([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-pthread", "-I", INCLUDE, "-I", host_dir, "-I", tools_dir,
                  "-idirafter", os.path.join(INCLUDE, "compat"), "-o", out, src,
                  "-L", PKG_DIR, "-lvstab_host", "-lvstab", "-Wl,-rpath,$ORIGIN/.."])
        outs.append(out)
    return outs


def build_oracle() -> None:
    _run(["make", "-s", "-C", os.path.join(REPO, "oracle")])
    if os.path.isdir("/root/reference") and os.path.exists(os.path.join(REPO, "oracle", "ref_shim")):
        _run(["make", "-s", "-C", os.path.join(REPO, "oracle"), "ref"])
        # the reference's own drivers against the product's headers and libraries
        _run(["make", "-s", "-C", os.path.join(REPO, "oracle"), "ref_drivers"])


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_host(force)
    build_tools(force)
    build_oracle()


if __name__ == "__main__":
    build_all(force=True)
    print("built:", LIB_CUDA, LIB_HOST if os.path.exists(LIB_HOST) else "(no host layer yet)")
