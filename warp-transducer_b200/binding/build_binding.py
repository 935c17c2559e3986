"""Builds the optional native extension module warp-transducer_b200/lib/warp_rnnt_native.so
(pybind11 + libtorch, plain g++; links libwarprnnt.so with an $ORIGIN rpath)."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "lib", "warp_rnnt_native.so")
SRC = os.path.join(HERE, "binding.cpp")


def build(force=False):
    lib = os.path.join(PKG, "lib", "libwarprnnt.so")
    if not os.path.exists(lib):
        raise RuntimeError("build libwarprnnt.so first (warp-transducer_b200/build.py)")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) > max(os.path.getmtime(SRC), os.path.getmtime(lib)):
        return OUT
    import pybind11
    import torch
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths() + [sysconfig.get_paths()["include"], "/usr/local/cuda/include",
                                os.path.join(os.path.dirname(PKG), "include"), pybind11.get_include()]
    cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=warp_rnnt_native",
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    cmd += ["-I" + i for i in inc] + [SRC, "-o", OUT, "-L" + os.path.join(PKG, "lib"), "-lwarprnnt",
                                      "-Wl,-rpath,$ORIGIN"]
    cmd += ["-L" + p for p in ce.library_paths()] + ["-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10", "-lc10_cuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("native binding build failed:\n" + r.stderr[-3000:])
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
