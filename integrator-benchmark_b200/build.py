"""Builds liblsi.so (hand-written sm_100a CUDA + the C ABI of include/lsi.h) in-tree with nvcc.

    python integrator-benchmark_b200/build.py [--force]

The shared object lands in integrator-benchmark_b200/lib/liblsi.so: git-ignored, but it travels
to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "liblsi.so")
SOURCES = ["program.cpp", "api.cu", "toy.cu", "reduce.cu", "particles.cu", "particles_f32.cu", "particles_f64.cu",
           "microbench.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3", "--expt-relaxed-constexpr",
    "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++",
]


def _nvcc():
    for c in ("/usr/local/cuda/bin/nvcc", "nvcc"):
        if c == "nvcc" or os.path.exists(c):
            return c


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "lsi.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.rsplit(".", 1)[0] + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get("LSI_EXTRA_NVCC_FLAGS", "").split() + (["-Xptxas", "-v"] if verbose else []) + \
              ["-x", "cu", "-dc" if False else "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("== %s\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building liblsi.so")
    subprocess.check_call([_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ccbin",
                          "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
