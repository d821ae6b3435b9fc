"""In-tree build of libsfft.so (hand-written CUDA for sm_100a + the C ABI).

    python -m sfft_b200.build          # or: from sfft_b200.build import build; build()

nvcc cross-compiles without a GPU; the resulting sfft_b200/libsfft.so travels to the
GPU box with the repo snapshot.  Rebuilds only when a source is newer than the .so.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsfft.so")
BUILD = os.path.join(HERE, "build")

CU_SOURCES = ["api.cu", "fft.cu", "plan_builder.cu", "plan_v12.cu", "v12_kernels.cu", "v3.cu", "shard.cu"]
C_SOURCES = ["cheb_host.c"]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"
HOST_CC = "/usr/bin/gcc"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                 # parity: no FMA contraction anywhere in the engine
    "-ccbin", HOST_CXX,
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fvisibility=default",
    "-Xptxas", "-v" if os.environ.get("SFFTB_PTXAS_V") else "-O3",
]


def _newer(src, dst):
    return (not os.path.exists(dst)) or os.path.getmtime(src) > os.path.getmtime(dst)


def _deps():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(HERE, "..", "include", "sfft.h"))
    files.append(os.path.abspath(__file__))
    return files


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    if not force and os.path.exists(OUT) and not any(_newer(f, OUT) for f in _deps()):
        return OUT
    objs = []
    procs = []
    for src in CU_SOURCES:
        obj = os.path.join(BUILD, src + ".o")
        cmd = [NVCC] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src in C_SOURCES:
        obj = os.path.join(BUILD, src + ".o")
        cmd = [HOST_CC, "-O2", "-fPIC", "-ffp-contract=off", "-std=gnu11", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- {src} ---\n{out}\n")
        elif verbose or os.environ.get("SFFTB_PTXAS_V"):
            sys.stderr.write(f"--- {src} ---\n{out}\n")
    if failed:
        raise RuntimeError("libsfft.so: compilation failed")
    link = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", HOST_CXX,
            "-o", OUT] + objs + ["-lm", "-lpthread"]
    subprocess.check_call(link)
    build_tools()
    return OUT


def build_tools():
    """The reference's three drivers over this library (tools/sfft_harness.cc) and the
    HBM random-gather microbenchmarks."""
    root = os.path.dirname(HERE)
    tools = os.path.join(root, "tools")
    src = os.path.join(tools, "sfft_harness.cc")
    for mode, name in ((0, "sfft-timing"), (1, "sfft-verification"), (2, "sfft-timing_many")):
        subprocess.check_call([HOST_CXX, "-O2", f"-DHARNESS_MODE={mode}", src, "-I", os.path.join(root, "include"),
                               "-L", HERE, "-lsfft", f"-Wl,-rpath,{HERE}", "-Wl,-rpath,$ORIGIN/../sfft_b200",
                               "-o", os.path.join(tools, name)])
    for mb in ("random_gather", "ld_variants", "bulk_copy", "pipe_overlap", "cluster_barrier"):
        cu = os.path.join(tools, "microbench", mb + ".cu")
        if os.path.exists(cu):
            subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
                                   "-ccbin", HOST_CXX, "-o", os.path.join(tools, "microbench", mb), cu])


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
