"""Build libflowtrack_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m flowtrack.pytorch_b200.build [--force]

The sources compile in parallel (one nvcc per .cu).  Next to the library the build writes `build_info.json`: whether this
call compiled or found the library up to date, the nvcc version, the flags and a SHA-256 over sources + header + flags,
so a stale or foreign .so is detected by content, not by file times.
"""
import concurrent.futures
import hashlib
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libflowtrack_b200.so")
INFO = os.path.join(HERE, "build_info.json")
HEADER = os.path.join(HERE, "..", "..", "include", "flowtrack_b200.h")
SOURCES = ["ft_ops.cu", "ft_conv.cu", "ft_corr.cu", "ft_track.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas=-v"]


def source_hash():
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(CSRC, f), "rb") as fh:
                h.update(f.encode() + b"\0" + fh.read())
    with open(HEADER, "rb") as fh:
        h.update(fh.read())
    return h.hexdigest()


def kernel_hash(names):
    """SHA-256 over the compile flags and the named files of csrc/ (a kernel's own translation unit plus the headers it includes):
    what an ncu capture of ONE kernel is stamped with, so that a change to an unrelated source file does not void it."""
    h = hashlib.sha256(" ".join(FLAGS).encode())
    for f in sorted(names):
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


CONV_SOURCES = ("ft_conv.cu", "ft_tc.cuh", "ft_common.cuh")     # conv_tc_kernel
CORR_SOURCES = ("ft_corr.cu", "ft_tc.cuh", "ft_common.cuh")     # corr_pair_kernel


def read_info():
    try:
        with open(INFO) as f:
            return json.load(f)
    except Exception:
        return {}


def _stale():
    return not os.path.exists(LIB) or read_info().get("source_sha256") != source_hash()


def _nvcc_version():
    try:
        out = subprocess.run([NVCC, "--version"], capture_output=True, text=True).stdout
        return out.strip().splitlines()[-2].strip() if out else None
    except Exception:
        return None


def build(force=False, verbose=False):
    """Compile every CUDA source of the package into one shared library.  Returns the .so path."""
    if not force and not _stale():
        info = read_info()
        info["last_call"] = "up to date (source hash matches the library's build record); nothing compiled"
        with open(INFO, "w") as f:
            json.dump(info, f, indent=1)
        return LIB
    t0 = time.time()
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(s):
        o = os.path.join(CSRC, os.path.basename(s)[:-3] + ".o")
        r = subprocess.run([NVCC] + FLAGS + ["-c", s, "-o", o], capture_output=True, text=True)
        return s, o, r

    objs, ptxas = [], {}
    with concurrent.futures.ThreadPoolExecutor(len(srcs)) as ex:
        for s, o, r in ex.map(compile_one, srcs):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed on %s" % s)
            objs.append(o)
            ptxas[os.path.basename(s)] = r.stderr.count("Compiling entry function")
    r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    info = {"compiled": True, "seconds": round(time.time() - t0, 1), "nvcc": _nvcc_version(), "flags": FLAGS, "sources": SOURCES,
            "kernels_compiled": ptxas, "source_sha256": source_hash(), "lib_bytes": os.path.getsize(LIB),
            "last_call": "compiled %d sources for sm_100a in %.1f s" % (len(srcs), time.time() - t0)}
    with open(INFO, "w") as f:
        json.dump(info, f, indent=1)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(read_info().get("last_call"))
