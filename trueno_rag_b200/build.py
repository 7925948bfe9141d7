"""In-tree build of libtrueno_rag_b200.so (sm_100a only).

Two flag sets:
  * STRICT translation units hold every kernel whose arithmetic must reproduce the reference's f32 results
    bit for bit: no FMA contraction, IEEE division and square root, no flush-to-zero;
  * the tensor-core fast pass (dense_gemm.cu) and the orchestration (capi.cu) use the default flags.
The shared object is written next to this file so that it travels with the repository snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "libtrueno_rag_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--extended-lambda"]
STRICT = ["-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false"]

CU_STRICT = ["dense_scan.cu", "bm25.cu", "fusion.cu", "synth.cu", "exchange.cu"]
CU_FAST = ["dense_gemm.cu", "capi.cu"]
CPP = ["host/host_mirror.cpp", "host/host_capi.cpp", "host/synth_host.cpp", "host/zstd_codec.cpp"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx() -> str:
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        for f in files:
            out.append(os.path.join(root, f))
    out.append(os.path.join(HERE, "..", "include", "trueno_rag_b200.h"))
    out.append(os.path.join(HERE, "..", "include", "trueno_rag.hpp"))
    out.append(os.path.join(HERE, "..", "include", "trueno_rag_host.h"))
    out.append(os.path.abspath(__file__))
    return sorted(out)


def _stamp() -> str:
    h = hashlib.sha256()
    for p in _sources():
        with open(p, "rb") as fh:
            h.update(os.path.relpath(p, HERE).encode())
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp_file = os.path.join(OBJ, "stamp")
    stamp = _stamp() + os.environ.get("TRR_BUILD_DEFS", "")
    if not force and os.path.exists(SO) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return SO
    nvcc, cxx = _nvcc(), _host_cxx()
    extra_defs = os.environ.get("TRR_BUILD_DEFS", "").split()  # experiments only, e.g. "-DTRR_GEMM_STAGES=2"
    jobs = []
    for src in CU_STRICT:
        jobs.append([nvcc, "-ccbin", cxx, *ARCH, *COMMON, *STRICT, *extra_defs, "-c", os.path.join(CSRC, src), "-o",
                     os.path.join(OBJ, src.replace("/", "_") + ".o")])
    for src in CU_FAST:
        jobs.append([nvcc, "-ccbin", cxx, *ARCH, *COMMON, *extra_defs, "-c", os.path.join(CSRC, src), "-o",
                     os.path.join(OBJ, src.replace("/", "_") + ".o")])
    for src in CPP:
        jobs.append([cxx, "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-fvisibility=hidden", "-fopenmp",
                     "-I/usr/local/cuda/include", "-c", os.path.join(CSRC, src), "-o",
                     os.path.join(OBJ, src.replace("/", "_") + ".o")])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        return r

    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        list(ex.map(run, jobs))
    objs = [j[-1] for j in jobs]
    run([nvcc, "-ccbin", cxx, *ARCH, "-shared", "-o", SO, *objs, "-Xcompiler", "-fopenmp", "-Xlinker", "--no-undefined"])
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
