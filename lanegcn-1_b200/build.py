"""In-tree build of the C-ABI CUDA library: ``csrc/*.cu`` -> ``liblgcn_b200.so`` (sm_100a only).

nvcc cross-compiles without a GPU, so this runs in the authoring container; the built ``.so`` is git-ignored
but travels to the GPU box with the repo snapshot.  ``python -m lanegcn_b200.build`` or ``build()``.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "liblgcn_b200.so")

SOURCES = ["api.cu", "host_pack.cu", "graph.cu", "dilate.cu", "laneconv.cu", "att.cu", "gemm_simt.cu", "gemm_tc.cu", "gemm_tc_wide.cu", "laneconv_fused.cu", "laneconv_v2.cu", "forward.cu", "actor_net.cu", "pred_net.cu", "preprocess.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
HAVE_TC = int(os.path.exists(os.path.join(CSRC, "gemm_tc.cu")))
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
         f"-DLGCN_HAVE_TC={HAVE_TC}", f"-DLGCN_DEFAULT_ENGINE={HAVE_TC}"]
FLAGS += os.environ.get("LGCN_NVCC_EXTRA", "").split()   # e.g. -DLGCN_TIMELINE for tools/timeline_fused.py


def _deps(src):
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "lgcn.h"))
    return [src] + hdrs + [os.path.abspath(__file__)]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(name, verbose):
    src = os.path.join(CSRC, name)
    obj = os.path.join(OBJ, name.replace(".cu", ".o"))
    if not _stale(obj, _deps(src)):
        return obj, ""
    cmd = [NVCC, *ARCH, *FLAGS, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{p.stdout}\n{p.stderr}")
    return obj, p.stderr


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "flags.txt")
    flags_now = " ".join([NVCC, *ARCH, *FLAGS])
    if not os.path.exists(stamp) or open(stamp).read() != flags_now:
        force = True  # objects compiled with other flags (e.g. before gemm_tc.cu existed) are stale
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    open(stamp, "w").write(flags_now)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in res]
    if verbose:
        for _, log in res:
            if log:
                print(log, file=sys.stderr)
    if _stale(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-ldl", "-lpthread", "-lrt"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
