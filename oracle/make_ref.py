"""ORACLE support — build recipe for ``oracle/_ref/``: the UNMODIFIED reference, compiled where it lies.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference is present)

The reference (leepaul009/LaneGCN-1) is pure Python, so "compiling it from its own source files" means byte-compiling
the four modules the forward path needs (lanegcn.py, layers.py, utils.py, data.py) straight from /root/reference into
sourceless ``oracle/_ref/<module>.refbc`` files (CPython bytecode, i.e. a .pyc under another extension: the gpurun
snapshot leaves *.pyc behind).  No reference source is copied into the repository: ``oracle/_ref/`` is
git-ignored build output (like our own ``.so``), but it is NOT gpurun-ignored, so it travels to the GPU box, where
/root/reference does not exist.  ``oracle/ref_loader.py`` imports the modules from there behind the same 4-item shim;
bench.py times them as the ``--impl reference`` arm (``cpu_baseline.kind == "reference"``) and as
``gpu_eager_reference`` (stock PyTorch eager on the B200).  The bytecode is tied to this image's CPython (3.12): the
GPU box runs the same image.
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("LGCN_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")
MODULES = ["lanegcn", "layers", "utils", "data"]
EXT = ".refbc"


def make(force: bool = False) -> bool:
    """Returns True if oracle/_ref is usable afterwards."""
    if not os.path.isfile(os.path.join(REF_SRC, "lanegcn.py")):
        return all(os.path.isfile(os.path.join(OUT, m + EXT)) for m in MODULES)
    os.makedirs(OUT, exist_ok=True)
    for m in MODULES:
        src, dst = os.path.join(REF_SRC, m + ".py"), os.path.join(OUT, m + EXT)
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, dfile=f"<reference>/{m}.py", doraise=True)
    with open(os.path.join(OUT, "BUILD_INFO.txt"), "w") as f:
        f.write(f"byte-compiled from {REF_SRC} by oracle/make_ref.py with CPython {sys.version.split()[0]}\n"
                f"modules: {', '.join(MODULES)}\n")
    return True


if __name__ == "__main__":
    print("oracle/_ref ready" if make(force="-f" in sys.argv) else "reference tree not found and oracle/_ref missing")
