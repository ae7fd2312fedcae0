"""Build recipe for libdge_b200.so (sm_100a only, in-tree, no JIT cache).

`python -m dge_b200.build` compiles every .cu under dge_b200/csrc with
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo and links them into
dge_b200/_build/libdge_b200.so. nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libdge_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "--extended-lambda",
    "-Xptxas", "-v",
]


def _newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(p) > t for p in (src, *extra))


def build(force=False, verbose=False, variant=None, defines=()):
    """variant/defines: an A/B build of the same library with extra -D flags, written to
    _build/var_<variant>/libdge_b200.so and selected at run time with DGE_B200_LIB=<that path>."""
    global OUT, LIB
    if variant:
        saved = OUT, LIB
        OUT = os.path.join(saved[0], "var_" + variant)
        LIB = os.path.join(OUT, "libdge_b200.so")
        try:
            return _build(True, verbose, tuple(defines))
        finally:
            OUT, LIB = saved
    return _build(force, verbose, ())


def _build(force, verbose, defines):
    os.makedirs(OUT, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "dge_b200.h"))
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    objs, jobs = [], []
    for f in srcs:
        src, obj = os.path.join(CSRC, f), os.path.join(OUT, f[:-3] + ".o")
        objs.append(obj)
        if force or _newer(src, obj, headers):
            jobs.append((src, obj))

    def cc(job):
        src, obj = job
        r = subprocess.run([NVCC, *FLAGS, *defines, "-c", src, "-o", obj], capture_output=True, text=True)
        with open(obj[:-2] + ".ptxas.log", "w") as fh:
            fh.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(cc, jobs))
    if jobs or force or not os.path.exists(LIB):
        r = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
                            "-o", LIB, *objs], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    # python -m dge_b200.build [--force] [-v] [--variant NAME -DX=1 -DY=2 ...]
    name = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=name,
                defines=[a for a in sys.argv if a.startswith("-D")]))
