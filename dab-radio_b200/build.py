"""Build recipe for libdab_b200.so: every CUDA translation unit in csrc/ compiled for sm_100a with nvcc, in-tree.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the resulting .so travels to the GPU box.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdab_b200.so")
OBJ = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libdab_b200.so cannot be built")
    return nvcc


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=()):
    """Compile csrc/*.cu -> libdab_b200.so (incremental per object). Returns the path of the library."""
    os.makedirs(OBJ, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    objs = []
    nvcc = _nvcc()
    procs = []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out, file=sys.stderr)
        if p.returncode != 0:
            failed.append((src, out))
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(f"{s}:\n{o}" for s, o in failed))
    if force or procs or _stale(OUT, objs):
        cmd = [nvcc, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-ldl"]
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    flags = [a for a in sys.argv[1:] if a not in ("-f", "-v")]
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv, extra_flags=flags))
