"""In-tree build of libcapf_b200.so (nvcc, sm_100a only).  ``python -m capf_b200.build`` or ``build()``."""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcapf_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(os.path.dirname(HERE), "include", "capf_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """One `nvcc -c` per translation unit, side by side (objects under csrc/_obj/, git-ignored), then one link."""
    if not force and not _stale():
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    objdir = os.path.join(CSRC, "_obj")
    os.makedirs(objdir, exist_ok=True)
    hdr_time = max(os.path.getmtime(d) for d in glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                   [os.path.join(os.path.dirname(HERE), "include", "capf_b200.h"), os.path.abspath(__file__)])
    flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.isfile(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_time):
            return obj, 0, ""
        r = subprocess.run([nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj], capture_output=True, text=True)
        return obj, r.returncode, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    for obj, rc, log in results:
        if rc != 0:
            sys.stderr.write(log)
            raise RuntimeError(f"nvcc failed compiling {obj}")
        if verbose:
            sys.stderr.write(log)
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + [o for o, _, _ in results] + ["-ldl", "-o", OUT + ".tmp"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libcapf_b200.so")
    os.replace(OUT + ".tmp", OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
