"""Build recipe of libclimaland_b200.so (hand-written CUDA for sm_100a, C ABI in
include/climaland_b200.h).  In-tree output next to this file so it travels with
`gpurun`; nvcc cross-compiles without a GPU."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "libclimaland_b200.so")
SOURCES = [os.path.join(CSRC, "clb_api.cu")]
DEPS = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "climaland_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libclimaland_b200.so cannot be built (there is no CPU fallback)")


def source_hash():
    """sha256 over the sources and the flags: what the built library is a function of"""
    import hashlib
    hsh = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in sorted(DEPS):
        hsh.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            hsh.update(f.read())
    return hsh.hexdigest()


STAMP = OUT + ".srchash"


def stale():
    """True when the library is missing or was built from other sources.  By CONTENT, not by mtime: a checkout or the
    copy to a GPU box changes file times, and a rebuild there would be 2 minutes of nvcc per process -- or, with one
    process per GPU, several of them writing the same file."""
    if not os.path.exists(OUT):
        return True
    try:
        with open(STAMP) as f:
            return f.read().strip() != source_hash()
    except OSError:
        return True


def build(force=False, verbose=False):
    """Compile the library if missing or built from other sources; return its path.  Safe with several processes
    (one per GPU): an exclusive file lock around the check and the build, output written aside and renamed."""
    if not force and not stale():
        return OUT
    import fcntl
    with open(OUT + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not stale():  # another process built it while this one waited
                return OUT
            tmp = OUT + f".tmp{os.getpid()}"
            cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES + ["-ldl"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
            os.replace(tmp, OUT)
            with open(STAMP, "w") as f:
                f.write(source_hash())
            if verbose:
                print(r.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return OUT


if __name__ == "__main__":
    import sys
    print(build(force=True, verbose="-v" in sys.argv))
