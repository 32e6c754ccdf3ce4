"""Build recipe for the CPU oracle (test infrastructure, see soil_oracle.h).

The reference (ClimaLand.jl) is 100% Julia and no Julia toolchain exists in the
image, so there is no ``oracle/_ref`` to compile; this builds OUR C restatement
into ``oracle/_build/libsoil_oracle.so`` (git-ignored, travels with gpurun).
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "libsoil_oracle.so")
SRC = os.path.join(HERE, "soil_oracle.c")
HDR = os.path.join(HERE, "soil_oracle.h")
BASE = ["-O2", "-std=c11", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wextra", "-shared"]


STAMP = OUT + ".srchash"


def _source_hash():
    import hashlib
    hsh = hashlib.sha256(" ".join(BASE).encode())
    for p in (SRC, HDR):
        with open(p, "rb") as f:
            hsh.update(f.read())
    return hsh.hexdigest()


def _stale():
    """by content, not by mtime (a checkout or the copy to a GPU box changes file times)"""
    if not os.path.exists(OUT):
        return True
    try:
        with open(STAMP) as f:
            return f.read().strip() != _source_hash()
    except OSError:
        return True


def build(force=False):
    """Compile the oracle if missing or built from other sources; return the .so path.  Safe with several processes:
    a file lock around the check and the build, output written aside and renamed."""
    if not force and not _stale():
        return OUT
    import fcntl
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():
                return OUT
            errors = []
            tmp = OUT + f".tmp{os.getpid()}"
            # /usr/bin/gcc first: the image's $CC wrapper cannot find libgomp.spec
            for cc in ("/usr/bin/gcc", shutil.which("gcc"), shutil.which("cc")):
                if not cc:
                    continue
                for omp in (["-fopenmp"], []):
                    cmd = [cc] + BASE + omp + ["-o", tmp, SRC, "-lm"]
                    r = subprocess.run(cmd, capture_output=True, text=True)
                    if r.returncode == 0:
                        os.replace(tmp, OUT)
                        with open(STAMP, "w") as f:
                            f.write(_source_hash())
                        return OUT
                    errors.append(" ".join(cmd) + "\n" + r.stderr)
            raise RuntimeError("oracle build failed:\n" + "\n".join(errors))
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build(force=True))
