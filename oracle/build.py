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


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(p) > t for p in (SRC, HDR))


def build(force=False):
    """Compile the oracle if missing or older than its sources; return the .so path."""
    if not force and not _stale():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    errors = []
    # /usr/bin/gcc first: the image's $CC wrapper cannot find libgomp.spec
    for cc in ("/usr/bin/gcc", shutil.which("gcc"), shutil.which("cc")):
        if not cc:
            continue
        for omp in (["-fopenmp"], []):
            cmd = [cc] + BASE + omp + ["-o", OUT, SRC, "-lm"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode == 0:
                return OUT
            errors.append(" ".join(cmd) + "\n" + r.stderr)
    raise RuntimeError("oracle build failed:\n" + "\n".join(errors))


if __name__ == "__main__":
    print(build(force=True))
