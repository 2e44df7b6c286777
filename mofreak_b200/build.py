"""Builds the product's native library in-tree: mofreak_b200/libmofreak_b200.so.

Hand-written CUDA for sm_100a only (no other arch, no JIT cache): the .so travels with the repository
snapshot to the GPU box.  -fmad=false because the BRISK refinement and FREAK orientation floats are part of
the bit-exact contract (a contracted multiply-add rounds differently from the reference's CPU arithmetic).
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmofreak_b200.so")

CU_SOURCES = ["context.cu", "kernels_front.cu", "kernels_dense.cu", "kernels_sparse.cu", "kernels_desc.cu", "kernels_bow.cu", "kernels_text.cu"]
CPP_SOURCES = [os.path.join(CSRC, "tables.cpp")]
SYNTH_SRC = os.path.join(HERE, "synth", "synth.cpp")
SYNTH_LIB = os.path.join(HERE, "synth", "libmofreak_synth.so")      # plain C++, no CUDA: frames for tests/bench

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off",
    "-cudart", "static",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, s) for s in CU_SOURCES] + CPP_SOURCES


def needs_build():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    deps.append(os.path.join(HERE, "..", "include", "mofreak_b200.h"))
    deps.append(os.path.join(HERE, "..", "include", "mofreak_b200_debug.h"))
    return any(os.path.getmtime(d) > built for d in deps)


def build_synth(force=False):
    if not force and os.path.exists(SYNTH_LIB) and os.path.getmtime(SYNTH_LIB) >= os.path.getmtime(SYNTH_SRC):
        return SYNTH_LIB
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([host_cxx, "-O2", "-std=c++14", "-fPIC", "-shared", "-o", SYNTH_LIB, SYNTH_SRC], check=True)
    return SYNTH_LIB


def build(force=False, verbose=False, extra=()):
    """Compile if sources are newer than the library.  Returns the library path."""
    build_synth(force)
    if not force and not needs_build():
        return LIB
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    extra = list(extra) + os.environ.get("MFK_NVCC_EXTRA", "").split()
    cmd = [nvcc_path(), "-ccbin", host_cxx] + NVCC_FLAGS + list(extra) + ["-shared", "-o", LIB] + sources() + ["-ldl"]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    import sys
    build(force="--force" in sys.argv, verbose=True, extra=[a for a in sys.argv[1:] if a.startswith("-X")])
