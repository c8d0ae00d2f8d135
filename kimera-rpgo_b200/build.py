"""In-tree build of the CUDA shared library (sm_100a only).  Called by __graft_entry__.build().

The library is a plain C-ABI .so (include/rpgo_b200.h); it is built next to this file so that it
travels to the GPU box with the repo snapshot."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librpgo_b200.so")
SOURCES = ["capi.cu", "comm.cu", "pcm_kernels.cu", "pcm_tiled.cu", "clique_kernels.cu", "clique_exact.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false",  # numerical contract: no implicit contraction; fused ops are explicit fma() calls
    "-Xcompiler", "-fPIC,-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build_variant(tag, extra_flags):
    """measurement builds (tools/k3_probe.py): the same sources with extra nvcc flags into variants/librpgo_b200_<tag>.so;
    loaded instead of the product library when RPGO_LIB_PATH points at it"""
    vdir = os.path.join(HERE, "variants")
    os.makedirs(os.path.join(vdir, tag), exist_ok=True)
    out = os.path.join(vdir, "librpgo_b200_%s.so" % tag)
    objs, procs = [], []
    for s in SOURCES:
        o = os.path.join(vdir, tag, s + ".o")
        objs.append(o)
        cmd = ["nvcc"] + NVCC_FLAGS + list(extra_flags) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        outp, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(outp)
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs + ["-lcudart", "-ldl"])
    return out


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest(deps):
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in srcs:
        o = os.path.join(HERE, "build", os.path.basename(s) + ".o")
        objs.append(o)
        cmd = ["nvcc"] + NVCC_FLAGS + os.environ.get("RPGO_EXTRA_NVCC_FLAGS", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart", "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
