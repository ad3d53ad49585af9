"""Build libmmvae_b200.so (the C-ABI library) in-tree with nvcc for sm_100a only.

    python multimodal-vae-comparison_b200/build.py [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  No JIT, no arch fallbacks.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmmvae_b200.so")
STAMP = os.path.join(HERE, "csrc", ".build_stamp")
SOURCES = ["loglik.cu", "catce.cu", "osigma.cu", "latent.cu", "moe.cu", "combine.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in SOURCES] + [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "peer.cuh"),
                                                         os.path.join(ROOT, "include", "mmvae_b200.h")]
    for f in files:
        h.update(open(f, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=True):
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    objs = []
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out.decode()))
        if verbose and out.strip():
            print(out.decode())
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    subprocess.check_call(cmd)
    open(STAMP, "w").write(dig)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
