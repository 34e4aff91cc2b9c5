"""In-tree build of the sm_100a CUDA library (libirb200.so) with nvcc.

`python -m image_restoration_models_b200.build` or `__graft_entry__.build()`.
The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libirb200.so")
STAMP = os.path.join(HERE, "csrc", ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-DIRB200_TESTING",          # per-kernel test hooks + hardware probe (include/irb200_testing.h); drop for deployment
]


# Sources compiled a second time as the bf16 flavour (csrc/bf16_build.h is force-included: float16 -> bfloat16, namespace
# irb -> irb_bf16, mode-taking entry points renamed; IR_MODE_BF16 of the primary build forwards to them).  Left out: kernels
# without a 16-bit path that the flavour never calls.
BF16_SKIP = {"probe_desc.cu", "tiling.cu", "metrics.cu"}
BF16_FLAGS = ["-include", os.path.join(CSRC, "bf16_build.h")]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _jobs(objdir, extra_flags=()):
    """(source, object, flags) of both flavours."""
    jobs = []
    for src in _sources():
        base = os.path.basename(src)
        jobs.append((src, os.path.join(objdir, base[:-3] + ".o"), [*NVCC_FLAGS, *extra_flags]))
        if base not in BF16_SKIP:
            jobs.append((src, os.path.join(objdir, base[:-3] + "_bf16.o"), [*NVCC_FLAGS, *extra_flags, *BF16_FLAGS]))
    return jobs


def _fingerprint():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):  # incl. bf16_build.h
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def find_nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; the CUDA library cannot be built")
    return nvcc


def build_variant(out_path: str, extra_flags, objdir: str) -> str:
    """A/B build of the library with extra nvcc flags (e.g. -DIRB_FUSED_EXPERIMENTS) next to the shipped one; select it at
    run time with IRB200_LIB=<out_path>.  Tuning only: never the product."""
    nvcc = find_nvcc()
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    for src, obj, flags in _jobs(objdir, extra_flags):
        procs.append((src, subprocess.Popen([nvcc, *flags, "-c", src, "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out_path, *objs,
                        "-Xlinker", "--no-undefined"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("CUDA link failed")
    return out_path


def build(force: bool = False, verbose: bool = False) -> str:
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == fp:
        return LIB
    nvcc = find_nvcc()
    objs = []
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src, obj, flags in _jobs(objdir):
        cmd = [nvcc, *flags, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}\n")
    if failed:
        raise RuntimeError("CUDA build failed")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs,
            "-Xlinker", "--no-undefined"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("CUDA link failed")
    with open(STAMP, "w") as f:
        f.write(fp)
    return LIB


if __name__ == "__main__":
    if "--experiments" in sys.argv:
        root = os.path.dirname(HERE)
        os.makedirs(os.path.join(root, "build_ab"), exist_ok=True)
        print(build_variant(os.path.join(root, "build_ab", "libirb200_dbg.so"), ["-DIRB_FUSED_EXPERIMENTS"],
                            os.path.join(root, "build_ab", "obj_dbg")))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
