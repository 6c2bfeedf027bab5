"""Builds libsodt_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The library links the CUDA runtime statically and has no dependency on torch; it is loaded
with ctypes by ``_capi.py``.  Usage: ``python small-object-detection-transformers_b200/build.py [--force] [--verbose]``.
"""
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libsodt_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path, extra=b""):
    h = hashlib.sha256(extra)
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def _headers_digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(PKG), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    return h.digest() + " ".join(FLAGS).encode()


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hd = _headers_digest()
    objs, procs = [], []
    for src in sources():
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src[:-3] + ".o")
        stamp = obj + ".sha"
        dig = _digest(path, hd)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        cmd = [NVCC, *FLAGS, "-c", path, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
        procs.append((src, stamp, dig, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, stamp, dig, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}:\n{out}\n")
        else:
            if verbose or "warning" in out:
                sys.stderr.write(out)
            with open(stamp, "w") as f:
                f.write(dig)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    if procs or force or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
