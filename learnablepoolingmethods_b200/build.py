"""In-tree build of liblpm_b200.so (nvcc, sm_100a only).  `python -m learnablepoolingmethods_b200.build`."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(CSRC, "build")
LIB_PATH = os.path.join(HERE, "liblpm_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "lpm_b200.h"))
    srcs = _sources()
    stamp_path = os.path.join(OBJ_DIR, "stamp.txt")
    digest = _digest(headers + [os.path.join(CSRC, s) for s in srcs])
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp_path):
        with open(stamp_path) as f:
            if f.read().strip() == digest:
                return LIB_PATH

    hdr_digest = _digest(headers)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        tag = os.path.join(OBJ_DIR, src[:-3] + ".tag")
        d = _digest([os.path.join(CSRC, src)]) + hdr_digest
        if not force and os.path.exists(obj) and os.path.exists(tag) and open(tag).read() == d:
            return src, 0, "(cached)"
        cmd = [NVCC] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode == 0:
            with open(tag, "w") as f:
                f.write(d)
        return src, r.returncode, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    for src, rc, out in results:
        if verbose or rc != 0:
            print(f"--- {src} ---\n{out}", file=sys.stderr)
        if rc != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    objs = [os.path.join(OBJ_DIR, s[:-3] + ".o") for s in srcs]
    cmd = [NVCC, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout + r.stderr, file=sys.stderr)
        raise RuntimeError("link failed")
    with open(stamp_path, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
