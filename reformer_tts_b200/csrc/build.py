"""Build libreformer_b200.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C ABI)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
LIB = HERE.parent / (os.environ.get("RTTS_LIB_NAME") or "libreformer_b200.so")      # RTTS_LIB_NAME + RTTS_DEFS: experiment builds beside the product library
OBJ_DIR = HERE / ("build" if not os.environ.get("RTTS_LIB_NAME") else "build_" + os.environ["RTTS_LIB_NAME"].replace(".so", ""))
SOURCES = ["api.cu", "lsh_bucket.cu", "lsh_hash_tc.cu", "lsh_attn_fwd.cu", "lsh_attn_fwd64.cu", "lsh_attn_fwd64p.cu", "lsh_attn_bwd.cu", "gemm.cu", "rowwise.cu", "xattn.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--use_fast_math",
         "-Xcompiler", "-fPIC", "-I", str(ROOT / "include"), "-I", str(HERE)]
FLAGS += os.environ.get("RTTS_DEFS", "").split()      # experiment builds only (e.g. RTTS_DEFS="-DRTTS_TIME_LD"), use with -f


def source_hash(names=None) -> str:
    """sha256 (16 hex digits) over the library's sources (all of them, or the given file names): what rtts_build_id() reports."""
    if names is None:
        files = sorted([*HERE.glob("*.cu"), *HERE.glob("*.cuh"), *HERE.glob("*.h"), ROOT / "include" / "rtts_b200.h"], key=lambda p: p.name)
    else:
        files = [HERE / n for n in names]
    h = hashlib.sha256()
    for f in files:
        h.update(f.name.encode() + b"\0" + f.read_bytes() + b"\0")
    return h.hexdigest()[:16]


def _stale(obj: Path, src: Path) -> bool:
    if not obj.exists():
        return True
    newest = max(p.stat().st_mtime for p in [src, *HERE.glob("*.cuh"), *HERE.glob("*.h"), ROOT / "include" / "rtts_b200.h"])
    return obj.stat().st_mtime < newest


def build(verbose: bool = False, force: bool = False) -> Path:
    objs = []
    procs = []
    build_id = source_hash()
    id_file = OBJ_DIR / "build_id.txt"
    id_changed = not id_file.exists() or id_file.read_text() != build_id
    for name in SOURCES:
        src = HERE / name
        if not src.exists():
            continue
        obj = OBJ_DIR / (name + ".o")
        obj.parent.mkdir(exist_ok=True)
        objs.append(obj)
        if force or _stale(obj, src) or (name == "api.cu" and id_changed):
            extra = [f'-DRTTS_BUILD_ID="{build_id}"'] if name == "api.cu" else []
            cmd = [NVCC, *FLAGS, *extra, *(["-Xptxas", "-v"] if verbose else []), "-c", str(src), "-o", str(obj)]
            procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for name, proc in procs:
        out, _ = proc.communicate()
        if verbose or proc.returncode:
            print(f"--- {name}\n{out}", file=sys.stderr)
        failed |= proc.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if procs or not LIB.exists():
        subprocess.check_call([NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"])
    id_file.write_text(build_id)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
