# -*- coding: utf-8 -*-
"""
Build libqmb200.so in-tree with nvcc for sm_100a (and nothing else).

    python -m quakemigrate_b200.build            # build if sources are newer than the library
    python -m quakemigrate_b200.build --force

The library lands next to this file so that it travels to the GPU box with the repository
snapshot; it is git-ignored (``*.so``).
"""

from __future__ import annotations

import pathlib
import shutil
import subprocess
import sys

HERE = pathlib.Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libqmb200.so"
SOURCES = ["qmb_api.cu", "qmb_lut.cu", "qmb_onset.cu", "qmb_scan.cu", "qmb_legacy.cu", "qmb_locate.cu",
           "qmb_output.cu", "qmb_preprocess.cu", "qmb_envelope.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and pathlib.Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libqmb200.so cannot be built (there is no CPU fallback)")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "qmb200.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> pathlib.Path:
    if not force and not needs_build():
        return LIB
    nvcc = nvcc_path()
    objdir = CSRC / "_obj"
    objdir.mkdir(exist_ok=True)
    host = "/usr/bin/g++" if pathlib.Path("/usr/bin/g++").exists() else "g++"
    common = [nvcc, "-ccbin", host, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v" if verbose else "-O3"]
    procs = []
    for src in SOURCES:
        obj = objdir / (src[:-3] + ".o")
        procs.append((src, subprocess.Popen(common + ["-c", str(CSRC / src), "-o", str(obj)],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(f"--- {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libqmb200.so")
    objs = [str(objdir / (s[:-3] + ".o")) for s in SOURCES]
    subprocess.run([nvcc, "-ccbin", host, *ARCH, "-shared", "-o", str(LIB), *objs, "-ldl"], check=True)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
