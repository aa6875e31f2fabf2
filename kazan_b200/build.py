"""Build recipe for the product library.

    python -m kazan_b200.build            # rebuild if out of date
    python -m kazan_b200.build --force

Built IN-TREE so the `.so` travels to the GPU box with the repo snapshot:
    kazan_b200/libkazan_b200.so   hand-written sm_100a kernels + the C ABI of include/kazan_b200.h
(The CPU checker used by the tests has its own recipe and is not part of this package.)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "kazan_b200" / "csrc"
LIB = ROOT / "kazan_b200" / "libkazan_b200.so"

NVCC_FLAGS = [
    "-std=c++17",
    "-O3",
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    # bit-exactness contract with the oracle: the compiler never contracts a*b+c; every fused
    # multiply-add in the kernels is an explicit __fmaf_rn
    "-fmad=false",
    "-prec-div=true",
    "-prec-sqrt=true",
    "-ftz=false",
    "-Xcompiler",
    "-fPIC,-fvisibility=hidden,-ffp-contract=off,-Wall,-Wno-unused-function",
    "-cudart",
    "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the kazan_b200 kernels cannot be built")


def _newer(target: Path, sources: list[Path]) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(s.stat().st_mtime <= t for s in sources)


def _run(cmd: list[str]) -> None:
    print("+", " ".join(str(c) for c in cmd), flush=True)
    subprocess.run(cmd, check=True)


def product_sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))


def build_product(force: bool = False, verbose_ptxas: bool = False) -> Path:
    srcs = product_sources()
    deps = srcs + sorted(CSRC.glob("*.h")) + sorted((ROOT / "include").glob("*.h"))
    if not force and _newer(LIB, deps):
        return LIB
    objdir = ROOT / "build" / "obj"
    objdir.mkdir(parents=True, exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = objdir / (s.stem + ".o")
        objs.append(o)
        cmd = [_nvcc(), *NVCC_FLAGS, *os.environ.get("KZN_NVCC_EXTRA", "").split(), "-x", "cu", "-c", str(s), "-o", str(o)]
        if verbose_ptxas:
            cmd[1:1] = ["-Xptxas", "-v"]
        print("+", " ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd)))
    for s, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    _run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-o", str(LIB), *map(str, objs), "-ldl", "-lpthread"])
    return LIB


def main(argv: list[str]) -> int:
    force = "--force" in argv
    build_product(force=force, verbose_ptxas="--ptxas-v" in argv)
    print(f"built {LIB.relative_to(ROOT)}")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
