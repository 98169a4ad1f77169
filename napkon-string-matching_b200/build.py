"""Builds csrc/*.cu into napkon_string_matching/gpu/libnsm_b200.so for sm_100a (in-tree, so the
library travels with the repository snapshot).  `python build.py [--force]`."""
from __future__ import annotations

import pathlib
import subprocess
import sys

HERE = pathlib.Path(__file__).resolve().parent
CSRC = HERE / "csrc"
INCLUDE = HERE.parent / "include"
OUT = HERE / "napkon_string_matching" / "gpu" / "libnsm_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-I", str(INCLUDE), "-I", str(CSRC),
]


def sources():
    return sorted(CSRC.glob("*.cu"))


def needs_build() -> bool:
    if not OUT.exists():
        return True
    newest = max(p.stat().st_mtime for p in [*CSRC.iterdir(), *INCLUDE.glob("*.h")])
    return newest > OUT.stat().st_mtime


def build(force: bool = False, verbose: bool = False, defines=(), out: pathlib.Path = OUT) -> pathlib.Path:
    """``defines`` / ``out``: experimental variants (-DNAME=VALUE) built next to the product library
    and selected at run time with NSM_B200_LIB."""
    if not force and not needs_build() and out == OUT:
        return OUT
    cmd = ["nvcc", *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", str(out), *map(str, sources())]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if out == OUT:
        (HERE / "build.log").write_text(res.stdout + res.stderr)
    else:
        sys.stderr.write("".join(l + "\n" for l in (res.stdout + res.stderr).splitlines()
                                 if "spill" in l and " 0 bytes spill stores" not in l))
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv or bool(defs), verbose=not defs, defines=defs,
                out=pathlib.Path(outs[0]).resolve() if outs else OUT))
