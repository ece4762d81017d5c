"""Build recipe of the in-tree CUDA library (sm_100a only).

    python -m dyros_robot_controller_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting libdrc_b200.so sits next to this file (git-ignored,
but it travels to the GPU box with the repository snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libdrc_b200.so"
SOURCES = [CSRC / "drc_lib.cu", CSRC / "drc_moma.cu", CSRC / "drc_mobile.cu", CSRC / "model.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
OBJDIR = PKG / "build"


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (needed to build libdrc_b200.so)")


HASH = PKG / "libdrc_b200.so.srchash"


def _source_hash() -> str:
    import hashlib
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*")) + [PKG.parent / "include" / "drc_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def stale() -> bool:
    """Content hash, not mtimes: the snapshot copied to the GPU box does not keep timestamps."""
    if not LIB.exists() or not HASH.exists():
        return True
    return HASH.read_text().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False, only: str | None = None, defines: tuple = ()) -> Path:
    """Compile every translation unit for sm_100a (in parallel), link libdrc_b200.so in-tree.
    `only` / `defines` are for kernel development (python -m ...build --dev): recompile everything but drc_moma.cu (its object
    is reused as it is -- valid while the shared struct layouts are unchanged), optionally with -DDRC_DEV_FR3_ONLY (7-dof
    instantiations only).  Such a build leaves the source hash stale, so the next plain build() is a full one."""
    if not force and not only and not stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    src_hash = _source_hash()   # of the sources as they are NOW (edits made while nvcc runs must leave the library stale)
    env = dict(os.environ)
    # the environment's CC/CXX wrappers are not nvcc host compilers; use the system gcc
    env.pop("CC", None), env.pop("CXX", None)
    OBJDIR.mkdir(exist_ok=True)

    def compile_one(src: Path) -> Path:
        obj = OBJDIR / (src.stem + ".o")
        if only and src.stem == "drc_moma" and only != "drc_moma" and obj.exists():
            return obj   # the slow unit that kernel development on the manipulator path does not touch
        cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], *(["-Xptxas", "-v"] if verbose else []), "-c", "-o", str(obj), str(src)]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", str(LIB), *map(str, objs)],
                       capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
    HASH.write_text("dev build" if only or defines else src_hash)
    return LIB


WRAP_SRC = PKG / "pywrap" / "py_wrapper.cpp"


def wrapper_path() -> Path:
    import sysconfig
    return PKG / ("dyros_robot_controller_cpp_wrapper" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_wrapper(force: bool = False) -> Path:
    """pybind11 extension module `dyros_robot_controller_cpp_wrapper` (the reference's module name, src/bindings.cpp:219) over the
    C ABI: g++ only, linked against libdrc_b200.so next to it (rpath $ORIGIN)."""
    import hashlib
    import sysconfig
    import pybind11
    out = wrapper_path()
    stamp = PKG / "dyros_robot_controller_cpp_wrapper.srchash"
    h = hashlib.sha256(WRAP_SRC.read_bytes() + (PKG.parent / "include" / "drc_b200.h").read_bytes()).hexdigest()
    if not force and out.exists() and stamp.exists() and stamp.read_text().strip() == h:
        return out
    env = dict(os.environ)
    env.pop("CC", None), env.pop("CXX", None)
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    cmd = [cxx, "-O2", "-shared", "-fPIC", "-std=c++17", "-fvisibility=hidden", f"-I{sysconfig.get_paths()['include']}", f"-I{pybind11.get_include()}",
           str(WRAP_SRC), "-o", str(out), f"-L{PKG}", "-ldrc_b200", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode != 0:
        raise RuntimeError("g++ failed on py_wrapper.cpp:\n" + r.stdout + r.stderr)
    stamp.write_text(h)
    return out


def check() -> None:
    """Host-code syntax pass in seconds: every translation unit with empty kernel bodies (-DDRC_SYNTAX_CHECK).  nvcc runs the
    device compilation (minutes for the real kernels) BEFORE the host compiler, so a typo in the C ABI otherwise shows up late."""
    env = dict(os.environ)
    env.pop("CC", None), env.pop("CXX", None)
    for src in SOURCES:
        r = subprocess.run([_nvcc(), *NVCC_FLAGS, "-DDRC_SYNTAX_CHECK", "-c", "-o", "/dev/null", str(src)], capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"syntax check failed on {src.name}:\n" + r.stdout + r.stderr)
    print("syntax check ok")


if __name__ == "__main__":
    if "--check" in sys.argv:
        check()
    elif "--dev" in sys.argv:
        extra = tuple(a[2:] for a in sys.argv if a.startswith("-D"))
        build(force=True, verbose="-v" in sys.argv, only="drc_lib", defines=("DRC_DEV_FR3_ONLY",) + extra)
        print("dev build (drc_lib.cu, 7-dof instantiations only)", LIB)
    else:
        build(force="--force" in sys.argv, verbose=True)
        print("built", LIB)
        print("built", build_wrapper(force="--force" in sys.argv))
