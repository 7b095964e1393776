"""In-tree build of libb200ldu.so (nvcc, sm_100a only) and of the CPU schedule emulator.

The library is built next to the package (``multiregionfoam_b200/lib/libb200ldu.so``) so that it
travels with the tree to a GPU box; nothing is installed into site-packages.  nvcc cross-compiles
without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from typing import List

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libb200ldu.so")
EMU_SRC = os.path.join(ROOT, "tests", "cpp", "schedule_emulate.cpp")
EMU_PATH = os.path.join(ROOT, "tests", "_build", "libschedule_emu.so")
ASM_EMU_SRC = os.path.join(ROOT, "tests", "cpp", "assemble_emulate.cpp")
ASM_EMU_PATH = os.path.join(ROOT, "tests", "_build", "libassemble_emu.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # the reference build has no FMA contraction (g++ -O3, x86-64 baseline): round every product
    "-fmad=false",
    "-Xcompiler", "-fPIC,-O3,-Wall",
    "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libb200ldu.so cannot be built (there is no CPU fallback)")


def _sources() -> List[str]:
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(ROOT, "include", "b200_ldu.h")]


def _stale(target: str, deps: List[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if force or _stale(LIB_PATH, _sources()):
        os.makedirs(LIB_DIR, exist_ok=True)
        cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB_PATH, os.path.join(CSRC, "b200_ldu.cu"), "-ldl"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    return LIB_PATH


def build_schedule_emulator(force: bool = False) -> str:
    deps = [EMU_SRC, os.path.join(CSRC, "schedule.hpp")]
    if force or _stale(EMU_PATH, deps):
        os.makedirs(os.path.dirname(EMU_PATH), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", EMU_PATH, EMU_SRC])
    return EMU_PATH


def build_assemble_emulator(force: bool = False) -> str:
    deps = [ASM_EMU_SRC, os.path.join(CSRC, "fv_assemble.hpp")]
    if force or _stale(ASM_EMU_PATH, deps):
        os.makedirs(os.path.dirname(ASM_EMU_PATH), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", ASM_EMU_PATH, ASM_EMU_SRC])
    return ASM_EMU_PATH


GGI_EMU_SRC = os.path.join(ROOT, "tests", "cpp", "ggi_emulate.cpp")
GGI_EMU_PATH = os.path.join(ROOT, "tests", "_build", "libggi_emu.so")


def build_ggi_emulator(force: bool = False) -> str:
    deps = [GGI_EMU_SRC, os.path.join(CSRC, "ggi_build.hpp")]
    if force or _stale(GGI_EMU_PATH, deps):
        os.makedirs(os.path.dirname(GGI_EMU_PATH), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", GGI_EMU_PATH, GGI_EMU_SRC])
    return GGI_EMU_PATH


MAP_EMU_SRC = os.path.join(ROOT, "tests", "cpp", "direct_map_emulate.cpp")
MAP_EMU_PATH = os.path.join(ROOT, "tests", "_build", "libdirect_map_emu.so")


def build_direct_map_emulator(force: bool = False) -> str:
    deps = [MAP_EMU_SRC, os.path.join(CSRC, "direct_map.hpp")]
    if force or _stale(MAP_EMU_PATH, deps):
        os.makedirs(os.path.dirname(MAP_EMU_PATH), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", MAP_EMU_PATH, MAP_EMU_SRC])
    return MAP_EMU_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
    print(build_schedule_emulator(force=True))
    print(build_assemble_emulator(force=True))
    print(build_ggi_emulator(force=True))
    print(build_direct_map_emulator(force=True))
