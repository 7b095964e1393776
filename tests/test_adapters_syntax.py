"""The foam-extend adapters (adapters/b200LduSolvers) cannot be built here - foam-extend 4.1 is not in the
container - but they must at least be consistent C++ against the interfaces they claim to use: every source
is type-checked with g++ -fsyntax-only against the stand-in declarations in adapters/foamStub (which restate
the foam-extend class interfaces and are NOT foam-extend) and against the real C ABI headers in include/."""
import glob
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPT = os.path.join(ROOT, "adapters", "b200LduSolvers")
SOURCES = sorted(glob.glob(os.path.join(ADAPT, "*.C")))


def test_make_files_lists_every_source():
    listed = [l.strip() for l in open(os.path.join(ADAPT, "Make", "files")) if l.strip().endswith(".C")]
    assert sorted(listed) == [os.path.basename(s) for s in SOURCES]


@pytest.mark.parametrize("src", SOURCES, ids=[os.path.basename(s) for s in SOURCES])
def test_adapter_source_type_checks(src):
    cmd = ["g++", "-std=c++11", "-fsyntax-only", "-Wall", "-Wextra", "-Wno-unused-parameter", "-Werror",
           "-I", os.path.join(ROOT, "adapters", "foamStub"), "-I", os.path.join(ROOT, "include"), src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_no_stub_left_in_the_adapters():
    """the regionCouple extraction and dump() were notImplemented."""
    for src in SOURCES + glob.glob(os.path.join(ADAPT, "*.H")):
        text = open(src).read()
        assert "notImplemented" not in text, src


def test_registered_names_match_the_python_mirror():
    """Every solver / preconditioner word the adapters register is a word the Python mirror of the selection
    tables (multiregionfoam_b200/solvers.py, blockldu.py) knows, and the other way round for the cuda* names."""
    from multiregionfoam_b200 import solvers
    words = set()
    for src in glob.glob(os.path.join(ADAPT, "*.H")):
        text = open(src).read()
        words |= set(re.findall(r'declareCuda(?!LduSmoother)\w+\((cuda\w+),', text))
        words |= set(re.findall(r'declareCudaCoupledLduSolver\(\w+, "(cuda\w+)"', text))
        words |= set(re.findall(r'declareCudaLduSmoother\(\w+, "(cuda\w+)"', text))      # smoother table words
        words |= set(re.findall(r'TypeName\("(cuda\w*GaussSeidel)"\)', text))
    words = {w for w in words if not w.startswith("cudaCoupled")}
    from multiregionfoam_b200 import smoother
    known = set(solvers.SOLVER_TABLE) | set(solvers.PRECOND_TABLE) | set(smoother.SMOOTHER_TABLE)
    assert {"cudaGaussSeidel", "cudaDICGaussSeidel"} <= words
    from multiregionfoam_b200 import blockldu
    known |= set(getattr(blockldu, "BLOCK_SOLVER_TABLE", {}))
    assert {"cudaPCG", "cudaPBiCGStab", "cudaPBiCG", "cudaDIC", "cudaDILU", "cudaBlockCG", "cudaBlockBiCGStab"} <= words
    missing = {w for w in words if w not in known}
    assert not missing, missing
