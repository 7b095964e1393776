"""Coupled patches of the block-coupled (vector4) path (BlockLduMatrix::initInterfaces /
updateInterfaces with coupleUpper): a pair of patches served inside one system on one GPU, and the decomposed solve on
two and three processes that share cuda:0 over the peer-to-peer transport - against oracle/pyblk_multi.py.
Bars as tests/test_gpu_block.py: Amul bit-exact, histories 1e-10 (20 iterations), fields 1e-8."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from block_helpers import box_addr, random_block_coeffs
from multiregionfoam_b200 import blockldu, ldu
from oracle.pyblk_multi import MultiBlockOracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cut_system(kinds, seed=5):
    """A box matrix whose faces across two planes are taken out of the LDU part and served as two pairs of patches of
    the same system (like a cyclic pair): sub dict for MultiBlockOracle with peer = 0."""
    dK, uK, sym = kinds
    n, l, u = box_addr(13, 8, 6)
    diag, upper, lower = random_block_coeffs(n, l, u, dK, uK, sym, seed=seed)
    zone = np.arange(n) * 3 // n
    cutA = np.nonzero((zone[l] == 0) & (zone[u] == 1))[0]
    cutB = np.nonzero((zone[l] == 1) & (zone[u] == 2))[0]
    inner = np.nonzero(zone[l] == zone[u])[0]
    low = (upper.transpose(0, 2, 1) if upper.ndim == 3 else upper) if lower is None else lower
    ifaces = []
    for cut in (cutA, cutB):
        k = len(ifaces)
        ifaces.append(dict(faceCells=l[cut].astype(np.int32), peer=0, peerIface=k + 1, coupleUpper=np.ascontiguousarray(-upper[cut])))
        ifaces.append(dict(faceCells=u[cut].astype(np.int32), peer=0, peerIface=k, coupleUpper=np.ascontiguousarray(-low[cut])))
    sub = dict(n=n, l=l[inner], u=u[inner], diag=diag, upper=np.ascontiguousarray(upper[inner]),
               lower=None if lower is None else np.ascontiguousarray(lower[inner]), ifaces=ifaces)
    return sub, (n, l, u, diag, upper, lower)


def device_system(ctx, sub):
    S = blockldu.BlockSystem(ctx, sub["l"], sub["u"], sub["n"])
    S.set_coeffs(sub["diag"], sub["upper"], sub["lower"])
    for I in sub["ifaces"]:
        k = S.add_interface(I["faceCells"], I["peer"], I["peerIface"])
        S.set_interface_coeffs(k, I["coupleUpper"])
    return S


@pytest.mark.parametrize("kinds", [(16, 16, False), (16, 16, True), (4, 4, False), (16, 1, False), (1, 16, True)])
def test_paired_patches_amul_bit_exact_and_equal_to_the_uncut_matrix(gpu_ctx, kinds):
    sub, (n, l, u, diag, upper, lower) = cut_system(kinds)
    S, M = device_system(gpu_ctx, sub), MultiBlockOracle([sub])
    W = blockldu.BlockSystem(gpu_ctx, l, u, n)
    W.set_coeffs(diag, upper, lower)
    try:
        x = np.random.default_rng(2).standard_normal((n, 4))
        y = S.amul(x)
        assert np.array_equal(y, M.amul([x])[0])
        whole = W.amul(x)
        assert np.max(np.abs(y - whole)) <= 1e-13 * np.max(np.abs(whole))
    finally:
        S.close()
        W.close()


@pytest.mark.parametrize("solver,pre,kinds", [("BiCGStab", "Cholesky", (16, 16, False)), ("BiCGStab", "diagonal", (16, 4, False)),
                                              ("CG", "Cholesky", (4, 1, True))])
def test_paired_patches_solve_history(gpu_ctx, solver, pre, kinds):
    sub, (n, *_rest) = cut_system(kinds, seed=8)
    S, M = device_system(gpu_ctx, sub), MultiBlockOracle([sub])
    try:
        rng = np.random.default_rng(6)
        x0, b = rng.standard_normal((n, 4)), rng.standard_normal((n, 4))
        sid = blockldu.SOLVER_CG if solver == "CG" else blockldu.SOLVER_BICGSTAB
        pid = ldu.PRECOND_CHOLESKY if pre == "Cholesky" else ldu.PRECOND_DIAGONAL
        xg, ig = S.solve(x0, b, sid, pid, tolerance=1e-11, maxIter=200)
        xo, io = M.solve([x0], [b], solver, pre, tolerance=1e-11, maxIter=200)
        assert abs(ig["normFactor"] - io["normFactor"]) <= 1e-12 * io["normFactor"]
        k = min(21, ig["history"].shape[0], io["history"].shape[0])
        assert k > 3
        assert np.max(np.abs(ig["history"][:k] - io["history"][:k]) / np.maximum(io["history"][:k], 1e-300)) < 1e-10
        assert np.linalg.norm(xg - xo[0]) <= 1e-8 * np.linalg.norm(xo[0])
    finally:
        S.close()


def test_interface_errors(gpu_ctx):
    n, l, u = box_addr(4, 3, 2)
    diag, upper, lower = random_block_coeffs(n, l, u, 4, 4, False)
    S = blockldu.BlockSystem(gpu_ctx, l, u, n)
    S.set_coeffs(diag, upper, lower)
    try:
        with pytest.raises(ldu.B200Error):
            S.add_interface(np.array([0, n], np.int32), 0, 1)          # face cell out of range
        with pytest.raises(ldu.B200Error):
            S.add_interface(np.array([0], np.int32), 1, 0)             # no rank 1 in this context
        a = S.add_interface(np.array([0, 1], np.int32), 0, 1)
        S.add_interface(np.array([2], np.int32), 0, a)                 # sizes do not match
        S.set_interface_coeffs(0, np.ones(2))
        S.set_interface_coeffs(1, np.ones(1))
        with pytest.raises(ldu.B200Error):
            S.amul(np.zeros((n, 4)))
    finally:
        S.close()


@pytest.mark.parametrize("world", [2, 3])
def test_selfpeer_decomposed_block_solve(world):
    uid = os.urandom(128).hex()
    with tempfile.TemporaryDirectory() as work:
        procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "scripts", "selfpeer_block.py"), work, str(rk), str(world), uid],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for rk in range(world)]
        outs = []
        try:
            for p in procs:
                outs.append(p.communicate(timeout=600)[0])
        finally:
            for p in procs:
                if p.poll() is None:
                    p.kill()
        assert all(p.returncode == 0 for p in procs), "\n".join(f"[rank {i} rc {p.returncode}] " + o[-2500:] for i, (p, o) in enumerate(zip(procs, outs)))
        assert "block_selfpeer_ok=True" in outs[0], outs[0][-2000:]
