"""Case builders shared by the tests (seeded, deterministic)."""
from __future__ import annotations

import numpy as np

from multiregionfoam_b200.assembly import cht_case, single_region_case, synthetic_coeffs
from multiregionfoam_b200.case import Case, Interface, RankSystem, REGION_COUPLE
from multiregionfoam_b200.decompose import decompose, decompose_cht_zslabs


def golden_region(golden_addr, key: str, symmetric: bool, seed: int = 12345):
    n = int(golden_addr[f"{key}_nCells"])
    return synthetic_coeffs(n, golden_addr[f"{key}_l"], golden_addr[f"{key}_u"], symmetric=symmetric, seed=seed, name=key)


def chain_region(n: int, symmetric: bool, seed: int = 1):
    """1-D chain (tridiagonal): the worst case for wavefront depth."""
    l = np.arange(n - 1, dtype=np.int32)
    return synthetic_coeffs(n, l, l + 1, symmetric=symmetric, seed=seed, name=f"chain{n}")


def ggi_case(golden_addr, seed: int = 7) -> Case:
    """Two unstructured regions (2dRisingBubble fluidA / fluidB addressings) coupled through a
    NON-conformal regionCouple pair: side A has 160 faces, side B only 96 of its 160 patch cells,
    with seeded ragged GGI addressing/weights (1-3 donors per face, one uncovered face)."""
    rng = np.random.default_rng(seed)
    A = golden_region(golden_addr, "bubbleA", symmetric=False, seed=seed)
    B = golden_region(golden_addr, "bubbleB", symmetric=True, seed=seed + 1)
    fcA = golden_addr["bubbleA_patch_interface"].astype(np.int32)
    fcB = golden_addr["bubbleB_patch_interfaceShadow"].astype(np.int32)[:96]
    nA, nB = fcA.size, fcB.size

    def csr(nTo, nFrom, uncovered):
        offs, addr, w = [0], [], []
        for i in range(nTo):
            k = 0 if i == uncovered else int(rng.integers(1, 4))
            donors = rng.choice(nFrom, size=k, replace=False)
            ww = rng.random(k) + 0.1
            ww /= ww.sum() if k else 1.0
            addr += list(np.sort(donors))
            w += list(ww)
            offs.append(len(addr))
        return np.array(offs, np.int32), np.array(addr, np.int32), np.array(w, np.float64)

    oA, aA, wA = csr(nA, nB, uncovered=5)    # maps B's face values onto A's faces
    oB, aB, wB = csr(nB, nA, uncovered=-1)
    cA, cB = rng.random(nA) + 0.5, rng.random(nB) + 0.5
    np.add.at(A.diag, fcA, cA)
    np.add.at(B.diag, fcB, cB)
    A.interfaces.append(Interface(REGION_COUPLE, fcA, cA, 0.9 * cA, 0, 1, 0, oA, aA, wA, "interface", nPeerFaces=nB))
    B.interfaces.append(Interface(REGION_COUPLE, fcB, cB, 1.1 * cB, 0, 0, 0, oB, aB, wB, "interfaceShadow", nPeerFaces=nA))
    return Case("ggi_bubble", [RankSystem(0, 1, [A, B])])


def random_vec(n: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).standard_normal(n)


def rel_l2(a, b) -> float:
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
