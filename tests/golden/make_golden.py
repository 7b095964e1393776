"""Generates the golden fixtures under tests/golden/ from the reference tree (run in the build
container only; /root/reference does not exist on the GPU box).

  python tests/golden/make_golden.py

* polymesh_addr.npz -- lowerAddr / upperAddr (+ coupled-patch faceCells) of shipped polyMesh
  directories: the known-answer test of the mesh generator and the
  unstructured addressing-only fixtures.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from multiregionfoam_b200.mesh import read_polymesh_addressing  # noqa: E402

REF = "/root/reference/tutorials"
MESHES = {
    "chtFluid": "conjugateHeatTransfer/flowOverHeatedPlate/constant/fluid/polyMesh",
    "chtSolid": "conjugateHeatTransfer/flowOverHeatedPlate/constant/solid/polyMesh",
    "bubbleA": "multiphaseFlow/2dRisingBubble/constant/fluidA/polyMesh",
    "bubbleB": "multiphaseFlow/2dRisingBubble/constant/fluidB/polyMesh",
    "duineveld0": "multiphaseFlow/3dDuineveldRisingBubble/polyBaseMeshes/constant/domain0/polyMesh",
    "duineveld1": "multiphaseFlow/3dDuineveldRisingBubble/polyBaseMeshes/constant/domain1/polyMesh",
}

out = {}
for key, rel in MESHES.items():
    nCells, l, u, patches = read_polymesh_addressing(os.path.join(REF, rel))
    out[f"{key}_nCells"] = np.int64(nCells)
    out[f"{key}_l"] = l.astype(np.int32)
    out[f"{key}_u"] = u.astype(np.int32)
    for p in patches:
        if p.get("type") in ("regionCouple", "ggi", "cyclicGgi") or p["name"] in ("interface", "top", "interfaceShadow"):
            out[f"{key}_patch_{p['name']}"] = p["faceCells"].astype(np.int32)
    print(key, nCells, l.size, [(p["name"], p.get("type"), p["nFaces"]) for p in patches])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "polymesh_addr.npz"), **out)
print("written", os.path.getsize(os.path.join(ROOT, "tests", "golden", "polymesh_addr.npz")), "bytes")
