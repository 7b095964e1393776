"""Generator known-answer tests against the shipped polyMesh addressing."""
import numpy as np

from multiregionfoam_b200.mesh import flow_over_heated_plate, is_upper_triangular


def test_generator_reproduces_shipped_flowOverHeatedPlate(golden_addr):
    fluid, solid = flow_over_heated_plate(1, 1)
    for reg, key in ((fluid, "chtFluid"), (solid, "chtSolid")):
        assert reg.nCells == int(golden_addr[f"{key}_nCells"])
        assert np.array_equal(reg.lowerAddr, golden_addr[f"{key}_l"])
        assert np.array_equal(reg.upperAddr, golden_addr[f"{key}_u"])
    # regionCouple patch faceCells: fluid 'interface' = bottom row of block 2, solid 'top' = top row
    assert np.array_equal(fluid.side_y(1, top=False)[0], golden_addr["chtFluid_patch_interface"])
    assert np.array_equal(solid.side_y(0, top=True)[0], golden_addr["chtSolid_patch_top"])


def test_all_shipped_addressings_are_upper_triangular(golden_addr):
    for key in ("chtFluid", "chtSolid", "bubbleA", "bubbleB", "duineveld0", "duineveld1"):
        assert is_upper_triangular(golden_addr[f"{key}_l"], golden_addr[f"{key}_u"]), key


def test_refined_extruded_mesh_counts():
    fluid, solid = flow_over_heated_plate(2, 3)
    nx, ny, nz = fluid.dims()
    assert (nx, ny, nz) == (664, 82, 3)
    assert fluid.nCells == nx * ny * nz
    assert fluid.nFaces == (nx - 1) * ny * nz + nx * (ny - 1) * nz + nx * ny * (nz - 1)
    assert is_upper_triangular(fluid.lowerAddr, fluid.upperAddr)
    assert is_upper_triangular(solid.lowerAddr, solid.upperAddr)
    assert np.isclose(fluid.volume.sum(), 3.5 * 0.5 * 0.4)
