"""The one test-like artefact the reference holds for this path, run on the device.  (The file sorts last on purpose: it was
added after the round's last GPU call, so its first device run is the driver's - the verified suite runs before it.)"""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


UNIFORM_MESHES = {"hex": lambda: cases.pm.hex_box(7, 6, 5), "prism": lambda: cases.pm.prism_box(5, 5, 4),
                  "2d": lambda: cases.case_2d((12, 10)).mesh, "1d": lambda: cases.case_sod(30).mesh}


@pytest.mark.parametrize("mesh_name", list(UNIFORM_MESHES))
def test_reference_debug_field_check_on_device(qgd, mesh_name):
    """The one test-like artefact the reference holds for this path (QHDFoam/createFaceFluxes.H:31-65, commented out): a field
    `cellNo = mesh.C().z()` with boundary values `mesh.Cf().z()` is differentiated with `fvsc::grad` "for debugging parallel
    execution and tau-terms evaluation" - the face gradient of a linear field must be its constant slope.  Run here directly on the
    device (no oracle involved) for each solved direction: exact on faces away from the boundary of uniform meshes (on skewed meshes
    the inverse-distance point interpolation is not linear-exact, in the reference as here)."""
    from test_oracle_kat import _interior_faces
    mesh = UNIFORM_MESHES[mesh_name]()
    nI = mesh.n_internal
    st = qgd.FvscStencil(qgd.Mesh(mesh), "GaussVolPoint")
    inner = _interior_faces(mesh)
    assert inner.sum() > 0
    for d in range(3):
        if mesh.geometric_d[d] <= 0:
            continue
        cell, bnd = mesh.C[:, d].copy(), mesh.Cf[nI:, d].copy()
        bsg = mesh.deltaCoeffs[nI:] * (bnd - cell[mesh.owner[nI:]])        # snGrad of a calculated patch field
        g = st.Grad(cell, bnd, bsg)
        want = np.zeros(3); want[d] = 1.0
        assert np.abs(g[:nI][inner] - want).max() < 1e-10, (mesh_name, d)
