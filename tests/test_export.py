"""exportSdfToVTI (reference src/DataExport/ExportToVTI.jl:22-67): header semantics and data round trip.  The host writer needs
no GPU; the device-streaming writer is checked against the downloaded field on the GPU box."""
import numpy as np
import pytest

from fixtures import load_mesh


def test_vti_host_round_trip(r2s, tmp_path):
    g = r2s.Grid(np.array([-1.0, -2.0, 0.5]), np.array([1.0, 0.0, 1.5]), 8, 3)
    for smooth, dt in ((None, np.float64), (2, np.float32)):
        sm = 1 if smooth is None else smooth
        dims = [int(v) * sm + 1 for v in g.N]
        rng = np.random.default_rng(5)
        vals = rng.standard_normal(dims[0] * dims[1] * dims[2]).astype(dt)
        path = r2s.exportSdfToVTI(str(tmp_path / ("f%d" % sm)), g, vals, "distance", smooth)
        assert path.endswith(".vti")
        d, origin, spacing, label, arr = r2s.read_vti(path)
        assert list(d) == dims and label == "distance" and arr.dtype == dt
        assert np.allclose(origin, g.AABB_min, rtol=0, atol=0)                        # origin = Float64.(grid.AABB_min)  (:30)
        assert np.allclose(spacing, [g.cell_size / sm] * 3, rtol=1e-16)               # spacing = cell_size / smooth     (:33-36)
        assert np.array_equal(arr.ravel(), vals)                                      # reshape(values, dims...) with x fastest (:55)
        head = open(path, "rb").read(400).decode(errors="ignore")
        assert 'type="ImageData"' in head and 'byte_order="LittleEndian"' in head
    with pytest.raises(r2s.R2SError, match="doesn't match grid dimensions"):          # :44-46
        r2s.exportSdfToVTI(str(tmp_path / "bad"), g, np.zeros(7), "distance")


@pytest.mark.gpu
def test_vti_device_streaming_matches_download(r2s, tmp_path):
    X, IEN, rho = load_mesh("sphere")
    opts = r2s.Rho2sdfOptions(threshold_density=0.5, sdf_grid_setup="automatic", rbf_grid="fine")
    mesh = r2s.Mesh(X, IEN, rho)
    grid = r2s.noninteractive_sdf_grid_setup(mesh)
    rn = r2s.DenseInNodes(mesh, rho)
    d, _ = r2s.evalDistances(mesh, grid, None, rn, 0.5, want_xp=False)
    s = r2s.Sign_Detection(mesh, grid, None, rn, 0.5)
    fine, fg = r2s.RBFs_smoothing(mesh, d * s, grid, True, 2, "t")
    p = r2s.export_device_result_to_vti(mesh, str(tmp_path / "fine"), "distance", fine=True)
    dims, origin, spacing, label, arr = r2s.read_vti(p)
    assert arr.shape == fine.shape and np.array_equal(arr, fine) and label == "distance"
    assert np.allclose(spacing, [grid.cell_size / 2] * 3) and np.allclose(origin, grid.AABB_min)
    mesh.ctx.close()
