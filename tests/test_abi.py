"""CPU tests (no GPU, no compute calls): libr2s.so loads and exports every entry point include/r2s.h declares, the
parameter/report structs the host mirror passes have the C layout, and the product fails loudly without a device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "r2s.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(r2s_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for n in ("r2s_create", "r2s_destroy", "r2s_set_mesh", "r2s_set_grid", "r2s_set_slab", "r2s_eval_distances", "r2s_sign_detection",
              "r2s_remove_artifacts", "r2s_rbf_smoothing", "r2s_pipeline", "r2s_pipeline_resident", "r2s_nodal_densities", "r2s_find_threshold"):
        assert n in names


def test_library_exports_every_declared_symbol(r2s):
    lib = C.CDLL(r2s.library_path())
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, "declared in include/r2s.h but not exported by libr2s.so: %s" % missing


def test_no_unexpected_dependencies(r2s):
    """The product library must not link the oracle or torch: only the CUDA runtime (and NCCL for the slab exchange)."""
    out = subprocess.run(["ldd", r2s.library_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out and "torch" not in out


def test_struct_layout_matches_header(r2s):
    """sizeof/offsets of r2s_params and r2s_report as a C compiler lays them out == the ctypes mirrors."""
    prog = r"""
#include <stdio.h>
#include <stddef.h>
#include "r2s.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(r2s_params), offsetof(r2s_params, remove_artifacts), offsetof(r2s_params, smooth), offsetof(r2s_params, target_volume),
         sizeof(r2s_report), offsetof(r2s_report, launches));
  return 0;
}
"""
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        vals = [int(v) for v in subprocess.check_output([os.path.join(d, "t")]).split()]
    P, R = r2s.Params, r2s.Report
    assert vals == [C.sizeof(P), P.remove_artifacts.offset, P.smooth.offset, P.target_volume.offset, C.sizeof(R), R.launches.offset]


def test_default_params_are_the_reference_constants(r2s):
    lib = r2s.load_library()
    p = r2s.Params()
    lib.r2s_default_params(C.byref(p))
    assert p.delta_factor == 1.1                     # sdfOnDensityField.jl:158
    assert p.artifact_min_ratio == 0.01              # RhoToSDF.jl:27
    assert p.artifact_threshold == 0.0               # RhoToSDF.jl:195
    assert p.rbf_cut == 1e-3                         # RBFs4Smoothing.jl:328
    assert p.rbf_interp == 1 and p.smooth == 1 and p.remove_artifacts == 1


def test_no_cpu_fallback_without_a_device(r2s):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(r2s.R2SError):
        r2s.Context(0)
    X = np.zeros((8, 3)); IEN = np.arange(1, 9, dtype=np.int64)[None, :]
    with pytest.raises(r2s.R2SError):
        r2s.Mesh(X, IEN, np.ones(1))


def test_options_validation_mirrors_reference(r2s):
    """Rho2sdfOptions warn-and-default behaviour (src/RhoToSDF.jl:33-77)."""
    with pytest.warns(UserWarning):
        o = r2s.Rho2sdfOptions(threshold_density=1.5)
    assert o.threshold_density is None
    with pytest.warns(UserWarning):
        o = r2s.Rho2sdfOptions(rbf_grid="coarse")
    assert o.rbf_grid == "same"
    with pytest.warns(UserWarning):
        o = r2s.Rho2sdfOptions(sdf_grid_setup="auto")
    assert o.sdf_grid_setup == "manual"
    o = r2s.Rho2sdfOptions()
    assert o.rbf_interp is True and o.remove_artifacts is True and o.artifact_min_component_ratio == 0.01 and o.element_type is r2s.HEX8


def test_grid_ctor_matches_reference_formula(r2s):
    """MeshGrid.Grid (src/MeshGrid/Grid.jl:10-34) on the sphere test's box: N = [16,16,16], 4913 points."""
    g = r2s.Grid(np.array([-1.0, -1.0, -1.0]), np.array([1.0, 1.0, 1.0]), 10, 3)
    assert list(g.N) == [16, 16, 16] and g.ngp == 4913 and abs(g.cell_size - 0.2) < 1e-15
    P = r2s.generateGridPoints(g)
    assert P.shape == (4913, 3) and np.allclose(P[1] - P[0], [g.cell_size, 0, 0])


def test_tuning_knobs_named_by_the_tests_exist_in_the_sources():
    """The GPU tests and tools select kernel paths through environment knobs that r2s_create reads once; a misspelt knob would silently
    test the default path.  Every knob they name must be parsed in csrc/, and every knob parsed in csrc/ must be exercised by a test."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = "".join(open(os.path.join(root, "rho2sdf.jl_b200", "csrc", f)).read() for f in os.listdir(os.path.join(root, "rho2sdf.jl_b200", "csrc")) if f.endswith((".cu", ".cuh")))
    read = set(re.findall(r'(?:getenv|knob)\("(R2S_[A-Z0-9_]+)"', src))
    named = set()
    for f in ("tests/test_gpu_parity.py", "tests/slab_parity_ranks.py", "tests/test_slabs.py", "tools/ab_variants.py"):
        named |= set(re.findall(r'"(R2S_[A-Z0-9_]+)"', open(os.path.join(root, f)).read()))
    named -= {"R2S_TEST_OPTIN"}
    assert named <= read, sorted(named - read)
    assert read <= named, "knobs without a test: %s" % sorted(read - named)


def _build_abi_smoke(tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.join(root, "rho2sdf.jl_b200")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c11", "-Wall", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "host", "abi_smoke.c"), "-o", exe,
                           "-L", libdir, "-lr2s", "-Wl,-rpath," + libdir, "-lm"])
    return exe


def test_c_host_program_links_against_the_header(r2s, tmp_path):
    """tests/host/abi_smoke.c: a plain C11 host compiles against include/r2s.h with -Werror and links libr2s.so (no Python, no torch)."""
    import subprocess
    exe = _build_abi_smoke(tmp_path)
    import torch
    if not torch.cuda.is_available():
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 2 and "no CPU fallback" in out.stderr      # fails loudly without a device


@pytest.mark.gpu
def test_c_host_program_runs_the_pipeline(r2s, tmp_path):
    """The same program on a GPU: single-context pipeline, then r2s_multi_pipeline on 3 slabs, compared with each other."""
    import subprocess
    exe = _build_abi_smoke(tmp_path)
    out = subprocess.run([exe, "16", "3"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ABI SMOKE OK" in out.stdout, out.stdout + out.stderr
