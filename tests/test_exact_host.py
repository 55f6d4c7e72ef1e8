"""Decision-critical device functions (rho2sdf.jl_b200/csrc/r2s_exact.cuh) compiled for the HOST and fuzzed against the CPU oracle bit for
bit: the sign field of the CUDA path is only reproducible if these agree in every bit, and a GPU is not needed to check that."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from fixtures import load_mesh

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SG = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], float)


@pytest.fixture(scope="module")
def host():
    src = os.path.join(HERE, "host", "exact_host.cpp")
    so = os.path.join(HERE, "host", "libexact_host.so")
    hdr = os.path.join(ROOT, "rho2sdf.jl_b200", "csrc", "r2s_exact.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I/usr/local/cuda/include", src, "-o", so])
    L = C.CDLL(so)
    dp = np.ctypeslib.ndpointer(np.float64, flags="C")
    L.exact_host_inverse_map_hex8.argtypes = [dp, dp, dp]
    L.exact_host_inverse_map_affine.argtypes = [dp, dp, dp]
    return L


def test_inverse_map_is_bit_identical_to_the_oracle(host):
    """Unstructured hexes of chapadlo.mat, boxes, distorted hexes and parallelepipeds; points inside and outside the element.  The
    per-element affine table of the sign kernel must give the very same bits as the general Newton path."""
    rng = np.random.default_rng(2)
    X, IEN, _ = load_mesh("chapadlo")
    n_aff = 0
    for t in range(6000):
        kind = t % 4
        c = rng.uniform(-5, 5, 3); h = rng.uniform(0.1, 1, 3)
        if kind == 0:
            Xe = X[IEN[rng.integers(IEN.shape[0])] - 1]
        elif kind == 1:
            Xe = c + SG * h
        elif kind == 2:
            Xe = c + SG * h + rng.uniform(-0.3, 0.3, (8, 3)) * h
        else:
            Xe = c + (SG * h) @ (np.eye(3) + rng.uniform(-0.4, 0.4, (3, 3))).T
        Xe = np.ascontiguousarray(Xe)
        x = np.ascontiguousarray(Xe.mean(0) + rng.uniform(-0.8, 0.8, 3) * (Xe.max(0) - Xe.min(0)))
        xi = np.zeros(3)
        ok = host.exact_host_inverse_map_hex8(Xe, x, xi)
        oko, xio = oracle.inverse_map_hex8(x, Xe)
        assert bool(ok) == oko and np.array_equal(xi, xio)
        xa = np.zeros(3)
        if host.exact_host_inverse_map_affine(Xe, x, xa) >= 0:
            n_aff += 1
            assert np.array_equal(xa, xi)
    assert n_aff > 1000
