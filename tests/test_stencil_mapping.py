"""Index model of the TMA stencil kernel's thread -> output mapping (k_stencil81_tma, r2s_rbf.cu), CPU only.

The kernel cannot run here, but its index arithmetic can be restated and checked exhaustively for small grids: which grid columns a
CTA's threads own, which tile columns their row window reads, which tile elements are written back to u_new, and which indices the
fused halo exchange stores into the neighbours' arrays.  The constants are parsed from the source, so that a change of the tile
geometry fails here first.  (Round 2: the fused halo stores tested the range with the index of a thread's FIRST output, which lies two
columns left of the row in the first CTA column -- two values per exchanged plane were never sent.  test_fused_halo_store_ranges
restates the condition as it is in the source and would have caught it.)"""
import os
import re
import numpy as np
import pytest

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rho2sdf.jl_b200", "csrc", "r2s_rbf.cu")


def consts():
    s = open(SRC).read()
    d = {k: int(re.search(r"#define\s+%s\s+(\d+)" % k, s).group(1)) for k in ("S3_X", "S3_Y", "S3_HX", "R2S_ST_NO")}
    d["S3_TX"] = d["S3_X"] + 2 * d["S3_HX"]; d["S3_TY"] = d["S3_Y"] + 4
    assert "constexpr int XS = NO == 4 ? 2 : 0;" in s and "const int gx = bx - XS + NO * tx, gy = by + ty;" in s
    assert "tile + (ty + 2 + dj) * TX + NO * tx + (S3_HX - 2 - XS)" in s
    assert "if (gi + o >= F.lo0 && gi + o < F.lo1 && in[o]) F.c_lower[gi + o]" in s and "if (gi + o >= F.hi0 && gi + o < F.hi1 && in[o]) F.c_upper[gi + o]" in s
    assert "dim3 sgrid(cdiv(nx + (R2S_ST_NO == 4 ? 2 : 0), S3_X), cdiv(ny, S3_Y)," in s
    return d


def threads(c, nx, ny):
    """all (CTA, thread) pairs of one plane: bx, by, tx, ty, gx, gy"""
    NO = c["R2S_ST_NO"]; XS = 2 if NO == 4 else 0
    gxn = -(-(nx + (2 if NO == 4 else 0)) // c["S3_X"]); gyn = -(-ny // c["S3_Y"])
    BX, BY, TXI, TYI = np.meshgrid(np.arange(gxn) * c["S3_X"], np.arange(gyn) * c["S3_Y"], np.arange(c["S3_X"] // NO), np.arange(c["S3_Y"]), indexing="ij")
    BX, BY, TXI, TYI = (a.ravel() for a in (BX, BY, TXI, TYI))
    return BX, BY, TXI, TYI, BX - XS + NO * TXI, BY + TYI, NO, XS


@pytest.mark.parametrize("nx,ny", [(1, 1), (2, 3), (30, 16), (31, 17), (32, 32), (33, 5), (62, 40), (99, 99), (519, 33)])
def test_outputs_cover_the_plane_once_and_windows_stay_in_the_tile(nx, ny):
    c = consts()
    BX, BY, TXI, TYI, GX, GY, NO, XS = threads(c, nx, ny)
    count = np.zeros((ny, nx), dtype=int)
    for o in range(NO):
        ok = (GX + o >= 0) & (GX + o < nx) & (GY < ny)
        np.add.at(count, (GY[ok], GX[ok] + o), 1)
    assert (count == 1).all()                                            # every output of the plane is produced by exactly one thread
    lo = NO * TXI + (c["S3_HX"] - 2 - XS)                                # first tile column of the row window x0 - 2 .. x0 + NO + 1
    assert (lo >= 0).all() and (lo + NO + 4 <= c["S3_TX"]).all()
    assert ((BX - c["S3_HX"]) + lo == GX - 2).all()                      # ... and it is the window of the thread's outputs
    if NO == 4:
        assert (lo % 4 == 0).all() and (GX % 2 == 0).all()               # two aligned 16-byte loads; 8-byte aligned output stores
    rows = TYI[:, None] + 2 + np.arange(-2, 3)[None, :]
    assert (rows >= 0).all() and (rows < c["S3_TY"]).all()


@pytest.mark.parametrize("nx,ny", [(2, 3), (31, 17), (33, 5), (99, 99)])
def test_unew_write_back_covers_the_plane_once(nx, ny):
    c = consts(); TX, TY, HX = c["S3_TX"], c["S3_TY"], c["S3_HX"]
    NO = c["R2S_ST_NO"]
    gxn = -(-(nx + (2 if NO == 4 else 0)) // c["S3_X"]); gyn = -(-ny // c["S3_Y"])
    count = np.zeros((ny, nx + 8), dtype=int)
    for bx in np.arange(gxn) * c["S3_X"]:
        for by in np.arange(gyn) * c["S3_Y"]:
            t = 4 * np.arange(TX * TY // 4); ly, lx = t // TX, t % TX; x, y = bx + lx - HX, by + ly - 2
            inner = (lx >= HX) & (lx < TX - HX) & (ly >= 2) & (ly < TY - 2) & (x < nx) & (y < ny)
            for q in range(4):                                           # a float4 starting at a valid x may run into the pad columns (zeros)
                np.add.at(count, (y[inner], x[inner] + q), 1)
    assert (count[:, :nx] == 1).all() and (count[:, nx + 3:] == 0).all()


@pytest.mark.parametrize("nx,ny,k0,k1,nz", [(5, 4, 3, 9, 12), (33, 17, 2, 4, 8), (99, 20, 0, 6, 12), (99, 20, 6, 12, 12)])
def test_fused_halo_store_ranges(nx, ny, k0, k1, nz):
    """the indices stored into the lower / upper neighbour's array are exactly my first / last two planes, valid columns only"""
    c = consts()
    BX, BY, TXI, TYI, GX, GY, NO, XS = threads(c, nx, ny)
    px = (nx + 3) & ~3; pl = px * ny; h = min(2, k1 - k0)
    lo0, lo1 = (k0 * pl, (k0 + h) * pl) if k0 > 0 else (0, 0)
    hi0, hi1 = ((k1 - h) * pl, k1 * pl) if k1 < nz else (0, 0)
    sent_lo, sent_hi = set(), set()
    for zo in range(k0, k1):
        gi = zo * pl + GY * px + GX
        for o in range(NO):
            inb = (GX + o >= 0) & (GX + o < nx) & (GY < ny)
            sent_lo.update((gi + o)[inb & (gi + o >= lo0) & (gi + o < lo1)].tolist())
            sent_hi.update((gi + o)[inb & (gi + o >= hi0) & (gi + o < hi1)].tolist())
    want = lambda planes: {z * pl + y * px + x for z in planes for y in range(ny) for x in range(nx)}
    assert sent_lo == (want(range(k0, k0 + h)) if k0 > 0 else set())
    assert sent_hi == (want(range(k1 - h, k1)) if k1 < nz else set())
