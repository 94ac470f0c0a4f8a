"""TEST INFRASTRUCTURE: geometric nested-dissection ordering of an nx*ny*nz grid (x fastest), written independently of
the product's b200s_grid_nd_perm so that the CPU reference arm of bench.py does not load libb200sparse.so.
perm[k] = grid index eliminated k-th (both halves first, the separator plane last, recursively)."""
import numpy as np


def grid_nd_perm(nx, ny, nz, leaf=64):
    out = []

    def rec(x0, x1, y0, y1, z0, z1):
        sx, sy, sz = x1 - x0, y1 - y0, z1 - z0
        if sx <= 0 or sy <= 0 or sz <= 0:
            return
        if sx * sy * sz <= leaf or max(sx, sy, sz) <= 2:
            xs, ys, zs = np.meshgrid(np.arange(x0, x1), np.arange(y0, y1), np.arange(z0, z1), indexing="ij")
            idx = (xs + nx * (ys + ny * zs)).transpose(2, 1, 0).reshape(-1)
            out.append(idx)
            return
        if sx >= sy and sx >= sz:
            m = x0 + sx // 2
            rec(x0, m, y0, y1, z0, z1); rec(m + 1, x1, y0, y1, z0, z1); sep(m, m + 1, y0, y1, z0, z1)
        elif sy >= sz:
            m = y0 + sy // 2
            rec(x0, x1, y0, m, z0, z1); rec(x0, x1, m + 1, y1, z0, z1); sep(x0, x1, m, m + 1, z0, z1)
        else:
            m = z0 + sz // 2
            rec(x0, x1, y0, y1, z0, m); rec(x0, x1, y0, y1, m + 1, z1); sep(x0, x1, y0, y1, m, m + 1)

    def sep(x0, x1, y0, y1, z0, z1):
        # the separator plane itself is a 2-D grid: dissect it as well (one of its extents is 1)
        rec(x0, x1, y0, y1, z0, z1)

    rec(0, nx, 0, ny, 0, nz)
    perm = np.concatenate(out).astype(np.int64)
    assert perm.size == nx * ny * nz
    return perm
