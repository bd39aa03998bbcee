"""Seeded synthetic triangle meshes for the benchmark configurations of BASELINE.json (gmsh is not
available, so the BWR / pin-lattice / quarter-core geometries of reference ``test/bwr-gmsh.jl:55-62``
are synthesised as a Delaunay pin-cell template tiled over the lattice).  Segmentation only needs a
valid conforming triangulation of the rectangle; material conformity is irrelevant to this path.
"""
from __future__ import annotations

import numpy as np

from .mesh import UnstructuredDiscreteModel


def jittered_triangle_mesh(nx: int, ny: int, lx: float = 1.0, ly: float = 1.0, jitter: float = 0.25,
                           seed: int = 1234, x0: float = 0.0, y0: float = 0.0) -> UnstructuredDiscreteModel:
    """nx x ny quads split into 2 triangles each; interior nodes jittered U(-jitter*h, jitter*h);
    boundary nodes stay on the rectangle (config 3 of BASELINE.json: 708 x 708 -> ~1.0 M triangles)."""
    rng = np.random.default_rng(seed)
    hx, hy = lx / nx, ly / ny
    ix, iy = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    x = x0 + ix * hx
    y = y0 + iy * hy
    x[:, -1] = x0 + lx
    y[-1, :] = y0 + ly
    jx = rng.uniform(-jitter * hx, jitter * hx, size=x.shape)
    jy = rng.uniform(-jitter * hy, jitter * hy, size=y.shape)
    interior = np.zeros_like(x, dtype=bool)
    interior[1:-1, 1:-1] = True
    x = np.where(interior, x + jx, x)
    y = np.where(interior, y + jy, y)
    xy = np.stack([x.reshape(-1), y.reshape(-1)], axis=1)
    nid = (iy * (nx + 1) + ix).astype(np.int64)
    a = nid[:-1, :-1].reshape(-1)
    b = nid[:-1, 1:].reshape(-1)
    c = nid[1:, :-1].reshape(-1)
    d = nid[1:, 1:].reshape(-1)
    # alternate the diagonal in a checkerboard so the mesh has no global directional bias
    par = ((ix[:-1, :-1] + iy[:-1, :-1]) % 2).reshape(-1).astype(bool)
    t1 = np.where(par[:, None], np.stack([a, b, d], 1), np.stack([a, b, c], 1))
    t2 = np.where(par[:, None], np.stack([a, d, c], 1), np.stack([b, d, c], 1))
    tri = np.empty((2 * a.size, 3), dtype=np.int64)
    tri[0::2] = t1
    tri[1::2] = t2
    return UnstructuredDiscreteModel.from_triangles(xy, tri)


def mixed_quad_triangle_mesh(nx: int, ny: int, lx: float = 1.0, ly: float = 1.0, jitter: float = 0.2, seed: int = 1234,
                             quad_fraction: float = 0.5, x0: float = 0.0, y0: float = 0.0) -> UnstructuredDiscreteModel:
    """nx x ny jittered quadrilaterals of which a seeded fraction stays a 4-node cell (nodes stored counter-clockwise, i.e. as
    a cycle) while the others are split into two triangles: the mixed mesh of SURVEY 8(f)-4."""
    rng = np.random.default_rng(seed)
    hx, hy = lx / nx, ly / ny
    ix, iy = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="xy")
    x = x0 + ix * hx
    y = y0 + iy * hy
    x[:, -1] = x0 + lx
    y[-1, :] = y0 + ly
    interior = np.zeros_like(x, dtype=bool)
    interior[1:-1, 1:-1] = True
    x = np.where(interior, x + rng.uniform(-jitter * hx, jitter * hx, size=x.shape), x)
    y = np.where(interior, y + rng.uniform(-jitter * hy, jitter * hy, size=y.shape), y)
    xy = np.stack([x.reshape(-1), y.reshape(-1)], axis=1)
    nid = (iy * (nx + 1) + ix).astype(np.int64)
    keep = rng.uniform(size=(ny, nx)) < quad_fraction
    cells = []
    for j in range(ny):
        for i in range(nx):
            a, b, c, d = int(nid[j, i]), int(nid[j, i + 1]), int(nid[j + 1, i + 1]), int(nid[j + 1, i])  # counter-clockwise
            if keep[j, i]:
                cells.append((a, b, c, d))
            elif (i + j) % 2:
                cells += [tuple(sorted((a, b, c))), tuple(sorted((a, c, d)))]
            else:
                cells += [tuple(sorted((a, b, d))), tuple(sorted((b, c, d)))]
    return UnstructuredDiscreteModel.from_cells(xy, cells)


def pin_cell_template(pitch: float, r_inner: float, clad: float, h: float, seed: int = 1234):
    """One pin cell [0,pitch]^2: rings of nodes inside r_inner + clad (two rings sit exactly on the pin
    and cladding radii), a jittered lattice in the moderator, uniformly spaced border nodes so that
    tiles conform.  Returns (xy, tri0, n_border_per_side)."""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(seed)
    c = pitch / 2.0
    r_out = r_inner + clad
    pts = [np.array([[c, c]])]
    radii = list(np.linspace(0.0, r_inner, max(2, int(round(r_inner / h)) + 1))[1:])
    n_clad = max(1, int(round(clad / h)))
    radii += list(np.linspace(r_inner, r_out, n_clad + 1)[1:])
    for r in radii:
        n_t = max(6, int(round(2 * np.pi * r / h)))
        th = rng.uniform(0, 2 * np.pi) + 2 * np.pi * np.arange(n_t) / n_t
        pts.append(np.stack([c + r * np.cos(th), c + r * np.sin(th)], 1))
    nb = max(2, int(round(pitch / h)))
    g = (np.arange(1, nb) / nb) * pitch
    gx, gy = np.meshgrid(g, g, indexing="xy")
    lat = np.stack([gx.reshape(-1), gy.reshape(-1)], 1)
    lat += rng.uniform(-0.2 * h, 0.2 * h, size=lat.shape)
    keep = np.hypot(lat[:, 0] - c, lat[:, 1] - c) > r_out + 0.6 * h
    keep &= (lat[:, 0] > 0.4 * h) & (lat[:, 0] < pitch - 0.4 * h) & (lat[:, 1] > 0.4 * h) & (lat[:, 1] < pitch - 0.4 * h)
    pts.append(lat[keep])
    s = (np.arange(nb + 1) / nb) * pitch
    s[-1] = pitch
    border = np.concatenate([
        np.stack([s, np.zeros_like(s)], 1), np.stack([s, np.full_like(s, pitch)], 1),
        np.stack([np.zeros_like(s[1:-1]), s[1:-1]], 1), np.stack([np.full_like(s[1:-1], pitch), s[1:-1]], 1)])
    pts.append(border)
    xy = np.concatenate(pts)
    tri = Delaunay(xy).simplices.astype(np.int64)
    p0, p1, p2 = xy[tri[:, 0]], xy[tri[:, 1]], xy[tri[:, 2]]
    area2 = (p1[:, 0] - p0[:, 0]) * (p2[:, 1] - p0[:, 1]) - (p2[:, 0] - p0[:, 0]) * (p1[:, 1] - p0[:, 1])
    tri = tri[np.abs(area2) > 1e-12 * pitch * pitch]
    return xy, tri, nb


def pin_lattice_mesh(n_pins: int, pitch: float = 1.26, r_inner: float = 0.4096, clad: float = 0.0655,
                     h: float = 0.05, seed: int = 1234, n_templates: int = 4) -> UnstructuredDiscreteModel:
    """n_pins x n_pins lattice of pin cells. ``n_templates`` differently seeded templates are cycled
    pseudo-randomly over the lattice so tracks do not see a perfectly periodic mesh."""
    temps = [pin_cell_template(pitch, r_inner, clad, h, seed + 17 * t) for t in range(n_templates)]
    rng = np.random.default_rng(seed)
    which = rng.integers(0, n_templates, size=(n_pins, n_pins))
    xs, ts = [], []
    off = 0
    for j in range(n_pins):
        for i in range(n_pins):
            xy, tri, _ = temps[which[j, i]]
            # border nodes are placed from exact lattice lines so shared nodes match after merging
            xs.append(xy + np.array([i * pitch, j * pitch]))
            ts.append(tri + off)
            off += xy.shape[0]
    xy = np.concatenate(xs)
    tri = np.concatenate(ts)
    # merge coincident border nodes via integer keys (grid far finer than any node distance)
    size = n_pins * pitch
    q = np.rint(xy / size * (2 ** 30)).astype(np.int64)
    key = q[:, 0] * (2 ** 31) + q[:, 1]
    _, first, inv = np.unique(key, return_index=True, return_inverse=True)
    xy_u = xy[first]
    # snap the outer border exactly onto the rectangle
    eps = size * 1e-9
    xy_u[np.abs(xy_u[:, 0]) < eps, 0] = 0.0
    xy_u[np.abs(xy_u[:, 1]) < eps, 1] = 0.0
    xy_u[np.abs(xy_u[:, 0] - size) < eps, 0] = size
    xy_u[np.abs(xy_u[:, 1] - size) < eps, 1] = size
    return UnstructuredDiscreteModel.from_triangles(xy_u, inv[tri])


def mesh_area(model: UnstructuredDiscreteModel) -> float:
    xy = model.node_coordinates
    if model.has_quads:  # shoelace sum over every cell's stored node cycle
        ptrs, data = model.cell_ptrs.astype(np.int64) - 1, model.cell_data.astype(np.int64) - 1
        nxt = np.arange(data.size) + 1
        last = ptrs[1:] - 1
        nxt[last] = ptrs[:-1]
        cross = xy[data, 0] * xy[data[nxt], 1] - xy[data[nxt], 0] * xy[data, 1]
        return float(np.abs(np.add.reduceat(cross, ptrs[:-1])).sum() / 2.0)
    t = model.triangles0()
    p0, p1, p2 = xy[t[:, 0]], xy[t[:, 1]], xy[t[:, 2]]
    a2 = (p1[:, 0] - p0[:, 0]) * (p2[:, 1] - p0[:, 1]) - (p2[:, 0] - p0[:, 0]) * (p1[:, 1] - p0[:, 1])
    return float(np.abs(a2).sum() / 2.0)


# Named workloads of BASELINE.json:configs. Returns (model, n_azim, delta).
def workload(name: str, scale: float = 1.0):
    """cfg2: BWR 4x4 (pitch 1.6, r_i 0.5, clad 0.1, lc 0.1; test/bwr-gmsh.jl:55-62), n_phi=16, delta=8e-2
    cfg3: unit square 708x708x2 jittered triangles, n_phi=64, delta=1e-3
    cfg4: 17x17 lattice pitch 1.26, ~4 M triangles, n_phi=128, delta=5e-4
    cfg5: 51x51 lattice (3x3 assemblies), ~20 M triangles, n_phi=256, delta=2e-4
    ``scale`` < 1 shrinks the mesh resolution (tests); 1.0 is the named size."""
    if name == "cfg2":
        return pin_lattice_mesh(4, 1.6, 0.5, 0.1, 0.1 / scale ** 0.5, 1234), 16, 8e-2
    if name == "cfg3":
        n = max(4, int(round(708 * scale ** 0.5)))
        return jittered_triangle_mesh(n, n, 1.0, 1.0, 0.25, 1234), 64, 1e-3
    if name == "cfg4":
        return pin_lattice_mesh(17, 1.26, 0.4096, 0.0655, 0.0158 / scale ** 0.5, 1234), 128, 5e-4
    if name == "cfg5":
        return pin_lattice_mesh(51, 1.26, 0.4096, 0.0655, 0.0212 / scale ** 0.5, 1234), 256, 2e-4
    raise ValueError(name)
