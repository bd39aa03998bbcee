"""Second, independent restatement (pure Python, brute-force nearest node) of the reference walk, used
only to cross-check the C oracle on tiny cases.  Follows SURVEY.md Appendix A / the reference files
src/track.jl:106-178, src/mesh.jl:91-176, src/intersection.jl:11-159, src/segment.jl:31-44."""
import math

import numpy as np

RTOL = math.sqrt(2.0 ** -52)


def isapprox(x, y, atol=0.0, rtol=None):
    if rtol is None:
        rtol = 0.0 if atol > 0 else RTOL
    return x == y or (math.isfinite(x) and math.isfinite(y) and abs(x - y) <= max(atol, rtol * max(abs(x), abs(y))))


def norm(a, b):
    return math.sqrt(a * a + b * b)


def isapprox_pt(p, q):
    return norm(p[0] - q[0], p[1] - q[1]) <= max(0.0, RTOL * max(norm(*p), norm(*q)))


def general_form(xi, xo):
    A = xi[1] - xo[1]
    B = xo[0] - xi[0]
    Cc = xi[0] * xo[1] - xo[0] * xi[1]
    n = math.sqrt(A * A + B * B + Cc * Cc)
    return (A / n, B / n, Cc / n)


def intersection(l1, l2):
    a = l1[1] * l2[0]
    b = l2[1] * l1[0]
    if isapprox(a, b):
        return True, (0.0, 0.0)
    det = a - b
    return False, ((l1[2] * l2[1] - l2[2] * l1[1]) / det, (l1[0] * l2[2] - l2[0] * l1[2]) / det)


def point_in_segment(p, q, x):
    return isapprox(norm(p[0] - x[0], p[1] - x[1]) + norm(q[0] - x[0], q[1] - x[1]), norm(p[0] - q[0], p[1] - q[1]))


class PyRef:
    def __init__(self, mesh):
        self.xy = mesh.model.node_coordinates
        self.cp, self.cd = mesh.cell_nodes
        self.np_, self.nd = mesh.node_cells
        self.bb_min, self.bb_max = mesh.bb_min, mesh.bb_max

    def cell_nodes(self, c):
        return self.cd[self.cp[c - 1] - 1:self.cp[c] - 1]

    def node_cells(self, n):
        return self.nd[self.np_[n - 1] - 1:self.np_[n] - 1]

    def pie(self, c, x):
        """point_in_element dispatched on the number of nodes (the reference's TODO at src/mesh.jl:148): a quadrilateral is tested
        with the four triangles (k1, k2, k3), k_j = mod1(i + j - 1, 4), of point_in_quadrangle (src/mesh.jl:184-201)"""
        n = self.cell_nodes(c)
        if len(n) == 4:
            return any(self.pit(c, x, [n[(i + j) % 4] for j in range(3)]) for i in range(4))
        return self.pit(c, x)

    def pit(self, c, x, n=None):
        if n is None:
            n = self.cell_nodes(c)
        (x1, y1), (x2, y2), (x3, y3) = (self.xy[n[0] - 1], self.xy[n[1] - 1], self.xy[n[2] - 1])
        x1, y1, x2, y2, x3, y3 = map(float, (x1, y1, x2, y2, x3, y3))
        d = x1 * (y2 - y3) + y1 * (x3 - x2) + (x2 * y3 - y2 * x3)
        if d == 0.0:
            return False
        l1 = ((y2 - y3) * x[0] + (x3 - x2) * x[1] + (x2 * y3 - x3 * y2)) / d
        l2 = ((y3 - y1) * x[0] + (x1 - x3) * x[1] + (x3 * y1 - x1 * y3)) / d
        l3 = ((y1 - y2) * x[0] + (x2 - x1) * x[1] + (x1 * y2 - x2 * y1)) / d
        return all(-RTOL <= v <= 1 + RTOL for v in (l1, l2, l3))

    def nearest(self, x, k):
        d2 = (x[0] - self.xy[:, 0]) ** 2 + (x[1] - self.xy[:, 1]) ** 2
        return np.lexsort((np.arange(d2.size), d2))[:k] + 1

    def find_element(self, x, k=2):
        ids = self.nearest(x, k + 1)
        for nid in ids:
            for c in self.node_cells(nid):
                if self.pie(c, x):
                    return int(c)
        return -1

    def inboundary(self, x, atol):
        return (isapprox(x[0], self.bb_max[0], atol) or isapprox(x[0], self.bb_min[0], atol)
                or isapprox(x[1], self.bb_max[1], atol) or isapprox(x[1], self.bb_min[1], atol))

    def order(self, phi, x1, x2):
        if phi < math.pi / 2:
            return (x1, x2) if x1[0] < x2[0] else (x2, x1)
        return (x1, x2) if x1[0] > x2[0] else (x2, x1)

    def intersections(self, c, abc, phi):
        n = self.cell_nodes(c)
        pts, par_found = [], False
        for i in range(len(n)):
            j = 0 if i == len(n) - 1 else i + 1
            p1 = tuple(map(float, self.xy[n[i] - 1]))
            p2 = tuple(map(float, self.xy[n[j] - 1]))
            par, X = intersection(abc, general_form(p1, p2))
            if par:
                par_found = True
                continue
            if not point_in_segment(p1, p2, X):
                continue
            pts.append(X)
        if len(pts) in (3, 4):
            best, sel = 0.0, None
            for i in range(2, len(pts) + 1):
                for j in range(i, len(pts) + 1):
                    x1, x2 = pts[i - 2], pts[j - 1]
                    li = norm(x1[0] - x2[0], x1[1] - x2[1])
                    if li > best:
                        best, sel = li, (x1, x2)
            return self.order(phi, *sel)
        if len(pts) == 2 and par_found:
            return self.order(phi, pts[0], pts[1])
        if len(pts) == 2:
            if isapprox_pt(pts[0], pts[1]):
                return pts[0], pts[1]
            return self.order(phi, pts[0], pts[1])
        return (0.0, 0.0), (0.0, 0.0)

    def walk(self, p, phi, abc, tiny=1e-8, k=5):
        sx, sy = tiny * math.cos(phi), tiny * math.sin(phi)
        segs = []
        xp = (p[0] + sx, p[1] + sy)
        i, prev = 0, -1
        while i < 10000:
            e = self.find_element(xp)
            if self.inboundary(xp, tiny):
                if not segs:
                    xp = (xp[0] + sx, xp[1] + sy)
                    continue
                break
            if e == -1:
                e = self.find_element(xp, k)
                if e == -1:
                    raise RuntimeError("Try increasing k")
            if e == prev:
                xp = (xp[0] + sx, xp[1] + sy)
                continue
            a, b = self.intersections(e, abc, phi)
            if isapprox_pt(a, b):
                xp = (xp[0] + sx, xp[1] + sy)
                continue
            segs.append((a[0], a[1], b[0], b[1], norm(a[0] - b[0], a[1] - b[1]), e))
            xp = (b[0] + sx, b[1] + sy)
            prev = e
            i += 1
        return segs
