"""ctypes wrapper around the CPU ORACLE (oracle/rt_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from the product package.  See rt_oracle.h for the
parity-pinning statement (segment-level parity is unpinned by any reference golden).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librt_oracle.so")

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")

STATUS = {0: "ok", 1: "try increasing k (track.jl:141)", 2: "length mismatch (track.jl:172)", 3: "runaway",
          4: "undefined x_int (intersection.jl:81-95)"}


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (IEEE: -ffp-contract=off, no fast-math)."""
    src = [os.path.join(_HERE, "rt_oracle.c"), os.path.join(_HERE, "rt_oracle.h")]
    if (not force) and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    base = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-std=gnu11", "-shared", "-o", _SO,
            src[0], "-lm"]
    try:
        subprocess.run(base[:6] + ["-fopenmp"] + base[6:], check=True, capture_output=True)
    except subprocess.CalledProcessError:
        subprocess.run(base, check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        L.orc_mesh_create.restype = vp
        L.orc_mesh_create.argtypes = [C.c_int32, _f64p, C.c_int32, _i32p, _i32p, _i32p, _i32p]
        L.orc_mesh_destroy.argtypes = [vp]
        L.orc_mesh_bbox.argtypes = [vp, _f64p, _f64p]
        L.orc_general_form.argtypes = [C.c_double] * 4 + [_f64p]
        L.orc_intersection.argtypes = [_f64p, _f64p, _f64p]
        L.orc_point_in_segment.argtypes = [C.c_double] * 6
        L.orc_isapprox_scalar.argtypes = [C.c_double] * 4
        L.orc_isapprox_point.argtypes = [C.c_double] * 4
        L.orc_nn.argtypes = [vp, C.c_double, C.c_double]
        L.orc_nn.restype = C.c_int32
        L.orc_nn_brute.argtypes = [vp, C.c_double, C.c_double]
        L.orc_nn_brute.restype = C.c_int32
        L.orc_knn.argtypes = [vp, C.c_double, C.c_double, C.c_int, C.c_int32, _i32p]
        L.orc_point_in_triangle.argtypes = [vp, C.c_int32, C.c_double, C.c_double]
        L.orc_point_in_element.argtypes = [vp, C.c_int32, C.c_double, C.c_double]
        L.orc_find_element.argtypes = [vp, C.c_double, C.c_double, C.c_int]
        L.orc_find_element.restype = C.c_int32
        L.orc_inboundary.argtypes = [vp, C.c_double, C.c_double, C.c_double]
        L.orc_intersections.argtypes = [vp, C.c_int32, _f64p, C.c_double, _f64p, _i32p, C.POINTER(C.c_int)]
        L.orc_tg_create.argtypes = [C.POINTER(vp), vp, C.c_int, C.c_double, _i32p, C.c_double]
        L.orc_tg_destroy.argtypes = [vp]
        L.orc_tg_nazim2.argtypes = [vp]
        L.orc_tg_n_total_tracks.argtypes = [vp]
        L.orc_tg_n_total_tracks.restype = C.c_int64
        L.orc_tg_counts.argtypes = [vp, _i64p, _i64p, _i64p]
        L.orc_trace.argtypes = [vp]
        L.orc_tg_quadrature.argtypes = [vp, _f64p, _f64p, _f64p]
        L.orc_tg_angle_tables.argtypes = [vp] + [_f64p] * 5
        L.orc_tg_tracks.argtypes = [vp] + [vp] * 13
        L.orc_segmentize.argtypes = [vp, C.c_int, C.c_double, C.c_int64, C.c_int64, C.c_int,
                                     C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_seg_counts.argtypes = [vp, C.c_int64, C.c_int64, _i64p, _i32p]
        L.orc_seg_copy.argtypes = [vp, C.c_int64, C.c_int64] + [vp] * 6
        L.orc_seg_stats.argtypes = [vp, _i64p]
        L.orc_set_diag_ties.argtypes = [C.c_int]
        L.orc_nn_is_tied.argtypes = [vp, C.c_double, C.c_double]
        L.orc_volumes.argtypes = [vp, _f64p]
        L.orc_seg_free.argtypes = [vp]
        _lib = L
    return _lib


RTOL = 1.4901161193847656e-8


def set_diag_ties(on: bool):
    """Count, in ``stats()["nn_ties"]``, the find_element queries whose two nearest nodes are at exactly the same distance -- the only
    queries whose answer depends on how an exact nearest-neighbour search breaks ties (diagnostic: slows the walk, keep it off
    in timed runs)."""
    lib().orc_set_diag_ties(1 if on else 0)

BC = {"Vacuum": 0, "Reflective": 1, "Periodic": 2}


class OracleError(RuntimeError):
    pass


class OracleMesh:
    def __init__(self, xy, cell_ptrs, cell_data, node_cell_ptrs, node_cell_data):
        self.xy = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        self.n_nodes = self.xy.shape[0]
        self.n_cells = len(cell_ptrs) - 1
        self._h = lib().orc_mesh_create(self.n_nodes, self.xy.reshape(-1), self.n_cells,
                                        np.ascontiguousarray(cell_ptrs, np.int32),
                                        np.ascontiguousarray(cell_data, np.int32),
                                        np.ascontiguousarray(node_cell_ptrs, np.int32),
                                        np.ascontiguousarray(node_cell_data, np.int32))

    @classmethod
    def from_mesh(cls, mesh):
        """from a raytracing_jl_b200.mesh.Mesh"""
        return cls(mesh.model.node_coordinates, mesh.cell_nodes[0], mesh.cell_nodes[1], mesh.node_cells[0],
                   mesh.node_cells[1])

    def bbox(self):
        a, b = np.zeros(2), np.zeros(2)
        lib().orc_mesh_bbox(self._h, a, b)
        return a, b

    def nn(self, x, y):
        return lib().orc_nn(self._h, x, y)

    def nn_brute(self, x, y):
        return lib().orc_nn_brute(self._h, x, y)

    def knn(self, x, y, k, skip):
        ids = np.zeros(32, np.int32)  # ORC_MAX_K
        n = lib().orc_knn(self._h, x, y, k, skip, ids)
        return ids[:n].copy()

    def find_element(self, x, y, k=2):
        return lib().orc_find_element(self._h, x, y, k)

    def point_in_triangle(self, cell, x, y):
        return bool(lib().orc_point_in_triangle(self._h, cell, x, y))

    def point_in_element(self, cell, x, y):
        return bool(lib().orc_point_in_element(self._h, cell, x, y))

    def inboundary(self, x, y, atol):
        return bool(lib().orc_inboundary(self._h, x, y, atol))

    def intersections(self, cell, abc, phi):
        pq = np.zeros(4)
        ed = np.zeros(2, np.int32)
        n = C.c_int(0)
        rc = lib().orc_intersections(self._h, cell, np.ascontiguousarray(abc, np.float64), phi, pq, ed, C.byref(n))
        return rc, pq, ed, n.value

    def __del__(self):
        try:
            lib().orc_mesh_destroy(self._h)
        except Exception:
            pass


class OracleTrackGenerator:
    """TrackGenerator(model, n_azim, delta; bcs, tiny_step) -> trace() -> segmentize()."""

    _ERR = {-1: "number of azimuthal angles must be positive.", -2: "number of azimuthal angles must be a multiple of 4.",
            -3: "azimuthal spacing must be positive.", -4: "could not found track exit point.",
            -5: "Boundaries do not match!", -6: "Point do not lie in the boundary.",
            -7: "Segmentation is intended after tracing. Please, call `trace!` first!"}

    def __init__(self, mesh: OracleMesh, n_azim: int, delta: float, bcs=(0, 0, 0, 0), tiny_step: float = 1e-8):
        self.mesh = mesh
        self.n_azim = n_azim
        h = C.c_void_p()
        rc = lib().orc_tg_create(C.byref(h), mesh._h, n_azim, float(delta), np.asarray(bcs, np.int32), tiny_step)
        if rc:
            raise OracleError(self._ERR[rc])
        self._h = h
        self.n2 = lib().orc_tg_nazim2(h)
        self.n_total_tracks = lib().orc_tg_n_total_tracks(h)
        self.n_tracks_x = np.zeros(self.n2, np.int64)
        self.n_tracks_y = np.zeros(self.n2, np.int64)
        self.n_tracks = np.zeros(self.n2, np.int64)
        lib().orc_tg_counts(h, self.n_tracks_x, self.n_tracks_y, self.n_tracks)

    def trace(self):
        rc = lib().orc_trace(self._h)
        if rc:
            raise OracleError(self._ERR[rc])
        n2, n = self.n2, self.n_total_tracks
        self.phis, self.deltas, self.weights = np.zeros(n2), np.zeros(n2), np.zeros(n2)
        lib().orc_tg_quadrature(self._h, self.phis, self.deltas, self.weights)
        self.sin_phi, self.cos_phi, self.tan_phi, self.dx_eff, self.dy_eff = (np.zeros(n2) for _ in range(5))
        lib().orc_tg_angle_tables(self._h, self.sin_phi, self.cos_phi, self.tan_phi, self.dx_eff, self.dy_eff)
        t = dict(azim_idx=np.zeros(n, np.int64), track_idx=np.zeros(n, np.int64), p=np.zeros((n, 2)),
                 q=np.zeros((n, 2)), phi=np.zeros(n), len=np.zeros(n), abc=np.zeros((n, 3)),
                 bc_fwd=np.zeros(n, np.int8), bc_bwd=np.zeros(n, np.int8), dir_fwd=np.zeros(n, np.int8),
                 dir_bwd=np.zeros(n, np.int8), next_fwd=np.zeros(n, np.int64), next_bwd=np.zeros(n, np.int64))
        lib().orc_tg_tracks(self._h, *[v.ctypes.data_as(C.c_void_p) for v in t.values()])
        self.tracks = t
        return self

    def segmentize(self, k: int = 5, rtol: float = RTOL, uid_begin: int = 1, uid_end: int | None = None,
                   nthreads: int = 1, fetch: bool = True, check: bool = True):
        if uid_end is None:
            uid_end = self.n_total_tracks + 1
        nseg, bad = C.c_int64(0), C.c_int64(0)
        rc = lib().orc_segmentize(self._h, k, rtol, uid_begin, uid_end, nthreads, C.byref(nseg), C.byref(bad))
        if rc == -7:
            raise OracleError(self._ERR[rc])
        if rc == -8:
            raise OracleError(f"k = {k} outside [1, 32] (ORC_MAX_K of the restatement; the reference has no limit)")
        self.n_segments = nseg.value
        self.first_bad_uid, self.bad_status = bad.value, rc
        nt = uid_end - uid_begin
        self.seg_counts = np.zeros(nt, np.int64)
        self.seg_status = np.zeros(nt, np.int32)
        lib().orc_seg_counts(self._h, uid_begin, uid_end, self.seg_counts, self.seg_status)
        self.seg_offsets = np.concatenate([[0], np.cumsum(self.seg_counts)])
        if fetch:
            S = self.n_segments
            self.seg = dict(px=np.zeros(S), py=np.zeros(S), qx=np.zeros(S), qy=np.zeros(S), len=np.zeros(S),
                            element=np.zeros(S, np.int32))
            lib().orc_seg_copy(self._h, uid_begin, uid_end, *[v.ctypes.data_as(C.c_void_p) for v in self.seg.values()])
        if check and rc:
            raise OracleError(f"track uid {bad.value}: {STATUS.get(rc, rc)}")
        return self

    def fetch(self, uid_begin: int, uid_end: int):
        """counts / status / Segment columns of the already segmentized uids [uid_begin, uid_end), concatenated in uid order"""
        nt = uid_end - uid_begin
        out = dict(counts=np.zeros(nt, np.int64), status=np.zeros(nt, np.int32))
        lib().orc_seg_counts(self._h, uid_begin, uid_end, out["counts"], out["status"])
        S = int(out["counts"].sum())
        out.update(px=np.zeros(S), py=np.zeros(S), qx=np.zeros(S), qy=np.zeros(S), len=np.zeros(S), element=np.zeros(S, np.int32))
        lib().orc_seg_copy(self._h, uid_begin, uid_end, *[out[k].ctypes.data_as(C.c_void_p) for k in ("px", "py", "qx", "qy", "len", "element")])
        return out

    def stats(self):
        s = np.zeros(8, np.int64)
        lib().orc_seg_stats(self._h, s)
        names = ["steps", "knn_fallbacks", "k_retries", "same_element_resteps", "boundary_start_steps",
                 "vertex_steps", "n_int_ge3", "nn_ties"]
        return dict(zip(names, s.tolist()))  # (nn_ties stays 0 unless set_diag_ties(True) was called before segmentize)

    def volumes(self):
        v = np.zeros(self.mesh.n_cells)
        lib().orc_volumes(self._h, v)
        return v

    def free_segments(self):
        lib().orc_seg_free(self._h)

    def __del__(self):
        try:
            lib().orc_tg_destroy(self._h)
        except Exception:
            pass
