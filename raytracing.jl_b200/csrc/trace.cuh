// trace.cuh -- trace!(t) (src/trackgenerator.jl:134-280) and next_tracks (:282-348) as ONE kernel over
// (phi, track) pairs: thread <-> uid.  Per-angle trigonometry arrives in host-computed tables (sin/cos/tan/atan
// are not reproducible across libms); everything else is IEEE +,-,*,/,sqrt in the reference's order.
#pragma once
#include "walk.cuh"

namespace rt {

struct TraceParams {
    int n2, n4;
    const long long *nx, *ny, *base;  // per angle: n_tracks_x, n_tracks_y, uid offset (base[i] = #tracks of angles < i)
    const double *phi, *tanp, *dxe, *dye;
    int bcs[4];  // top, bottom, right, left
    double bbmin[2], bbmax[2];
    long long uid_begin;  // 1-based uid of the first track of the shard
    long long n;          // tracks in the shard
    TrackSoA t;
    double *len_only;             // if non-null: only write the track's COST here (shard planning): its length plus cost_w times
                                  // the reciprocal sines of the angles at which it leaves and reaches the bounding box
    double cost_w;
    unsigned long long *err;      // min over failing tracks of (uid << 4 | code)
};

enum { TRACE_E_NO_EXIT = 1, TRACE_E_NOT_ON_BOUNDARY = 2, TRACE_E_BC_MISMATCH = 3 };

// boundary_condition(x, sides, bcs)  src/boundary.jl:48-63 with sides from src/trackgenerator.jl:172-177
__device__ __forceinline__ int boundary_condition(const TraceParams &P, P2 x, int &bc) {
    const double *mn = P.bbmin, *mx = P.bbmax;
    P2 p1{mn[0], mn[1]}, p2{mn[0], mx[1]}, p3{mx[0], mx[1]}, p4{mx[0], mn[1]};
    if (point_in_segment(p2, p3, norm2(p2.x - p3.x, p2.y - p3.y), x))
        bc = P.bcs[0];
    else if (point_in_segment(p4, p1, norm2(p4.x - p1.x, p4.y - p1.y), x))
        bc = P.bcs[1];
    else if (point_in_segment(p3, p4, norm2(p3.x - p4.x, p3.y - p4.y), x))
        bc = P.bcs[2];
    else if (point_in_segment(p1, p2, norm2(p1.x - p2.x, p1.y - p2.y), x))
        bc = P.bcs[3];
    else
        return TRACE_E_NOT_ON_BOUNDARY;
    return 0;
}

__global__ void __launch_bounds__(256) k_trace(const __grid_constant__ TraceParams P) {
    long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (idx >= P.n) return;
    long long uid = P.uid_begin + idx;  // 1-based
    // angle i (1-based) with base[i-1] < uid <= base[i]
    int lo = 0, hi = P.n2;  // invariant: base[lo] < uid <= base[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (P.base[mid] < uid)
            lo = mid;
        else
            hi = mid;
    }
    int i = hi;  // 1-based angle
    long long j = uid - P.base[i - 1];
    long long nx = P.nx[i - 1], ny = P.ny[i - 1], nt = nx + ny;
    bool right = i <= P.n4;  // points_right, src/azimuthal_quad.jl:62
    double dxe = P.dxe[i - 1], dye = P.dye[i - 1];
    double Dx = P.bbmax[0] - P.bbmin[0], Dy = P.bbmax[1] - P.bbmin[1];
    double px, py, qx, qy;
    if (j <= nx) {  // src/trackgenerator.jl:188-200
        px = right ? dxe * ((double)(nx - j) + 1.0 / 2) : dxe * ((double)j - 1.0 / 2);
        py = 0.0;
    } else {
        px = right ? 0.0 : Dx;
        py = dye * ((double)(j - nx) - 1.0 / 2);
    }
    double mm = P.tanp[i - 1];  // m = tan(phi), :203
    qx = px - (py - Dy) / mm;
    qy = Dy;
    int code = 0;
    if (!(0 <= qx && qx <= Dx)) {
        if (right) {
            qx = Dx;
            qy = py + mm * (Dx - px);
        } else {
            qx = 0.0;
            qy = py - mm * px;
        }
        if (!(0 <= qy && qy <= Dy)) code = TRACE_E_NO_EXIT;
    }
    px += P.bbmin[0];
    py += P.bbmin[1];
    qx += P.bbmin[0];
    qy += P.bbmin[1];
    double len = norm2(px - qx, py - qy);
    if (P.len_only) {
        // Shard planning.  A track costs its segments (~ its length) plus what the boundary band costs at both ends: within one
        // cell of the bounding box every transition is examined exactly on the slow side of the walk, and a track that meets the
        // boundary at an angle theta stays in that band for ~1/sin(theta) cells ("chunk layout and the boundary band").
        const double t = fabs(mm), c = 1.0 / sqrt(1.0 + t * t), sn = t * c;  // |cos phi|, sin phi
        const double s_in = (j <= nx) ? sn : c;            // starts on the bottom side / on a lateral side
        const double s_out = (qy == Dy + P.bbmin[1]) ? sn : c;  // ends on the top side / on a lateral side
        P.len_only[idx] = len + P.cost_w * (1.0 / fmax(s_in, 1e-6) + 1.0 / fmax(s_out, 1e-6));
        return;
    }
    Line abc = general_form(P2{px, py}, P2{qx, qy});
    int bcf = 0, bcb = 0;
    if (!code) code = boundary_condition(P, P2{qx, qy}, bcf);
    if (!code) code = boundary_condition(P, P2{px, py}, bcb);
    int bcf1, bcb1;  // the index rule, :235-241
    if (right) {
        bcf1 = j <= ny ? P.bcs[2] : P.bcs[0];
        bcb1 = j <= nx ? P.bcs[1] : P.bcs[3];
    } else {
        bcf1 = j <= ny ? P.bcs[3] : P.bcs[0];
        bcb1 = j <= nx ? P.bcs[1] : P.bcs[2];
    }
    if (!code && (bcf != bcf1 || bcb != bcb1)) code = TRACE_E_BC_MISMATCH;
    if (code) {
        atomicMin(P.err, ((unsigned long long)uid << 4) | (unsigned long long)code);
        bcf = bcf1;
        bcb = bcb1;
    }
    int df = (j <= ny) ? 0 : (bcf == 2 ? 0 : 1);  // :247-255
    int db = (j <= nx) ? (bcb == 2 ? 1 : 0) : 1;  // :257-265
    // next_track_fwd / next_track_bwd, :294-348 ; k = supplementary angle
    int k = P.n2 - i + 1;
    long long ai, aj;
    if (j <= ny) {
        ai = (bcf == 2) ? i : k;
        aj = j + nx;
    } else if (bcf == 2) {
        ai = i;
        aj = j - ny;
    } else {
        ai = k;
        aj = nt + ny - j + 1;
    }
    long long nfwd = P.base[ai - 1] + aj;
    if (j <= nx) {
        if (bcb == 2) {
            ai = i;
            aj = j + ny;
        } else {
            ai = k;
            aj = nx - j + 1;
        }
    } else {
        ai = (bcb == 2) ? i : k;
        aj = j - nx;
    }
    long long nbwd = P.base[ai - 1] + aj;

    const TrackSoA &t = P.t;
    t.px[idx] = px;
    t.py[idx] = py;
    t.qx[idx] = qx;
    t.qy[idx] = qy;
    t.len[idx] = len;
    t.a[idx] = abc.a;
    t.b[idx] = abc.b;
    t.c[idx] = abc.c;
    t.azim[idx] = i - 1;
    t.track_idx[idx] = j;
    t.bc_fwd[idx] = (signed char)bcf;
    t.bc_bwd[idx] = (signed char)bcb;
    t.dir_fwd[idx] = (signed char)df;
    t.dir_bwd[idx] = (signed char)db;
    t.next_fwd[idx] = nfwd;
    t.next_bwd[idx] = nbwd;
}

}  // namespace rt
