// walk.cuh -- the segmentation walk: _segmentize_track! (src/track.jl:106-178) as a count pass and a
// fill pass over (track -> thread).  Two ways to take a step, producing identical results:
//
//   LITERAL  re-locate xp = q + tiny*(cos phi, sin phi) exactly like the reference (find_element through the
//            nearest node, tolerant barycentric test, inboundary, intersections over all three edges);
//   FAST     when the exit point of the current cell is provably generic ("fast-path
//            equivalence"), the reference's next accepted cell is the neighbour across the exit edge and its
//            chord is (shared-edge hit, hit on the one other crossed edge); only that one line/line
//            intersection is evaluated, with the reference's own formula so p, q, len are bit-identical.
//
// Threads of a warp alternate between a batch of FAST transitions and one LITERAL run-until-push, so the
// rare literal steps of different tracks execute together instead of serialising the warp.
#pragma once
#include "mesh_dev.cuh"

namespace rt {

struct AngleTabs {
    const double *phi, *sinp, *cosp, *delta_eff;  // per azimuthal index (0-based), n_azim_2 entries
};

struct TrackSoA {
    double *px, *py, *qx, *qy, *len, *a, *b, *c;
    int *azim;  // 0-based azimuthal index
    long long *track_idx;
    signed char *bc_fwd, *bc_bwd, *dir_fwd, *dir_bwd;
    long long *next_fwd, *next_bwd;
};

struct WalkParams {
    DevMesh m;
    long long n_tracks;   // tracks handled by this launch
    long long trk_begin;  // first track (shard-local index) of this launch
    TrackSoA t;
    AngleTabs ang;
    double tiny, rtol, lmin;
    int k, max_iter;
    unsigned flags;
    // count pass
    int *count;
    int *status;
    // fill pass
    const long long *offsets;  // shard-local exclusive scan of count
    long long offset_base;
    double *opx, *opy, *oqx, *oqy, *olen;
    int *oelem;
    double *vol;                   // unnormalised sum(delta_eff*len) per element, or nullptr
    unsigned long long *counters;  // [fast transitions, literal iterations, nn queries, knn queries] or nullptr
};

enum { MODE_FAST = 0, MODE_SLOW = 1, MODE_DONE = 2 };
constexpr int kFastBatch = 16;
constexpr long long kRunaway = 4000000;

template <bool FILL>
__global__ void __launch_bounds__(128) k_walk(const __grid_constant__ WalkParams P) {
    const unsigned FULL = 0xffffffffu;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int mode = MODE_DONE;
    const DevMesh &m = P.m;

    // track state
    Line trk{0, 0, 0};
    double tlen = 0, sx = 0, sy = 0, g = 0, delta = 0;
    bool right = true;
    double xpx = 0, xpy = 0;  // literal walk position
    int prev = -1, nseg = 0, status = 0;
    double sum = 0.0;
    long long out = 0, t = 0;
    // fast state: current cell, its exit edge, exit point, clearance bookkeeping
    int cur = -1, kexit = -1;
    double qx = 0, qy = 0, clearA = INFINITY;
    bool clean = false;
    long long slow_iters = 0;
    unsigned long long cnt[4] = {0, 0, 0, 0};
    const bool literal_only = (P.flags & 1u) != 0;

    if (i < P.n_tracks) {
        t = P.trk_begin + i;
        int az = P.t.azim[t];
        trk.a = P.t.a[t];
        trk.b = P.t.b[t];
        trk.c = P.t.c[t];
        tlen = P.t.len[t];
        double phi = P.ang.phi[az];
        right = phi < kPi / 2;  // isless(phi, pi/2), src/intersection.jl:153
        sx = P.tiny * P.ang.cosp[az];  // advance_step: x + step*Point2D(cos phi, sin phi), src/point.jl:43
        sy = P.tiny * P.ang.sinp[az];
        delta = P.vol ? P.ang.delta_eff[az] : 0.0;
        g = sqrt(trk.a * trk.a + trk.b * trk.b);
        xpx = P.t.px[t] + sx;  // src/track.jl:114
        xpy = P.t.py[t] + sy;
        if (FILL) out = P.offsets[t] - P.offset_base;
        mode = MODE_SLOW;
    }

    auto push = [&](int e, double ax, double ay, double bx, double by, double l) {
        if (FILL) {
            long long o = out + nseg;
            P.opx[o] = ax;
            P.opy[o] = ay;
            P.oqx[o] = bx;
            P.oqy[o] = by;
            P.olen[o] = l;
            P.oelem[o] = e + 1;
            if (P.vol) atomicAdd(&P.vol[e], delta * l);  // volumes[i] += delta_s[a]*l, src/trackgenerator.jl:382
        }
        sum += l;
        nseg += 1;
    };

    while (__any_sync(FULL, mode != MODE_DONE)) {
        // ------------------------------------------------------------------ FAST phase
#pragma unroll 1
        for (int it = 0; it < kFastBatch; ++it) {
            if (!__any_sync(FULL, mode == MODE_FAST)) break;
            if (mode != MODE_FAST) continue;
            if (nseg >= P.max_iter) {  // while i < MAX_ITER, src/track.jl:119
                mode = MODE_DONE;
                continue;
            }
            bool ok = false;
            int B = m.cells[cur].nbr[kexit];
            if (B >= 0 && clean) {
                const CellRec rb = m.cells[B];
                double s0 = trk.a * rb.vx[0] + trk.b * rb.vy[0] + trk.c;
                double s1 = trk.a * rb.vx[1] + trk.b * rb.vy[1] + trk.c;
                double s2 = trk.a * rb.vx[2] + trk.b * rb.vy[2] + trk.c;
                double clearB = (double)rb.clear;
                double thr = g * fmax(clearA, clearB);
                bool clear_ok = (fabs(s0) >= thr) && (fabs(s1) >= thr) && (fabs(s2) >= thr);
                // edges k = (k, k+1): crossed iff the end points lie on opposite sides of the track line
                bool c0 = (s0 > 0) != (s1 > 0), c1 = (s1 > 0) != (s2 > 0), c2 = (s2 > 0) != (s0 > 0);
                int kin = (rb.nbr[0] == cur) ? 0 : ((rb.nbr[1] == cur) ? 1 : ((rb.nbr[2] == cur) ? 2 : -1));
                int ncross = (int)c0 + (int)c1 + (int)c2;
                if (clear_ok && ncross == 2 && kin >= 0) {
                    bool cin = kin == 0 ? c0 : (kin == 1 ? c1 : c2);
                    int kout = (c0 && kin != 0) ? 0 : ((c1 && kin != 1) ? 1 : 2);
                    if (cin) {
                        const EdgeRec e = m.edges[3 * B + kout];
                        Line L{e.a, e.b, e.c};
                        P2 X;
                        bool par = intersection(trk, L, X);  // same formula as src/intersection.jl:127-138
                        if (!par) {
                            P2 Xin{qx, qy};
                            // int_points are stored in edge order; order_intersection_points picks the first
                            bool in_first = kin < kout ? order_first(right, Xin, X) : !order_first(right, X, Xin);
                            double l = norm2(qx - X.x, qy - X.y);  // Segment(p, q): norm(p - q), src/segment.jl:32
                            if (in_first && l > P.lmin) {
                                push(B, qx, qy, X.x, X.y, l);
                                cur = B;
                                kexit = kout;
                                qx = X.x;
                                qy = X.y;
                                clearA = clearB;
                                ok = true;
                                cnt[0]++;
                            }
                        }
                    }
                }
            }
            if (!ok) {
                // fall back to the literal walk from xp = advance_step(q, tiny, phi), src/track.jl:165-166
                xpx = qx + sx;
                xpy = qy + sy;
                prev = cur;
                mode = MODE_SLOW;
            }
        }
        // ------------------------------------------------------------------ LITERAL phase (until one push)
        if (mode == MODE_SLOW) {
            while (true) {
                if (nseg >= P.max_iter) {
                    mode = MODE_DONE;
                    break;
                }
                if (++slow_iters > kRunaway) {
                    status = 3;
                    mode = MODE_DONE;
                    break;
                }
                cnt[1]++;
                // find_element's result is discarded on boundary steps (src/track.jl:122-134): test the boundary first
                if (inboundary(m, xpx, xpy, P.tiny)) {
                    if (nseg == 0) {
                        xpx = xpx + sx;
                        xpy = xpy + sy;
                        continue;
                    }
                    mode = MODE_DONE;
                    break;
                }
                int e = find_element(m, xpx, xpy, 2, &cnt[2]);
                if (e < 0) {
                    e = find_element(m, xpx, xpy, P.k, &cnt[2]);
                    if (e < 0) {
                        status = 1;  // "Try increasing `k`", src/track.jl:141
                        mode = MODE_DONE;
                        break;
                    }
                }
                if (e == prev) {
                    xpx = xpx + sx;
                    xpy = xpy + sy;
                    continue;
                }
                P2 p, q;
                int e_p, e_q;
                int rc = intersections(m, e, trk, right, p, q, e_p, e_q);
                if (rc) {
                    status = rc;
                    mode = MODE_DONE;
                    break;
                }
                if (isapprox_pt(p, q)) {  // src/track.jl:156-159
                    xpx = xpx + sx;
                    xpy = xpy + sy;
                    continue;
                }
                push(e, p.x, p.y, q.x, q.y, norm2(p.x - q.x, p.y - q.y));
                xpx = q.x + sx;
                xpy = q.y + sy;
                prev = e;
                if (literal_only || e_q < 0) break;  // stay literal
                // arm the fast path: current cell, exit edge, and whether its vertices are clear of the track
                const CellRec &rc_ = m.cells[e];
                cur = e;
                kexit = e_q;
                qx = q.x;
                qy = q.y;
                clearA = (double)rc_.clear;
                double thr = g * clearA;
                clean = (fabs(trk.a * rc_.vx[0] + trk.b * rc_.vy[0] + trk.c) >= thr) &&
                        (fabs(trk.a * rc_.vx[1] + trk.b * rc_.vy[1] + trk.c) >= thr) &&
                        (fabs(trk.a * rc_.vx[2] + trk.b * rc_.vy[2] + trk.c) >= thr);
                // even when this cell is not clean the neighbour test needs clean==true to pass, so only go FAST if so
                mode = clean ? MODE_FAST : MODE_SLOW;
                break;
            }
        }
    }

    if (i < P.n_tracks) {
        if (!FILL) {
            // sum(l.(segments)) ~ track.l, src/track.jl:171-175 (errors thrown earlier skip this check)
            if (status == 0 && !isapprox(tlen, sum, 0.0, P.rtol)) status = 2;
            P.count[t] = nseg;
            P.status[t] = status;
        }
    }
    if (P.counters) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            unsigned long long v = cnt[q];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
            if ((threadIdx.x & 31) == 0 && v) atomicAdd(&P.counters[q], v);
        }
    }
}

}  // namespace rt
