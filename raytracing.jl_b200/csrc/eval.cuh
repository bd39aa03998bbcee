// eval.cuh -- stage 2 of the two-stage fill pass: ONE THREAD PER SEGMENT (see topo.cuh for stage 1).
//
// Input: one 4-byte record per segment, already at the segment's final position (uid order, then walk order):
//      fast    (h << 2) | (exit1 << 1)      h = entry half-edge 3*cell + k_in, exit through edge k_in+1 (exit1) or k_in+2
//      literal (cell << 2) | 1              the Segment was written by the walk itself, nothing to do here
// Output: the Segment columns px, py, qx, qy, len, element (src/segment.jl:23-33), per-element sum(delta*len)
// (src/trackgenerator.jl:382), per-track sum(len) for the length check (src/track.jl:171-175).
//
// A segment's exit point q is the reference's intersection(track.ABC, general_form(edge)) (src/intersection.jl:11-18,
// 127-138) evaluated from the node coordinates in the cell's stored orientation; its entry point p is the exit point of the
// previous segment of the same track (the shared edge gives the same line up to an exact sign flip, and intersection() is
// invariant under that flip bit for bit), handed over by a warp shuffle.  A warp evaluates groups of 31 consecutive
// segments: lane 0 re-evaluates the exit of the segment just before the group only to hand it to lane 1.
//
// The kernel is a gather (record -> cell's node ids -> node coordinates; group -> track -> track line) followed by ~140
// FP64 instructions, six coalesced streaming stores and one RED per segment.  The node and cell tables (20 B/cell) are kept in
// L2 (evict_last) while records and Segment columns stream through it (evict_first); the gathers of the next three groups are
// in flight while a group is evaluated (see k_eval2).
#pragma once
#include "topo.cuh"

namespace rt {

// per track: the line and the signed effective spacing (one 32-byte sector per thread instead of five scattered loads)
struct __align__(32) TrackRec {
    double a, b, c;   // track.ABC
    double sdelta;    // delta_eff[azim] (1.0 when volumes are off), NEGATIVE when phi >= pi/2 (src/intersection.jl:153)
};

__global__ void k_track_recs(const __grid_constant__ EvalParams P, long long n_tracks, TrackRec *out) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_tracks) return;
    const int az = P.t.azim[t];
    const bool right = P.ang.phi[az] < kPi / 2;
    const double d = P.vol ? P.ang.delta_eff[az] : 1.0;
    TrackRec r;
    r.a = P.t.a[t];
    r.b = P.t.b[t];
    r.c = P.t.c[t];
    r.sdelta = right ? d : -d;
    out[t] = r;
}

constexpr int kEvalGroup = 31;  // segments per warp group (lane 0 is the hand-over lane)
constexpr int kEval2Threads = 256;
#ifndef RT_EVAL_MIN_BLOCKS
#define RT_EVAL_MIN_BLOCKS 2
#endif

// track of the first owned segment (31*g, batch relative) of every group: bisection of the offsets table
__global__ void k_eval_groups(const __grid_constant__ EvalParams P, long long n_groups, int *grp_track) {
    long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const long long s0 = g * kEvalGroup + P.offset_base;
    long long lo = P.trk_begin, hi = P.trk_end;  // invariant: offsets[lo] <= s0 < offsets[hi]
    while (hi - lo > 1) {
        long long mid = (lo + hi) >> 1;
        if (P.offsets[mid] <= s0)
            lo = mid;
        else
            hi = mid;
    }
    grp_track[g] = (int)(lo - P.trk_begin);
}

// the plain IEEE evaluation of one edge crossing (operands outside the range of the shared-reciprocal sequence)
__device__ __noinline__ bool eval_cold(const Line &trk, double2 a, double2 b, P2 &q) {
    return intersection(trk, general_form(P2{a.x, a.y}, P2{b.x, b.y}), q);
}

// pipeline registers of one group (31 segments, one per lane) on its way through the three gather levels
struct EvA {  // level 1 issued: the record and the group's first track
    int rec, t0;
};
struct EvB {  // level 2 issued: node ids of the exit edge, begin/end of the group's first track
    int rec, t0, na, nb;
    long long tbeg, tend;
};
struct EvC {  // level 3 issued: node coordinates, track line
    int rec, t;
    bool exists;
    double2 pa, pb;
    TrackRec tr;
};

// Persistent, software-pipelined: warp w evaluates groups w, w + W, w + 2W, ... (W = warps in the grid).  In every iteration it
// issues the level-1 loads of group k+3, the level-2 loads of group k+2, the level-3 loads of group k+1 -- each using operands
// that were requested one iteration earlier -- and then evaluates group k, so the latency of every gather level hides behind one
// full iteration of arithmetic of the resident warps.
__global__ void __launch_bounds__(kEval2Threads, RT_EVAL_MIN_BLOCKS)
    k_eval2(const __grid_constant__ EvalParams P, const TrackRec *__restrict__ trk_recs, const int *__restrict__ grp_track,
            long long n_groups) {
    const unsigned FULL = 0xffffffffu;
    const DevMesh &m = P.m;
    const int lane = threadIdx.x & 31;
    const unsigned long long pol_keep = l2_policy_keep(), pol_stream = l2_policy_stream();
    const long long W = (long long)gridDim.x * (kEval2Threads / 32);
    const long long gw = blockIdx.x * (long long)(kEval2Threads / 32) + (threadIdx.x >> 5);

    auto seg_of = [&](long long g) { return g * kEvalGroup + lane - 1; };  // lane 0: the segment before the group
    auto issueA = [&](long long g) {
        EvA a;
        g = g < n_groups ? g : n_groups - 1;
        const long long s = seg_of(g);
        const long long sc = s < 0 ? 0 : (s >= P.n_seg ? P.n_seg - 1 : s);
        a.rec = ldg_i32_pol(P.rec + sc, pol_stream);
        a.t0 = __ldg(grp_track + g);
        return a;
    };
    auto issueB = [&](const EvA &a) {
        EvB b;
        b.rec = a.rec;
        b.t0 = a.t0;
        const int r = a.rec;
        const bool fastrec = (r & 1) == 0;
        const int h = r >> 2;
        const int cell = fastrec ? h / 3 : h;
        const int kin = fastrec ? h - 3 * cell : 0;
        int kout = kin + ((r & 2) ? 1 : 2);
        kout = kout >= 3 ? kout - 3 : kout;
        const int kn = kout == 2 ? 0 : kout + 1;
        const int *cn = m.cell_nodes + 3 * (long long)cell;
        b.na = ldg_i32_pol(cn + kout, pol_keep);
        b.nb = ldg_i32_pol(cn + kn, pol_keep);
        const long long *off = P.offsets + P.trk_begin + a.t0;
        b.tbeg = __ldg(off);
        b.tend = __ldg(off + 1);
        return b;
    };
    auto issueC = [&](const EvB &b, long long g) {
        EvC c;
        c.rec = b.rec;
        const long long s = seg_of(g);
        c.exists = g < n_groups && s >= 0 && s < P.n_seg;
        int t = b.t0;
        if (lane > 0 && c.exists) {  // lanes beyond the end of the group's first track (rare) walk the offsets table
            long long tend = b.tend - P.offset_base;
            while (tend <= s) {
                ++t;
                tend = __ldg(P.offsets + P.trk_begin + t + 1) - P.offset_base;
            }
        }
        // lane 0 only matters when its segment belongs to the same track as lane 1's
        if (lane == 0 && s < b.tbeg - P.offset_base) c.exists = false;
        c.t = t;
        c.pa = ldg_f64x2_pol(m.xy + b.na, pol_keep);
        c.pb = ldg_f64x2_pol(m.xy + b.nb, pol_keep);
        const double *tp = reinterpret_cast<const double *>(trk_recs + (P.trk_begin + t));
        ldg256_keep(tp, pol_keep, c.tr.a, c.tr.b, c.tr.c, c.tr.sdelta);
        return c;
    };

    if (gw >= n_groups) return;  // whole warp
    // Two register sets used alternately (the loop is unrolled by two): copying a set at the end of an iteration would wait for
    // the loads that were just issued into it.
    struct EvSet {
        EvA a;
        EvB b;
        EvC c;
    };
    EvSet X, Y;
    // prologue: fill the pipeline
    X.a = issueA(gw + 2 * W);
    X.b = issueB(issueA(gw + W));
    X.c = issueC(issueB(issueA(gw)), gw);

    auto step = [&](long long g, const EvSet &in, EvSet &out) {
        if (g >= n_groups) return;  // whole warp
        // ---- issue the gathers of the following groups (operands requested one iteration ago)
        out.a = issueA(g + 3 * W);
        out.b = issueB(in.a);
        out.c = issueC(in.b, g + W);
        const EvC &c = in.c;
        // ---- evaluate group g
        const int r = c.rec;
        const bool fast = (r & 1) == 0 && c.exists;
        const bool live = c.exists && lane > 0;
        const int h = r >> 2;
        const int cell = ((r & 1) == 0) ? h / 3 : h;
        const int kin = ((r & 1) == 0) ? h - 3 * cell : 0;
        const Line trk{c.tr.a, c.tr.b, c.tr.c};
        const bool right = c.tr.sdelta > 0.0;
        P2 q;
        bool ok = true;
        bool par_out = intersection_try(trk, general_form_try(P2{c.pa.x, c.pa.y}, P2{c.pb.x, c.pb.y}, ok), q, ok);
        if (!ok) par_out = eval_cold(trk, c.pa, c.pb, q);
        // hand-over: lane i-1's exit is my entry iff it is the previous segment of the same track and a fast record (consecutive
        // fast records of one track are always edge-adjacent)
        const double upx = __shfl_up_sync(FULL, q.x, 1), upy = __shfl_up_sync(FULL, q.y, 1);
        const int up_key = __shfl_up_sync(FULL, fast ? c.t : -1, 1);
        P2 p{upx, upy};
        const bool own_entry = fast && live && up_key != c.t;
        if (own_entry) {  // first segment of a track, or the previous record is literal
            const int *cn = m.cell_nodes + 3 * (long long)cell;
            const int ki2 = kin == 2 ? 0 : kin + 1;
            const double2 ea = ldg_f64x2_pol(m.xy + ldg_i32_pol(cn + kin, pol_keep), pol_keep);
            const double2 eb = ldg_f64x2_pol(m.xy + ldg_i32_pol(cn + ki2, pol_keep), pol_keep);
            eval_cold(trk, ea, eb, p);  // never parallel: the previous chord ended here
        }
        double l = 0.0;
        if (fast && live) {
            l = norm2(p.x - q.x, p.y - q.y);  // Segment(p, q): norm(p - q), src/segment.jl:32
            const long long so = seg_of(g);
            stg_f64_pol(P.opx + so, p.x, pol_stream);
            stg_f64_pol(P.opy + so, p.y, pol_stream);
            stg_f64_pol(P.oqx + so, q.x, pol_stream);
            stg_f64_pol(P.oqy + so, q.y, pol_stream);
            stg_f64_pol(P.olen + so, l, pol_stream);
            stg_i32_pol(P.oelem + so, cell + 1, pol_stream);
            if (P.vol) atomicAdd(&P.vol[cell], fabs(c.tr.sdelta) * l);  // volumes[i] += delta_s[a]*l, src/trackgenerator.jl:382
            // the geometric conditions of the sequential fast path (walk.cuh); k_topo's filters make them hold
            const bool exit1 = (r & 2) != 0;
            const bool kin_lt_kout = (kin == 0) || (kin == 1 && exit1);
            const bool in_first = right ? (kin_lt_kout ? (p.x < q.x) : !(q.x < p.x)) : (kin_lt_kout ? (p.x > q.x) : !(q.x > p.x));
            if (par_out || !in_first || !(l > P.lmin)) atomicExch(P.verify_fail, 1);
        }
        // per-track sum of lengths (decides the reference's length check up to a margin, see k_track_status)
        if (P.tsum) {
            const int t1 = __shfl_sync(FULL, c.t, 1);
            if (__all_sync(FULL, c.t == t1 || lane == 0)) {
                double v = l;
                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
                if (lane == 0 && v != 0.0) atomicAdd(&P.tsum[P.trk_begin + t1], v);
            } else if (fast && live) {
                atomicAdd(&P.tsum[P.trk_begin + c.t], l);
            }
        }
    };
    for (long long g = gw; g < n_groups; g += 2 * W) {
        step(g, X, Y);
        step(g + W, Y, X);
    }
}

}  // namespace rt
