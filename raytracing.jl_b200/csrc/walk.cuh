// walk.cuh -- the segmentation walk: _segmentize_track! (src/track.jl:106-178) as a count pass and a
// fill pass.  Two ways to take a step, producing identical results:
//
//   LITERAL  re-locate xp = q + tiny*(cos phi, sin phi) exactly like the reference (find_element through the
//            nearest node, tolerant barycentric test, inboundary, intersections over all three edges);
//   FAST     when the exit point of the current cell is provably generic ("fast-path
//            equivalence"), the reference's next accepted cell is the neighbour across the exit edge and its
//            chord is (shared-edge hit, hit on the one other crossed edge); only that one line/line
//            intersection is evaluated, with the reference's own formula so p, q, len are bit-identical.
//
// Parallel decomposition: a track is cut into CHUNKS at seed points x_j = p + (j/n)*len*(cos, sin).  A seed is
// valid when x_j lies in a cell C whose chord is generic; the state "C has just been pushed" is then history
// independent (intersections() only depends on the cell and the track), so walker j starts from it while walker
// j-1 stops at its first push of C ("chunk hand-off").  Invalid seeds are void: the previous walker
// simply continues.  One thread per (track, chunk); a warp holds the same chunk index of 32 consecutive tracks,
// i.e. 32 adjacent parallel rays marching through the same cells.  Threads of a warp alternate between a batch
// of FAST transitions and one LITERAL run-until-push, so the rare literal steps execute together.
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

// chunk bookkeeping; per-chunk arrays are indexed cidx = unit*32 + lane, consecutive chunks of a track are 32 apart
struct ChunkPlan {
    int *nch;               // per track: number of chunks (>= 1)
    int *unit_block;        // per warp-unit: which block of 32 consecutive tracks
    long long *unit_base;   // per block: first unit; [n_blocks] = n_units
    long long n_units;
    int *seed_cell;         // >= 0: valid seed (cell already pushed by the previous walker); -1: void
    int *seed_kexit;
    double *seed_qx, *seed_qy;
    int *count;             // segments pushed by this chunk (count pass; final after fix-up)
    double *sum;            // sum of their lengths
    int *endcode;           // END_* | status << 8
    int *prefix;            // exclusive prefix of final counts inside the track
};

enum { END_HANDOFF = 0, END_TRACK = 1, END_ERROR = 2, END_CAP = 3 };

struct WalkParams {
    DevMesh m;
    long long n_tracks;     // tracks of the shard
    long long trk_begin, trk_end;  // fill pass: only tracks in [trk_begin, trk_end) (batching); count pass: all
    long long unit_begin, unit_end;
    TrackSoA t;
    AngleTabs ang;
    ChunkPlan ch;
    double tiny, rtol, lmin;
    int k, max_iter;
    unsigned flags;
    int *count;   // per track (fix-up output)
    int *status;  // per track
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
constexpr int kMaxChunksPerTrack = 4096;

// ---- chunk planning -----------------------------------------------------------------------------------
__global__ void k_plan_chunks(long long n_tracks, const double *len, double chunk_len, int *nch, int *blk_chunks) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int n = 0;
    if (t < n_tracks) {
        double r = ceil(len[t] / chunk_len);
        n = (r >= 1.0) ? (r > (double)kMaxChunksPerTrack ? kMaxChunksPerTrack : (int)r) : 1;
        nch[t] = n;
    }
    int mx = __reduce_max_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0 && t < n_tracks) blk_chunks[t >> 5] = mx;
}

__global__ void k_fill_units(long long n_blocks, const long long *unit_base, int *unit_block) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    for (long long u = unit_base[b]; u < unit_base[b + 1]; ++u) unit_block[u] = (int)b;
}

struct TrackCtx {
    Line trk;
    double g, sx, sy, tlen, delta;
    bool right;
};

__device__ __forceinline__ bool cell_clean(const CellRec &r, const Line &trk, double g, double clear) {
    double thr = g * clear;
    return (fabs(trk.a * r.vx[0] + trk.b * r.vy[0] + trk.c) >= thr) && (fabs(trk.a * r.vx[1] + trk.b * r.vy[1] + trk.c) >= thr) &&
           (fabs(trk.a * r.vx[2] + trk.b * r.vy[2] + trk.c) >= thr);
}

// distance of a point from the nearest bounding-box line
__device__ __forceinline__ double bbox_dist(const DevMesh &m, double x, double y) {
    return fmin(fmin(fabs(x - m.bbmin[0]), fabs(x - m.bbmax[0])), fmin(fabs(y - m.bbmin[1]), fabs(y - m.bbmax[1])));
}

// ---- seeds: one thread per chunk j >= 1 ------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_seed(const __grid_constant__ WalkParams P) {
    long long gw = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    long long unit = P.unit_begin + gw;
    if (unit >= P.unit_end) return;
    const DevMesh &m = P.m;
    int b = P.ch.unit_block[unit];
    int j = (int)(unit - P.ch.unit_base[b]);
    long long t = 32LL * b + lane;
    long long cidx = unit * 32 + lane;
    if (t >= P.n_tracks) return;
    int n = P.ch.nch[t];
    if (j >= n) return;
    P.ch.seed_cell[cidx] = -1;
    if (j == 0) return;
    int az = P.t.azim[t];
    Line trk{P.t.a[t], P.t.b[t], P.t.c[t]};
    double s = P.t.len[t] * ((double)j / (double)n);
    double x = P.t.px[t] + s * P.ang.cosp[az], y = P.t.py[t] + s * P.ang.sinp[az];
    bool right = P.ang.phi[az] < kPi / 2;
    int c = find_element(m, x, y, 2, nullptr);
    if (c < 0) return;
    const CellRec &r = m.cells[c];
    double clear = (double)r.clear;
    if (!(clear >= 0.0) || !isfinite(clear)) return;  // boundary-band or degenerate cells never seed
    double g = sqrt(trk.a * trk.a + trk.b * trk.b);
    if (!cell_clean(r, trk, g, clear)) return;
    // the seed point itself must be well inside (not merely tolerantly inside) the cell
    {
        double x1 = r.vx[0], y1 = r.vy[0], x2 = r.vx[1], y2 = r.vy[1], x3 = r.vx[2], y3 = r.vy[2];
        double d = x1 * (y2 - y3) + y1 * (x3 - x2) + (x2 * y3 - y2 * x3);
        double l1 = ((y2 - y3) * x + (x3 - x2) * y + (x2 * y3 - x3 * y2)) / d;
        double l2 = ((y3 - y1) * x + (x1 - x3) * y + (x3 * y1 - x1 * y3)) / d;
        double l3 = ((y1 - y2) * x + (x2 - x1) * y + (x1 * y2 - x2 * y1)) / d;
        const double mrg = 1e-6;
        if (!(l1 > mrg && l2 > mrg && l3 > mrg)) return;
    }
    P2 p, q;
    int e_p, e_q;
    if (intersections(m, c, trk, right, p, q, e_p, e_q) != 0) return;
    if (e_q < 0 || e_p < 0) return;
    double l = norm2(p.x - q.x, p.y - q.y);
    if (!(l > P.lmin)) return;
    P.ch.seed_cell[cidx] = c;
    P.ch.seed_kexit[cidx] = e_q;
    P.ch.seed_qx[cidx] = q.x;
    P.ch.seed_qy[cidx] = q.y;
}

// ---- 256-bit read-only loads / 256-bit stores (LDG.E.ENL2.256 / STG.E.ENL2.256 on sm_100a) with L2 eviction policies:
// the half-edge table is re-read by every track that crosses a cell (keep: evict_last), the segment records are
// written once and never read by this kernel (stream: evict_first), so the output does not flush the mesh out of L2.
__device__ __forceinline__ unsigned long long l2_policy_keep() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_stream() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void ldg256_keep(const void *p, unsigned long long pol, double &a, double &b, double &c, double &d) {
    asm volatile("ld.global.nc.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p), "l"(pol));
}
__device__ __forceinline__ void stg256_stream(void *p, unsigned long long pol, double a, double b, double c, double d) {
    asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg128_stream(void *p, unsigned long long pol, int a, int b, int c, int d) {
    asm volatile("st.global.L2::cache_hint.v4.s32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "l"(pol) : "memory");
}

constexpr int kWalkThreads = 128;

// ---- the walk ----------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(kWalkThreads) k_walk(const __grid_constant__ WalkParams P) {
    const unsigned FULL = 0xffffffffu;
    const DevMesh &m = P.m;
    // fill pass: every thread stages 4 consecutive segments and writes them as full, aligned 32-byte sectors
    __shared__ double s_buf[FILL ? 5 * 4 * kWalkThreads : 1];
    __shared__ int s_el[FILL ? 4 * kWalkThreads : 1];
    const int tid = threadIdx.x;
    long long gw = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    long long unit = P.unit_begin + gw;
    if (unit >= P.unit_end) return;  // whole warp
    int blk = P.ch.unit_block[unit];
    int j = (int)(unit - P.ch.unit_base[blk]);
    long long t = 32LL * blk + lane;
    long long cidx = unit * 32 + lane;

    int mode = MODE_DONE;
    Line trk{0, 0, 0};
    double sx = 0, sy = 0, g = 0, delta = 0;
    bool right = true;
    double xpx = 0, xpy = 0;  // literal walk position
    int prev = -1, nseg = 0, status = 0, endcode = END_TRACK;
    double sum = 0.0;
    long long out = 0;
    // fast state: last pushed cell, the half-edge through which the next cell is entered (-1: boundary), the end
    // points (u, w) of that edge in the half-edge's order with their signed distances (sp, sq) from the track line,
    // and the exit point
    int cur = -1, hB = -1;
    double ux = 0, uy = 0, wx = 0, wy = 0;
    double sp = 0, sq = 0, qx = 0, qy = 0, clearA = INFINITY;
    bool clean = false;
    int stop_cell = -1;  // count pass: hand-off cell of the next valid chunk
    int limit = P.max_iter;
    long long slow_iters = 0;
    unsigned long long cnt[4] = {0, 0, 0, 0};
    const bool literal_only = (P.flags & 1u) != 0;
    bool active = false;
    const unsigned long long pol_keep = l2_policy_keep();
    const unsigned long long pol_stream = FILL ? l2_policy_stream() : 0ull;

    // arm the fast path after cell e was pushed with exit edge e_q (local index) and exit point (ex, ey)
    auto arm = [&](int e, int e_q, double ex, double ey) {
        const CellRec &r = m.cells[e];
        cur = e;
        qx = ex;
        qy = ey;
        clearA = fabs((double)r.clear);
        double s[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) s[q] = trk.a * r.vx[q] + trk.b * r.vy[q] + trk.c;
        double thr = g * clearA;
        int e_n = e_q == 2 ? 0 : e_q + 1;
        double sa = s[e_q], sb = s[e_n];
        clean = (fabs(s[0]) >= thr) && (fabs(s[1]) >= thr) && (fabs(s[2]) >= thr) && ((sa > 0) != (sb > 0));
        int enc = m.twin[3 * e + e_q];
        hB = enc < 0 ? -1 : (enc >> 1);
        bool flip = (enc & 1) != 0;
        sp = flip ? sb : sa;
        sq = flip ? sa : sb;
        ux = flip ? r.vx[e_n] : r.vx[e_q];
        uy = flip ? r.vy[e_n] : r.vy[e_q];
        wx = flip ? r.vx[e_q] : r.vx[e_n];
        wy = flip ? r.vy[e_q] : r.vy[e_n];
    };

    if (t < P.n_tracks && t >= P.trk_begin && t < P.trk_end) {
        int n = P.ch.nch[t];
        int seed = (j < n) ? (j == 0 ? -2 : P.ch.seed_cell[cidx]) : -1;
        active = (seed != -1);
        if (FILL && active) {
            limit = P.ch.count[cidx];
            active = limit > 0;
        }
        if (active) {
            int az = P.t.azim[t];
            trk.a = P.t.a[t];
            trk.b = P.t.b[t];
            trk.c = P.t.c[t];
            double phi = P.ang.phi[az];
            right = phi < kPi / 2;         // isless(phi, pi/2), src/intersection.jl:153
            sx = P.tiny * P.ang.cosp[az];  // advance_step: x + step*Point2D(cos phi, sin phi), src/point.jl:43
            sy = P.tiny * P.ang.sinp[az];
            delta = P.vol ? P.ang.delta_eff[az] : 0.0;
            g = sqrt(trk.a * trk.a + trk.b * trk.b);
            if (FILL) out = P.offsets[t] - P.offset_base + P.ch.prefix[cidx];
            if (j == 0) {
                xpx = P.t.px[t] + sx;  // src/track.jl:114
                xpy = P.t.py[t] + sy;
                mode = MODE_SLOW;
            } else {
                arm(seed, P.ch.seed_kexit[cidx], P.ch.seed_qx[cidx], P.ch.seed_qy[cidx]);  // k_seed verified `clean`
                prev = cur;
                xpx = qx + sx;
                xpy = qy + sy;
                mode = (literal_only || !clean) ? MODE_SLOW : MODE_FAST;
            }
            if (!FILL) {
                for (int jj = j + 1; jj < n; ++jj) {
                    int sc = P.ch.seed_cell[cidx + 32LL * (jj - j)];
                    if (sc >= 0) {
                        stop_cell = sc;
                        break;
                    }
                }
                if (j > 0 && stop_cell == cur) {  // next seed sits in the same cell: this chunk is empty
                    endcode = END_HANDOFF;
                    mode = MODE_DONE;
                }
                if (limit <= 0) {  // while i < MAX_ITER never runs
                    endcode = END_CAP;
                    mode = MODE_DONE;
                }
            }
        }
    }

    auto push = [&](int e, double ax, double ay, double bx, double by, double l) {
        if (FILL) {
            long long o = out + nseg;
            int k = (int)(o & 3);
            s_buf[(0 * 4 + k) * kWalkThreads + tid] = ax;
            s_buf[(1 * 4 + k) * kWalkThreads + tid] = ay;
            s_buf[(2 * 4 + k) * kWalkThreads + tid] = bx;
            s_buf[(3 * 4 + k) * kWalkThreads + tid] = by;
            s_buf[(4 * 4 + k) * kWalkThreads + tid] = l;
            s_el[k * kWalkThreads + tid] = e + 1;
            bool last = nseg + 1 >= limit;
            if (k == 3 || last) {
                long long g0 = o & ~3LL;
                int kf = (int)((g0 > out ? g0 : out) - g0);
                if (kf == 0 && k == 3) {
                    double *dst[5] = {P.opx, P.opy, P.oqx, P.oqy, P.olen};
#pragma unroll
                    for (int a = 0; a < 5; ++a)
                        stg256_stream(dst[a] + g0, pol_stream, s_buf[(a * 4 + 0) * kWalkThreads + tid], s_buf[(a * 4 + 1) * kWalkThreads + tid],
                                      s_buf[(a * 4 + 2) * kWalkThreads + tid], s_buf[(a * 4 + 3) * kWalkThreads + tid]);
                    stg128_stream(P.oelem + g0, pol_stream, s_el[0 * kWalkThreads + tid], s_el[1 * kWalkThreads + tid],
                                  s_el[2 * kWalkThreads + tid], s_el[3 * kWalkThreads + tid]);
                } else {
                    for (int kk = kf; kk <= k; ++kk) {
                        P.opx[g0 + kk] = s_buf[(0 * 4 + kk) * kWalkThreads + tid];
                        P.opy[g0 + kk] = s_buf[(1 * 4 + kk) * kWalkThreads + tid];
                        P.oqx[g0 + kk] = s_buf[(2 * 4 + kk) * kWalkThreads + tid];
                        P.oqy[g0 + kk] = s_buf[(3 * 4 + kk) * kWalkThreads + tid];
                        P.olen[g0 + kk] = s_buf[(4 * 4 + kk) * kWalkThreads + tid];
                        P.oelem[g0 + kk] = s_el[kk * kWalkThreads + tid];
                    }
                }
            }
        }
        if (P.vol) atomicAdd(&P.vol[e], delta * l);  // volumes[i] += delta_s[a]*l, src/trackgenerator.jl:382
        sum += l;
        nseg += 1;
        if (nseg >= limit) {  // while i < MAX_ITER (src/track.jl:119) / this chunk's final count in the fill pass
            endcode = END_CAP;
            mode = MODE_DONE;
        } else if (!FILL && e == stop_cell) {
            endcode = END_HANDOFF;
            mode = MODE_DONE;
        }
    };

    while (__any_sync(FULL, mode != MODE_DONE)) {
        // ------------------------------------------------------------------ FAST phase
#pragma unroll 1
        for (int it = 0; it < kFastBatch; ++it) {
            if (!__any_sync(FULL, mode == MODE_FAST)) break;
            if (mode != MODE_FAST) continue;
            bool ok = false;
            if (hB >= 0 && clean) {
                double ax, ay, w0, w1;
                ldg256_keep(m.he + hB, pol_keep, ax, ay, w0, w1);
                int tw1 = __double2loint(w0), tw2 = __double2hiint(w0);
                float clearf = __int_as_float(__double2loint(w1));
                double sa = trk.a * ax + trk.b * ay + trk.c;
                double clearB = fabs((double)clearf);
                double thr = g * fmax(clearA, clearB);
                bool clear_ok = (fabs(sa) >= thr) && (fabs(sp) >= thr) && (fabs(sq) >= thr);
                // the entry edge (v_k, v_k+1) = (u, w) is crossed (sp, sq of opposite sign); the exit is the other edge whose
                // non-apex end lies on the opposite side of the apex: edge k+1 = (w, apex) or k+2 = (apex, u)
                bool exit1 = (sa > 0) != (sq > 0);
                if (clear_ok) {
                    // general_form(P_i, P_j) of the exit edge in the cell's stored orientation, src/intersection.jl:11-18,57
                    P2 xi{exit1 ? wx : ax, exit1 ? wy : ay}, xo{exit1 ? ax : ux, exit1 ? ay : uy};
                    Line L = general_form(xi, xo);
                    P2 X;
                    bool par = intersection(trk, L, X);  // same formula as src/intersection.jl:127-138
                    if (!par) {
                        P2 Xin{qx, qy};
                        // int_points are stored in edge-index order; order_intersection_points picks the first
                        int cellB = hB / 3;
                        int kin = hB - 3 * cellB;
                        bool kin_lt_kout = (kin == 0) || (kin == 1 && exit1);
                        bool in_first = kin_lt_kout ? order_first(right, Xin, X) : !order_first(right, X, Xin);
                        double l = norm2(qx - X.x, qy - X.y);  // Segment(p, q): norm(p - q), src/segment.jl:32
                        bool accept = in_first && l > P.lmin;
                        // cells touching the bounding-box band: the re-location points must not be `inboundary`
                        if (accept && clearf < 0.0f) accept = bbox_dist(m, qx, qy) > 0.25 * l + 8.0 * P.tiny;
                        if (accept) {
                            double pxx = qx, pyy = qy;
                            int enc = exit1 ? tw1 : tw2;
                            double si = exit1 ? sq : sa, so = exit1 ? sa : sp;  // s at the exit edge's ordered end points
                            bool flip = (enc & 1) != 0;
                            sp = flip ? so : si;
                            sq = flip ? si : so;
                            ux = flip ? xo.x : xi.x;
                            uy = flip ? xo.y : xi.y;
                            wx = flip ? xi.x : xo.x;
                            wy = flip ? xi.y : xo.y;
                            hB = enc < 0 ? -1 : (enc >> 1);
                            cur = cellB;
                            qx = X.x;
                            qy = X.y;
                            clearA = clearB;
                            ok = true;
                            cnt[0]++;
                            push(cellB, pxx, pyy, X.x, X.y, l);
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
                if (++slow_iters > kRunaway) {
                    status = 3;
                    endcode = END_ERROR;
                    mode = MODE_DONE;
                    break;
                }
                cnt[1]++;
                // find_element's result is discarded on boundary steps (src/track.jl:122-134): test the boundary first
                if (inboundary(m, xpx, xpy, P.tiny)) {
                    if (nseg == 0 && j == 0) {
                        xpx = xpx + sx;
                        xpy = xpy + sy;
                        continue;
                    }
                    endcode = END_TRACK;
                    mode = MODE_DONE;
                    break;
                }
                int e = find_element(m, xpx, xpy, 2, &cnt[2]);
                if (e < 0) {
                    e = find_element(m, xpx, xpy, P.k, &cnt[2]);
                    if (e < 0) {
                        status = 1;  // "Try increasing `k`", src/track.jl:141
                        endcode = END_ERROR;
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
                    endcode = END_ERROR;
                    mode = MODE_DONE;
                    break;
                }
                if (isapprox_pt(p, q)) {  // src/track.jl:156-159
                    xpx = xpx + sx;
                    xpy = xpy + sy;
                    continue;
                }
                xpx = q.x + sx;
                xpy = q.y + sy;
                prev = e;
                mode = MODE_SLOW;
                push(e, p.x, p.y, q.x, q.y, norm2(p.x - q.x, p.y - q.y));
                if (mode == MODE_DONE || literal_only || e_q < 0) break;
                arm(e, e_q, q.x, q.y);
                mode = clean ? MODE_FAST : MODE_SLOW;
                break;
            }
        }
    }

    if (!FILL && t < P.n_tracks && j < P.ch.nch[t]) {
        P.ch.count[cidx] = active ? nseg : 0;
        P.ch.sum[cidx] = sum;
        P.ch.endcode[cidx] = active ? (endcode | (status << 8)) : (END_HANDOFF | (0 << 8));
    }
    if (P.counters) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            unsigned long long v = cnt[q];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
            if (lane == 0 && v) atomicAdd(&P.counters[q], v);
        }
    }
}

// ---- per-track fix-up: combine the chunks of a track exactly like one serial walk would have ended ----------
//  * a chunk that ended with an error / at the track end / at the MAX_ITER cap ends the track: later chunks are dropped
//  * the total is capped at max_iter (src/track.jl:119); the length check (src/track.jl:171-175) uses the chunk sums
__global__ void k_fixup_tracks(const __grid_constant__ WalkParams P) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= P.n_tracks) return;
    int n = P.ch.nch[t];
    long long blk = t >> 5;
    int lane = (int)(t & 31);
    long long c0 = P.ch.unit_base[blk] * 32 + lane;
    int total = 0, status = 0;
    double sum = 0.0;
    bool open = true, truncated = false;
    for (int j = 0; j < n; ++j) {
        long long cidx = c0 + 32LL * j;
        bool valid = (j == 0) || P.ch.seed_cell[cidx] >= 0;
        int cnt = (valid && open) ? P.ch.count[cidx] : 0;
        if (valid && open) {
            int ec = P.ch.endcode[cidx];
            int end = ec & 255, st = ec >> 8;
            if (total + cnt >= P.max_iter) {
                if (total + cnt > P.max_iter) truncated = true;
                cnt = P.max_iter - total;
                open = false;
            }
            sum += P.ch.sum[cidx];
            if (end == END_ERROR) {
                status = st;
                open = false;
            } else if (end != END_HANDOFF) {
                open = false;  // END_TRACK, or END_CAP inside one chunk
            }
        }
        P.ch.prefix[cidx] = total;
        P.ch.count[cidx] = cnt;
        total += cnt;
    }
    if (status == 0 && (truncated || !isapprox(P.t.len[t], sum, 0.0, P.rtol))) status = 2;
    P.count[t] = total;
    P.status[t] = status;
}

}  // namespace rt
