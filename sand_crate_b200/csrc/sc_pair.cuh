// sc_pair.cuh - the two fused pair kernels (density, force+integrate), templated on the arithmetic type.
//
//   k_density  (K4)  collision_detector.py:52-121 pair discovery fused with crate.py:161-175 (populate_colliders),
//                    261-275 (pressure) and 337-342 (surface normals)
//   k_force    (K5)  crate.py:343-353 tension, 309-310 gravity, 295-307 pressure, 316-323 viscosity,
//                    245-259 wall bounce, 177-200 continuous collision, 360-361 integration
//
// Both walk, for the particle at sorted index s, the three contiguous sorted ranges that cover the 3x3 cells
// around it and visit accepted neighbors in EXACTLY the reference's list order:
//   same row ascending from s+1, next row ascending, same row descending from s-1, previous row descending,
// first 20 only.  Every neighbor quantity read is a start-of-tick snapshot (Jacobi), so there are no atomics.
#pragma once
#include "sc_common.cuh"

namespace sc {

template <typename Real> struct Vec2;
template <> struct Vec2<double> { typedef double2 type; };
template <> struct Vec2<float> { typedef float2 type; };

__device__ inline double sc_sqrt(double v) { return sqrt(v); }
__device__ inline float sc_sqrt(float v) { return sqrtf(v); }

// Per-thread neighbor list kept in shared memory, one column per thread (conflict-free: slot k of thread t is
// word k * SC_BLOCK + t).
struct NbrList {
    uint32_t *col;
    __device__ __forceinline__ uint32_t get(int k) const { return col[k * SC_BLOCK]; }
    __device__ __forceinline__ void set(int k, uint32_t j) { col[k * SC_BLOCK] = j; }
};

// The reference accepts a pair iff sqrt(dx*dx + dy*dy) <= d in fp64 (collision_detector.py:77-79).  q = dx*dx+dy*dy
// is formed exactly as NumPy forms it; sqrt is correctly rounded and monotone, so outside a 2^-40 relative band
// around d*d the outcome is decided by q alone and the fp64 sqrt (about 30 instructions) is only evaluated inside
// the band - bit-identical decisions, including the lattice cases that sit exactly on q == d*d.
struct AcceptBand {
    double d, d2_lo, d2_hi;
    __device__ __forceinline__ explicit AcceptBand(double d_) : d(d_) {
        const double d2 = d_ * d_;
        d2_lo = d2 * (1.0 - 9.094947017729282e-13);  // 2^-40
        d2_hi = d2 * (1.0 + 9.094947017729282e-13);
    }
    __device__ __forceinline__ bool operator()(double dx, double dy) const {
        const double q = dx * dx + dy * dy;
        if (q > d2_hi) return false;
        if (q < d2_lo) return true;
        return in_band(q, d);  // also where NaN ends up: false, like the reference
    }
    // out of line on purpose: inlined, ptxas hoists the fp64 sqrt sequence in front of the two fast exits
    static __device__ __noinline__ bool in_band(double q, double d) { return sqrt(q) <= d; }
};

// Phase 1 of both pair kernels: collects the neighbors of sorted particle s in EXACTLY the reference's list order
// into `lst` and returns their number (<= 20).  Acceptance is evaluated in fp64 in both precision modes: the
// x-window from the LOWER-sorted particle (collision_detector.py:106-119) and the distance test above.
__device__ __forceinline__ int collect_neighbors(uint32_t s, double2 ps, uint32_t c, const Grid &g,
                                                 const uint32_t *__restrict__ cell_start,
                                                 const double2 *__restrict__ pos, NbrList lst) {
    const double d = g.d;
    const AcceptBand accept(d);
    int count = 0;
    const uint32_t a1 = cell_start[c + 2];
    const uint32_t cn = c + (uint32_t)g.ncols, cp = c - (uint32_t)g.ncols;
    const double xs_hi = ps.x + d, xs_lo = ps.x - d;
    // same row, ascending from s + 1
    for (uint32_t j = s + 1; j < a1 && count < SC_MAX_NEIGHBORS; ++j) {
        const double2 pj = pos[j];
        if (!(pj.x <= xs_hi)) break;  // sorted by x inside the row: nothing further can pass the window
        if (accept(pj.x - ps.x, pj.y - ps.y)) lst.set(count++, j);
    }
    // next row, ascending
    {
        const uint32_t b0 = cell_start[cn - 1], b1 = cell_start[cn + 2];
        for (uint32_t j = b0; j < b1 && count < SC_MAX_NEIGHBORS; ++j) {
            const double2 pj = pos[j];
            if (!(xs_lo <= pj.x && pj.x <= xs_hi)) continue;
            if (accept(pj.x - ps.x, pj.y - ps.y)) lst.set(count++, j);
        }
    }
    // same row, descending from s - 1: j is the lower-sorted one, so the window is evaluated from j
    {
        const uint32_t a0 = cell_start[c - 1];
        for (uint32_t j = s; j > a0 && count < SC_MAX_NEIGHBORS;) {
            --j;
            const double2 pj = pos[j];
            if (!(ps.x <= pj.x + d)) break;
            if (accept(ps.x - pj.x, ps.y - pj.y)) lst.set(count++, j);
        }
    }
    // previous row, descending
    {
        const uint32_t c0 = cell_start[cp - 1], c1 = cell_start[cp + 2];
        for (uint32_t j = c1; j > c0 && count < SC_MAX_NEIGHBORS;) {
            --j;
            const double2 pj = pos[j];
            if (!(pj.x - d <= ps.x && ps.x <= pj.x + d)) continue;
            if (accept(ps.x - pj.x, ps.y - pj.y)) lst.set(count++, j);
        }
    }
    return count;
}

// crate.py:167-174 for one directed pair (i <- j): unit vector from the (noised) neighbor to i and the weight
// w = 1 - clip(dist / d, 0, 1) (crate.py:270).
template <typename Real> struct PairGeom { Real nx, ny, w; };

template <typename Real>
__device__ __forceinline__ PairGeom<Real> pair_geom(const DevParams &P, double2 pi, double2 pj, uint32_t uid_i,
                                                    uint32_t uid_j, const double *__restrict__ host_noise,
                                                    uint32_t noise_index);

template <>
__device__ __forceinline__ PairGeom<double> pair_geom<double>(const DevParams &P, double2 pi, double2 pj,
                                                              uint32_t uid_i, uint32_t uid_j,
                                                              const double *__restrict__ host_noise,
                                                              uint32_t noise_index) {
    double qx = pj.x, qy = pj.y;
    if (P.noise_mode != SC_NOISE_NONE) {
        double ux, uy;
        if (P.noise_mode == SC_NOISE_HOST) {
            ux = host_noise[2 * (size_t)noise_index];
            uy = host_noise[2 * (size_t)noise_index + 1];
        } else {
            uint32_t hx, hy;
            pair_noise_bits(P.tick_key, uid_i, uid_j, hx, hy);
            ux = (double)hx * (1.0 / 4294967296.0);
            uy = (double)hy * (1.0 / 4294967296.0);
        }
        qx += (ux - 0.5) * P.d * P.level;
        qy += (uy - 0.5) * P.d * P.level;
    }
    const double rx = pi.x - qx, ry = pi.y - qy;
    const double dist = sqrt(rx * rx + ry * ry);
    PairGeom<double> g;
    g.nx = rx / dist;
    g.ny = ry / dist;
    double cl = dist / P.d;
    if (cl < 0) cl = 0;
    if (cl > 1) cl = 1;
    g.w = 1 - cl;
    return g;
}

template <>
__device__ __forceinline__ PairGeom<float> pair_geom<float>(const DevParams &P, double2 pi, double2 pj,
                                                            uint32_t uid_i, uint32_t uid_j,
                                                            const double *__restrict__ host_noise,
                                                            uint32_t noise_index) {
    // the difference is formed in fp64 (absolute fp32 coordinates lose 1e-4-level precision in the weights at
    // d ~ 1e-4), everything after it is fp32
    float rx = (float)(pi.x - pj.x), ry = (float)(pi.y - pj.y);
    if (P.noise_mode != SC_NOISE_NONE) {
        float ux, uy;
        if (P.noise_mode == SC_NOISE_HOST) {
            ux = (float)host_noise[2 * (size_t)noise_index];
            uy = (float)host_noise[2 * (size_t)noise_index + 1];
        } else {
            uint32_t hx, hy;
            pair_noise_bits(P.tick_key, uid_i, uid_j, hx, hy);
            ux = (float)hx * (1.0f / 4294967296.0f);
            uy = (float)hy * (1.0f / 4294967296.0f);
        }
        const float amp = (float)(P.d * P.level);
        rx = fmaf(0.5f - ux, amp, rx);
        ry = fmaf(0.5f - uy, amp, ry);
    }
    const float d2 = fmaf(rx, rx, ry * ry);
    const float inv = rsqrtf(d2);
    const float dist = d2 * inv;
    PairGeom<float> g;
    g.nx = rx * inv;
    g.ny = ry * inv;
    float cl = dist * (float)(1.0 / P.d);
    cl = fminf(fmaxf(cl, 0.0f), 1.0f);
    g.w = 1.0f - cl;
    return g;
}

// crate.py:272 np.sum over a 1-D array: NumPy pairwise_sum (sequential below 8; 8 lanes + tail up to 20)
__device__ inline double np_sum_1d(const double *a, int n) {
    if (n < 8) {
        if (n == 0) return 0.0;
        double r = a[0];
        for (int i = 1; i < n; ++i) r += a[i];
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    if (n >= 16) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += a[8 + j];
        i = 16;
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

// ------------------------------------------------------------------------------------------------------------
// One record per directed pair (i <- j), produced by K4 and consumed by K5, so the pair geometry - including the
// noise hash and the reciprocal square root - is evaluated once per tick instead of twice.  HBM is the idle
// resource on this path (the kernels are issue-bound), so 16 bytes per pair are a good trade.
template <typename Real> struct PairRec;
template <> struct __align__(16) PairRec<float> { uint32_t j; float nx, ny, w; };
template <> struct __align__(16) PairRec<double> { double nx, ny, w; uint32_t j, pad_; };

// K4: neighbor discovery, pair geometry, pressure p_i and surface normal s_i
template <typename Real>
__global__ void __launch_bounds__(SC_BLOCK)
k_density(Counters *__restrict__ cnt, Grid g, DevParams P, const uint32_t *__restrict__ cell_start,
          const double2 *__restrict__ pos, const uint32_t *__restrict__ cell_key, const uint32_t *__restrict__ uid,
          const double *__restrict__ host_noise, const uint32_t *__restrict__ noise_off,
          const uint32_t *__restrict__ rank_of_uid, PairRec<Real> *__restrict__ pairs,
          uint32_t *__restrict__ pair_off, uint8_t *__restrict__ pair_cnt, Real *__restrict__ pressure,
          typename Vec2<Real>::type *__restrict__ tension) {
    __shared__ uint32_t s_list[SC_MAX_NEIGHBORS * SC_BLOCK];
    __shared__ uint32_t s_base;
    const uint32_t n = cell_start[g.ncells];
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = s < n;
    NbrList lst{s_list + threadIdx.x};
    double2 ps = make_double2(0, 0);
    int K = 0;
    if (live) {
        ps = pos[s];
        K = collect_neighbors(s, ps, cell_key[s], g, cell_start, pos, lst);
    }
    // the block's records go to one contiguous chunk of the pair buffer (one atomic per block; where the chunk
    // lands is arbitrary, but it is only ever reached through pair_off, so results do not depend on it)
    uint32_t total;
    const uint32_t before = block_exclusive_scan((uint32_t)K, total);
    if (threadIdx.x == 0) s_base = total ? atomicAdd(&cnt->pair_cursor, total) : 0u;
    __syncthreads();
    if (!live) return;
    const uint32_t off = s_base + before;
    pair_off[s] = off;
    pair_cnt[s] = (uint8_t)K;
    const uint32_t uid_s = uid[s];
    const uint32_t nbase = (P.noise_mode == SC_NOISE_HOST) ? noise_off[rank_of_uid[uid_s]] : 0u;
    Real ax = 0, ay = 0;
    Real psum = 0;
    double wl[sizeof(Real) == 8 ? SC_MAX_NEIGHBORS : 1];  // fp64 only: np.sum's pairwise order needs the list
    for (int k = 0; k < K; ++k) {
        const uint32_t j = lst.get(k);
        const PairGeom<Real> pg = pair_geom<Real>(P, ps, pos[j], uid_s, uid[j], host_noise, nbase + (uint32_t)k);
        PairRec<Real> rec;
        rec.j = j; rec.nx = pg.nx; rec.ny = pg.ny; rec.w = pg.w;
        if constexpr (sizeof(Real) == 8) rec.pad_ = 0;
        pairs[(size_t)off + k] = rec;
        if constexpr (sizeof(Real) == 8) wl[k] = (double)pg.w; else psum += pg.w;
        const Real c = (1 - pg.w) * pg.w;
        const Real tx = c * pg.nx, ty = c * pg.ny;
        if (k == 0) { ax = tx; ay = ty; } else { ax += tx; ay += ty; }
    }
    Real p = 0;
    if (K > 0) {
        Real pr;
        if constexpr (sizeof(Real) == 8) pr = (Real)np_sum_1d(wl, K) - (Real)P.ignored;
        else pr = psum - (Real)P.ignored;
        p = (pr > 0 || pr != pr) ? pr : (Real)0;  // np.maximum(0, pr), crate.py:273
    }
    pressure[s] = p;
    typename Vec2<Real>::type t;
    t.x = ax; t.y = ay;
    tension[s] = t;
}

// ------------------------------------------------------------------------------------------------------------
// K5: all forces, wall bounce, continuous collision and integration for particle s
template <typename Real>
__global__ void __launch_bounds__(SC_BLOCK)
k_force(const uint32_t *__restrict__ n_ptr, DevParams P, const __grid_constant__ WallParams W,
        const double2 *__restrict__ pos, const typename Vec2<Real>::type *__restrict__ vel,
        const PairRec<Real> *__restrict__ pairs, const uint32_t *__restrict__ pair_off,
        const uint8_t *__restrict__ pair_cnt, const Real *__restrict__ pressure,
        const typename Vec2<Real>::type *__restrict__ tension, const uint32_t *__restrict__ wall_bits,
        const uint32_t *__restrict__ wall_slot, const double2 *__restrict__ wall_pre,
        double2 *__restrict__ pos_out, typename Vec2<Real>::type *__restrict__ vel_out) {
    typedef typename Vec2<Real>::type R2;
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= *n_ptr) return;
    const double2 ps = pos[s];
    const Real p_i = pressure[s];
    const R2 s_i = tension[s];
    const R2 v0 = vel[s];
    const PairRec<Real> *__restrict__ mine = pairs + pair_off[s];
    const int K = pair_cnt[s];
    const Real smooth = (Real)P.smooth, two_target = (Real)(2 * P.target);
    Real tx = 0, ty = 0;  // F3 sum
    Real qx = 0, qy = 0;  // F5 sum
    Real sum_vx = 0, sum_vy = 0;  // fp32 mode: sum of neighbor velocities
    for (int k = 0; k < K; ++k) {
        const PairRec<Real> rec = mine[k];
        const Real p_j = pressure[rec.j];
        const R2 s_j = tension[rec.j];
        // F3 pass 2, crate.py:347-353
        const Real ddx = s_i.x - s_j.x, ddy = s_i.y - s_j.y;
        const Real align = (ddx * rec.nx + ddy * rec.ny) * smooth;
        const Real fix = p_j + p_i - two_target;
        const Real cc = align + fix;
        const Real ex = cc * rec.nx, ey = cc * rec.ny;
        // F5, crate.py:301-306
        const Real ps_ = p_i + p_j;
        const Real fx = rec.nx * ps_, fy = rec.ny * ps_;
        if (k == 0) { tx = ex; ty = ey; qx = fx; qy = fy; } else { tx += ex; ty += ey; qx += fx; qy += fy; }
        if constexpr (sizeof(Real) == 4) { const R2 vj = vel[rec.j]; sum_vx += vj.x; sum_vy += vj.y; }
    }

    // walls: contacts are recomputed from the position the particle had BEFORE apply_hard_wall_fix
    // (crate.py:216 runs before 202-211 and the vectors are never refreshed)
    int V = 0;
    double wnx = 0, wny = 0, wux = 0, wuy = 0;  // sequential sums for np.mean (crate.py:249-250)
    const bool touching = (wall_bits[s >> 5] >> (s & 31)) & 1u;
    if (touching) {
        const double2 pre = wall_pre[wall_slot[s]];
        int nb[SC_MAX_BODIES];
        for (int b = 0; b < W.nbodies; ++b) nb[b] = 0;
        for (int q = 0; q < W.S; ++q) {
            double cx, cy;
            if (point_segment(pre.x, pre.y, W.seg[q][0], W.seg[q][1], W.seg[q][2], W.seg[q][3], cx, cy) <= P.touch)
                nb[W.seg_body[q]]++;
        }
        for (int q = 0; q < W.S; ++q) {
            double cx, cy;
            if (!(point_segment(pre.x, pre.y, W.seg[q][0], W.seg[q][1], W.seg[q][2], W.seg[q][3], cx, cy) <= P.touch))
                continue;
            const double vcx = (pre.x - cx) * 2, vcy = (pre.y - cy) * 2;  // crate.py:234, not normalised
            // W1b as written (crate.py:73-85): row V is overwritten by every body with more than V contacts
            double ux = 0, uy = 0;
            for (int b = 0; b < W.nbodies; ++b)
                if (nb[b] > V) {
                    const double rx = cx - W.kin[b][3], ry = cy - W.kin[b][4];
                    ux = W.kin[b][0] + ry * W.kin[b][2];
                    uy = W.kin[b][1] + (-rx) * W.kin[b][2];
                }
            // F5 virtual rows: n_k = vc_k, p_k = 0 (crate.py:286-293, 301-306)
            const Real ps_ = p_i + (Real)0;
            const Real fx = (Real)vcx * ps_, fy = (Real)vcy * ps_;
            if (K == 0 && V == 0) { qx = fx; qy = fy; } else { qx += fx; qy += fy; }
            if (V == 0) { wnx = vcx; wny = vcy; wux = ux; wuy = uy; }
            else { wnx += vcx; wny += vcy; wux += ux; wuy += uy; }
            ++V;
        }
    }

    const Real dt = (Real)P.dt;
    Real vx = v0.x, vy = v0.y;
    if (K > 0) { vx += dt * tx; vy += dt * ty; }                       // F3, crate.py:352
    vx += (Real)(P.dt * P.gx); vy += (Real)(P.dt * P.gy);               // F4, crate.py:310
    if (K + V > 0) {                                                     // F5, crate.py:297, 306
        const Real c = (Real)(P.dt * P.amp);
        vx += c * qx; vy += c * qy;
    }
    {                                                                    // F6, crate.py:319-323
        Real ax = 0, ay = 0;
        if constexpr (sizeof(Real) == 8) {
            for (int q = 0; q < K; ++q) {
                const R2 vj = vel[mine[q].j];
                const Real ex = vj.x - vx, ey = vj.y - vy;
                if (q == 0) { ax = ex; ay = ey; } else { ax += ex; ay += ey; }
            }
        } else {
            ax = sum_vx - (Real)K * vx;
            ay = sum_vy - (Real)K * vy;
        }
        const Real c = (Real)(P.dt * P.visc);
        vx += c * ax; vy += c * ay;
    }
    double dvx = (double)vx, dvy = (double)vy;
    if (V > 0) {                                                         // B1, crate.py:245-259
        const double Nx = wnx / (double)V, Ny = wny / (double)V;
        const double Ux = wux / (double)V, Uy = wuy / (double)V;
        const double nrm = sqrt(fma(Ny, Ny, Nx * Nx));                   // np.linalg.norm 1-D = sqrt(ddot)
        const double hx = Nx / nrm, hy = Ny / nrm;
        const double rvx = dvx - Ux, rvy = dvy - Uy;
        const double dot = fma(rvy, hy, rvx * hx);                       // np.dot = ddot
        if (dot < 0) {
            const double cx = -1 * dot * hx, cy = -1 * dot * hy;
            dvx += cx; dvy += cy;
            dvx += cx * P.decay; dvy += cy * P.decay;
        }
    }
    {                                                                    // B2, crate.py:177-200
        const double mvx = dvx * P.dt, mvy = dvy * P.dt;
        const double bx = ps.x + mvx, by = ps.y + mvy;
        double f = 1.0;
        const double mxlo = fmin(ps.x, bx), mxhi = fmax(ps.x, bx), mylo = fmin(ps.y, by), myhi = fmax(ps.y, by);
        // one test for the bulk of the liquid: the movement stays inside a rectangle no padded segment reaches
        const bool clear = mxlo > W.safe_ccd[0] && mxhi < W.safe_ccd[1] && mylo > W.safe_ccd[2] && myhi < W.safe_ccd[3];
        if (!clear) {
            const double bax = bx - ps.x, bay = by - ps.y;
            for (int q = 0; q < 2 * W.S; ++q) {
                if (mxhi < W.pad_box[q][0] || mxlo > W.pad_box[q][1] || myhi < W.pad_box[q][2] || mylo > W.pad_box[q][3])
                    continue;  // the movement cannot reach this padded segment
                const double cx = W.pad[q][0], cy = W.pad[q][1], ex = W.pad[q][2], ey = W.pad[q][3];
                const double cdx = ex - cx, cdy = ey - cy;
                const bool opposite = (cdy * bax + (-cdx) * bay) < 0;    // geometry_utils.py:205
                if (!opposite) continue;
                const bool c1 = orientation(ps.x, ps.y, bx, by, cx, cy) != orientation(ps.x, ps.y, bx, by, ex, ey);
                const bool c2 = orientation(cx, cy, ex, ey, ps.x, ps.y) != orientation(cx, cy, ex, ey, bx, by);
                if (c1 && c2) {
                    const double acx = ps.x - cx, acy = ps.y - cy;
                    const double t = (acx * cdy - acy * cdx) / (cdx * mvy - cdy * mvx);  // geometry_utils.py:141-143
                    if (t < f) f = t;  // Python min(): NaN never wins (crate.py:199)
                }
            }
        }
        dvx *= f; dvy *= f;
    }
    R2 vo;
    vo.x = (Real)dvx; vo.y = (Real)dvy;
    vel_out[s] = vo;
    double2 po;                                                          // I, crate.py:361
    po.x = ps.x + P.dt * (double)vo.x;
    po.y = ps.y + P.dt * (double)vo.y;
    pos_out[s] = po;
}

}  // namespace sc
