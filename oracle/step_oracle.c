/*
 * TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C, scalar fp64) of the SandCrate particle step
 * - David-Taub/sand_crate, `Crate.physics_tick()` src/crate/crate.py:91-129 and everything it calls in
 * src/crate/collision_detector.py and src/crate/utils/geometry_utils.py - used as the parity oracle for the
 * CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it; the product (sand_crate_b200/) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function here bit-for-bit against golden
 * vectors recorded from the unmodified reference executed in the build container (oracle/make_golden.py,
 * fixtures under tests/golden/), and against the 8 known-answer cases of the reference's own
 * tests/test_distance.py.
 *
 * The reference is NumPy; NumPy never fuses multiply-add, so this file must be compiled with
 * -ffp-contract=off.  The two BLAS call sites on the path (1-D np.linalg.norm crate.py:251 and np.dot
 * crate.py:253) evaluate as fma(x1, y1, x0*y0) with OpenBLAS and are written
 * with fma() below.
 *
 * Everything is expressed per particle in the order the reference applies it:
 *   W1/W1b/W2 -> N -> F1 -> F2 -> F3 -> F4 -> F5 -> F6 -> B1 -> B2 -> I
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OC_MAX_NEIGHBORS 20 /* collision_detector.py:6 */

typedef struct {
    double dt, radius, wall_collision_decay, pressure_amplifier, ignored_pressure, collider_noise_level,
        viscosity, surface_smoothing, target_pressure, gx, gy;
} oc_params;

/* ------------------------------------------------------------------------------------------------------ */
/* counter-based pair noise: the PRODUCT's definition (sand_crate_b200/csrc/sc_common.cuh), restated here so an */
/* oracle run can consume exactly the same uniforms as the GPU's production mode.                            */
static inline uint64_t oc_mix64(uint64_t z) {
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27; z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
uint64_t oc_tick_key(uint64_t seed, uint64_t tick) { return oc_mix64(seed * 0x9E3779B97F4A7C15ULL + tick); }
static inline uint32_t oc_lowbias32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}
/* counter-based particle sources: the PRODUCT's production definition (sc_emit_particles, csrc/sc_sort.cuh k_emit,  */
/* csrc/sc_common.cuh source_key / source_uniform), restated.  What it models is the reference's                     */
/* create_new_particles + ParticleSource.generate_particles (crate.py:138-147, particle_source.py:17-24) with the     */
/* global MT19937 stream replaced by a counter stream: element 0 of a source's stream is the uniform behind its       */
/* Binomial(flow, dt) emission count (inverse CDF, evaluated by the caller), elements 1 + 4k + {0,1,2,3} are the      */
/* position / velocity jitters of its k-th particle of the tick.                                                      */
uint64_t oc_source_key(uint64_t tick_key, uint32_t q) {
    return oc_mix64(tick_key ^ (0xD1B54A32D192ED03ULL * (uint64_t)(q + 1u)));
}
double oc_source_uniform(uint64_t skey, uint64_t j) {
    return (double)(oc_mix64(skey + (j + 1ull) * 0x9E3779B97F4A7C15ULL) >> 11) * (1.0 / 9007199254740992.0);
}
/* src: nsrc x 6 (position x, y, radius, velocity x, y, noise); count / index per source; P = particle count before the */
/* tick's removal.  Appends rows in source order; returns how many.  pos_out / vel_out: room for sum(count) rows.      */
int64_t oc_emit_counter(uint64_t tick_key, int nsrc, const double *src, const int32_t *count, const uint32_t *index,
                        int64_t P, int64_t max_particles, double *pos_out, double *vel_out) {
    int64_t out = 0;
    for (int q = 0; q < nsrc; ++q) {
        const double *S = src + 6 * q;
        int64_t room = max_particles - P;               /* crate.py:143: max_particles - particle_count */
        if (room < 0) room = 0;
        const int64_t a = count[q] < room ? count[q] : room;   /* particle_source.py:18 */
        const uint64_t key = oc_source_key(tick_key, index[q]);
        for (int64_t k = 0; k < a; ++k) {
            const double ux = oc_source_uniform(key, 1 + 4 * (uint64_t)k), uy = oc_source_uniform(key, 2 + 4 * (uint64_t)k);
            const double wx = oc_source_uniform(key, 3 + 4 * (uint64_t)k), wy = oc_source_uniform(key, 4 + 4 * (uint64_t)k);
            pos_out[2 * out] = (ux - 0.5) * S[2] + S[0];        /* particle_source.py:21 */
            pos_out[2 * out + 1] = (uy - 0.5) * S[2] + S[1];
            vel_out[2 * out] = S[3] + (wx - 0.5) * S[5];        /* particle_source.py:22-23 */
            vel_out[2 * out + 1] = S[4] + (wy - 0.5) * S[5];
            ++out;
        }
        P += a;
    }
    return out;
}

void oc_pair_noise(uint64_t tick_key, uint32_t uid_i, uint32_t uid_j, double *ux, double *uy) {
    const uint32_t a = uid_i & 0x7FFFFFFFu, b = uid_j & 0x7FFFFFFFu;
    const uint32_t h = oc_lowbias32((a * 0x9E3779B1u) ^ (b * 0x85EBCA77u) ^
                                    ((uint32_t)tick_key ^ (uint32_t)(tick_key >> 32)));
    *ux = (double)(h >> 16) * (1.0 / 65536.0);   /* high half: x uniform, low half: y uniform */
    *uy = (double)(h & 0xFFFFu) * (1.0 / 65536.0);
}

/* ------------------------------------------------------------------------------------------------------ */
/* C2  remove_particles  crate.py:149-159: stable delete of rows with any coordinate < -r or > 1 + r.        */
int64_t oc_remove_particles(double *pos, double *vel, int64_t P, double radius, uint8_t *removed_mask) {
    int64_t w = 0;
    const double lo = -radius, hi = 1 + radius;
    for (int64_t i = 0; i < P; ++i) {
        const double x = pos[2 * i], y = pos[2 * i + 1];
        const int out = (x < lo) | (x > hi) | (y < lo) | (y > hi);
        if (removed_mask) removed_mask[i] = (uint8_t)out;
        if (out) continue;
        pos[2 * w] = x; pos[2 * w + 1] = y;
        vel[2 * w] = vel[2 * i]; vel[2 * w + 1] = vel[2 * i + 1];
        ++w;
    }
    return w;
}

/* ------------------------------------------------------------------------------------------------------ */
/* geometry_utils.py:7-39  points_to_segments_distance for one (point, segment).                            */
static inline double point_segment(double px, double py, const double *seg, double *cx, double *cy) {
    const double ax = seg[0], ay = seg[1], bx = seg[2], by = seg[3];
    const double abx = bx - ax, aby = by - ay;
    const double apx = px - ax, apy = py - ay;
    const double rate = (apx * abx + apy * aby) / (abx * abx + aby * aby);
    double t = rate;                       /* np.clip(rate, 0, 1); NaN propagates */
    if (t < 0) t = 0;
    if (t > 1) t = 1;
    *cx = abx * t + ax;
    *cy = aby * t + ay;
    const double dx = *cx - px, dy = *cy - py;
    return sqrt(dx * dx + dy * dy);
}

void oc_points_to_segments_distance(const double *pos, int64_t P, const double *segments, int S,
                                    double *nearest /* P*S*2 */, double *dist /* P*S */) {
    for (int64_t i = 0; i < P; ++i)
        for (int k = 0; k < S; ++k) {
            double cx, cy;
            dist[i * S + k] = point_segment(pos[2 * i], pos[2 * i + 1], segments + 4 * k, &cx, &cy);
            nearest[(i * S + k) * 2] = cx;
            nearest[(i * S + k) * 2 + 1] = cy;
        }
}

/* geometry_utils.py:146-172  pad_segments: first S are (a+o, b+o), next S are (b-o, a-o).                  */
void oc_pad_segments(const double *segments, int S, double pad, double *padded /* 2S*4 */) {
    for (int k = 0; k < S; ++k) {
        const double ax = segments[4 * k], ay = segments[4 * k + 1], bx = segments[4 * k + 2], by = segments[4 * k + 3];
        const double abx = bx - ax, aby = by - ay;
        const double nx = aby, ny = -abx;                 /* rot90cw: (x, y) -> (y, -x) */
        const double norm = sqrt(nx * nx + ny * ny);
        const double ox = nx * pad / norm, oy = ny * pad / norm;
        double *p1 = padded + 4 * k, *p2 = padded + 4 * (S + k);
        p1[0] = ax + ox; p1[1] = ay + oy; p1[2] = bx + ox; p1[3] = by + oy;
        p2[0] = bx - ox; p2[1] = by - oy; p2[2] = ax - ox; p2[3] = ay - oy;
    }
}

static inline double sign_np(double v) { /* np.sign: -1, 0, 1, NaN */
    if (v > 0) return 1.0;
    if (v < 0) return -1.0;
    if (v == 0) return 0.0;
    return v;
}
/* geometry_utils.py:212-222 */
static inline double orientation(double px, double py, double qx, double qy, double rx, double ry) {
    return sign_np(((qy - py) * (rx - qx)) - ((qx - px) * (ry - qy)));
}

/* ------------------------------------------------------------------------------------------------------ */
/* N  detect_particle_collisions  collision_detector.py:9-128                                               */
typedef struct { int64_t row; double x; int64_t idx; } sort_rec;
static int cmp_rec(const void *a, const void *b) {
    const sort_rec *p = (const sort_rec *)a, *q = (const sort_rec *)b;
    if (p->row != q->row) return p->row < q->row ? -1 : 1;
    if (p->x != q->x) return p->x < q->x ? -1 : 1;
    return p->idx < q->idx ? -1 : (p->idx > q->idx ? 1 : 0); /* np.lexsort is stable */
}

static int64_t upper_bound(const double *a, int64_t n, double v) { /* searchsorted side="right" */
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (a[m] <= v) lo = m + 1; else hi = m; }
    return lo;
}
static int64_t lower_bound(const double *a, int64_t n, double v) { /* searchsorted side="left" */
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (a[m] < v) lo = m + 1; else hi = m; }
    return lo;
}

/* outputs: rows_sorted[P], order[P] (collision_detector.py:124-128), counts[P] and idx[P*20] in ORIGINAL
 * particle index order holding ORIGINAL indices (collision_detector.py:46-48).  Returns 0 / -1 (alloc).   */
int oc_detect_particle_collisions(const double *pos, int64_t P, double diameter, int64_t *rows_sorted,
                                  int64_t *order, int32_t *counts, int32_t *idx) {
    if (P == 0) return 0;
    sort_rec *rec = (sort_rec *)malloc(sizeof(sort_rec) * (size_t)P);
    double *sx = (double *)malloc(sizeof(double) * (size_t)P);
    double *sy = (double *)malloc(sizeof(double) * (size_t)P);
    int64_t *strip_start = (int64_t *)malloc(sizeof(int64_t) * (size_t)(P + 3));
    int64_t *fwd_off = (int64_t *)calloc((size_t)P + 1, sizeof(int64_t));
    int64_t *bwd_cnt = (int64_t *)calloc((size_t)P + 1, sizeof(int64_t));
    if (!rec || !sx || !sy || !strip_start || !fwd_off || !bwd_cnt) return -1;
    for (int64_t i = 0; i < P; ++i) {
        rec[i].row = (int64_t)floor(pos[2 * i + 1] / diameter);
        rec[i].x = pos[2 * i];
        rec[i].idx = i;
    }
    qsort(rec, (size_t)P, sizeof(sort_rec), cmp_rec);
    for (int64_t s = 0; s < P; ++s) {
        sx[s] = pos[2 * rec[s].idx];
        sy[s] = pos[2 * rec[s].idx + 1];
        if (rows_sorted) rows_sorted[s] = rec[s].row;
        if (order) order[s] = rec[s].idx;
    }
    /* strips = runs of equal row; the "next strip" is the next NON-EMPTY one (collision_detector.py:34-40) */
    int64_t nstrips = 0;
    for (int64_t s = 0; s < P; ++s)
        if (s == 0 || rec[s].row != rec[s - 1].row) strip_start[nstrips++] = s;
    strip_start[nstrips] = P;
    strip_start[nstrips + 1] = P;

    /* two passes over the forward search: count, then fill */
    int64_t *fwd = NULL;
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            int64_t tot = 0;
            for (int64_t s = 0; s < P; ++s) { int64_t c = fwd_off[s]; fwd_off[s] = tot; tot += c; }
            fwd_off[P] = tot;
            fwd = (int64_t *)malloc(sizeof(int64_t) * (size_t)(tot ? tot : 1));
            if (!fwd) return -1;
        }
        for (int64_t st = 0; st < nstrips; ++st) {
            const int64_t a = strip_start[st], b = strip_start[st + 1], c = strip_start[st + 2];
            for (int64_t s = a; s < b; ++s) {
                const double x = sx[s], y = sy[s];
                int64_t n = 0;
                /* current strip: (s, a + searchsorted(strip_x, x + d, right)) */
                const int64_t end_in = a + upper_bound(sx + a, b - a, x + diameter);
                /* next strip: [b + searchsorted(next_x, x - d, left), b + searchsorted(next_x, x + d, right)) */
                const int64_t beg_nx = b + lower_bound(sx + b, c - b, x - diameter);
                const int64_t end_nx = b + upper_bound(sx + b, c - b, x + diameter);
                for (int seg = 0; seg < 2; ++seg) {
                    const int64_t j0 = seg == 0 ? s + 1 : beg_nx, j1 = seg == 0 ? end_in : end_nx;
                    for (int64_t j = j0; j < j1; ++j) {
                        const double dx = sx[j] - x, dy = sy[j] - y;
                        if (sqrt(dx * dx + dy * dy) <= diameter) { /* collision_detector.py:77-79 */
                            if (pass == 1) fwd[fwd_off[s] + n] = j; else bwd_cnt[j]++;
                            ++n;
                        }
                    }
                }
                if (pass == 0) fwd_off[s] = n;
            }
        }
    }
    /* add_reverse_collisions (85-88) + trim (91-93): list(s) = fwd(s) ascending ++ {lo : s in fwd(lo)} descending */
    int64_t *bwd_off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(P + 1));
    if (!bwd_off) return -1;
    { int64_t tot = 0; for (int64_t s = 0; s < P; ++s) { bwd_off[s] = tot; tot += bwd_cnt[s]; } bwd_off[P] = tot; }
    int64_t *bwd = (int64_t *)malloc(sizeof(int64_t) * (size_t)(bwd_off[P] ? bwd_off[P] : 1));
    int64_t *bfill = (int64_t *)calloc((size_t)P + 1, sizeof(int64_t));
    if (!bwd || !bfill) return -1;
    for (int64_t s = P - 1; s >= 0; --s)
        for (int64_t k = fwd_off[s + 1] - 1; k >= fwd_off[s]; --k) {
            const int64_t j = fwd[k];
            bwd[bwd_off[j] + bfill[j]++] = s;
        }
    for (int64_t s = 0; s < P; ++s) {
        const int64_t o = rec[s].idx;
        int32_t n = 0;
        for (int64_t k = fwd_off[s]; k < fwd_off[s + 1] && n < OC_MAX_NEIGHBORS; ++k)
            idx[o * OC_MAX_NEIGHBORS + n++] = (int32_t)rec[fwd[k]].idx;
        for (int64_t k = bwd_off[s]; k < bwd_off[s + 1] && n < OC_MAX_NEIGHBORS; ++k)
            idx[o * OC_MAX_NEIGHBORS + n++] = (int32_t)rec[bwd[k]].idx;
        counts[o] = n;
    }
    free(rec); free(sx); free(sy); free(strip_start); free(fwd_off); free(bwd_cnt);
    free(fwd); free(bwd_off); free(bwd); free(bfill);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------ */
/* crate.py:272 `np.sum(collider_overlaps, 0)` on a 1-D contiguous array: NumPy's pairwise_sum - sequential */
/* for n < 8, otherwise 8 accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential    */
/* tail (n <= 20 here, so the 128-block recursion is never reached).                                        */
static double np_sum_1d(const double *a, int n) {
    if (n < 8) {
        double r = 0.0;               /* NumPy starts from a[0]; 0.0 + a[0] == a[0] for the w >= 0 summed here */
        if (n == 0) return 0.0;
        r = a[0];
        for (int i = 1; i < n; ++i) r += a[i];
        return r;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

/* per-particle wall contact record (W1 + W1b): contacts ordered by segment index                            */
typedef struct {
    int n;            /* V_i */
    double sum_vcx, sum_vcy;   /* sequential sums for np.mean (crate.py:249) */
    double sum_ux, sum_uy;     /* sequential sums of contact velocities (crate.py:250) */
} wall_summary;

/*
 * The step proper, on inputs taken AFTER create/remove/apply_bodies_velocity (crate.py:92-95):
 *   pos, vel      P x 2, updated in place
 *   segments      S x 4 (ax, ay, bx, by), bodies concatenated in order (crate.py:69-71)
 *   body_len      segments per body; body_kin = nbodies x 5 (vcx, vcy, omega, posx, posy)
 *   noise_mode    0: noise term skipped (exact when collider_noise_level == 0)
 *                 1: counter-based (tick_key, uid) - the product's production definition
 *                 2: `noise` holds sum(K_i) x 2 uniforms in CSR order, i.e. the reference's own
 *                    np.random.rand stream (crate.py:168-170)
 *   uid           P ids for noise_mode 1 (NULL -> index)
 * optional outputs (may be NULL): pos_search (P x 2, after W2), counts/idx (P, P x 20), pressure (P),
 *   tension_vec (P x 2), ccd_factor (P), wall_count (P), monitor (6)
 */
int oc_step(const oc_params *prm, int64_t P, double *pos, double *vel, const double *segments, int S,
            const int32_t *body_len, const double *body_kin, int nbodies, int noise_mode, const double *noise,
            uint64_t tick_key, const uint32_t *uid, double *pos_search, int32_t *counts_out, int32_t *idx_out,
            double *pressure_out, double *tension_out, double *ccd_out, int32_t *wall_count_out,
            double *monitor_out /* 6: mean |dv| per force section, utils/force_monitor.py:27-33 */) {
    const double dt = prm->dt, r = prm->radius;
    const double d = r * 2; /* crate.py:65-67 */
    if (P == 0) return 0;

    /* ---- W1 calc_virtual_colliders crate.py:213-243 (+ W1b 73-85), W2 apply_hard_wall_fix 202-211 ------- */
    int *seg_body = (int *)malloc(sizeof(int) * (size_t)(S ? S : 1));
    { int k = 0; for (int b = 0; b < nbodies; ++b) for (int q = 0; q < body_len[b]; ++q) seg_body[k++] = b; }
    int32_t *wall_n = (int32_t *)calloc((size_t)P, sizeof(int32_t));
    /* ragged contact storage, capacity S per particle only for touching particles -> CSR via two passes */
    int64_t *woff = (int64_t *)malloc(sizeof(int64_t) * (size_t)(P + 1));
    const double touch = r * 1.2; /* crate.py:229 */
    int64_t wtot = 0;
    for (int64_t i = 0; i < P; ++i) {
        woff[i] = wtot;
        for (int k = 0; k < S; ++k) {
            double cx, cy;
            if (point_segment(pos[2 * i], pos[2 * i + 1], segments + 4 * k, &cx, &cy) <= touch) { wall_n[i]++; wtot++; }
        }
    }
    woff[P] = wtot;
    double *vc = (double *)malloc(sizeof(double) * 2 * (size_t)(wtot ? wtot : 1));    /* virtual_colliders */
    double *vcv = (double *)malloc(sizeof(double) * 2 * (size_t)(wtot ? wtot : 1));   /* virtual_colliders_velocity */
    double *cpt = (double *)malloc(sizeof(double) * 2 * (size_t)(S ? S : 1));
    int *cseg = (int *)malloc(sizeof(int) * (size_t)(S ? S : 1));
    for (int64_t i = 0; i < P; ++i) {
        if (!wall_n[i]) continue;
        const double px = pos[2 * i], py = pos[2 * i + 1];
        int n = 0;
        for (int k = 0; k < S; ++k) {
            double cx, cy;
            if (point_segment(px, py, segments + 4 * k, &cx, &cy) <= touch) {
                cpt[2 * n] = cx; cpt[2 * n + 1] = cy; cseg[n] = k;
                vc[2 * (woff[i] + n)] = (px - cx) * 2;       /* crate.py:234: not normalised */
                vc[2 * (woff[i] + n) + 1] = (py - cy) * 2;
                vcv[2 * (woff[i] + n)] = 0.0;
                vcv[2 * (woff[i] + n) + 1] = 0.0;
                ++n;
            }
        }
        /* W1b rigid_bodies_points_velocities crate.py:73-85 AS WRITTEN: `calculated_points` stays 0, so each
         * body with n_b contacts overwrites rows [0, n_b) using points [0, n_b) (rigid_body.py:28-34).     */
        for (int b = 0; b < nbodies; ++b) {
            int nb = 0;
            for (int q = 0; q < n; ++q) nb += (seg_body[cseg[q]] == b);
            if (!nb) continue;
            const double *kin = body_kin + 5 * b;
            for (int q = 0; q < nb; ++q) {
                const double cx = cpt[2 * q] - kin[3], cy = cpt[2 * q + 1] - kin[4];
                const double tx = cy, ty = -cx;                       /* rot90cw */
                vcv[2 * (woff[i] + q)] = kin[0] + tx * kin[2];
                vcv[2 * (woff[i] + q) + 1] = kin[1] + ty * kin[2];
            }
        }
        /* W2: pos += sum_k vc_k * (max(r / |vc_k|, 0.5) - 0.5) */
        double sx_ = 0, sy_ = 0;
        for (int q = 0; q < n; ++q) {
            const double vx = vc[2 * (woff[i] + q)], vy = vc[2 * (woff[i] + q) + 1];
            double rel = r / sqrt(vx * vx + vy * vy);
            if (rel < 0.5) rel = 0.5;
            const double cxq = vx * (rel - 0.5), cyq = vy * (rel - 0.5);
            if (q == 0) { sx_ = cxq; sy_ = cyq; } else { sx_ += cxq; sy_ += cyq; }
        }
        pos[2 * i] += sx_;
        pos[2 * i + 1] += sy_;
    }
    if (pos_search) memcpy(pos_search, pos, sizeof(double) * 2 * (size_t)P);
    if (wall_count_out) memcpy(wall_count_out, wall_n, sizeof(int32_t) * (size_t)P);

    /* ---- N neighbor search on the corrected positions ------------------------------------------------- */
    int32_t *counts = (int32_t *)malloc(sizeof(int32_t) * (size_t)P);
    int32_t *idx = (int32_t *)malloc(sizeof(int32_t) * (size_t)P * OC_MAX_NEIGHBORS);
    if (!counts || !idx) return -1;
    memset(idx, 0xFF, sizeof(int32_t) * (size_t)P * OC_MAX_NEIGHBORS); /* -1 padding */
    if (oc_detect_particle_collisions(pos, P, d, NULL, NULL, counts, idx)) return -1;
    if (counts_out) memcpy(counts_out, counts, sizeof(int32_t) * (size_t)P);
    if (idx_out) memcpy(idx_out, idx, sizeof(int32_t) * (size_t)P * OC_MAX_NEIGHBORS);
    int64_t *off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(P + 1));
    { int64_t t = 0; for (int64_t i = 0; i < P; ++i) { off[i] = t; t += counts[i]; } off[P] = t; }
    const int64_t npairs = off[P];

    /* ---- F1 populate_colliders crate.py:161-175; F2 pressures 261-284 ----------------------------------- */
    double *nx = (double *)malloc(sizeof(double) * (size_t)(npairs ? npairs : 1));
    double *ny = (double *)malloc(sizeof(double) * (size_t)(npairs ? npairs : 1));
    double *w = (double *)malloc(sizeof(double) * (size_t)(npairs ? npairs : 1));
    double *p = (double *)malloc(sizeof(double) * (size_t)P);
    double *sv = (double *)calloc((size_t)P * 2, sizeof(double));
    double *v0 = (double *)malloc(sizeof(double) * 2 * (size_t)P);
    memcpy(v0, vel, sizeof(double) * 2 * (size_t)P); /* collider_velocities are start-of-tick copies (175) */
    const double level = prm->collider_noise_level;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < P; ++i) {
        const int K = counts[i];
        for (int k = 0; k < K; ++k) {
            const int32_t j = idx[i * OC_MAX_NEIGHBORS + k];
            double qx = pos[2 * j], qy = pos[2 * j + 1];
            if (noise_mode != 0) {
                double ux, uy;
                if (noise_mode == 2) { ux = noise[2 * (off[i] + k)]; uy = noise[2 * (off[i] + k) + 1]; }
                else oc_pair_noise(tick_key, uid ? uid[i] : (uint32_t)i, uid ? uid[j] : (uint32_t)j, &ux, &uy);
                qx += (ux - 0.5) * d * level;      /* crate.py:168-170 */
                qy += (uy - 0.5) * d * level;
            }
            const double rx = pos[2 * i] - qx, ry = pos[2 * i + 1] - qy;
            const double dist = sqrt(rx * rx + ry * ry);
            nx[off[i] + k] = rx / dist;
            ny[off[i] + k] = ry / dist;
            double c = dist / d;                    /* crate.py:270 */
            if (c < 0) c = 0;
            if (c > 1) c = 1;
            w[off[i] + k] = 1 - c;
        }
        if (K == 0) { p[i] = 0.0; continue; }
        double pr = np_sum_1d(w + off[i], K) - prm->ignored_pressure;
        p[i] = (pr > 0.0 || pr != pr) ? pr : 0.0; /* np.maximum(0, pr) crate.py:273: NaN propagates */
    }
    if (pressure_out) memcpy(pressure_out, p, sizeof(double) * (size_t)P);

    /* ---- F3 apply_tension pass 1 crate.py:337-342 ------------------------------------------------------ */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < P; ++i) {
        const int K = counts[i];
        if (K == 0) continue;
        double ax = 0, ay = 0;
        for (int k = 0; k < K; ++k) {
            const double wk = w[off[i] + k];
            const double c = (1 - wk) * wk;
            const double tx = c * nx[off[i] + k], ty = c * ny[off[i] + k];
            if (k == 0) { ax = tx; ay = ty; } else { ax += tx; ay += ty; }
        }
        sv[2 * i] = ax; sv[2 * i + 1] = ay;
    }
    if (tension_out) memcpy(tension_out, sv, sizeof(double) * 2 * (size_t)P);

    double *padded = (double *)malloc(sizeof(double) * 8 * (size_t)(S ? S : 1));
    oc_pad_segments(segments, S, r, padded); /* crate.py:182 */

    double m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0, m5 = 0;
#define OC_STAGE(ACC) { const double ex_ = vx - mvx_, ey_ = vy - mvy_; ACC += sqrt(ex_ * ex_ + ey_ * ey_); mvx_ = vx; mvy_ = vy; }
#pragma omp parallel for schedule(static) reduction(+ : m0, m1, m2, m3, m4, m5)
    for (int64_t i = 0; i < P; ++i) {
        const int K = counts[i];
        const int V = wall_n[i];
        double vx = vel[2 * i], vy = vel[2 * i + 1];
        double mvx_ = vx, mvy_ = vy;
        const double pi_ = p[i];
        /* ---- F3 pass 2 crate.py:343-353 ---- */
        if (K > 0) {
            double ax = 0, ay = 0;
            for (int k = 0; k < K; ++k) {
                const int32_t j = idx[i * OC_MAX_NEIGHBORS + k];
                const double nkx = nx[off[i] + k], nky = ny[off[i] + k];
                const double ddx = sv[2 * i] - sv[2 * j], ddy = sv[2 * i + 1] - sv[2 * j + 1];
                const double align = (ddx * nkx + ddy * nky) * prm->surface_smoothing;
                const double fix = p[j] + pi_ - 2 * prm->target_pressure;
                const double c = align + fix;
                const double tx = c * nkx, ty = c * nky;
                if (k == 0) { ax = tx; ay = ty; } else { ax += tx; ay += ty; }
            }
            vx += dt * ax; vy += dt * ay;
        }
        OC_STAGE(m0)
        /* ---- F4 apply_gravity crate.py:309-310 ---- */
        vx += dt * prm->gx; vy += dt * prm->gy;
        OC_STAGE(m1)
        /* ---- F5 apply_pressure crate.py:295-307: real rows then virtual rows (p_k = 0, n_k = vc_k) ---- */
        if (K + V > 0) {
            double ax = 0, ay = 0;
            int first = 1;
            for (int k = 0; k < K; ++k) {
                const int32_t j = idx[i * OC_MAX_NEIGHBORS + k];
                const double s_ = pi_ + p[j];
                const double tx = nx[off[i] + k] * s_, ty = ny[off[i] + k] * s_;
                if (first) { ax = tx; ay = ty; first = 0; } else { ax += tx; ay += ty; }
            }
            for (int q = 0; q < V; ++q) {
                const double s_ = pi_ + 0.0;
                const double tx = vc[2 * (woff[i] + q)] * s_, ty = vc[2 * (woff[i] + q) + 1] * s_;
                if (first) { ax = tx; ay = ty; first = 0; } else { ax += tx; ay += ty; }
            }
            const double c = dt * prm->pressure_amplifier;
            vx += c * ax; vy += c * ay;
        }
        OC_STAGE(m2)
        /* ---- F6 apply_viscosity crate.py:316-323: snapshot v_j, CURRENT v_i; runs for every particle ---- */
        {
            double ax = 0, ay = 0;
            for (int k = 0; k < K; ++k) {
                const int32_t j = idx[i * OC_MAX_NEIGHBORS + k];
                const double tx = v0[2 * j] - vx, ty = v0[2 * j + 1] - vy;
                if (k == 0) { ax = tx; ay = ty; } else { ax += tx; ay += ty; }
            }
            const double c = dt * prm->viscosity;
            vx += c * ax; vy += c * ay;
        }
        OC_STAGE(m3)
        /* ---- B1 apply_wall_bounce crate.py:245-259 ---- */
        if (V > 0) {
            double sx_ = 0, sy_ = 0, ux = 0, uy = 0;
            for (int q = 0; q < V; ++q) {
                if (q == 0) {
                    sx_ = vc[2 * woff[i]]; sy_ = vc[2 * woff[i] + 1]; ux = vcv[2 * woff[i]]; uy = vcv[2 * woff[i] + 1];
                } else {
                    sx_ += vc[2 * (woff[i] + q)]; sy_ += vc[2 * (woff[i] + q) + 1];
                    ux += vcv[2 * (woff[i] + q)]; uy += vcv[2 * (woff[i] + q) + 1];
                }
            }
            const double Nx = sx_ / (double)V, Ny = sy_ / (double)V;
            const double Ux = ux / (double)V, Uy = uy / (double)V;
            const double nrm = sqrt(fma(Ny, Ny, Nx * Nx));        /* BLAS ddot form */
            const double hx = Nx / nrm, hy = Ny / nrm;
            const double rvx = vx - Ux, rvy = vy - Uy;
            const double dot = fma(rvy, hy, rvx * hx);            /* BLAS ddot form */
            if (dot < 0) {
                const double cx = -1 * dot * hx, cy = -1 * dot * hy;
                vx += cx; vy += cy;
                vx += cx * prm->wall_collision_decay; vy += cy * prm->wall_collision_decay;
            }
        }
        OC_STAGE(m4)
        /* ---- B2 apply_continuous_collision_velocity_fix crate.py:177-200 ---- */
        {
            const double ax_ = pos[2 * i], ay_ = pos[2 * i + 1];
            const double mvx = vx * dt, mvy = vy * dt;
            const double bx_ = ax_ + mvx, by_ = ay_ + mvy;
            double f = 1.0;
            for (int k = 0; k < 2 * S; ++k) {
                const double cx = padded[4 * k], cy = padded[4 * k + 1], dx_ = padded[4 * k + 2], dy_ = padded[4 * k + 3];
                const double cdx = dx_ - cx, cdy = dy_ - cy;
                /* rot90cw(d - c) . (b - a) < 0   geometry_utils.py:205 */
                const double bax = bx_ - ax_, bay = by_ - ay_;
                const int opposite = (cdy * bax + (-cdx) * bay) < 0;
                const int c1 = orientation(ax_, ay_, bx_, by_, cx, cy) != orientation(ax_, ay_, bx_, by_, dx_, dy_);
                const int c2 = orientation(cx, cy, dx_, dy_, ax_, ay_) != orientation(cx, cy, dx_, dy_, bx_, by_);
                if (c1 && c2 && opposite) {
                    /* calc_collision_point(a, ab = v*dt, c, cd)  geometry_utils.py:141-143 */
                    const double acx = ax_ - cx, acy = ay_ - cy;
                    const double t = (acx * cdy - acy * cdx) / (cdx * mvy - cdy * mvx);
                    if (t < f) f = t;                               /* Python min(): NaN never wins */
                }
            }
            if (ccd_out) ccd_out[i] = f;
            vx *= f; vy *= f;
        }
        OC_STAGE(m5)
        vel[2 * i] = vx; vel[2 * i + 1] = vy;
    }
#undef OC_STAGE
    if (monitor_out) {
        monitor_out[0] = m0 / (double)P; monitor_out[1] = m1 / (double)P; monitor_out[2] = m2 / (double)P;
        monitor_out[3] = m3 / (double)P; monitor_out[4] = m4 / (double)P; monitor_out[5] = m5 / (double)P;
    }
    /* ---- I apply_particles_velocity crate.py:360-361 ---- */
    for (int64_t i = 0; i < 2 * P; ++i) pos[i] += dt * vel[i];

    free(seg_body); free(wall_n); free(woff); free(vc); free(vcv); free(cpt); free(cseg);
    free(counts); free(idx); free(off); free(nx); free(ny); free(w); free(p); free(sv); free(v0); free(padded);
    return 0;
}

int oc_max_neighbors(void) { return OC_MAX_NEIGHBORS; }
int oc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
