// sc_dist.cuh - strip decomposition across GPUs: pack / unpack kernels of the per-tick boundary exchange.
//
// The reference's own neighbor search is a 1-D strip decomposition in y with strip height one diameter
// (collision_detector.py:10-31, 124-128); the same rows are the multi-GPU partition unit.  Rank k owns the cell rows
// [row_lo, row_hi).  Every tick, BEFORE the step:
//   pack    drops last tick's ghosts, keeps particles whose row is still owned, turns particles whose row left the
//           strip into MIGRANT records for the neighbor (and keeps them here as ghosts for this tick), and copies
//           owned particles within `halo` rows of a cut into HALO records for that neighbor
//   (host)  exchanges the two fixed-capacity buffers with rank-1 / rank+1 (NCCL send/recv over NVLink)
//   unpack  appends received migrants as owned particles and received halos as ghosts
// and the ordinary single-GPU step then runs on owned + ghosts.  A ghost is a full particle (it is sorted, gets
// its own pressure and normal, is integrated) whose results are simply discarded by the next pack; with a halo of
// 4 rows every owned particle sees exactly the neighbors, pressures and normals it would see in the global
// computation, in the same order, so fp64 results are bit-identical to a single-GPU run.
#pragma once
#include "sc_sort.cuh"

namespace sc {

#define SC_GHOST_BIT 0x80000000u
#define SC_WIRE_MIGRANT 0u
#define SC_WIRE_HALO 1u

// 16-byte header followed by `count` records
struct WireHeader { uint32_t count, overflow, too_far, pad_; };
struct __align__(8) WireRec { double px, py, vx, vy; uint32_t uid, kind; };  // 40 bytes

struct DistCfg {
    long long row_lo, row_hi;  // owned rows: row_lo <= floor(y / d) < row_hi
    int halo;                  // rows
    int has_lo, has_hi;        // neighbors exist
    uint32_t cap;              // records per wire buffer
};

__device__ __forceinline__ void wire_push(WireHeader *h, WireRec *recs, uint32_t cap, double2 p, double vx, double vy,
                                          uint32_t uid, uint32_t kind) {
    const uint32_t k = atomicAdd(&h->count, 1u);
    if (k >= cap) { h->overflow = 1u; return; }
    WireRec r;
    r.px = p.x; r.py = p.y; r.vx = vx; r.vy = vy; r.uid = uid; r.kind = kind;
    recs[k] = r;
}

// counters used: cnt->n (in: particles in the *_in arrays; out: particles in the *_out arrays)
template <typename Real>
__global__ void __launch_bounds__(SC_BLOCK)
k_dist_pack(Counters *__restrict__ cnt, const uint32_t *__restrict__ n_in_ptr, Grid g, DistCfg D,
            const double2 *__restrict__ pos_in, const typename Vec2<Real>::type *__restrict__ vel_in,
            const uint32_t *__restrict__ uid_in, double2 *__restrict__ pos_out,
            typename Vec2<Real>::type *__restrict__ vel_out, uint32_t *__restrict__ uid_out,
            uint32_t *__restrict__ n_out, uint32_t out_cap, WireHeader *__restrict__ lo_hdr,
            WireHeader *__restrict__ hi_hdr) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_in_ptr) return;
    (void)cnt;
    const uint32_t u = uid_in[i];
    if (u & SC_GHOST_BIT) return;  // last tick's ghost: its owner has the authoritative copy
    const double2 p = pos_in[i];
    const typename Vec2<Real>::type v = vel_in[i];
    const double fr = floor_div(p.y, g);
    const long long row = (fr >= -9.0e18 && fr <= 9.0e18) ? (long long)fr : 0;  // NaN: stays where it is
    WireRec *lo_recs = reinterpret_cast<WireRec *>(lo_hdr + 1), *hi_recs = reinterpret_cast<WireRec *>(hi_hdr + 1);
    uint32_t tag = u;
    const bool below = row < D.row_lo && D.has_lo, above = row >= D.row_hi && D.has_hi;
    if (below || above) {
        // the particle's row now belongs to a neighbor: hand it over, keep it here as a ghost for this tick
        if (below) {
            if (row < D.row_lo - D.halo) lo_hdr->too_far = 1u;
            wire_push(lo_hdr, lo_recs, D.cap, p, (double)v.x, (double)v.y, u, SC_WIRE_MIGRANT);
        } else {
            if (row >= D.row_hi + D.halo) hi_hdr->too_far = 1u;
            wire_push(hi_hdr, hi_recs, D.cap, p, (double)v.x, (double)v.y, u, SC_WIRE_MIGRANT);
        }
        tag = u | SC_GHOST_BIT;
    } else {
        if (D.has_lo && row < D.row_lo + D.halo) wire_push(lo_hdr, lo_recs, D.cap, p, (double)v.x, (double)v.y, u, SC_WIRE_HALO);
        if (D.has_hi && row >= D.row_hi - D.halo) wire_push(hi_hdr, hi_recs, D.cap, p, (double)v.x, (double)v.y, u, SC_WIRE_HALO);
    }
    const uint32_t k = atomicAdd(n_out, 1u);
    if (k >= out_cap) { lo_hdr->overflow = 1u; return; }
    pos_out[k] = p;
    vel_out[k] = v;
    uid_out[k] = tag;
}

template <typename Real>
__global__ void __launch_bounds__(SC_BLOCK)
k_dist_unpack(const WireHeader *__restrict__ hdr, uint32_t wire_cap, double2 *__restrict__ pos,
              typename Vec2<Real>::type *__restrict__ vel, uint32_t *__restrict__ uid, uint32_t *__restrict__ n,
              uint32_t cap, uint32_t *__restrict__ overflow) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t count = hdr->count < wire_cap ? hdr->count : wire_cap;
    if (i >= count) return;
    const WireRec r = reinterpret_cast<const WireRec *>(hdr + 1)[i];
    const uint32_t k = atomicAdd(n, 1u);
    if (k >= cap) { *overflow = 1u; return; }
    pos[k] = make_double2(r.px, r.py);
    typename Vec2<Real>::type v;
    v.x = (Real)r.vx; v.y = (Real)r.vy;
    vel[k] = v;
    uid[k] = r.kind == SC_WIRE_HALO ? (r.uid | SC_GHOST_BIT) : r.uid;
}

// owned particles (no ghosts) compacted into staging arrays for readback; order is arbitrary, uids identify rows
template <typename Real>
__global__ void __launch_bounds__(SC_BLOCK)
k_dist_collect_owned(const uint32_t *__restrict__ n_ptr, const double2 *__restrict__ pos,
                     const typename Vec2<Real>::type *__restrict__ vel, const uint32_t *__restrict__ uid,
                     double2 *__restrict__ pos_out, double2 *__restrict__ vel_out, uint32_t *__restrict__ uid_out,
                     uint32_t *__restrict__ n_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *n_ptr) return;
    const uint32_t u = uid[i];
    if (u & SC_GHOST_BIT) return;
    const uint32_t k = atomicAdd(n_out, 1u);
    pos_out[k] = pos[i];
    const typename Vec2<Real>::type v = vel[i];
    vel_out[k] = make_double2((double)v.x, (double)v.y);
    uid_out[k] = u;
}

__global__ void k_wire_reset(WireHeader *a, WireHeader *b, uint32_t *n_out) {
    a->count = 0; a->overflow = 0; a->too_far = 0; a->pad_ = 0;
    b->count = 0; b->overflow = 0; b->too_far = 0; b->pad_ = 0;
    *n_out = 0;
}

}  // namespace sc
