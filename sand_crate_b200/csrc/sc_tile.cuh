// sc_tile.cuh - the production (mixed precision, device noise) density kernel: K4 with the neighborhood of a block
// staged in shared memory.
//
// A block owns SC_TILE consecutive particles of the sorted set.  Because the sort is cell-major (row, col, x), every
// neighbor of those particles lies in one of three CONTIGUOUS windows of the sorted set - the block's own stretch of
// its cell row(s) widened by one cell at each end, and the same stretch of cells one row below / one row above:
//
//     W0 = [cell_start[c_lo - 1],         cell_start[c_hi + 2])           same row(s)
//     W1 = [cell_start[c_lo + ncols - 1], cell_start[c_hi + ncols + 2])   next row(s)
//     W2 = [cell_start[c_lo - ncols - 1], cell_start[c_hi - ncols + 2])   previous row(s)
//
// (c_lo, c_hi = cells of the block's first / last particle; k_rank_gather leaves the six bounds in a per-block
// descriptor, so the kernel starts with ONE dependent load instead of a chain of four).  The windows and the block's
// slice of the cell boundaries are copied into shared memory with coalesced loads; the candidate loop then reads
// 16-byte records (cell-relative position, cell column, uid) from shared memory instead of issuing a dependent global
// load per candidate.  A staged particle is addressed by its position in the concatenation W0 | W1 | W2 ("local
// index", 11 bits), so the per-thread neighbor lists are 16-bit.
//
// When the three windows do not fit the staging buffer (a block of spray that spans many sparse rows), the block
// runs the same code with a pass-through accessor: local index = sorted index, reads go to global memory.  The
// arithmetic is identical in both modes, so a particle's result does not depend on which mode its block ran in
// (the strip decomposition relies on that: a ghost and its owner must compute the same bits).
//
// What the kernel computes is what k_density in sc_pair.cuh computes (same reference lines, same list order, same
// 20-trim, same fp32-screen / fp64-replay acceptance); only the data movement differs.  Measured on B200 (dam-break
// 1M): 69 -> 59 us.  The same staging for K5 (neighbor pressure / normal / velocity from shared memory instead of
// gathers) was built and measured SLOWER than the gathering kernel (60 vs 46 us: K5 does too little work per staged
// byte to pay for the extra barrier and the 3x staging traffic), so K5 stays untiled and reads sorted indices.
#pragma once
#include "sc_pair.cuh"

namespace sc {

#define SC_TILE_CAP (5 * SC_TILE)              // staged particles per block (3 windows of ~SC_TILE + a few cells each)
#define SC_TILE_CELLS (SC_TILE + SC_TILE / 2 + 64)  // staged cell boundaries per row (blocks that wrap around a row end
                                               // read them from global)

// search record of one sorted particle, written by k_rank_gather in mixed mode:
//   x, y = cell-relative position (see collect_neighbors), z = (float)cell column, w = uid bits
// The column travels as a float so that the x offset between two cells is one subtraction and one FMA:
//   x_j - x_i = (rel_j.x - rel_i.x) + (col_j - col_i) * d     with col_j - col_i in {-1, 0, 1}, exact in fp32.
typedef float4 SearchRec;

// Per block of SC_TILE sorted particles, written by k_rank_gather (the threads that place the block's first and last
// particle): the three windows and the block's first / last cell.  One 32-byte read replaces a chain of dependent
// loads (particle count -> cell keys -> cell boundaries) at the head of K4 and K5.
struct __align__(16) BlockDesc {
    uint32_t base[3], c_lo;  // first sorted index of W0, W1, W2
    uint32_t end[3], c_hi;   // one past the last
};

struct TileWindows {
    uint32_t base[3];  // first sorted index of W0, W1, W2
    uint32_t off[3];   // first local index of W0, W1, W2 (staged) / = base (pass-through)
    uint32_t total;    // staged particles
    uint32_t c_lo, ncw;  // first cell of the block, cells per row slice (c_lo - 1 .. c_hi + 2)
    bool staged, cells_staged;
};

__device__ __forceinline__ TileWindows tile_windows(const BlockDesc *desc) {
    const uint4 lo = reinterpret_cast<const uint4 *>(desc)[0], hi = reinterpret_cast<const uint4 *>(desc)[1];
    TileWindows w;
    w.base[0] = lo.x; w.base[1] = lo.y; w.base[2] = lo.z; w.c_lo = lo.w;
    const uint32_t n0 = hi.x - lo.x, n1 = hi.y - lo.y, n2 = hi.z - lo.z;
    w.total = n0 + n1 + n2;
    w.ncw = hi.w - lo.w + 4u;
    w.staged = w.total <= SC_TILE_CAP;
    w.cells_staged = w.staged && w.ncw <= SC_TILE_CELLS;
    if (w.staged) { w.off[0] = 0u; w.off[1] = n0; w.off[2] = n0 + n1; }
    else { w.off[0] = w.base[0]; w.off[1] = w.base[1]; w.off[2] = w.base[2]; }
    return w;
}

// sorted index of the staged particle at local index t (used only while staging)
__device__ __forceinline__ uint32_t tile_source(const TileWindows &w, uint32_t t) {
    return t < w.off[1] ? w.base[0] + t : (t < w.off[2] ? w.base[1] + (t - w.off[1]) : w.base[2] + (t - w.off[2]));
}

// accessors: shared memory by 32-bit shared address (keeps the address arithmetic to one instruction), or global
template <typename T> struct SmemAcc {
    uint32_t addr;
    __device__ __forceinline__ T get(uint32_t L) const;
};
template <> __device__ __forceinline__ float4 SmemAcc<float4>::get(uint32_t L) const {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr + L * 16u));
    return v;
}
template <> __device__ __forceinline__ float2 SmemAcc<float2>::get(uint32_t L) const {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr + L * 8u));
    return v;
}
template <typename T> struct GmemAcc {
    const T *p;
    __device__ __forceinline__ T get(uint32_t L) const { return p[L]; }
};
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// per-thread neighbor list, one column per thread; 16-bit entries when the indices are local (11 bits + row code)
template <typename E, int kShift> struct TileList {
    E *col;
    __device__ __forceinline__ void set(int k, uint32_t L, uint32_t code) { col[k * SC_TILE] = (E)(L | (code << kShift)); }
    __device__ __forceinline__ void get(int k, uint32_t &L, uint32_t &code) const {
        const uint32_t e = col[k * SC_TILE];
        L = e & ((1u << kShift) - 1u); code = e >> kShift;
    }
};

__device__ __forceinline__ float rsqrt_ftz(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ------------------------------------------------------------------------------------------------------------
// K4 body for one particle.  s = sorted index; b[] = the six boundaries of its four candidate ranges as LOCAL
// indices (m0, m3 | n0, n3 | p0, p3 of collect_neighbors, shifted into the staged windows).  List order = reference
// list order.  The pair records carry SORTED indices (K5 gathers from global memory).
template <int kNoise, class Acc, class List>
__device__ __forceinline__ void density_particle(const Acc &A, List lst, const TileWindows &w, bool live, uint32_t s,
                                                 const uint32_t (&b)[6], const Grid &g, const DevParams &P,
                                                 Counters *cnt, const double2 *pos,
                                                 uint2 *pair_rec, uint32_t *pair_off,
                                                 uint8_t *pair_cnt, PS<float> *ps_out) {
    const float df = (float)g.d;
    const uint32_t d0 = w.off[0] - w.base[0], d1 = w.off[1] - w.base[1], d2 = w.off[2] - w.base[2];
    int K = 0;
    SearchRec me = make_float4(0, 0, 0, 0);
    if (live) {
        const float hi = (df * df) * (1.0f + 4e-6f), lo = (df * df) * (1.0f - 4e-6f);
        const uint32_t Ls = s + d0;
        me = A.get(Ls);
        int count = 0;
        // FIRST .. STOP (exclusive) in local indices, walking by STEP; BY = y of this particle in the frame of the
        // range's cell row; DELTA = local - sorted index of the range's window.  count < 20 in the loop condition =
        // trim_collisions, collision_detector.py:91-93 (no `break`: it keeps the warp from reconverging).
        // (Ending the walk once a candidate is more than d away in x - rows are sorted by x - was measured: the extra
        // predicate costs what the shorter walks save.)
#define SC_TILE_RANGE(FIRST, STOP, STEP, BY, DR, DELTA, ASC)                                                     \
        {                                                                                                        \
            const float by = (BY);                                                                               \
            for (uint32_t L = (FIRST); L != (STOP) && count < SC_MAX_NEIGHBORS; L += (STEP)) {                   \
                const SearchRec r = A.get(L);                                                                    \
                const float dx = fmaf(r.z - me.z, df, r.x - me.x), dy = r.y - by;                                \
                const float qd = fmaf(dx, dx, dy * dy);                                                          \
                /* qd > hi: surely farther than d (NaN too: the reference rejects NaN); qd < lo: surely inside */ \
                if (qd <= hi && (qd < lo || accept_exact(pos[s], pos[L - (DELTA)], g.d, (DR), (ASC)))) {         \
                    lst.set(count, L, (uint32_t)((DR) + 1));                                                     \
                    ++count;                                                                                     \
                }                                                                                                \
            }                                                                                                    \
        }
        SC_TILE_RANGE(Ls + 1u, b[1], 1u, me.y, 0, d0, true)
        SC_TILE_RANGE(b[2], b[3], 1u, me.y - df, 1, d1, true)
        SC_TILE_RANGE(Ls - 1u, b[0] - 1u, 0xFFFFFFFFu, me.y, 0, d0, false)
        SC_TILE_RANGE(b[5] - 1u, b[4] - 1u, 0xFFFFFFFFu, me.y + df, -1, d2, false)
#undef SC_TILE_RANGE
        K = count;
    }
    // the warp's records go to one contiguous chunk of the pair buffer (one atomic per warp); where the chunk lands
    // is arbitrary, but it is only ever reached through pair_off
    const int lane = threadIdx.x & 31;
    uint32_t inc = (uint32_t)K;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    uint32_t base = 0;
    if (lane == 31 && total) base = atomicAdd(&cnt->pair_cursor, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (!live) return;
    const uint32_t off = base + inc - (uint32_t)K;
    pair_off[s] = off;
    pair_cnt[s] = (uint8_t)K;
    const uint32_t uid_s = __float_as_uint(me.w);
    const float inv_d = (float)(1.0 / P.d);
    const float amp = (float)(P.d * P.level);
    float ax = 0, ay = 0, psum = 0;
    uint2 *out = pair_rec + off;
    for (int k = 0; k < K; ++k) {
        uint32_t L, code;  // code = dr + 1
        lst.get(k, L, code);
        const SearchRec r = A.get(L);
        float rx = fmaf(me.z - r.z, df, me.x - r.x);
        float ry = (me.y - r.y) - fmaf((float)code, df, -df);  // the neighbor's row is dr * d further down
        // crate.py:167-174: the neighbor's position is noised, then the unit vector from it to i and the distance
        if constexpr (kNoise == SC_NOISE_COUNTER) {
            float fx, fy;  // uniforms + 1
            pair_noise_f32_1to2(pair_noise_bits(P.tick_key, uid_s, __float_as_uint(r.w)), fx, fy);
            rx = fmaf(1.5f - fx, amp, rx);  // 0.5 - u, exact
            ry = fmaf(1.5f - fy, amp, ry);
        }
        const float q = fmaf(rx, rx, ry * ry);
        const float inv = rsqrt_ftz(q);
        const float nx = rx * inv, ny = ry * inv;
        const float cl = __saturatef((q * inv) * inv_d);  // np.clip(dist / d, 0, 1), crate.py:270
        const float wgt = 1.0f - cl;
        // one 8-byte record per directed pair: neighbor index + the unit vector as two signed 16-bit fractions
        const int ix = __float2int_rn(nx * 32767.0f), iy = __float2int_rn(ny * 32767.0f);
        const uint32_t jrec = L - (code == 1u ? d0 : (code == 2u ? d1 : d2));
        out[k] = make_uint2(jrec, __byte_perm((uint32_t)ix, (uint32_t)iy, 0x5410));
        psum += wgt;
        const float c = cl * wgt;  // (1 - w) w, crate.py:340
        ax = fmaf(c, nx, ax);
        ay = fmaf(c, ny, ay);
    }
    float p = 0;
    if (K > 0) {
        const float pr = psum - (float)P.ignored;
        p = (pr > 0 || pr != pr) ? pr : 0.0f;  // np.maximum(0, pr), crate.py:273
    }
    PS<float> o;
    o.p = p; o.sx = ax; o.sy = ay; o.pad_ = 0;
    ps_out[s] = o;
}

#define SC_TILE_SMEM_K4 (SC_TILE_CAP * 16 + 3 * SC_TILE_CELLS * 4 + SC_MAX_NEIGHBORS * SC_TILE * 2)

template <int kNoise, int kRepeat = 1>
#ifndef SC_TILE_RESIDENT
#define SC_TILE_RESIDENT 1536  // threads per SM the register allocation is held to (1536 = 40 registers)
#endif
__global__ void __launch_bounds__(SC_TILE, SC_TILE_RESIDENT / SC_TILE)
k_density_tile(Counters *cnt, Grid g, DevParams P, const uint32_t *cell_start,
               const BlockDesc *desc, const double2 *pos,
               const SearchRec *rec, const uint32_t *cell_key,
               uint2 *pair_rec, uint32_t *pair_off, uint8_t *pair_cnt,
               PS<float> *ps_out) {
    pdl_enter();
    // staged: [records 20 KB | cell boundaries 5.25 KB | 16-bit lists 10 KB]; pass-through: [32-bit lists 20 KB]
    __shared__ __align__(16) unsigned char s_raw[SC_TILE_SMEM_K4];
    static_assert(SC_TILE_SMEM_K4 >= SC_MAX_NEIGHBORS * SC_TILE * 4, "pass-through lists must fit");
    const uint32_t b0 = blockIdx.x * SC_TILE;
    const uint32_t s = b0 + threadIdx.x;
    // first round of loads, all independent: live count, block descriptor, own cell
    const uint32_t n = cell_start[g.ncells];
    if (b0 >= n) return;
    const TileWindows w = tile_windows(desc + blockIdx.x);
    const bool live = s < n;
    const uint32_t c = live ? cell_key[s] : w.c_lo;
    const uint32_t nc = (uint32_t)g.ncols;
    uint32_t b[6];
    if (w.staged) {
        SearchRec *s_rec = reinterpret_cast<SearchRec *>(s_raw);
        uint32_t *s_cs = reinterpret_cast<uint32_t *>(s_raw + SC_TILE_CAP * 16);
        uint16_t *s_list = reinterpret_cast<uint16_t *>(s_raw + SC_TILE_CAP * 16 + 3 * SC_TILE_CELLS * 4);
        // second round: the three windows and (unless the block wraps around a row end) its cell boundaries, stored
        // as local indices
        for (uint32_t t = threadIdx.x; t < w.total; t += SC_TILE) s_rec[t] = rec[tile_source(w, t)];
        const uint32_t d0 = w.off[0] - w.base[0], d1 = w.off[1] - w.base[1], d2 = w.off[2] - w.base[2];
        if (w.cells_staged) {
            const uint32_t *src = cell_start + w.c_lo - 1u;
            for (uint32_t t = threadIdx.x; t < w.ncw; t += SC_TILE) {
                s_cs[t] = src[t] + d0;
                s_cs[SC_TILE_CELLS + t] = src[nc + t] + d1;
                s_cs[2 * SC_TILE_CELLS + t] = (src - nc)[t] + d2;
            }
        } else if (live) {
            const uint32_t *cs0 = cell_start + c - 1u;
            b[0] = cs0[0] + d0; b[1] = cs0[3] + d0;
            b[2] = cs0[nc] + d1; b[3] = cs0[nc + 3u] + d1;
            b[4] = (cs0 - nc)[0] + d2; b[5] = (cs0 - nc)[3] + d2;
        }
        __syncthreads();
        if (w.cells_staged) {
            const uint32_t q = c - w.c_lo;
            b[0] = s_cs[q]; b[1] = s_cs[q + 3u];
            b[2] = s_cs[SC_TILE_CELLS + q]; b[3] = s_cs[SC_TILE_CELLS + q + 3u];
            b[4] = s_cs[2 * SC_TILE_CELLS + q]; b[5] = s_cs[2 * SC_TILE_CELLS + q + 3u];
        }
#pragma unroll 1
        for (int rep = 0; rep < kRepeat; ++rep)  // kRepeat > 1: developer timing aid (cost of a pass without its prologue)
        density_particle<kNoise>(SmemAcc<SearchRec>{smem_addr(s_rec)},
                                             TileList<uint16_t, 13>{s_list + threadIdx.x}, w, live, s, b, g, P, cnt, pos,
                                             pair_rec, pair_off, pair_cnt, ps_out);
    } else {
        if (threadIdx.x == 0) atomicAdd(&cnt->n_untiled, 1u);  // rare; lets a test prove this path ran
        if (live) {
            const uint32_t *cs0 = cell_start + c - 1u;
            b[0] = cs0[0]; b[1] = cs0[3];
            b[2] = cs0[nc]; b[3] = cs0[nc + 3u];
            b[4] = (cs0 - nc)[0]; b[5] = (cs0 - nc)[3];
        }
        density_particle<kNoise>(GmemAcc<SearchRec>{rec},
                                             TileList<uint32_t, 28>{reinterpret_cast<uint32_t *>(s_raw) + threadIdx.x}, w,
                                             live, s, b, g, P, cnt, pos, pair_rec, pair_off, pair_cnt, ps_out);
    }
}

}  // namespace sc
