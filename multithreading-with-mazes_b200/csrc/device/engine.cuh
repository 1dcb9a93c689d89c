// engine.cuh -- the BFS hot path, B200-first.
//
// What the reference does (solvers/bfs_threads.cc:35-85,144-185; painters/distance.cc:129-153):
// a FIFO breadth-first search per thread over the packed Square grid, one cell per pop, parents
// and levels in an unordered_map.  What this file does instead:
//
//   * the maze is cut into 64x64-square TILES.  A tile's passable squares are ONE BIT each
//     (64 x uint64 per tile, tile-major so a warp reads a tile as 512 contiguous bytes);
//   * one WARP owns a tile while it is being relaxed.  Lane i keeps rows 2i and 2i+1 of the
//     tile's path mask, visited mask and current frontier IN REGISTERS; one BFS level is
//         next = ((F<<1) | (F>>1) | F_row_above | F_row_below) & path & ~visited
//     with the two neighbouring rows fetched by warp shuffles.  No shared memory, no
//     __syncthreads, no atomics inside a tile: a level costs a few dozen instructions whether
//     it settles one square of a corridor or sixty-four squares of an open wavefront;
//   * a tile is relaxed to its LOCAL FIXPOINT in one go (all levels, seeds from the four
//     neighbouring tiles injected at their own levels), so a long corridor costs one warp a few
//     tens of nanoseconds per square instead of one kernel launch / grid sync per BFS level;
//   * tiles talk through per-tile EDGE arrays (the levels of their boundary rows/columns) and a
//     per-tile state word; a tile whose neighbour's edge improved is marked dirty and pushed on
//     a global ticket queue drained by a PERSISTENT kernel (one launch per search).  Relaxation
//     is label-correcting: levels only ever decrease, any processing order converges to the
//     exact BFS levels, and on spanning-tree mazes no square is ever settled twice;
//   * the warp that finishes a tile walks straight into one of the tiles it just woke
//     (no queue round trip on the critical path) and queues the rest for idle warps.
//
// Every kernel body is a plain function of (warp id, lane) written against simt.cuh, so
// tests/simt/ can execute the very same code on a CPU warp emulator and compare it with the
// oracle; kernels.cu wraps each body in a __global__.
#pragma once
#include "simt.cuh"

namespace lab {

constexpr uint32_t INF = 0xFFFFFFFFu;
constexpr int TS = 64;              // tile side, squares
constexpr int MAX_PLANES = 4;       // independent BFS fields searched by one launch (corners: 4)
constexpr int MAX_TARGETS = 4;
constexpr uint32_t Q_EMPTY = 0xFFFFFFFFu;
constexpr uint32_t ITEM_EXIT = 0xFFFFFFFEu;
constexpr uint32_t ITEM_NONE = 0xFFFFFFFDu;

enum Side : int { SIDE_N = 0, SIDE_S = 1, SIDE_W = 2, SIDE_E = 3 };
constexpr uint32_t ST_ACTIVE = 1u; // queued or being relaxed
constexpr uint32_t ST_DIRTY = 2u;  // a neighbour's edge changed since the tile last looked

enum Bound_mode : uint32_t {
    BOUND_NONE = 0,    // search everything (distance map)
    BOUND_MIN_ANY = 1, // stop at the nearest target (hunt, corners: first colour home ends the game)
    BOUND_MAX_ALL = 2  // stop once every target is reached (gather)
};

// Square bits (maze/maze.cc:170-173, solvers/solve_utilities.cc:59-79, painters/rgb.cc:21-22)
constexpr uint16_t SQ_PATH = 0x2000, SQ_START = 0x4000, SQ_FINISH = 0x8000;
constexpr uint16_t SQ_PAINT_MASK = 0x00F0, SQ_CACHE_MASK = 0x0F00, SQ_MEASURE = 0x0200;
constexpr int SQ_PAINT_SHIFT = 4, SQ_CACHE_SHIFT = 8;

struct Geom {
    int rows, cols;       // squares (of THIS band when sharded)
    int tiles_y, tiles_x; // ceil(rows/64), ceil(cols/64)
    int ntiles;
};

// One BFS field.
struct Plane {
    uint64_t *visited; // [ntiles][64]   settled at least once
    uint32_t *edges;   // [(tiles_y+2)*tiles_x][4][64]  levels of each tile's boundary squares; one ghost
                       //                               tile-row above and below (other bands' edges)
    uint32_t *state;   // [ntiles]
    uint32_t *dist;    // [rows*cols]    dense output, INF = wall / unreached
};

// Run control block, lives in global memory.
struct Control {
    unsigned long long head, tail; // ticket queue
    uint32_t pending;              // tiles ACTIVE
    uint32_t done;
    uint32_t error;                // 1 = deadline passed, 2 = queue slot clobbered
    uint32_t bound;                // frontiers at level >= bound are not expanded
    uint32_t bound_mode;
    uint32_t n_targets;
    uint32_t target_plane[MAX_TARGETS];
    uint32_t target_tile[MAX_TARGETS];
    unsigned long long target_cell[MAX_TARGETS]; // r*cols+c
    uint32_t target_val[MAX_TARGETS];
    uint32_t max_level;            // max level ever written (exact iff n_resettles == 0)
    uint32_t pad0;
    uint32_t src_tile[MAX_PLANES]; // the tile holding plane k's level-0 square (0xFFFFFFFF: none here)
    uint32_t src_rc[MAX_PLANES];   // (row_in_tile << 8) | col_in_tile
    unsigned long long deadline_ns; // %globaltimer value after which the workers give up (error 1)
    // statistics
    unsigned long long n_activations, n_empty, n_levels, n_rechecks, n_resettles, n_settled, n_pushes, n_direct;
};

struct Params {
    Geom g;
    uint64_t const *path; // [ntiles][64]
    uint64_t const *cross; // [ntiles][4] boundary squares passable on BOTH sides of the tile border
    Plane plane[MAX_PLANES];
    int nplanes;
    Control *ctl;
    uint32_t *queue;
    uint32_t qmask; // capacity-1 (power of two > planes*ntiles + #warps)
};

LAB_DEV uint64_t lab_shfl_up64(uint64_t v, int d) {
    uint32_t lo = lab_shfl_up(static_cast<uint32_t>(v), d);
    uint32_t hi = lab_shfl_up(static_cast<uint32_t>(v >> 32), d);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
LAB_DEV uint64_t lab_shfl_down64(uint64_t v, int d) {
    uint32_t lo = lab_shfl_down(static_cast<uint32_t>(v), d);
    uint32_t hi = lab_shfl_down(static_cast<uint32_t>(v >> 32), d);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
LAB_DEV uint64_t lab_shfl64(uint64_t v, int src) {
    uint32_t lo = lab_shfl(static_cast<uint32_t>(v), src);
    uint32_t hi = lab_shfl(static_cast<uint32_t>(v >> 32), src);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}
LAB_DEV uint64_t lab_ballot64(bool lo, bool hi) {
    uint32_t a = lab_ballot(lo);
    uint32_t b = lab_ballot(hi);
    return (static_cast<uint64_t>(b) << 32) | a;
}
LAB_DEV uint32_t lab_min3(uint32_t a, uint32_t b, uint32_t c) { return a < b ? (a < c ? a : c) : (b < c ? b : c); }
LAB_DEV uint32_t lab_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }

// ---------------------------------------------------------------------------------------------
// pack: Square grid -> one passable bit per square, tile-major.
// passable  <=>  (square & mask) == value     (solvers: path_bit, bfs_threads.cc:71-72;
//                                              distance: path_bit set AND measure clear, distance.cc:145-148)
// One warp-item = 64 squares of one row (lane reads squares lane and lane+32).
LAB_DEV void pack_body(Geom g, uint16_t const *grid, uint64_t *path, uint16_t mask, uint16_t value,
                       long long gwarp, long long nwarps, int lane) {
    constexpr int U = 4;
    long long const rows_pad = static_cast<long long>(g.tiles_y) * TS;
    long long const total = rows_pad * g.tiles_x;
    for (long long base = gwarp * U; base < total; base += nwarps * U) {
        uint16_t a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            long long const item = base + u;
            long long const r = item / g.tiles_x;
            int const j = static_cast<int>(item % g.tiles_x);
            int const c0 = j * TS + lane;
            int const c1 = c0 + 32;
            bool const in_r = item < total && r < g.rows;
            a[u] = (in_r && c0 < g.cols) ? lab_ld_ro(grid + r * g.cols + c0) : static_cast<uint16_t>(~value);
            b[u] = (in_r && c1 < g.cols) ? lab_ld_ro(grid + r * g.cols + c1) : static_cast<uint16_t>(~value);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            long long const item = base + u;
            long long const r = item / g.tiles_x;
            int const j = static_cast<int>(item % g.tiles_x);
            bool const in_r = item < total && r < g.rows;
            int const c0 = j * TS + lane;
            uint64_t const w = lab_ballot64(in_r && c0 < g.cols && (a[u] & mask) == value,
                                            in_r && c0 + 32 < g.cols && (b[u] & mask) == value);
            if (lane == 0 && item < total) {
                path[((r / TS) * g.tiles_x + j) * TS + (r % TS)] = w;
            }
        }
    }
}

// cross masks: bit i of cross[t][side] <=> boundary square i of tile t on that side AND the square
// facing it in the neighbouring tile are both passable.  Warp per tile.  ghost_n / ghost_s are the
// path rows of the neighbouring BANDS' facing tile rows when sharded (else null).
LAB_DEV void cross_body(Geom g, uint64_t const *path, uint64_t *cross, uint64_t const *ghost_n,
                        uint64_t const *ghost_s, long long gwarp, long long nwarps, int lane) {
    for (long long t = gwarp; t < g.ntiles; t += nwarps) {
        int const ty = static_cast<int>(t / g.tiles_x);
        int const tx = static_cast<int>(t % g.tiles_x);
        uint64_t const *me = path + t * TS;
        uint64_t xn = 0, xs = 0;
        if (ty > 0) {
            xn = me[0] & path[(t - g.tiles_x) * TS + 63];
        } else if (ghost_n) {
            xn = me[0] & ghost_n[tx];
        }
        if (ty < g.tiles_y - 1) {
            xs = me[63] & path[(t + g.tiles_x) * TS + 0];
        } else if (ghost_s) {
            xs = me[63] & ghost_s[tx];
        }
        uint64_t const r_lo = me[lane], r_hi = me[lane + 32];
        bool w_lo = false, w_hi = false, e_lo = false, e_hi = false;
        if (tx > 0) {
            uint64_t const *w = path + (t - 1) * TS;
            w_lo = (r_lo & 1) && (w[lane] >> 63);
            w_hi = (r_hi & 1) && (w[lane + 32] >> 63);
        }
        if (tx < g.tiles_x - 1) {
            uint64_t const *e = path + (t + 1) * TS;
            e_lo = (r_lo >> 63) && (e[lane] & 1);
            e_hi = (r_hi >> 63) && (e[lane + 32] & 1);
        }
        uint64_t const xw = lab_ballot64(w_lo, w_hi);
        uint64_t const xe = lab_ballot64(e_lo, e_hi);
        if (lane == 0) {
            cross[t * 4 + SIDE_N] = xn;
            cross[t * 4 + SIDE_S] = xs;
            cross[t * 4 + SIDE_W] = xw;
            cross[t * 4 + SIDE_E] = xe;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// queue
LAB_DEV void queue_push(Params const &p, uint32_t item) {
    unsigned long long const t = lab_atomic_add(reinterpret_cast<uint64_t *>(&p.ctl->tail), 1ull);
    uint32_t const old = lab_atomic_exch(&p.queue[t & p.qmask], item);
    if (old != Q_EMPTY) {
        lab_atomic_max(&p.ctl->error, 2u);
    }
}

// Mark (plane, tile) dirty; returns true if THIS call made it active (caller must run or queue it).
LAB_DEV bool tile_wake(Params const &p, int plane, long long t) {
    uint32_t const old = lab_atomic_or(&p.plane[plane].state[t], ST_ACTIVE | ST_DIRTY);
    if (old & ST_ACTIVE) {
        return false;
    }
    lab_atomic_add(&p.ctl->pending, 1u);
    return true;
}

// Seed one source square: forces it passable (the reference pushes its start without looking at
// path_bit: bfs_threads.cc:41-43, distance.cc:135-138) and makes its tile active.  Call with one
// thread, BEFORE cross_body.
LAB_DEV void source_force_passable(Geom g, uint64_t *path, int r, int c) {
    long long const t = static_cast<long long>(r / TS) * g.tiles_x + c / TS;
    path[t * TS + (r % TS)] |= 1ull << (c % TS);
}

// ---------------------------------------------------------------------------------------------
// One activation of one tile by one warp.
struct Tile_ctx {
    uint32_t *dist_row0; // &dist[(row0 + 2*lane) * cols + col0]
    uint32_t *edges;     // this tile's [4][64]
    long long cols;
    int lane;
};

LAB_DEV void settle_row(Tile_ctx const &tc, uint64_t bits, int which, int row_in_tile, uint32_t lvl) {
    uint32_t *drow = tc.dist_row0 + which * tc.cols;
    uint64_t w = bits;
    while (w) {
        int const c = lab_ffs64(w) - 1;
        w &= w - 1;
        drow[c] = lvl;
    }
    if (bits & 1ull) {
        tc.edges[SIDE_W * TS + row_in_tile] = lvl;
    }
    if (bits >> 63) {
        tc.edges[SIDE_E * TS + row_in_tile] = lvl;
    }
}

LAB_DEV void settle_edge_row(uint32_t *edge, uint64_t bits, uint32_t lvl) {
    uint64_t w = bits;
    while (w) {
        int const c = lab_ffs64(w) - 1;
        w &= w - 1;
        edge[c] = lvl;
    }
}

// Which neighbours could use the squares just settled at `lvl`?  (bit per Side)
// A neighbour is woken only if lvl + 1 beats the level it already has across the border, which
// also stops the wave echoing back into the tile it came from.
LAB_DEV uint32_t wake_mask(uint64_t b0, uint64_t b1, uint32_t lvl, int lane, uint64_t XN, uint64_t XS,
                           uint32_t fNmax, uint32_t fSmax, uint32_t fW0, uint32_t fW1, uint32_t fE0,
                           uint32_t fE1) {
    uint32_t const l1 = lvl + 1;
    uint32_t m = 0;
    m |= (lane == 0 && (b0 & XN) && l1 < fNmax) ? (1u << SIDE_N) : 0u;
    m |= (lane == 31 && (b1 & XS) && l1 < fSmax) ? (1u << SIDE_S) : 0u;
    m |= (((b0 & 1ull) && l1 < fW0) || ((b1 & 1ull) && l1 < fW1)) ? (1u << SIDE_W) : 0u;
    m |= (((b0 >> 63) && l1 < fE0) || ((b1 >> 63) && l1 < fE1)) ? (1u << SIDE_E) : 0u;
    return m;
}

// One border square: `seed` = facing level + 1 if that beats what this tile holds, else INF.
// `f` = the facing level (0 when the border cannot be crossed here).
LAB_DEV void load_seed(bool crossing, uint32_t const *facing, uint32_t const *own, uint32_t &seed, uint32_t &f) {
    seed = INF;
    f = 0;
    if (crossing) {
        f = lab_ld_cg(facing);
        uint32_t const o = lab_ld_cg(own);
        if (f != INF && f + 1 < o) {
            seed = f + 1;
        }
    }
}

// returns the next item this warp should relax itself (ITEM_NONE if none)
LAB_DEV uint32_t relax_tile(Params const &p, uint32_t item, int lane) {
    Geom const &g = p.g;
    int const plane = static_cast<int>(item / static_cast<uint32_t>(g.ntiles));
    long long const t = item % static_cast<uint32_t>(g.ntiles);
    int const ty = static_cast<int>(t / g.tiles_x);
    int const tx = static_cast<int>(t % g.tiles_x);
    Plane const &pl = p.plane[plane];
    Control *ctl = p.ctl;
    int const r0 = 2 * lane, r1 = 2 * lane + 1;

    uint64_t const P0 = lab_ld_ro(p.path + t * TS + r0);
    uint64_t const P1 = lab_ld_ro(p.path + t * TS + r1);
    uint64_t const XN = lab_ld_ro(p.cross + t * 4 + SIDE_N);
    uint64_t const XS = lab_ld_ro(p.cross + t * 4 + SIDE_S);
    uint64_t const XW = lab_ld_ro(p.cross + t * 4 + SIDE_W);
    uint64_t const XE = lab_ld_ro(p.cross + t * 4 + SIDE_E);
    // edges has one ghost tile-row above and below
    uint32_t *const my_edges = pl.edges + (t + g.tiles_x) * (4 * TS);
    uint32_t const *const n_edges = my_edges - static_cast<long long>(g.tiles_x) * (4 * TS);
    uint32_t const *const s_edges = my_edges + static_cast<long long>(g.tiles_x) * (4 * TS);
    uint32_t const *const w_edges = my_edges - 4 * TS;
    uint32_t const *const e_edges = my_edges + 4 * TS;

    Tile_ctx tc;
    tc.cols = g.cols;
    tc.lane = lane;
    tc.edges = my_edges;
    tc.dist_row0 = pl.dist + (static_cast<long long>(ty) * TS + r0) * g.cols + static_cast<long long>(tx) * TS;

    uint32_t st_levels = 0, st_rechecks = 0, st_resettles = 0, st_settled = 0;
    uint32_t st_passes = 0, st_empty = 0;
    uint32_t woke_total = 0; // sides on which a neighbour must look again (accumulated over reruns)
    uint32_t max_lvl = 0;

    for (;;) { // rerun while a neighbour dirtied us during the pass
        if (lane == 0) {
            lab_atomic_and(&pl.state[t], ~ST_DIRTY);
        }
        lab_threadfence();
        lab_syncwarp();

        uint64_t V0 = lab_ld_cg(pl.visited + t * TS + r0);
        uint64_t V1 = lab_ld_cg(pl.visited + t * TS + r1);
        uint64_t O0 = V0, O1 = V1; // settled by an EARLIER activation and not re-settled by this one
        uint64_t const V0_in = V0, V1_in = V1;

        // ---- seeds: facing edge level + 1 where that beats what this tile has ------------------
        // N/S seeds: lane holds columns lane and lane+32.  W/E seeds: lane holds rows 2*lane, 2*lane+1.
        // f* = the facing square's level where the border can be crossed, else 0 ("never worth waking").
        uint32_t sN0, sN1, sS0, sS1, sW0, sW1, sE0, sE1;
        uint32_t fN0, fN1, fS0, fS1, fW0, fW1, fE0, fE1;
        load_seed((XN >> lane) & 1, n_edges + SIDE_S * TS + lane, my_edges + SIDE_N * TS + lane, sN0, fN0);
        load_seed((XN >> (lane + 32)) & 1, n_edges + SIDE_S * TS + lane + 32, my_edges + SIDE_N * TS + lane + 32, sN1, fN1);
        load_seed((XS >> lane) & 1, s_edges + SIDE_N * TS + lane, my_edges + SIDE_S * TS + lane, sS0, fS0);
        load_seed((XS >> (lane + 32)) & 1, s_edges + SIDE_N * TS + lane + 32, my_edges + SIDE_S * TS + lane + 32, sS1, fS1);
        load_seed((XW >> r0) & 1, w_edges + SIDE_E * TS + r0, my_edges + SIDE_W * TS + r0, sW0, fW0);
        load_seed((XW >> r1) & 1, w_edges + SIDE_E * TS + r1, my_edges + SIDE_W * TS + r1, sW1, fW1);
        load_seed((XE >> r0) & 1, e_edges + SIDE_W * TS + r0, my_edges + SIDE_E * TS + r0, sE0, fE0);
        load_seed((XE >> r1) & 1, e_edges + SIDE_W * TS + r1, my_edges + SIDE_E * TS + r1, sE1, fE1);
        // a row-0 / row-63 square settled at level L can only help the tile above / below if
        // L + 1 beats the WORST level over there (conservative: exact per-column test not needed)
        uint32_t const fNmax = lab_reduce_max(fN0 > fN1 ? fN0 : fN1);
        uint32_t const fSmax = lab_reduce_max(fS0 > fS1 ? fS0 : fS1);
        // The plane's source square, if it lies in this tile and is not settled yet, is a level-0 seed.
        uint64_t src0 = 0, src1 = 0;
        if (ctl->src_tile[plane] == static_cast<uint32_t>(t)) {
            uint32_t const rc = ctl->src_rc[plane];
            int const sr = static_cast<int>(rc >> 8), sc = static_cast<int>(rc & 0xFF);
            if (sr == r0 && !((V0 >> sc) & 1)) src0 = 1ull << sc;
            if (sr == r1 && !((V1 >> sc) & 1)) src1 = 1ull << sc;
        }
        uint32_t sSrc = (src0 | src1) ? 0u : INF;

        uint32_t pending_min = lab_umin(lab_umin(lab_min3(sN0, sN1, sS0), lab_min3(sS1, sW0, sW1)),
                                        lab_min3(sE0, sE1, sSrc));
        pending_min = lab_reduce_min(pending_min);
        st_passes++;

        uint32_t chg = 0;
        if (pending_min == INF) {
            st_empty++;
        } else {
            uint32_t lvl = pending_min;
            uint64_t F0 = 0, F1 = 0;
            uint32_t bound = lab_ld_volatile(&ctl->bound);
            for (;;) {
                // ---- inject the seeds whose level is `lvl` ---------------------------------------
                if (lvl == pending_min) {
                    uint64_t const nrow = lab_ballot64(sN0 == lvl, sN1 == lvl); // -> row 0  (lane 0, r0)
                    uint64_t const srow = lab_ballot64(sS0 == lvl, sS1 == lvl); // -> row 63 (lane 31, r1)
                    uint64_t i0 = (sW0 == lvl ? 1ull : 0ull) | (sE0 == lvl ? (1ull << 63) : 0ull);
                    uint64_t i1 = (sW1 == lvl ? 1ull : 0ull) | (sE1 == lvl ? (1ull << 63) : 0ull);
                    if (lane == 0) i0 |= nrow;
                    if (lane == 31) i1 |= srow;
                    if (sSrc == lvl) {
                        i0 |= src0;
                        i1 |= src1;
                    }
                    // not if this activation already settled the square at a level <= lvl
                    i0 &= ~(V0 & ~O0);
                    i1 &= ~(V1 & ~O1);
                    st_resettles += lab_popc64(i0 & O0) + lab_popc64(i1 & O1);
                    V0 |= i0;
                    V1 |= i1;
                    O0 &= ~i0;
                    O1 &= ~i1;
                    F0 |= i0;
                    F1 |= i1;
                    if (i0 | i1) {
                        settle_row(tc, i0, 0, r0, lvl);
                        settle_row(tc, i1, 1, r1, lvl);
                        if (lane == 0 && i0) settle_edge_row(my_edges + SIDE_N * TS, i0, lvl);
                        if (lane == 31 && i1) settle_edge_row(my_edges + SIDE_S * TS, i1, lvl);
                        st_settled += lab_popc64(i0) + lab_popc64(i1);
                        if (lvl > max_lvl) max_lvl = lvl;
                        chg |= wake_mask(i0, i1, lvl, lane, XN, XS, fNmax, fSmax, fW0, fW1, fE0, fE1);
                    }
                    if (sN0 == lvl) sN0 = INF;
                    if (sN1 == lvl) sN1 = INF;
                    if (sS0 == lvl) sS0 = INF;
                    if (sS1 == lvl) sS1 = INF;
                    if (sW0 == lvl) sW0 = INF;
                    if (sW1 == lvl) sW1 = INF;
                    if (sE0 == lvl) sE0 = INF;
                    if (sE1 == lvl) sE1 = INF;
                    if (sSrc == lvl) sSrc = INF;
                    pending_min = lab_umin(lab_umin(lab_min3(sN0, sN1, sS0), lab_min3(sS1, sW0, sW1)),
                                           lab_min3(sE0, sE1, sSrc));
                    pending_min = lab_reduce_min(pending_min);
                }
                if (lvl >= bound) {
                    break; // nothing at or beyond the bound needs expanding
                }
                // ---- one BFS level ----------------------------------------------------------------
                uint64_t up = lab_shfl_up64(F1, 1);     // row 2*lane-1
                uint64_t down = lab_shfl_down64(F0, 1); // row 2*lane+2
                if (lane == 0) up = 0;
                if (lane == 31) down = 0;
                uint64_t const C0 = ((F0 << 1) | (F0 >> 1) | up | F1) & P0;
                uint64_t const C1 = ((F1 << 1) | (F1 >> 1) | F0 | down) & P1;
                uint64_t N0 = C0 & ~V0;
                uint64_t N1 = C1 & ~V1;
                uint64_t const R0 = C0 & O0;
                uint64_t const R1 = C1 & O1;
                uint32_t const nl = lvl + 1;
                st_levels++;
                if (lab_any((R0 | R1) != 0)) {
                    // label-correcting slow path: the wave touches squares an earlier activation
                    // settled; they are re-settled only if this wave is strictly better.
                    uint64_t w = R0;
                    uint32_t *d0 = tc.dist_row0;
                    while (w) {
                        int const c = lab_ffs64(w) - 1;
                        w &= w - 1;
                        st_rechecks++;
                        if (lab_ld_cg(d0 + c) > nl) {
                            N0 |= 1ull << c;
                            O0 &= ~(1ull << c);
                            st_resettles++;
                        }
                    }
                    w = R1;
                    uint32_t *d1 = tc.dist_row0 + tc.cols;
                    while (w) {
                        int const c = lab_ffs64(w) - 1;
                        w &= w - 1;
                        st_rechecks++;
                        if (lab_ld_cg(d1 + c) > nl) {
                            N1 |= 1ull << c;
                            O1 &= ~(1ull << c);
                            st_resettles++;
                        }
                    }
                }
                V0 |= N0;
                V1 |= N1;
                if (N0 | N1) {
                    settle_row(tc, N0, 0, r0, nl);
                    settle_row(tc, N1, 1, r1, nl);
                    if (lane == 0 && N0) settle_edge_row(my_edges + SIDE_N * TS, N0, nl);
                    if (lane == 31 && N1) settle_edge_row(my_edges + SIDE_S * TS, N1, nl);
                    st_settled += lab_popc64(N0) + lab_popc64(N1);
                    if (nl > max_lvl) max_lvl = nl;
                    chg |= wake_mask(N0, N1, nl, lane, XN, XS, fNmax, fSmax, fW0, fW1, fE0, fE1);
                }
                F0 = N0;
                F1 = N1;
                lvl = nl;
                if (!lab_any((F0 | F1) != 0)) {
                    if (pending_min == INF) {
                        break;
                    }
                    lvl = pending_min; // nothing moving: jump to the next seed level
                    bound = lab_ld_volatile(&ctl->bound);
                    if (lvl >= bound) {
                        break;
                    }
                }
            }
            if ((V0 != V0_in) | (V1 != V1_in)) {
                lab_st_cg(pl.visited + t * TS + r0, V0);
                lab_st_cg(pl.visited + t * TS + r1, V1);
            }
        }
        chg = lab_reduce_or(chg);
        woke_total |= chg;

        // ---- publish: everything above must be visible before any flag below -------------------
        lab_threadfence();
        lab_syncwarp();

        // targets in this tile: refresh their values and the search bound
        uint32_t const ntg = ctl->n_targets;
        if (ntg && lane == 0) {
            bool touched = false;
            for (uint32_t j = 0; j < ntg; j++) {
                if (ctl->target_plane[j] == static_cast<uint32_t>(plane) && ctl->target_tile[j] == static_cast<uint32_t>(t)) {
                    uint32_t const v = lab_ld_cg(pl.dist + ctl->target_cell[j]);
                    if (v != INF) {
                        lab_atomic_min(&ctl->target_val[j], v);
                        touched = true;
                    }
                }
            }
            if (touched) {
                uint32_t b = ctl->bound_mode == BOUND_MIN_ANY ? INF : 0u;
                for (uint32_t j = 0; j < ntg; j++) {
                    uint32_t const v = lab_ld_volatile(&ctl->target_val[j]);
                    if (ctl->bound_mode == BOUND_MIN_ANY) {
                        b = v < b ? v : b;
                    } else {
                        b = v > b ? v : b; // INF while any target is unreached
                    }
                }
                if (ctl->bound_mode != BOUND_NONE && b != INF) {
                    lab_atomic_min(&ctl->bound, b);
                }
            }
        }

        // try to go idle; if somebody dirtied us meanwhile, go round again
        uint32_t again = 0;
        if (lane == 0) {
            uint32_t const old = lab_atomic_cas(&pl.state[t], ST_ACTIVE, 0u);
            again = (old != ST_ACTIVE) ? 1u : 0u;
        }
        again = lab_shfl(again, 0);
        if (!again) {
            break;
        }
    }

    // ---- wake the neighbours whose facing edge changed; keep one for ourselves ------------------
    uint32_t next = ITEM_NONE;
    if (lane == 0) {
        uint32_t const base = static_cast<uint32_t>(plane) * static_cast<uint32_t>(g.ntiles);
        long long nb[4];
        nb[SIDE_N] = ty > 0 ? t - g.tiles_x : -1;
        nb[SIDE_S] = ty < g.tiles_y - 1 ? t + g.tiles_x : -1;
        nb[SIDE_W] = tx > 0 ? t - 1 : -1;
        nb[SIDE_E] = tx < g.tiles_x - 1 ? t + 1 : -1;
        unsigned pushes = 0, direct = 0;
        for (int s = 0; s < 4; s++) {
            if (((woke_total >> s) & 1u) && nb[s] >= 0 && tile_wake(p, plane, nb[s])) {
                uint32_t const it = base + static_cast<uint32_t>(nb[s]);
                if (next == ITEM_NONE) {
                    next = it;
                    direct++;
                } else {
                    queue_push(p, it);
                    pushes++;
                }
            }
        }
        lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->n_activations), static_cast<uint64_t>(st_passes));
        lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->n_empty), static_cast<uint64_t>(st_empty));
        lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->n_levels), static_cast<uint64_t>(st_levels));
        if (pushes) lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->n_pushes), static_cast<uint64_t>(pushes));
        if (direct) lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->n_direct), static_cast<uint64_t>(direct));
        // this tile is idle now
        uint32_t const left = lab_atomic_sub(&ctl->pending, 1u) - 1u;
        if (left == 0u) {
            lab_st_volatile(&ctl->done, 1u);
        }
    }
    // per-lane statistics
    {
        uint32_t const s_set = lab_reduce_add(static_cast<uint32_t>(st_settled));
        uint32_t const s_rec = lab_reduce_add(static_cast<uint32_t>(st_rechecks));
        uint32_t const s_res = lab_reduce_add(static_cast<uint32_t>(st_resettles));
        uint32_t const s_max = lab_reduce_max(max_lvl);
        if (lane == 0) {
            if (s_set) lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->n_settled), static_cast<uint64_t>(s_set));
            if (s_rec) lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->n_rechecks), static_cast<uint64_t>(s_rec));
            if (s_res) lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->n_resettles), static_cast<uint64_t>(s_res));
            if (s_max) lab_atomic_max(&ctl->max_level, s_max);
        }
    }
    next = lab_shfl(next, 0);
    return next;
}

// The persistent worker: every warp of the launch runs this until the search is done.
LAB_DEV void bfs_worker_body(Params const &p, int lane) {
    Control *ctl = p.ctl;
    uint32_t item = ITEM_NONE;
    for (;;) {
        if (item == ITEM_NONE) {
            if (lane == 0) {
                unsigned long long const h = lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->head), 1ull);
                uint32_t *slot = &p.queue[h & p.qmask];
                unsigned backoff = 32;
                unsigned polls = 0;
                for (;;) {
                    item = lab_ld_volatile(slot);
                    if (item != Q_EMPTY) {
                        lab_st_volatile(slot, Q_EMPTY);
                        break;
                    }
                    if ((polls & 3u) == 0u && lab_ld_volatile(&ctl->done)) {
                        item = ITEM_EXIT;
                        break;
                    }
                    if ((polls & 255u) == 255u && lab_clock_ns() > ctl->deadline_ns) {
                        lab_atomic_max(&ctl->error, 1u);
                        lab_st_volatile(&ctl->done, 1u);
                        item = ITEM_EXIT;
                        break;
                    }
                    lab_nanosleep(backoff);
                    if (backoff < 2048) backoff *= 2;
                    polls++;
                }
            }
            item = lab_shfl(item, 0);
            if (item == ITEM_EXIT) {
                return;
            }
        }
        item = relax_tile(p, item, lane);
        // a direct hand-off still honours the deadline
        if (item != ITEM_NONE && lab_ld_volatile(&ctl->done)) {
            // only possible after an error: drop the tile so the launch can end
            if (lane == 0) lab_atomic_sub(&ctl->pending, 1u);
            return;
        }
    }
}

} // namespace lab
