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
    // SM cycles summed over warps (LAB_PROFILE builds only): loading a tile, the level loop, publishing, idle
    unsigned long long cyc_load, cyc_loop, cyc_publish, cyc_idle, cyc_step, cyc_emit, n_batches;
    // the open room (device/room.cuh; decided by tree_init_body): the passable squares are exactly the
    // rectangle [room_r0, room_r1] x [room_c0, room_c1], so plane 0's level of a square is its Manhattan
    // distance from the source (room_sr, room_sc) and no field has to be searched
    // distance map on a tree: the fix-up has set the measure bit on every square it gave a level (n_marked of them),
    // measure_body only adds the count
    uint32_t measured;
    uint32_t spec_mark; // Square bits pack_body set speculatively on every passable square (see measure_body)
    unsigned long long n_marked;
    uint32_t room;
    int32_t room_r0, room_r1, room_c0, room_c1, room_sr, room_sc; // (row bands: relative to the band, see room.cuh)
    uint32_t painted; // the lattice road has applied the paint rules while writing the levels (lattice.cuh): paint_body has nothing to do
    uint32_t walked;  // ... and has marked the winner's path from the tour's ranks (lat_path_*): hunt's walk has nothing to do
    uint32_t walk_fallback; // lat_path_begin_body met a finish it leaves to the one-warp walk (unreachable, or on a pillar): `walked` is not set
    uint32_t fronts_done; // ... and has found every colour's path.front() by probing (lat_fronts_body; the field was not written): gather's one-step walks have nothing to do
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

#if defined(LAB_PROFILE)
#define LAB_PROF(x) x
#else
#define LAB_PROF(x)
#endif

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
// A value other warps may be changing, read ONCE per warp: lane 0 loads, everybody gets its copy.
// (32 separate loads could return different values and send lanes down different paths around
// the *_sync primitives.)
LAB_DEV uint32_t lab_ld_uniform(uint32_t const *p, int lane) {
    uint32_t v = 0;
    if (lane == 0) {
        v = lab_ld_volatile(p);
    }
    return lab_shfl(v, 0);
}
LAB_DEV uint32_t lab_min3(uint32_t a, uint32_t b, uint32_t c) { return a < b ? (a < c ? a : c) : (b < c ? b : c); }
LAB_DEV uint32_t lab_umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
LAB_DEV uint32_t lab_umax(uint32_t a, uint32_t b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------------------------
// pack: Square grid -> one passable bit per square, tile-major.
// passable  <=>  (square & mask) == value     (solvers: path_bit, bfs_threads.cc:71-72;
//                                              distance: path_bit set AND measure clear, distance.cc:145-148)
// One warp per row (grid stride), 256 squares per step: every lane loads 16 bytes (8 squares) and makes
// an 8-bit mask; three shuffles join eight neighbouring masks into the 64 bits of a tile word.  Rows of
// an odd-sized maze do not start on a 16-byte boundary, so a row is read as  head (the a < 8 squares
// before the first boundary, one scalar load each) + aligned 16-byte loads, and every word is put
// together from two pieces: (64 - a bits of this step) | (a bits carried over from the one before).
LAB_DEV uint32_t pack_mask8(lab_u32x4 v, uint32_t mask2, uint32_t value2) {
    // halfword h of x passes  <=>  ((x & mask) ^ value) has no bit in that halfword
    uint32_t const w[4] = {v.x, v.y, v.z, v.w};
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t const d = (w[k] & mask2) ^ value2;
        m |= ((d & 0xFFFFu) == 0 ? 1u : 0u) << (2 * k);
        m |= ((d >> 16) == 0 ? 1u : 0u) << (2 * k + 1);
    }
    return m;
}

// mark != 0 (the distance map on a maze that may be a lattice, device/lattice.cuh): the squares that pass also get
// `mark` ORed in on the way -- the measure bit of painters/distance.cc:138,149, set SPECULATIVELY for every passable
// square while the grid is being read anyway.  One tree: every passable square is reached and the bits are final;
// otherwise measure_body clears them again on the squares the search did not reach.
LAB_DEV void pack_body(Geom g, uint16_t *grid, uint64_t *path, uint16_t mask, uint16_t value, uint16_t mark,
                       long long gwarp, long long nwarps, int lane) {
    constexpr int U = 4; // steps whose loads are in flight together (2 KB per warp)
    int const rows_pad = g.tiles_y * TS;
    uint32_t const mask2 = static_cast<uint32_t>(mask) * 0x10001u, value2 = static_cast<uint32_t>(value) * 0x10001u;
    unsigned long long const total = static_cast<unsigned long long>(g.rows) * static_cast<unsigned long long>(g.cols);
    int const nsteps = (g.tiles_x + 3) / 4; // a step makes four words
    for (long long rr = gwarp; rr < rows_pad; rr += nwarps) {
        int const r = static_cast<int>(rr);
        uint64_t *out = path + (static_cast<long long>(r / TS) * g.tiles_x) * TS + (r % TS);
        if (r >= g.rows) { // padding rows of the last tile row: not passable
            for (int j = lane; j < g.tiles_x; j += 32) {
                out[static_cast<long long>(j) * TS] = 0;
            }
            continue;
        }
        unsigned long long const row0 = static_cast<unsigned long long>(r) * static_cast<unsigned long long>(g.cols);
        uint16_t *row = grid + row0;
        int a = static_cast<int>(((16u - (reinterpret_cast<uintptr_t>(row) & 15u)) & 15u) >> 1); // squares before the boundary
        a = a < g.cols ? a : g.cols;
        // head: bits [0, a) of word 0, parked in the top bits of the carry
        uint16_t const hsq = lab_ld_ro(row + (lane < a ? lane : 0));
        bool const hpass = lane < a && (hsq & mask) == value;
        if (mark && hpass) {
            row[lane] = static_cast<uint16_t>(hsq | mark);
        }
        uint32_t const head = lab_ballot(hpass);
        uint64_t carry = a ? static_cast<uint64_t>(head) << (64 - a) : 0ull;
        for (int s0 = 0; s0 < nsteps; s0 += U) {
            lab_u32x4 v[U];
            uint32_t valid[U]; // which of the lane's 8 squares are inside the row
#pragma unroll
            for (int u = 0; u < U; u++) {
                int const c = a + 256 * (s0 + u) + 8 * lane;
                int const left = g.cols - c;
                valid[u] = left >= 8 ? 0xFFu : (left > 0 ? (1u << left) - 1u : 0u);
                v[u] = lab_u32x4{~value2, ~value2, ~value2, ~value2};
                if (left > 0) {
                    if (row0 + static_cast<unsigned long long>(c) + 8ull <= total) { // (may run into the next row: masked below)
                        v[u] = lab_ld_stream_v4(row + c); // (read once; not the read-only path: pack may write the squares back)
                    } else { // the last few squares of the grid
                        uint32_t w[4] = {~value2, ~value2, ~value2, ~value2};
                        for (int k = 0; k < left && k < 8; k++) {
                            uint32_t const sq = lab_ld_ro(row + c + k);
                            w[k >> 1] = (k & 1) ? ((w[k >> 1] & 0x0000FFFFu) | (sq << 16)) : ((w[k >> 1] & 0xFFFF0000u) | sq);
                        }
                        v[u] = lab_u32x4{w[0], w[1], w[2], w[3]};
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                uint32_t x = pack_mask8(v[u], mask2, value2) & valid[u];
                if (mark && x) {
                    int const c = a + 256 * (s0 + u) + 8 * lane;
                    if (valid[u] == 0xFFu && row0 + static_cast<unsigned long long>(c) + 8ull <= total) { // the whole vector is this row's
                        uint32_t const m2 = static_cast<uint32_t>(mark);
                        lab_u32x4 o = v[u];
                        o.x |= ((x & 1u) ? m2 : 0u) | ((x & 2u) ? m2 << 16 : 0u);
                        o.y |= ((x & 4u) ? m2 : 0u) | ((x & 8u) ? m2 << 16 : 0u);
                        o.z |= ((x & 16u) ? m2 : 0u) | ((x & 32u) ? m2 << 16 : 0u);
                        o.w |= ((x & 64u) ? m2 : 0u) | ((x & 128u) ? m2 << 16 : 0u);
                        lab_st_v4(row + c, o);
                    } else { // the row's last few squares
                        for (int k = 0; k < 8; k++) {
                            if ((x >> k) & 1u) {
                                row[c + k] = static_cast<uint16_t>(row[c + k] | mark);
                            }
                        }
                    }
                }
                x |= lab_shfl_down(x, 1) << 8;
                x |= lab_shfl_down(x, 2) << 16;
                uint32_t const hi = lab_shfl_down(x, 4);
                uint64_t const R = static_cast<uint64_t>(x) | (static_cast<uint64_t>(hi) << 32); // squares 8*lane .. 8*lane+63 of the step (lanes 0, 8, 16, 24)
                uint64_t const before = lab_shfl_up64(R, 8);
                uint64_t const prev = lane == 0 ? carry : before;
                carry = lab_shfl64(R, 24);
                uint64_t const word = a ? ((R << a) | (prev >> (64 - a))) : R;
                int const j = 4 * (s0 + u) + (lane >> 3);
                if ((lane & 7) == 0 && j < g.tiles_x) {
                    out[static_cast<long long>(j) * TS] = word;
                }
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

// Seed one source square: forces it passable (the reference pushes its start without looking at
// path_bit: bfs_threads.cc:41-43, distance.cc:135-138) and makes its tile active.  Call with one
// thread, BEFORE cross_body.
LAB_DEV void source_force_passable(Geom g, uint64_t *path, int r, int c) {
    long long const t = static_cast<long long>(r / TS) * g.tiles_x + c / TS;
    path[t * TS + (r % TS)] |= 1ull << (c % TS);
}

// ---------------------------------------------------------------------------------------------
// One activation of one tile by one warp.
//
// A pass = one multi-source BFS over the tile from the SEEDS (border squares whose neighbour
// across the border offers a better level, and the plane's source square), run level by level
// in register bitboards.  The level loop does nothing but step: what was settled when is kept
// BIT-SLICED (six 64-bit planes per row = a 6-bit level counter per square), so a batch of up
// to 64 levels costs one shuffle + a few LOP3 per level on the critical path, and the levels are
// written out afterwards square by square (emit_row).  Seeds are injected from bit-sliced planes
// the same way.  Squares settled by an earlier activation are NOT blocked: the wave may run over
// them and emit_row keeps the smaller level (atomicMin), which is what makes the relaxation
// label-correcting without any per-level look-up.  On spanning-tree mazes a wave never meets
// an already settled square, so nothing is ever written twice.

constexpr int WINDOW = 64; // levels per batch
constexpr uint32_t DENSE_BATCH = 384; // squares settled in a batch from which the row-wise write-out pays

// Per-warp shared-memory stage for writing a batch out (see emit_batch).
// layout (uint64 units): [0,6*64) level planes, [6*64,7*64) visited-before, [7*64,8*64) facing levels
// of the W and E borders (2 x 64 uint32), then the cell list (4096 x uint16).
constexpr int STAGE_FIXED_U64 = 8 * TS;
constexpr int STAGE_U64 = STAGE_FIXED_U64 + (TS * TS * 2) / 8;
constexpr int STAGE_BYTES = STAGE_U64 * 8;

struct Tile_ctx {
    uint32_t *dist0; // &dist[row0 * cols + col0]
    uint32_t *edges; // this tile's [4][64]
    long long cols;
};

struct Emit_acc {
    uint32_t wake;      // Side bits: a neighbour can use what was just settled
    uint32_t max_lvl;
    uint32_t settled;
    uint32_t resettled; // squares that already had a level
};

// One border square of this tile settled at `lvl`: record it in the edge array; wake the
// neighbour if lvl + 1 beats what it holds across the border (f, 0 = border not crossable here).
LAB_DEV void edge_update(uint32_t *e, uint32_t lvl, bool old, uint32_t f, int side, Emit_acc &acc) {
    bool improved = true;
    if (!old) {
        *e = lvl;
    } else {
        improved = lab_atomic_min(e, lvl) > lvl;
    }
    if (improved && lvl + 1 < f) {
        acc.wake |= 1u << side;
    }
}

// The state of one wave inside a tile: lane i holds rows 2i (suffix 0) and 2i+1 (suffix 1).
struct Wave {
    uint64_t A0, A1;       // squares this pass may still settle
    uint64_t F0, F1;       // frontier: squares settled at the last level
    uint64_t B0[6], B1[6]; // level of every square settled in this batch, relative to its base, bit-sliced
};

// Seeds of the current batch, relative to its base (anything outside [0, 63] never matches).
struct Seeds {
    uint32_t rW0, rW1, rE0, rE1, rSrc;
    uint64_t src0, src1;
    uint64_t nsv, nsp[6]; // N seeds -> row 0 (lane 0), S seeds -> row 63 (lane 31), bit-sliced
};

// Eight BFS levels.  INJECT = some seed's level falls into this group.
// Critical path per level: shuffle -> LOP3 -> shuffle.  Everything else (planes, seeds) hangs off it.
// NS = some of those seeds come through the north / south border (bit-sliced planes to match).
template <bool INJECT, bool NS> LAB_DEV void step_group(Wave &w, Seeds const &sd, int gi, int lane) {
    uint64_t gm = 0;
    if (INJECT && NS) {
        gm = sd.nsv;
        gm &= (gi & 1) ? sd.nsp[3] : ~sd.nsp[3];
        gm &= (gi & 2) ? sd.nsp[4] : ~sd.nsp[4];
        gm &= (gi & 4) ? sd.nsp[5] : ~sd.nsp[5];
    }
    uint64_t G0 = 0, G1 = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        // lanes 0 / 31 get their own row back from the out-of-range shuffle: harmless, that row is
        // already OR-ed in below
        uint64_t const up = lab_shfl_up64(w.F1, 1);     // row 2*lane-1
        uint64_t const down = lab_shfl_down64(w.F0, 1); // row 2*lane+2
        uint64_t i0 = 0, i1 = 0;
        if (INJECT) {
            uint32_t const rel = 8u * static_cast<uint32_t>(gi) + static_cast<uint32_t>(j);
            uint64_t m = 0;
            if (NS) {
                m = gm;
                m &= (j & 1) ? sd.nsp[0] : ~sd.nsp[0];
                m &= (j & 2) ? sd.nsp[1] : ~sd.nsp[1];
                m &= (j & 4) ? sd.nsp[2] : ~sd.nsp[2];
            }
            i0 = (sd.rW0 == rel ? 1ull : 0ull) | (sd.rE0 == rel ? (1ull << 63) : 0ull) | (sd.rSrc == rel ? sd.src0 : 0ull);
            i1 = (sd.rW1 == rel ? 1ull : 0ull) | (sd.rE1 == rel ? (1ull << 63) : 0ull) | (sd.rSrc == rel ? sd.src1 : 0ull);
            i0 |= lane == 0 ? m : 0ull;
            i1 |= lane == 31 ? m : 0ull;
        }
        uint64_t const N0 = ((w.F0 << 1) | (w.F0 >> 1) | w.F1 | up | i0) & w.A0;
        uint64_t const N1 = ((w.F1 << 1) | (w.F1 >> 1) | w.F0 | down | i1) & w.A1;
        w.A0 &= ~N0;
        w.A1 &= ~N1;
        if (j & 1) { w.B0[0] |= N0; w.B1[0] |= N1; }
        if (j & 2) { w.B0[1] |= N0; w.B1[1] |= N1; }
        if (j & 4) { w.B0[2] |= N0; w.B1[2] |= N1; }
        G0 |= N0;
        G1 |= N1;
        w.F0 = N0;
        w.F1 = N1;
    }
    if (gi & 1) { w.B0[3] |= G0; w.B1[3] |= G1; }
    if (gi & 2) { w.B0[4] |= G0; w.B1[4] |= G1; }
    if (gi & 4) { w.B0[5] |= G0; w.B1[5] |= G1; }
}

// level (relative to the batch base) of square c of a row whose bit-sliced planes are B
LAB_DEV uint32_t plane_level(uint64_t const (&B)[6], int c) {
    uint32_t k = 0;
#pragma unroll
    for (int b = 0; b < 6; b++) {
        k |= static_cast<uint32_t>((B[b] >> c) & 1ull) << b;
    }
    return k;
}
// largest relative level among the squares S of a row
LAB_DEV uint32_t plane_max_level(uint64_t const (&B)[6], uint64_t S) {
    uint32_t k = 0;
#pragma unroll
    for (int b = 5; b >= 0; b--) {
        uint64_t const t = S & B[b];
        if (t) {
            k |= 1u << b;
            S = t;
        } else {
            S &= ~B[b];
        }
    }
    return k;
}

// Writing a batch out happens in two steps so that the neighbouring tiles can start as early as
// possible:
//   emit_borders   stages the batch (bit-sliced level planes, visited-before) in the warp's shared
//                  memory, brings the tile's EDGE arrays up to date for the border squares settled in
//                  the batch (columns 0/63: the lane that owns the row, from its registers; rows
//                  0/63: the whole warp) and tells which neighbours can use them.  Cheap.
//   -- the caller wakes those neighbours here --
//   emit_interior  writes the levels of all settled squares into the dense field:
//                  * dense batch (open areas): one tile ROW at a time, lane l writes columns l and
//                    l+32: two coalesced 128-byte stores per row, no data-dependent branches;
//                  * sparse batch (corridors): every lane lists the squares of ITS two rows as
//                    (row<<6|col) in a shared-memory cell list at a prefix-sum offset, then the list
//                    is consumed 32 squares at a time, so no lane holds a long corridor alone.
// S0/S1 = squares settled in the batch (rows 2*lane, 2*lane+1), V0/V1 = squares that already had a
// level before this pass (they get atomicMin instead of a store).
// Returns the number of squares settled in the batch (warp-wide).
// EDGES = false: only stage the batch and keep the statistics (tree.cuh's final flood has no neighbours to tell).
template <bool EDGES>
LAB_DEV uint32_t emit_borders(Tile_ctx const &tc, uint64_t *stage, Wave const &w, uint64_t S0, uint64_t S1, uint64_t V0,
                              uint64_t V1, uint32_t base, int lane, uint32_t fW0, uint32_t fW1, uint32_t fE0, uint32_t fE1,
                              uint32_t fNmax, uint32_t fSmax, Emit_acc &acc) {
    uint32_t const mine = static_cast<uint32_t>(lab_popc64(S0) + lab_popc64(S1));
    uint32_t const total = lab_reduce_add(mine);
    if (total == 0) {
        return 0;
    }
    int const r0 = 2 * lane, r1 = 2 * lane + 1;
    // stage[wd * 64 + row]: level planes and visited-before
#pragma unroll
    for (int b = 0; b < 6; b++) {
        stage[b * TS + r0] = w.B0[b];
        stage[b * TS + r1] = w.B1[b];
    }
    stage[6 * TS + r0] = V0;
    stage[6 * TS + r1] = V1;
    acc.settled += mine;
    acc.resettled += static_cast<uint32_t>(lab_popc64(S0 & V0) + lab_popc64(S1 & V1));
    {
        uint32_t const k0 = S0 ? plane_max_level(w.B0, S0) : 0u, k1 = S1 ? plane_max_level(w.B1, S1) : 0u;
        uint32_t const top = base + (k0 > k1 ? k0 : k1);
        if (mine && top > acc.max_lvl) {
            acc.max_lvl = top;
        }
    }
    if (!EDGES) {
        lab_syncwarp(); // the stage is complete
        return total;
    }
    // columns 0 and 63: the lane that owns the row, straight from its registers
    uint64_t const edge_cols = (1ull << 63) | 1ull;
    if ((S0 | S1) & edge_cols) {
        if (S0 & 1ull) edge_update(tc.edges + SIDE_W * TS + r0, base + plane_level(w.B0, 0), V0 & 1ull, fW0, SIDE_W, acc);
        if (S1 & 1ull) edge_update(tc.edges + SIDE_W * TS + r1, base + plane_level(w.B1, 0), V1 & 1ull, fW1, SIDE_W, acc);
        if (S0 >> 63) edge_update(tc.edges + SIDE_E * TS + r0, base + plane_level(w.B0, 63), V0 >> 63, fE0, SIDE_E, acc);
        if (S1 >> 63) edge_update(tc.edges + SIDE_E * TS + r1, base + plane_level(w.B1, 63), V1 >> 63, fE1, SIDE_E, acc);
    }
    // rows 0 and 63: the whole warp, lane l looks after columns l and l+32
    uint64_t const Sn = lab_shfl64(S0, 0), Ss = lab_shfl64(S1, 31);
    lab_syncwarp(); // the stage is complete (also needed by emit_interior)
    if (Sn | Ss) {
#pragma unroll
        for (int side = 0; side < 2; side++) {
            uint64_t const S = side == 0 ? Sn : Ss;
            if (S == 0) continue;
            int const r = side == 0 ? 0 : TS - 1;
            uint64_t Bp[6];
#pragma unroll
            for (int b = 0; b < 6; b++) {
                Bp[b] = stage[b * TS + r];
            }
            uint64_t const Vold = stage[6 * TS + r];
#pragma unroll
            for (int half = 0; half < 2; half++) {
                int const c = lane + 32 * half;
                if ((S >> c) & 1ull) {
                    edge_update(tc.edges + (side == 0 ? SIDE_N : SIDE_S) * TS + c, base + plane_level(Bp, c), (Vold >> c) & 1ull,
                                side == 0 ? fNmax : fSmax, side == 0 ? SIDE_N : SIDE_S, acc);
                }
            }
        }
    }
    return total;
}

LAB_DEV void emit_interior(Tile_ctx const &tc, uint64_t *stage, uint64_t S0, uint64_t S1, uint32_t base, int lane, uint32_t total) {
    if (total == 0) {
        return;
    }
    int const r0 = 2 * lane, r1 = 2 * lane + 1;
    if (total >= DENSE_BATCH) {
        uint64_t *Srow = stage + STAGE_FIXED_U64; // the cell-list area is free in this mode
        Srow[r0] = S0;
        Srow[r1] = S1;
        lab_syncwarp();
        // two rows (2i, 2i+1) per iteration, branch free: an empty row simply stores nothing
        uint32_t pairs = lab_ballot((S0 | S1) != 0); // bit i <-> rows 2i, 2i+1
        while (pairs) {
            int const r = 2 * (lab_ffs64(pairs) - 1);
            pairs &= pairs - 1;
            uint64_t Sx[2], Vx[2], Bp[2][6];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                Sx[j] = Srow[r + j];
                Vx[j] = stage[6 * TS + r + j];
#pragma unroll
                for (int b = 0; b < 6; b++) {
                    Bp[j][b] = stage[b * TS + r + j];
                }
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {
                uint32_t *drow = tc.dist0 + (r + j) * tc.cols;
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    uint32_t const sbit = (static_cast<uint32_t>(Sx[j] >> (32 * half)) >> lane) & 1u;
                    uint32_t const old = (static_cast<uint32_t>(Vx[j] >> (32 * half)) >> lane) & 1u;
                    uint32_t k = 0;
#pragma unroll
                    for (int b = 0; b < 6; b++) {
                        k |= ((static_cast<uint32_t>(Bp[j][b] >> (32 * half)) >> lane) & 1u) << b;
                    }
                    if (sbit & ~old) {
                        drow[lane + 32 * half] = base + k;
                    }
                    if (sbit & old) {
                        lab_atomic_min(drow + lane + 32 * half, base + k);
                    }
                }
            }
        }
    } else {
        uint32_t const mine = static_cast<uint32_t>(lab_popc64(S0) + lab_popc64(S1));
        uint32_t incl = mine; // inclusive prefix sum of the per-lane counts
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t const v = lab_shfl_up(incl, d);
            if (lane >= d) {
                incl += v;
            }
        }
        uint16_t *cells = reinterpret_cast<uint16_t *>(stage + STAGE_FIXED_U64);
        {
            uint32_t q = incl - mine;
            uint64_t m = S0;
            while (m) {
                int const c = lab_ffs64(m) - 1;
                m &= m - 1;
                cells[q++] = static_cast<uint16_t>((r0 << 6) | c);
            }
            m = S1;
            while (m) {
                int const c = lab_ffs64(m) - 1;
                m &= m - 1;
                cells[q++] = static_cast<uint16_t>((r1 << 6) | c);
            }
        }
        lab_syncwarp();
        for (uint32_t q = static_cast<uint32_t>(lane); q < total; q += 64) {
            // two squares per lane and iteration (independent: their latencies overlap)
            bool const two = q + 32 < total;
            uint32_t const rcA = cells[q], rcB = two ? cells[q + 32] : rcA;
            int const rA = static_cast<int>(rcA >> 6), cA = static_cast<int>(rcA & 63u);
            int const rB = static_cast<int>(rcB >> 6), cB = static_cast<int>(rcB & 63u);
            uint32_t kA = 0, kB = 0;
#pragma unroll
            for (int b = 0; b < 6; b++) {
                kA |= static_cast<uint32_t>((stage[b * TS + rA] >> cA) & 1ull) << b;
                kB |= static_cast<uint32_t>((stage[b * TS + rB] >> cB) & 1ull) << b;
            }
            bool const oldA = (stage[6 * TS + rA] >> cA) & 1ull;
            bool const oldB = (stage[6 * TS + rB] >> cB) & 1ull;
            uint32_t *dA = tc.dist0 + rA * tc.cols + cA;
            uint32_t *dB = tc.dist0 + rB * tc.cols + cB;
            if (!oldA) {
                *dA = base + kA;
            } else {
                lab_atomic_min(dA, base + kA);
            }
            if (two) {
                if (!oldB) {
                    *dB = base + kB;
                } else {
                    lab_atomic_min(dB, base + kB);
                }
            }
        }
    }
    lab_syncwarp(); // the stage is reused by the next batch
}

// One border square: `seed` = facing level + 1 if that beats what this tile holds, else INF.
// f and o were loaded unconditionally; `crossing` says whether the border can be crossed here.
LAB_DEV void make_seed(bool crossing, uint32_t &f, uint32_t o, uint32_t &seed) {
    seed = INF;
    if (!crossing) {
        f = 0; // "never worth waking"
    } else if (f != INF && f + 1 < o) {
        seed = f + 1;
    }
}

LAB_DEV uint32_t seed_after(uint32_t s, uint32_t lvl) { return (s != INF && s > lvl) ? s : INF; }
LAB_DEV uint32_t seed_group_bit(uint32_t s, uint32_t base) {
    uint32_t const rel = s - base;
    return rel < static_cast<uint32_t>(WINDOW) ? (1u << (rel >> 3)) : 0u;
}

// Work counters a warp keeps in registers and adds to the Control block once, when it retires.
struct Warp_stats {
    uint32_t passes = 0, empty = 0, levels = 0, pushes = 0, direct = 0;
    uint32_t settled = 0, resettled = 0, max_lvl = 0; // per lane
    LAB_PROF(uint64_t c_load = 0; uint64_t c_loop = 0; uint64_t c_pub = 0; uint64_t c_step = 0; uint64_t c_emit = 0; uint64_t c_idle = 0; uint64_t batches = 0;)
};

LAB_DEV void flush_stats(Control *ctl, Warp_stats const &ws, int lane) {
    uint32_t const s_set = lab_reduce_add(ws.settled);
    uint32_t const s_res = lab_reduce_add(ws.resettled);
    uint32_t const s_max = lab_reduce_max(ws.max_lvl);
    if (lane == 0) {
        lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->n_activations), static_cast<uint64_t>(ws.passes));
        lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->n_empty), static_cast<uint64_t>(ws.empty));
        lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->n_levels), static_cast<uint64_t>(ws.levels));
        lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->n_pushes), static_cast<uint64_t>(ws.pushes));
        lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->n_direct), static_cast<uint64_t>(ws.direct));
        lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->n_settled), static_cast<uint64_t>(s_set));
        lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->n_resettles), static_cast<uint64_t>(s_res));
        lab_atomic_max(&ctl->max_level, s_max);
        LAB_PROF(lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->cyc_load), ws.c_load);
                 lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->cyc_loop), ws.c_loop);
                 lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->cyc_publish), ws.c_pub);
                 lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->cyc_step), ws.c_step);
                 lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->cyc_emit), ws.c_emit);
                 lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->cyc_idle), ws.c_idle);
                 lab_atomic_add(reinterpret_cast<uint64_t *>(&ctl->n_batches), ws.batches);)
    }
}

// Wake the neighbours flagged in `woke` (Side bits; lane s < 4 looks after side s, its tile is nb).
// A neighbour is taken with CAS(idle -> ACTIVE), so it starts with DIRTY clear and needs no clearing
// round trip before it reads the edges; one that is already active is marked DIRTY instead and
// whoever runs it goes round again.  Returns true in the lanes that now OWN their neighbour.
LAB_DEV bool wake_neighbours(Plane const &pl, Control *ctl, uint32_t woke, int lane, long long nb, bool skip) {
    bool const attempt = lane < 4 && ((woke >> lane) & 1u) && nb >= 0 && !skip;
    uint32_t const n_attempt = static_cast<uint32_t>(lab_popc64(lab_ballot(attempt)));
    if (n_attempt == 0) {
        return false;
    }
    if (lane == 0) {
        lab_atomic_add(&ctl->pending, n_attempt); // over-counts until the failures are taken back
    }
    lab_threadfence(); // the edge levels written so far are visible before any flag below
    lab_syncwarp();
    uint32_t r = 0;
    if (attempt) {
        r = lab_atomic_cas(&pl.state[nb], 0u, ST_ACTIVE);
    }
    bool const failed = attempt && r != 0u;
    bool got = attempt && !failed;
    if (lab_any(failed)) {
        if (failed) {
            uint32_t const old = lab_atomic_or(&pl.state[nb], ST_ACTIVE | ST_DIRTY);
            if (!(old & ST_ACTIVE)) {
                got = true; // it went idle in between: ours after all (it will do one spurious rerun)
            } else {
                lab_atomic_sub(&ctl->pending, 1u);
            }
        }
    }
    return got;
}

// returns the next item this warp should relax itself (ITEM_NONE if none)
LAB_DEV uint32_t relax_tile(Params const &p, uint32_t item, int lane, uint64_t *stage, Warp_stats &ws) {
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
    // edges has one ghost tile-row above and below, so every neighbour pointer stays inside the
    // allocation and the edge loads below can be issued unconditionally (one round trip)
    uint32_t *const my_edges = pl.edges + (t + g.tiles_x) * (4 * TS);
    uint32_t const *const n_edges = my_edges - static_cast<long long>(g.tiles_x) * (4 * TS);
    uint32_t const *const s_edges = my_edges + static_cast<long long>(g.tiles_x) * (4 * TS);
    uint32_t const *const w_edges = my_edges - 4 * TS;
    uint32_t const *const e_edges = my_edges + 4 * TS;
    bool const bounded = ctl->bound_mode != BOUND_NONE;

    Tile_ctx tc;
    tc.cols = g.cols;
    tc.edges = my_edges;
    tc.dist0 = pl.dist + (static_cast<long long>(ty) * TS) * g.cols + static_cast<long long>(tx) * TS;

    // neighbour this lane looks after when waking (lanes 0..3 = Side)
    long long nb = -1;
    if (lane == SIDE_N) nb = ty > 0 ? t - g.tiles_x : -1;
    if (lane == SIDE_S) nb = ty < g.tiles_y - 1 ? t + g.tiles_x : -1;
    if (lane == SIDE_W) nb = tx > 0 ? t - 1 : -1;
    if (lane == SIDE_E) nb = tx < g.tiles_x - 1 ? t + 1 : -1;

    Emit_acc acc{0u, 0u, 0u, 0u};
    bool first_pass = true;
    uint32_t guard = 0; // belt and braces: a pass that cannot end must become an error, never a hang

    for (;;) { // rerun while a neighbour dirtied us during the pass
        LAB_PROF(uint64_t const pt0 = lab_cycles();)
        if (!first_pass) {
            // the tile arrives with DIRTY clear (see the wake protocol below); on a rerun it must be
            // cleared BEFORE the edges are read again
            if (lane == 0) {
                lab_atomic_and(&pl.state[t], ~ST_DIRTY);
            }
            lab_threadfence();
            lab_syncwarp();
        }
        first_pass = false;

        uint64_t const V0 = lab_ld_cg(pl.visited + t * TS + r0);
        uint64_t const V1 = lab_ld_cg(pl.visited + t * TS + r1);

        // ---- seeds: facing edge level + 1 where that beats what this tile has ------------------
        // N/S seeds: lane holds columns lane and lane+32.  W/E seeds: lane holds rows 2*lane, 2*lane+1.
        uint32_t fN0 = lab_ld_cg(n_edges + SIDE_S * TS + lane), fN1 = lab_ld_cg(n_edges + SIDE_S * TS + lane + 32);
        uint32_t fS0 = lab_ld_cg(s_edges + SIDE_N * TS + lane), fS1 = lab_ld_cg(s_edges + SIDE_N * TS + lane + 32);
        uint32_t fW0 = lab_ld_cg(w_edges + SIDE_E * TS + r0), fW1 = lab_ld_cg(w_edges + SIDE_E * TS + r1);
        uint32_t fE0 = lab_ld_cg(e_edges + SIDE_W * TS + r0), fE1 = lab_ld_cg(e_edges + SIDE_W * TS + r1);
        uint32_t const oN0 = lab_ld_cg(my_edges + SIDE_N * TS + lane), oN1 = lab_ld_cg(my_edges + SIDE_N * TS + lane + 32);
        uint32_t const oS0 = lab_ld_cg(my_edges + SIDE_S * TS + lane), oS1 = lab_ld_cg(my_edges + SIDE_S * TS + lane + 32);
        uint32_t const oW0 = lab_ld_cg(my_edges + SIDE_W * TS + r0), oW1 = lab_ld_cg(my_edges + SIDE_W * TS + r1);
        uint32_t const oE0 = lab_ld_cg(my_edges + SIDE_E * TS + r0), oE1 = lab_ld_cg(my_edges + SIDE_E * TS + r1);
        uint32_t sN0, sN1, sS0, sS1, sW0, sW1, sE0, sE1;
        make_seed((XN >> lane) & 1, fN0, oN0, sN0);
        make_seed((XN >> (lane + 32)) & 1, fN1, oN1, sN1);
        make_seed((XS >> lane) & 1, fS0, oS0, sS0);
        make_seed((XS >> (lane + 32)) & 1, fS1, oS1, sS1);
        make_seed((XW >> r0) & 1, fW0, oW0, sW0);
        make_seed((XW >> r1) & 1, fW1, oW1, sW1);
        make_seed((XE >> r0) & 1, fE0, oE0, sE0);
        make_seed((XE >> r1) & 1, fE1, oE1, sE1);
        // The plane's source square, if it lies in this tile and is not settled yet, is a level-0 seed.
        Seeds sd;
        sd.src0 = 0;
        sd.src1 = 0;
        if (ctl->src_tile[plane] == static_cast<uint32_t>(t)) {
            uint32_t const rc = ctl->src_rc[plane];
            int const sr = static_cast<int>(rc >> 8), sc = static_cast<int>(rc & 0xFF);
            if (sr == r0 && !((V0 >> sc) & 1)) sd.src0 = 1ull << sc;
            if (sr == r1 && !((V1 >> sc) & 1)) sd.src1 = 1ull << sc;
        }
        uint32_t sSrc = (sd.src0 | sd.src1) ? 0u : INF;

        uint32_t pending_min = lab_umin(lab_umin(lab_min3(sN0, sN1, sS0), lab_min3(sS1, sW0, sW1)), lab_min3(sE0, sE1, sSrc));
        pending_min = lab_reduce_min(pending_min);
        ws.passes++;
        LAB_PROF(uint64_t const pt1 = lab_cycles(); ws.c_load += pt1 - pt0;)

        acc.wake = 0;
        if (pending_min == INF) {
            ws.empty++;
        } else {
            // a row-0 / row-63 square settled at level L can only help the tile above / below if L + 1
            // beats the WORST level over there (conservative; the exact per-column test is not needed)
            uint32_t const fNmax = XN ? lab_reduce_max(fN0 > fN1 ? fN0 : fN1) : 0u;
            uint32_t const fSmax = XS ? lab_reduce_max(fS0 > fS1 ? fS0 : fS1) : 0u;
            bool const any_ns = lab_any((sN0 & sN1 & sS0 & sS1) != INF);
            Wave w;
            w.A0 = P0;
            w.A1 = P1;
            w.F0 = 0;
            w.F1 = 0;
            uint64_t Vn0 = V0, Vn1 = V1;
            uint32_t base = pending_min; // level of the first iteration of the batch
            uint32_t bound = bounded ? lab_ld_uniform(&ctl->bound, lane) : INF;
            uint32_t batches = 0;
            while (base <= bound) { // ---- one batch of up to WINDOW levels ------------------------
                if (++batches > (1u << 22)) {
                    lab_atomic_max(&ctl->error, 4u);
                    lab_st_volatile(&ctl->done, 1u);
                    break;
                }
#pragma unroll
                for (int b = 0; b < 6; b++) {
                    w.B0[b] = 0;
                    w.B1[b] = 0;
                }
                uint64_t const A0s = w.A0, A1s = w.A1;
                sd.rW0 = sW0 - base;
                sd.rW1 = sW1 - base;
                sd.rE0 = sE0 - base;
                sd.rE1 = sE1 - base;
                sd.rSrc = sSrc - base;
                // which groups of 8 levels of this window hold a seed (uniform 8-bit mask)
                uint32_t grp = seed_group_bit(sW0, base) | seed_group_bit(sW1, base) | seed_group_bit(sE0, base)
                               | seed_group_bit(sE1, base) | seed_group_bit(sSrc, base);
                sd.nsv = 0;
#pragma unroll
                for (int b = 0; b < 6; b++) {
                    sd.nsp[b] = 0;
                }
                bool ns_window = false; // some N/S seed falls into this batch's window
                if (any_ns) {
                    // N seeds land in row 0 (lane 0, r0), S seeds in row 63 (lane 31, r1): bit-sliced by ballots
                    uint32_t const rN0 = sN0 - base, rN1 = sN1 - base, rS0 = sS0 - base, rS1 = sS1 - base;
                    bool const vN0 = rN0 < WINDOW, vN1 = rN1 < WINDOW, vS0 = rS0 < WINDOW, vS1 = rS1 < WINDOW;
                    uint64_t const wn = lab_ballot64(vN0, vN1), ws = lab_ballot64(vS0, vS1);
                    sd.nsv = lane == 0 ? wn : (lane == 31 ? ws : 0ull);
                    if (wn | ws) {
                        ns_window = true;
                        grp |= seed_group_bit(sN0, base) | seed_group_bit(sN1, base) | seed_group_bit(sS0, base)
                               | seed_group_bit(sS1, base);
#pragma unroll
                        for (int b = 0; b < 6; b++) {
                            uint64_t const pn = lab_ballot64(vN0 && ((rN0 >> b) & 1), vN1 && ((rN1 >> b) & 1));
                            uint64_t const ps = lab_ballot64(vS0 && ((rS0 >> b) & 1), vS1 && ((rS1 >> b) & 1));
                            sd.nsp[b] = lane == 0 ? pn : (lane == 31 ? ps : 0ull);
                        }
                    }
                }
                grp = lab_reduce_or(grp);
                LAB_PROF(uint64_t const ps0 = lab_cycles(); ws.batches++;)
                int gi = 0; // groups of 8 levels completed
                for (; gi < WINDOW / 8; gi++) {
                    if (base + 8u * static_cast<uint32_t>(gi) > bound) {
                        break;
                    }
                    if (!((grp >> gi) & 1u)) {
                        step_group<false, false>(w, sd, gi, lane);
                    } else if (ns_window) {
                        step_group<true, true>(w, sd, gi, lane);
                    } else {
                        step_group<true, false>(w, sd, gi, lane);
                    }
                    ws.levels += 8;
                    if (!lab_any((w.F0 | w.F1) != 0)) {
                        // nothing moving: skip to the next group of this window that holds a seed, if any
                        uint32_t const later = grp & ~((2u << gi) - 1u);
                        if (later == 0 || base + 8u * static_cast<uint32_t>(lab_ffs64(later) - 1) > bound) {
                            gi++;
                            break;
                        }
                        gi = lab_ffs64(later) - 2; // the loop increment lands on that group
                    }
                }
                // ---- write the batch out ---------------------------------------------------------
                LAB_PROF(uint64_t const ps1 = lab_cycles(); ws.c_step += ps1 - ps0;)
                uint64_t const S0 = A0s & ~w.A0, S1 = A1s & ~w.A1;
                uint32_t const n_batch = emit_borders<true>(tc, stage, w, S0, S1, V0, V1, base, lane, fW0, fW1, fE0, fE1, fNmax, fSmax, acc);
                // let the neighbours this batch reached start NOW, on other warps: they only read our
                // edge arrays, which are complete; the interior levels are written out afterwards
                uint32_t const woke_now = lab_reduce_or(acc.wake);
                if (woke_now) {
                    acc.wake = 0;
                    bool const got = wake_neighbours(pl, ctl, woke_now, lane, nb, false);
                    if (got) {
                        queue_push(p, static_cast<uint32_t>(plane) * static_cast<uint32_t>(g.ntiles) + static_cast<uint32_t>(nb));
                    }
                    ws.pushes += static_cast<uint32_t>(lab_popc64(lab_ballot(got)));
                }
                emit_interior(tc, stage, S0, S1, base, lane, n_batch);
                Vn0 |= S0;
                Vn1 |= S1;
                LAB_PROF(ws.c_emit += lab_cycles() - ps1;)
                if (gi == 0) {
                    break; // bound reached before anything ran
                }
                uint32_t const done_lvl = base + 8u * static_cast<uint32_t>(gi) - 1u;
                sN0 = seed_after(sN0, done_lvl);
                sN1 = seed_after(sN1, done_lvl);
                sS0 = seed_after(sS0, done_lvl);
                sS1 = seed_after(sS1, done_lvl);
                sW0 = seed_after(sW0, done_lvl);
                sW1 = seed_after(sW1, done_lvl);
                sE0 = seed_after(sE0, done_lvl);
                sE1 = seed_after(sE1, done_lvl);
                sSrc = seed_after(sSrc, done_lvl);
                if (bounded) {
                    bound = lab_ld_uniform(&ctl->bound, lane);
                }
                if (lab_any((w.F0 | w.F1) != 0)) {
                    base = done_lvl + 1; // the wave is still moving: next batch continues from it
                } else {
                    pending_min = lab_umin(lab_umin(lab_min3(sN0, sN1, sS0), lab_min3(sS1, sW0, sW1)), lab_min3(sE0, sE1, sSrc));
                    pending_min = lab_reduce_min(pending_min);
                    if (pending_min == INF) {
                        break;
                    }
                    base = pending_min; // jump to the next seed level
                }
                if (base > bound) {
                    break;
                }
            }
            if ((Vn0 != V0) | (Vn1 != V1)) {
                lab_st_cg(pl.visited + t * TS + r0, Vn0);
                lab_st_cg(pl.visited + t * TS + r1, Vn1);
            }
        }
        bool const settled_any = lab_any(acc.settled != 0);
        LAB_PROF(uint64_t const pt2 = lab_cycles(); ws.c_loop += pt2 - pt1;)

        // ---- publish: the levels and visited bits above must be visible before the tile is let go ---
        if (settled_any) {
            lab_threadfence();
        }
        lab_syncwarp();

        // targets in this tile: refresh their values and the search bound
        uint32_t const ntg = ctl->n_targets;
        if (ntg && lane == 0 && settled_any) {
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

        // let this tile go idle; if somebody marked it DIRTY meanwhile, go round again
        uint32_t r = 0;
        if (lane == 4) {
            r = lab_atomic_cas(&pl.state[t], ST_ACTIVE, 0u);
        }
        uint32_t const r4 = lab_shfl(r, 4); // what this tile's state word held when we tried to let go
        bool const again = r4 != ST_ACTIVE;
        LAB_PROF(ws.c_pub += lab_cycles() - pt2;)
        if (!again) {
            break;
        }
        if (++guard > 4096u || lab_ld_uniform(&ctl->done, lane)) {
            if (lane == 0) {
                lab_atomic_max(&ctl->error, 3u + (r4 << 8));
                lab_st_volatile(&ctl->done, 1u);
            }
            break;
        }
    }
    LAB_PROF(uint64_t const pt3 = lab_cycles();)

    ws.settled += acc.settled;
    ws.resettled += acc.resettled;
    ws.max_lvl = acc.max_lvl > ws.max_lvl ? acc.max_lvl : ws.max_lvl;
    if (lane == 0) {
        lab_atomic_sub(&ctl->pending, 1u); // this tile is idle now
    }
    LAB_PROF(ws.c_pub += lab_cycles() - pt3;)
    return ITEM_NONE; // every woken neighbour went through the queue
}

// The persistent worker: every warp of the launch runs this until no tile is active any more.
// `stage` = this warp's STAGE_BYTES of shared memory.
LAB_DEV void bfs_worker_body(Params const &p, int lane, uint64_t *stage) {
    Control *ctl = p.ctl;
    uint32_t item = ITEM_NONE;
    Warp_stats ws;
    for (;;) {
        if (item == ITEM_NONE) {
            LAB_PROF(uint64_t const pi0 = lab_cycles();)
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
                    // pending only ever OVER-counts the active tiles, so zero means the search is over
                    if ((polls & 3u) == 0u && (lab_ld_volatile(&ctl->pending) == 0u || lab_ld_volatile(&ctl->done))) {
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
                    if (backoff < 512) backoff *= 2;
                    polls++;
                }
            }
            item = lab_shfl(item, 0);
            LAB_PROF(ws.c_idle += lab_cycles() - pi0;)
            if (item == ITEM_EXIT) {
                break;
            }
        }
        item = relax_tile(p, item, lane, stage, ws);
        if (item != ITEM_NONE && lab_ld_uniform(&ctl->done, lane)) {
            break; // only after an error (deadline): give up, the host reports it
        }
    }
    flush_stats(ctl, ws, lane);
}

} // namespace lab
