// post.cuh -- what happens to the BFS fields after (and just before) the tile search:
// run initialisation, the Square-grid side effects of each entry point, and path extraction.
//
//   distance  painters/distance.cc:138,149     Rgb::measure (bit 9) on every reached square
//   hunt      solvers/bfs_threads.cc:61,263-274 paint bits, winner's path repainted in one colour
//   gather    solvers/bfs_threads.cc:158-164,355-364  paint+cache bits, claims, path.front() recolour
//   corners   solvers/bfs_threads.cc:61          paint bits only (static game does not repaint)
//
// The racy parts of the reference are replaced by the canonical lock-step model of
// SURVEY.md App. A.4 (restated on the CPU in oracle/restate.cc orc_canon_*): colour k paints
// exactly the squares whose level in k's field is below the level at which k's game ends.
#pragma once
#include "engine.cuh"
#include "room.cuh"

namespace lab {

enum Game : int { GAME_DISTANCE = 0, GAME_HUNT = 1, GAME_GATHER = 2, GAME_CORNERS = 3 };

struct Paint_rule {
    uint32_t plane; // which BFS field
    uint32_t below; // paint squares whose level is < below
    uint32_t bits;  // Square bits to OR in
    uint32_t pad;
};

// Per-run results, device resident; the host reads it once at the end.
struct Result {
    int32_t winner;             // hunt/corners: winning colour, -1 none
    int32_t claimed_by[4];      // gather: colour that claims finish j, -1 unreachable
    int32_t finish_of[4];       // gather: finish index colour k claims, -1 none
    uint32_t game_level[4];     // level at which colour k's game ends (D_k), INF if never
    unsigned long long path_len[4];
    unsigned long long settled; // squares with a level (distance: map size)
    unsigned long long max_level;
    uint32_t n_rules;
    uint32_t pad;
    Paint_rule rule[4];
    // where walk `slot` stopped, and the level it still had to go (0: it reached the start).  Row bands: a walk that
    // runs out of its band stops at the band's border with walk_left > 0 and goes on in the neighbouring band.
    int32_t walk_r[4], walk_c[4];
    uint32_t walk_left[4];
};

// Host-filled description of one search.
struct Run_desc {
    int game;
    int nplanes;
    int32_t src_r[MAX_PLANES], src_c[MAX_PLANES]; // -1: this band has no source for plane k
    int n_targets;
    int32_t tgt_plane[MAX_TARGETS], tgt_r[MAX_TARGETS], tgt_c[MAX_TARGETS];
    uint32_t bound_mode;
    uint32_t spec_mark; // Square bits pack_body has set speculatively on every passable square (0: none)
};

// One thread.  Control and the queue were memset by the host (0 / 0xFF); this seeds the sources.
LAB_DEV void init_run_body(Params p, Run_desc d, unsigned long long budget_ns) {
    Control *ctl = p.ctl;
    ctl->bound = INF;
    ctl->spec_mark = d.spec_mark;
    ctl->bound_mode = d.bound_mode;
    ctl->n_targets = static_cast<uint32_t>(d.n_targets);
    for (int j = 0; j < d.n_targets; j++) {
        ctl->target_plane[j] = static_cast<uint32_t>(d.tgt_plane[j]);
        ctl->target_tile[j] = static_cast<uint32_t>((d.tgt_r[j] / TS) * p.g.tiles_x + d.tgt_c[j] / TS);
        ctl->target_cell[j] = static_cast<unsigned long long>(d.tgt_r[j]) * p.g.cols + d.tgt_c[j];
        ctl->target_val[j] = INF;
    }
    for (int k = 0; k < MAX_PLANES; k++) {
        ctl->src_tile[k] = 0xFFFFFFFFu;
        ctl->src_rc[k] = 0;
    }
    uint32_t pending = 0;
    for (int k = 0; k < d.nplanes; k++) {
        if (d.src_r[k] < 0) {
            continue;
        }
        uint32_t const t = static_cast<uint32_t>((d.src_r[k] / TS) * p.g.tiles_x + d.src_c[k] / TS);
        ctl->src_tile[k] = t;
        ctl->src_rc[k] = static_cast<uint32_t>(((d.src_r[k] % TS) << 8) | (d.src_c[k] % TS));
        p.plane[k].state[t] = ST_ACTIVE; // arrives like any woken tile: active, DIRTY clear
        p.queue[ctl->tail & p.qmask] = static_cast<uint32_t>(k) * static_cast<uint32_t>(p.g.ntiles) + t;
        ctl->tail++;
        pending++;
    }
    ctl->pending = pending;
    ctl->done = pending == 0 ? 1u : 0u;
    ctl->deadline_ns = lab_clock_ns() + budget_ns;
}

// Sources are passable whatever their Square says (the reference pushes its start unconditionally).
// One thread, between pack and cross.
LAB_DEV void force_sources_body(Geom g, uint64_t *path, Run_desc d, uint16_t *grid) {
    for (int k = 0; k < d.nplanes; k++) {
        if (d.src_r[k] >= 0) {
            source_force_passable(g, path, d.src_r[k], d.src_c[k]);
            if (d.spec_mark) { // (the start is marked whatever its bits: distance.cc:138)
                grid[static_cast<long long>(d.src_r[k]) * g.cols + d.src_c[k]] |= static_cast<uint16_t>(d.spec_mark);
            }
        }
    }
}

// distance epilogue: measure bit on every visited square + exact count of visited squares.
// Same row/word item space as pack_body.
LAB_DEV void measure_body(Geom g, uint16_t *grid, uint64_t const *visited, uint64_t const *path, Control const *ctl, Result *res,
                          long long gwarp, long long nwarps, int lane) {
    if (ctl->room) {
        return; // an open room: room_distance_body has set the bits and counted the squares
    }
    if (ctl->measured) { // one tree: tree_fixup_body (or pack_body, speculatively) has set the bits, this only reports how many
        if (gwarp == 0 && lane == 0) {
            lab_atomic_add(reinterpret_cast<uint64_t *>(&res->settled), static_cast<uint64_t>(ctl->n_marked));
        }
        return;
    }
    // pack_body set the bit on every passable square in advance (spec_mark): it has to go again where the search
    // did not come (parts of the maze cut off from the start)
    bool const spec = ctl->spec_mark != 0;
    // one warp per row (grid stride); a 64-square word with no visited square costs one 8-byte load
    unsigned long long count = 0;
    constexpr int U = 4; // 64-square words of a row handled together: independent read-modify-writes in flight
    long long const chunks = (g.tiles_x + U - 1) / U;
    for (long long it = gwarp; it < static_cast<long long>(g.rows) * chunks; it += nwarps) {
        int const r = static_cast<int>(it / chunks);
        int const j0 = static_cast<int>(it % chunks) * U;
        long long const w0 = (static_cast<long long>(r / TS) * g.tiles_x) * TS + (r % TS);
        uint16_t *row = grid + static_cast<long long>(r) * g.cols;
        uint64_t v[U], x[U]; // reached / passable but not reached
#pragma unroll
        for (int u = 0; u < U; u++) {
            v[u] = j0 + u < g.tiles_x ? visited[w0 + static_cast<long long>(j0 + u) * TS] : 0ull;
            x[u] = (spec && j0 + u < g.tiles_x) ? (path[w0 + static_cast<long long>(j0 + u) * TS] & ~v[u]) : 0ull;
        }
        uint16_t a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            int const c0 = (j0 + u) * TS + lane;
            a[u] = (((v[u] | x[u]) >> lane) & 1) ? row[c0] : static_cast<uint16_t>(0);
            b[u] = (((v[u] | x[u]) >> (lane + 32)) & 1) ? row[c0 + 32] : static_cast<uint16_t>(0);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            int const c0 = (j0 + u) * TS + lane;
            if ((v[u] >> lane) & 1) {
                row[c0] = a[u] | SQ_MEASURE;
            } else if ((x[u] >> lane) & 1) {
                row[c0] = a[u] & static_cast<uint16_t>(~SQ_MEASURE);
            }
            if ((v[u] >> (lane + 32)) & 1) {
                row[c0 + 32] = b[u] | SQ_MEASURE;
            } else if ((x[u] >> (lane + 32)) & 1) {
                row[c0 + 32] = b[u] & static_cast<uint16_t>(~SQ_MEASURE);
            }
            count += static_cast<unsigned long long>(lab_popc64(v[u]));
        }
    }
    if (lane == 0 && count) {
        lab_atomic_add(reinterpret_cast<uint64_t *>(&res->settled), static_cast<uint64_t>(count));
    }
}

// exact maximum level (only needed when some square was settled more than once)
// (res->max_level was zeroed by max_level_prepare_body when the exact pass is needed)
LAB_DEV void max_level_body(Geom g, uint32_t const *dist, Control const *ctl, Result *res, long long gthread,
                            long long nthreads) {
    if (ctl->n_resettles == 0) {
        return; // every square was settled once: the running maximum is already exact
    }
    long long const n = static_cast<long long>(g.rows) * g.cols;
    uint32_t m = 0;
    for (long long i = gthread; i < n; i += nthreads) {
        uint32_t const d = dist[i];
        if (d != INF && d > m) {
            m = d;
        }
    }
    m = lab_reduce_max(m);
    if ((gthread & 31) == 0 && m) {
        lab_atomic_max(reinterpret_cast<uint32_t *>(&res->max_level), m); // low word of the u64 (little endian)
    }
}

LAB_DEV void max_level_prepare_body(Control const *ctl, Result *res) {
    res->max_level = ctl->n_resettles == 0 ? ctl->max_level : 0ull;
}

// One thread: turn the targets' levels into the paint rules of the game (orc_canon_* in the oracle).
LAB_DEV void resolve_body(Params p, Run_desc d, Result *res, int32_t const *finish_rc /* n_targets (r,c) */) {
    Control const *ctl = p.ctl;
    res->winner = -1;
    res->n_rules = 0;
    for (int k = 0; k < 4; k++) {
        res->claimed_by[k] = -1;
        res->finish_of[k] = -1;
        res->game_level[k] = INF;
        res->path_len[k] = 0;
    }
    if (d.game == GAME_DISTANCE) {
        max_level_prepare_body(ctl, res);
        return;
    }
    if (d.game == GAME_HUNT) {
        uint32_t const T = ctl->target_val[0];
        res->winner = T != INF ? 0 : -1;
        res->game_level[0] = T;
        res->n_rules = 1;
        res->rule[0] = {0u, T, SQ_PAINT_MASK, 0u};
        return;
    }
    if (d.game == GAME_CORNERS) {
        uint32_t T = INF;
        int w = -1;
        for (int k = 0; k < 4; k++) {
            if (ctl->target_val[k] < T) {
                T = ctl->target_val[k];
                w = k;
            }
        }
        res->winner = w;
        res->n_rules = 4;
        for (int k = 0; k < 4; k++) {
            res->game_level[k] = T;
            res->rule[k] = {static_cast<uint32_t>(k), T, (1u << k) << SQ_PAINT_SHIFT, 0u};
        }
        return;
    }
    // gather: finishes sorted by (level, row, col); colour k claims the k-th
    int order[4] = {0, 1, 2, 3};
    for (int a = 1; a < 4; a++) {
        for (int b = a; b > 0; b--) {
            int const x = order[b - 1], y = order[b];
            uint32_t const dx = ctl->target_val[x], dy = ctl->target_val[y];
            bool less = dy < dx;
            if (dy == dx) {
                less = finish_rc[2 * y] < finish_rc[2 * x]
                       || (finish_rc[2 * y] == finish_rc[2 * x] && finish_rc[2 * y + 1] < finish_rc[2 * x + 1]);
            }
            if (!less) {
                break;
            }
            order[b - 1] = y;
            order[b] = x;
        }
    }
    res->n_rules = 4;
    for (int k = 0; k < 4; k++) {
        uint32_t const D = ctl->target_val[order[k]];
        res->game_level[k] = D;
        if (D != INF) {
            res->claimed_by[order[k]] = k;
            res->finish_of[k] = order[k];
        }
        res->rule[k] = {0u, D, ((1u << k) << SQ_PAINT_SHIFT) | ((1u << k) << SQ_CACHE_SHIFT), 0u};
    }
}

// Elementwise: OR each rule's bits into the squares whose level in the rule's field is below the
// rule's threshold; counts the squares that received any bit into res->settled.
LAB_DEV void paint_body(Params p, uint16_t *grid, Result *res, long long gthread, long long nthreads) {
    if (p.ctl->painted) { // lat_levels_body<LAT_PAINT> applied the rules on its way
        return;
    }
    long long const n = static_cast<long long>(p.g.rows) * p.g.cols;
    uint32_t const nr = res->n_rules;
    Paint_rule r[4];
    for (uint32_t k = 0; k < 4; k++) {
        r[k] = res->rule[k < nr ? k : 0];
    }
    if (p.ctl->room) { // an open room: no field was searched, the levels are a closed form (room.cuh)
        Room_rules q{};
        q.n = nr;
        for (uint32_t k = 0; k < 4; k++) {
            q.below[k] = r[k].below;
            q.bits[k] = r[k].bits;
        }
        room_paint_body(p, q, grid, &res->settled, gthread, nthreads);
        return;
    }
    uint32_t painted = 0;
    bool const same_plane = nr <= 1 || (r[0].plane == r[1].plane && r[0].plane == r[2].plane && r[0].plane == r[3].plane);
    // All arrays on 16-byte boundaries: eight squares per thread and step -- per field two 16-byte loads of levels (one field
    // for hunt and gather, one per colour for corners), one 16-byte read-modify-write of Squares -- and a group whose levels
    // are all at or beyond the thresholds is not touched at all.
    bool vec_ok = nr >= 1 && (reinterpret_cast<uintptr_t>(grid) & 15u) == 0;
    for (uint32_t k = 0; k < nr; k++) {
        vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(p.plane[r[k].plane].dist) & 15u) == 0;
    }
    if (vec_ok) {
        uint32_t const nfields = same_plane ? 1u : nr; // (same_plane: every rule looks at field r[0].plane)
        long long const groups = n / 8;
        for (long long gi = gthread; gi < groups; gi += nthreads) {
            uint32_t bits[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            lab_u32x4 a[4], b[4];
            for (uint32_t f = 0; f < nfields; f++) { // (all the loads first: independent)
                uint32_t const *d = p.plane[r[f].plane].dist;
                a[f] = lab_ld_stream_v4(d + gi * 8);
                b[f] = lab_ld_stream_v4(d + gi * 8 + 4);
            }
            for (uint32_t f = 0; f < nfields; f++) {
                uint32_t const lv[8] = {a[f].x, a[f].y, a[f].z, a[f].w, b[f].x, b[f].y, b[f].z, b[f].w};
                for (uint32_t k = same_plane ? 0u : f; k < (same_plane ? nr : f + 1u); k++) {
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        bits[q] |= lv[q] < r[k].below ? r[k].bits : 0u;
                    }
                }
            }
            uint32_t const any = bits[0] | bits[1] | bits[2] | bits[3] | bits[4] | bits[5] | bits[6] | bits[7];
            if (!any) {
                continue;
            }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                painted += bits[q] ? 1u : 0u;
            }
            lab_u32x4 sq = lab_ld_stream_v4(grid + gi * 8);
            sq.x |= bits[0] | (bits[1] << 16);
            sq.y |= bits[2] | (bits[3] << 16);
            sq.z |= bits[4] | (bits[5] << 16);
            sq.w |= bits[6] | (bits[7] << 16);
            lab_st_v4(grid + gi * 8, sq);
        }
        for (long long i = groups * 8 + gthread; i < n; i += nthreads) { // the last few squares
            uint32_t bits1 = 0;
            for (uint32_t k = 0; k < nr; k++) {
                bits1 |= p.plane[r[k].plane].dist[i] < r[k].below ? r[k].bits : 0u;
            }
            if (bits1) {
                grid[i] |= static_cast<uint16_t>(bits1);
                painted++;
            }
        }
        painted = lab_reduce_add(painted);
        if ((gthread & 31) == 0 && painted) {
            lab_atomic_add(reinterpret_cast<uint64_t *>(&res->settled), static_cast<uint64_t>(painted));
        }
        return;
    }
    constexpr int U = 8; // independent loads in flight per thread (a streaming pass: HBM latency x bandwidth)
    for (long long i0 = gthread; i0 < n; i0 += nthreads * U) {
        uint32_t dv[U];
        uint32_t bits[U];
        if (same_plane) {
            uint32_t const *d = p.plane[r[0].plane].dist;
#pragma unroll
            for (int u = 0; u < U; u++) {
                long long const i = i0 + u * nthreads;
                dv[u] = i < n ? d[i] : INF;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                bits[u] = 0;
                for (uint32_t k = 0; k < nr; k++) {
                    bits[u] |= dv[u] < r[k].below ? r[k].bits : 0u;
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; u++) {
                long long const i = i0 + u * nthreads;
                bits[u] = 0;
                for (uint32_t k = 0; k < nr; k++) {
                    bits[u] |= (i < n && p.plane[r[k].plane].dist[i] < r[k].below) ? r[k].bits : 0u;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (bits[u]) { // (levels < threshold only exist inside the grid)
                grid[i0 + u * nthreads] |= static_cast<uint16_t>(bits[u]);
                painted++;
            }
        }
    }
    painted = lab_reduce_add(painted);
    if ((gthread & 31) == 0 && painted) {
        lab_atomic_add(reinterpret_cast<uint64_t *>(&res->settled), static_cast<uint64_t>(painted));
    }
}

// Canonical shortest path (orc canonical_path): from the finish step to the FIRST neighbour in
// N,E,S,W order whose level is one less.  One warp per path; lanes 0..3 look at the 4 neighbours.
// out (may be null) receives (row,col) pairs finish-side first, the start last.
// If repaint_bits != 0 every path square gets its paint nibble replaced by repaint_bits (hunt).
// Returns via res->path_len[slot].  `first_only` stops after the first step (gather's path.front()).
//
// The walk is one dependent chain, 10^4 .. 10^6 steps long on the reference's mazes; with a global load (and the
// Square's read-modify-write) on that chain a step costs a DRAM round trip, and the walk used to take 7 x as long
// as the search that made the field.  So: the 64x64 tile of levels under the walker is kept in shared memory (a
// step inside it is four LDS + a vote; only border squares look outside), and nothing is written on the chain --
// path squares are collected in shared memory and flushed WALK_BATCH at a time with all lanes in flight.
constexpr int WALK_BATCH = 1024;
constexpr int WALK_W = TS + 2;                            // the cached tile with a one-square ring around it
constexpr int WALK_SM_U32 = WALK_W * WALK_W + WALK_BATCH; // per warp: cached levels | collected path squares (flat indices)

LAB_DEV void walk_flush(Geom const &g, uint32_t const *batch, uint32_t nb, unsigned long long base, uint16_t *grid, int32_t *out,
                        unsigned long long max_out, uint32_t repaint_bits, int lane) {
    lab_syncwarp();
    constexpr uint32_t U = 8; // squares per lane whose Squares are in flight together
    for (uint32_t i0 = static_cast<uint32_t>(lane); i0 < nb; i0 += 32u * U) {
        uint32_t idx[U];
        uint16_t sq[U];
#pragma unroll
        for (uint32_t u = 0; u < U; u++) {
            uint32_t const i = i0 + 32u * u;
            idx[u] = i < nb ? batch[i] : 0xFFFFFFFFu;
            sq[u] = (repaint_bits && idx[u] != 0xFFFFFFFFu) ? grid[idx[u]] : static_cast<uint16_t>(0);
        }
#pragma unroll
        for (uint32_t u = 0; u < U; u++) {
            uint32_t const i = i0 + 32u * u;
            if (idx[u] == 0xFFFFFFFFu) {
                continue;
            }
            if (out && base + i < max_out) {
                uint32_t const r = idx[u] / static_cast<uint32_t>(g.cols);
                out[2 * (base + i)] = static_cast<int32_t>(r);
                out[2 * (base + i) + 1] = static_cast<int32_t>(idx[u] - r * static_cast<uint32_t>(g.cols));
            }
            if (repaint_bits) {
                grid[idx[u]] = static_cast<uint16_t>((sq[u] & ~SQ_PAINT_MASK) | repaint_bits);
            }
        }
    }
    lab_syncwarp();
}

// include_first: (fr, fc) is itself a square of the path (a walk taken over from the neighbouring band), not the finish.
// confine: ONE SEGMENT of a path (lattice.cuh, the winner's path without the walk): the walk stays inside the 64x64 tile of
// (fr, fc) and stops where its next step would leave it; its squares go to out[out_base ...]; `res` is not touched.
LAB_DEV void walk_body(Params p, int plane, int fr, int fc, uint16_t *grid, int32_t *out,
                       unsigned long long max_out, uint32_t repaint_bits, int first_only, Result *res, int slot,
                       int lane, uint32_t *sm, int include_first = 0, unsigned long long out_base = 0, int confine = 0) {
    Geom const &g = p.g;
    uint32_t const *dist = p.plane[plane].dist;
    bool const room = p.ctl->room != 0; // (single-field games only: the levels are room_level, no field exists)
    Room const m = room_load(p.ctl);
    uint32_t *cache = sm, *batch = sm + WALK_W * WALK_W;
    int r = fr, c = fc; // (rows, cols < 2^16: make_geom)
    // (in a room the finish may lie in another band, outside this one's rows: the closed form does not care)
    uint32_t const d0 = room ? room_absdiff(r, m.sr) + room_absdiff(c, m.sc) : dist[static_cast<long long>(r) * g.cols + c];
    uint32_t d = d0;
    unsigned long long flushed = out_base;
    uint32_t nb = 0;
    if (d == INF) {
        if (lane == 0 && !confine) {
            res->path_len[slot] = 0;
            res->walk_r[slot] = r;
            res->walk_c[slot] = c;
            res->walk_left[slot] = INF;
        }
        return;
    }
    // Row bands: the squares just above / below the band are the neighbouring bands' border rows; where the tile wave made
    // the field their levels are in the ghost tile rows of the edge arrays (band_import_edges_body).  The canonical step
    // (first of N, E, S, W one level lower) must see them: on a maze with loops a square of the border row can have a lower
    // neighbour across the border AND one inside the band.  (Anywhere else the ghosts hold INF and change nothing.)
    uint32_t const *edges = p.plane[plane].edges;
    auto halo = [&](long long rr, long long cc) -> uint32_t {
        if (!edges || cc < 0 || cc >= g.cols) return INF;
        if (rr == -1) return lab_ld_cg(edges + (cc >> 6) * (4 * TS) + SIDE_S * TS + (cc & 63));
        if (rr == g.rows) return lab_ld_cg(edges + (static_cast<long long>(g.ntiles) + g.tiles_x + (cc >> 6)) * (4 * TS) + SIDE_N * TS + (cc & 63));
        return INF;
    };
    bool handover = false; // the canonical step leads into the neighbouring band: the walk stops here (walk_left = d)
    // N, E, S, W (maze/maze.cc:213-218) without a table in local memory: k -> (k odd ? 0 : k - 1, k odd ? 2 - k : 0)
    bool const voter = lane < 4;
    int const my_dr = (lane & 1) ? 0 : (lane & 3) - 1, my_dc = (lane & 1) ? 2 - (lane & 3) : 0;
    if (room) {
        // In an open room the canonical path is two straight segments, known without walking: N is tried before E before
        // S before W, so the walker goes up first if the start is above it, else sideways first if the start is to its
        // right, else down and then left.  Every lane writes its share of the squares.
        int const sr = m.sr, sc = m.sc;
        int const vr = r > sr ? -1 : (r < sr ? 1 : 0), hc = c < sc ? 1 : (c > sc ? -1 : 0);
        uint32_t const nv = room_absdiff(r, sr), nh = room_absdiff(c, sc);
        bool const vertical_first = vr < 0 || (vr > 0 && hc <= 0);
        uint32_t const limit = first_only ? (d0 ? 1u : 0u) : d0;
        for (uint32_t i = static_cast<uint32_t>(lane); i < limit; i += 32u) { // the (i + 1)-th square after the finish
            int rr, cc;
            if (vertical_first) {
                rr = i < nv ? r + vr * static_cast<int>(i + 1) : sr;
                cc = i < nv ? c : c + hc * static_cast<int>(i + 1 - nv);
            } else {
                rr = i < nh ? r : r + vr * static_cast<int>(i + 1 - nh);
                cc = i < nh ? c + hc * static_cast<int>(i + 1) : sc;
            }
            if (rr < 0 || rr >= g.rows) {
                continue; // (row bands: another band's part of the path)
            }
            if (out && i < max_out) {
                out[2 * static_cast<unsigned long long>(i)] = rr;
                out[2 * static_cast<unsigned long long>(i) + 1] = cc;
            }
            if (repaint_bits) {
                long long const at = static_cast<long long>(rr) * g.cols + cc;
                grid[at] = static_cast<uint16_t>((grid[at] & ~SQ_PAINT_MASK) | repaint_bits);
            }
        }
        if (lane == 0) {
            res->path_len[slot] = d0;
            res->walk_r[slot] = sr;
            res->walk_c[slot] = sc;
            res->walk_left[slot] = 0;
        }
        return;
    }
    if (include_first) {
        if (lane == 0) {
            batch[0] = static_cast<uint32_t>(r) * static_cast<uint32_t>(g.cols) + static_cast<uint32_t>(c);
        }
        nb = 1;
    }
    if (first_only) {
        // A single step (gather's path.front()): the plain step on the field.
        while (d > 0) {
            bool hit = false;
            if (voter) {
                int const nr = r + my_dr, nc = c + my_dc;
                if (nr >= 0 && nr < g.rows && nc >= 0 && nc < g.cols) {
                    hit = dist[static_cast<long long>(nr) * g.cols + nc] == d - 1;
                } else {
                    hit = halo(nr, nc) == d - 1;
                }
            }
            uint32_t const mk = lab_ballot(hit);
            if (mk == 0) {
                break; // cannot happen on a converged field
            }
            int const k = lab_ffs32(mk) - 1;
            if (r + ((k & 1) ? 0 : k - 1) < 0 || r + ((k & 1) ? 0 : k - 1) >= g.rows) {
                handover = true;
                break;
            }
            r += (k & 1) ? 0 : k - 1;
            c += (k & 1) ? 2 - k : 0;
            d--;
            if (lane == 0) {
                batch[nb] = static_cast<uint32_t>(r) * static_cast<uint32_t>(g.cols) + static_cast<uint32_t>(c);
            }
            nb++;
            if (nb == static_cast<uint32_t>(WALK_BATCH)) {
                walk_flush(g, batch, nb, flushed, grid, out, max_out, repaint_bits, lane);
                flushed += nb;
                nb = 0;
            }
            if (first_only) {
                break;
            }
        }
    } else {
        // The 64x64 tile under the walker plus a ring of one square, in shared memory: every step taken from a square of
        // the tile is an LDS, a vote and a shuffle (a single warp issues one dependent instruction every ~5 cycles, so the
        // step's instruction count IS the walk's speed); stepping onto the ring makes that square's tile the next one.
        int const my_off = voter ? my_dr * WALK_W + my_dc : 0; // lanes 0..3 look N, E, S, W of the walker; the others at it
        bool lost = false;
        while (d > 0 && !lost && !handover) {
            int const cty = r >> 6, ctx = c >> 6;
            lab_syncwarp(); // (everybody is done with the old tile)
            {
                // all loads of a lane are issued before the first is used: the tile costs ONE round trip to memory
                constexpr int K = TS * TS / 32;
                uint32_t v[K], h[8];
                long long const row0 = static_cast<long long>(cty) * TS, col0 = static_cast<long long>(ctx) * TS;
                uint32_t const *src = dist + row0 * g.cols + col0 + lane;
                bool const in0 = col0 + lane < g.cols, in1 = col0 + lane + 32 < g.cols;
#pragma unroll
                for (int k = 0; k < K; k++) { // entry lane + 32 k of the tile: row k / 2, column lane + 32 (k % 2)
                    bool const ok = row0 + (k >> 1) < g.rows && ((k & 1) ? in1 : in0);
                    v[k] = ok ? lab_ld_ro(src + static_cast<long long>(k >> 1) * g.cols + 32 * (k & 1)) : INF;
                }
                // the ring: rows -1 and 64 (columns lane, lane + 32), columns -1 and 64 (rows lane, lane + 32); corners unused
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    int const q = lane + 32 * (k & 1);
                    long long const rr = k < 4 ? row0 + ((k & 2) ? TS : -1) : row0 + q;
                    long long const cc = k < 4 ? col0 + q : col0 + ((k & 2) ? TS : -1);
                    bool const ok = rr >= 0 && rr < g.rows && cc >= 0 && cc < g.cols;
                    h[k] = ok ? lab_ld_ro(dist + rr * g.cols + cc) : halo(rr, cc);
                }
#pragma unroll
                for (int k = 0; k < K; k++) {
                    cache[((k >> 1) + 1) * WALK_W + 1 + lane + 32 * (k & 1)] = v[k];
                }
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    int const q = lane + 32 * (k & 1);
                    int const at = k < 4 ? ((k & 2) ? (TS + 1) * WALK_W : 0) + 1 + q : (q + 1) * WALK_W + ((k & 2) ? TS + 1 : 0);
                    cache[at] = h[k];
                }
            }
            lab_syncwarp();
            uint32_t const origin = static_cast<uint32_t>(cty << 6) * static_cast<uint32_t>(g.cols) + static_cast<uint32_t>(ctx << 6);
            int lr = r & 63, lc = c & 63;
            int idx = (lr + 1) * WALK_W + lc + 1;
            while (d > 0 && static_cast<unsigned>(lr) < 64u && static_cast<unsigned>(lc) < 64u) {
                int const cand = idx + my_off;
                uint32_t const mk = lab_ballot(voter && cache[cand] == d - 1);
                if (mk == 0) {
                    lost = true; // cannot happen on a converged field
                    break;
                }
                int const k = lab_ffs32(mk) - 1; // first of N, E, S, W one level lower
                if ((cty << 6) + lr + ((k & 1) ? 0 : k - 1) < 0 || (cty << 6) + lr + ((k & 1) ? 0 : k - 1) >= g.rows) {
                    handover = true; // (into the neighbouring band)
                    break;
                }
                if (confine && (static_cast<unsigned>(lr + ((k & 1) ? 0 : k - 1)) >= 64u || static_cast<unsigned>(lc + ((k & 1) ? 2 - k : 0)) >= 64u)) {
                    handover = true; // (the segment ends at its tile's border)
                    break;
                }
                idx = static_cast<int>(lab_shfl(static_cast<uint32_t>(cand), k));
                lr += (k & 1) ? 0 : k - 1;
                lc += (k & 1) ? 2 - k : 0;
                d--;
                if (lane == 0) { // (the ring's squares too: lr, lc in -1 .. 64; the maze's edge is never on a path)
                    batch[nb] = origin + static_cast<uint32_t>(lr * g.cols + lc);
                }
                nb++;
                if (nb == static_cast<uint32_t>(WALK_BATCH)) {
                    walk_flush(g, batch, nb, flushed, grid, out, max_out, repaint_bits, lane);
                    flushed += nb;
                    nb = 0;
                }
            }
            r = (cty << 6) + lr;
            c = (ctx << 6) + lc;
        }
    }
    unsigned long long const n = flushed + nb; // squares walked
    walk_flush(g, batch, nb, flushed, grid, out, max_out, repaint_bits, lane);
    if (lane == 0 && !confine) {
        res->path_len[slot] = first_only ? static_cast<unsigned long long>(d0) : n;
        res->walk_r[slot] = r;
        res->walk_c[slot] = c;
        res->walk_left[slot] = d;
    }
}

// gather fix-ups, one thread, after paint: the claims (bfs_threads.cc:158-161) and the
// path.front() recolour in colour order (:355-364).  front_rc = first path square of colour k.
LAB_DEV void gather_fixup_body(Geom g, uint16_t *grid, Result const *res, int32_t const *finish_rc,
                               int32_t const *front_rc) {
    for (int k = 0; k < 4; k++) {
        int const j = res->finish_of[k];
        if (j >= 0) {
            grid[static_cast<long long>(finish_rc[2 * j]) * g.cols + finish_rc[2 * j + 1]]
                |= static_cast<uint16_t>((1u << k) << SQ_CACHE_SHIFT);
        }
    }
    for (int k = 0; k < 4; k++) {
        if (res->finish_of[k] >= 0 && res->game_level[k] > 0 && front_rc[2 * k] >= 0) {
            long long const i = static_cast<long long>(front_rc[2 * k]) * g.cols + front_rc[2 * k + 1];
            grid[i] = static_cast<uint16_t>((grid[i] & ~SQ_PAINT_MASK) | ((1u << k) << SQ_PAINT_SHIFT));
        }
    }
}


// ---------------------------------------------------------------------------------------------
// The heat map (painters/distance.cc:45-70, `painter`): the colour of every square of the distance map.
//   intensity = double(max - level) / double(max);  dark = uint8(255.0 * intensity);
//   bright = 128 + uint8(127.0 * intensity);  one channel (the reference draws it once per picture,
//   `channel` here) gets bright, the other two dark.
// Same IEEE double arithmetic as the reference, evaluated per square (div.rn.f64 / mul.rn.f64 and a
// truncating conversion: no fused or approximate operation is involved).  One 32-bit pixel per square:
// R | G << 8 | B << 16 | 0xFF << 24; squares without a map entry (the reference prints a wall glyph
// there) are 0.  max == 0 (a map of the start square alone): 0.0 / 0.0, which the reference's x86 casts
// turn into dark 0, bright 128.  Streaming, 4 squares per thread and step: 4 B read + 4 B written per square.
LAB_DEV uint32_t heat_pixel(uint32_t level, uint32_t max, double dmax, int channel) {
    if (level == INF) {
        return 0u;
    }
    uint32_t dark = 0, bright = 128;
    if (max != 0) {
        double const intensity = static_cast<double>(max - level) / dmax;
        dark = static_cast<uint32_t>(255.0 * intensity) & 0xFFu;
        bright = 128u + (static_cast<uint32_t>(127.0 * intensity) & 0xFFu);
    }
    uint32_t const grey = dark * 0x010101u;
    return ((grey & ~(0xFFu << (8 * channel))) | (bright << (8 * channel))) | 0xFF000000u;
}

LAB_DEV void heatmap_body(Geom g, uint32_t const *dist, Result const *res, int channel, uint32_t *pixels, long long gthread,
                          long long nthreads) {
    unsigned long long const n = static_cast<unsigned long long>(g.rows) * static_cast<unsigned long long>(g.cols);
    uint32_t const max = static_cast<uint32_t>(res->max_level);
    double const dmax = static_cast<double>(res->max_level);
    bool const aligned = ((reinterpret_cast<uintptr_t>(dist) | reinterpret_cast<uintptr_t>(pixels)) & 15u) == 0;
    unsigned long long const nvec = aligned ? n / 4 : 0;
    constexpr int U = 2;
    unsigned long long const stride = static_cast<unsigned long long>(nthreads);
    for (unsigned long long k0 = static_cast<unsigned long long>(gthread); k0 < nvec; k0 += stride * U) {
        lab_u32x4 in[U];
#pragma unroll
        for (int j = 0; j < U; j++) {
            unsigned long long const k = k0 + j * stride;
            in[j] = k < nvec ? lab_ld_stream_v4(dist + k * 4) : lab_u32x4{INF, INF, INF, INF};
        }
#pragma unroll
        for (int j = 0; j < U; j++) {
            unsigned long long const k = k0 + j * stride;
            if (k < nvec) {
                lab_st_stream_v4(pixels + k * 4, {heat_pixel(in[j].x, max, dmax, channel), heat_pixel(in[j].y, max, dmax, channel),
                                                  heat_pixel(in[j].z, max, dmax, channel), heat_pixel(in[j].w, max, dmax, channel)});
            }
        }
    }
    for (unsigned long long i = nvec * 4 + static_cast<unsigned long long>(gthread); i < n; i += stride) {
        pixels[i] = heat_pixel(dist[i], max, dmax, channel);
    }
}

// ---------------------------------------------------------------------------------------------
// Runs::paint_runs (painters/runs.cc:130-161): the length of the straight run the search's tree path ends with at
// every square -- `len = turn ? 1 : len(parent) + 1`, 0 at the start (:148-151).  On ONE TREE (every perfect maze) the
// FIFO search's parent of a square is simply its one neighbour a level lower, so the runs follow from the level field:
// walk back from the square while the parent lies in the same direction.  Runs in mazes are a few squares long, the
// walk is a handful of loads.  (A maze with loops needs the FIFO order among equal levels to name a parent: not here.)
// runs[i] = INF where the map has no entry; *max_out = the longest run.  One thread per square.
LAB_DEV void runs_body(Geom g, uint32_t const *dist, uint32_t *runs, uint32_t *max_out, long long gthread, long long nthreads) {
    long long const n = static_cast<long long>(g.rows) * g.cols;
    long long const step[4] = {-static_cast<long long>(g.cols), 1, static_cast<long long>(g.cols), -1}; // N, E, S, W (maze.cc:213-218)
    uint32_t mx = 0;
    for (long long i = gthread; i < n; i += nthreads) {
        uint32_t const d = dist[i];
        uint32_t len = INF;
        if (d == 0u) {
            len = 0u;
        } else if (d != INF) {
            long long const r = i / g.cols, c = i % g.cols;
            int dir = -1;
            for (int k = 0; k < 4 && dir < 0; k++) {
                bool const inside = k == 0 ? r > 0 : (k == 1 ? c + 1 < g.cols : (k == 2 ? r + 1 < g.rows : c > 0));
                if (inside && dist[i + step[k]] == d - 1u) {
                    dir = k;
                }
            }
            len = 1u;
            if (dir >= 0) {
                long long p = i + step[dir]; // the parent; its own parent in the same direction continues the run
                long long pr = r + (dir == 0 ? -1 : (dir == 2 ? 1 : 0)), pc = c + (dir == 1 ? 1 : (dir == 3 ? -1 : 0));
                for (uint32_t dp = d - 1u; dp > 0u; dp--) {
                    pr += dir == 0 ? -1 : (dir == 2 ? 1 : 0);
                    pc += dir == 1 ? 1 : (dir == 3 ? -1 : 0);
                    if (pr < 0 || pr >= g.rows || pc < 0 || pc >= g.cols || dist[p + step[dir]] != dp - 1u) {
                        break;
                    }
                    p += step[dir];
                    len++;
                }
            }
            mx = len > mx ? len : mx;
        }
        runs[i] = len;
    }
    mx = lab_reduce_max(mx);
    if ((gthread & 31) == 0 && mx) {
        lab_atomic_max(max_out, mx);
    }
}

// ---------------------------------------------------------------------------------------------
// Row bands (one band of a taller maze per GPU).  Between two local searches the ranks exchange
// the levels of their first / last ROW; these bodies translate between those plain rows
// (tiles_x*64 entries) and the tile-major structures of one band.

// passable flags (one byte per column) of the band's first and last row.  One thread per column.
LAB_DEV void band_path_rows_body(Geom g, uint64_t const *path, uint8_t *top, uint8_t *bottom, long long i) {
    long long const n = static_cast<long long>(g.tiles_x) * TS;
    if (i >= n) return;
    long long const tx = i / TS;
    int const c = static_cast<int>(i % TS);
    int const last_row = (g.rows - 1) % TS;
    top[i] = static_cast<uint8_t>((path[tx * TS + 0] >> c) & 1ull);
    bottom[i] = static_cast<uint8_t>((path[((static_cast<long long>(g.tiles_y) - 1) * g.tiles_x + tx) * TS + last_row] >> c) & 1ull);
}

// neighbours' border rows (bytes) -> one 64-bit word per tile column, as cross_body wants them.
// One warp per tile column; a null pointer means "no band there".
LAB_DEV void band_ghost_words_body(Geom g, uint8_t const *above_bottom, uint8_t const *below_top, uint64_t *ghost_n,
                                   uint64_t *ghost_s, long long gwarp, long long nwarps, int lane) {
    for (long long tx = gwarp; tx < g.tiles_x; tx += nwarps) {
        uint64_t const wn = above_bottom ? lab_ballot64(above_bottom[tx * TS + lane] != 0, above_bottom[tx * TS + lane + 32] != 0) : 0ull;
        uint64_t const wsb = below_top ? lab_ballot64(below_top[tx * TS + lane] != 0, below_top[tx * TS + lane + 32] != 0) : 0ull;
        if (lane == 0) {
            ghost_n[tx] = wn;
            ghost_s[tx] = wsb;
        }
    }
}

// levels of the band's first row (top tiles' N edges) and last row (bottom tiles' S edges).
LAB_DEV void band_export_edges_body(Geom g, uint32_t const *edges, uint32_t *top, uint32_t *bottom, long long i) {
    long long const n = static_cast<long long>(g.tiles_x) * TS;
    if (i >= n) return;
    long long const tx = i / TS;
    int const c = static_cast<int>(i % TS);
    long long const t_top = tx, t_bot = (static_cast<long long>(g.tiles_y) - 1) * g.tiles_x + tx;
    top[i] = edges[(t_top + g.tiles_x) * (4 * TS) + SIDE_N * TS + c];
    bottom[i] = edges[(t_bot + g.tiles_x) * (4 * TS) + SIDE_S * TS + c];
}

// Take the neighbours' border levels into the ghost tile rows; a border tile whose ghost improved
// where the border can be crossed becomes active and is queued.  One warp per tile column.
// (Runs between two launches of the search kernel: every tile is idle, the queue is empty.)
LAB_DEV void band_import_edges_body(Params p, int plane, uint32_t const *above_bottom, uint32_t const *below_top, uint32_t *activated,
                                    long long gwarp, long long nwarps, int lane) {
    Geom const &g = p.g;
    Plane const &pl = p.plane[plane];
    for (long long tx = gwarp; tx < g.tiles_x; tx += nwarps) {
        for (int side = 0; side < 2; side++) {
            uint32_t const *src = side == 0 ? above_bottom : below_top;
            if (!src) continue;
            long long const t = side == 0 ? tx : (static_cast<long long>(g.tiles_y) - 1) * g.tiles_x + tx;
            // ghost tile: the one above the top row / below the bottom row in the padded edge array
            uint32_t *ghost = pl.edges + (side == 0 ? tx : (static_cast<long long>(g.ntiles) + g.tiles_x + tx)) * (4 * TS)
                              + (side == 0 ? SIDE_S : SIDE_N) * TS;
            uint32_t const *own = pl.edges + (t + g.tiles_x) * (4 * TS) + (side == 0 ? SIDE_N : SIDE_S) * TS;
            uint64_t const X = p.cross[t * 4 + (side == 0 ? SIDE_N : SIDE_S)];
            bool improved = false;
            for (int h = 0; h < 2; h++) {
                int const c = lane + 32 * h;
                uint32_t const v = src[tx * TS + c];
                if (v < ghost[c]) {
                    ghost[c] = v;
                    if (((X >> c) & 1ull) && v + 1 < own[c]) improved = true;
                }
            }
            if (lab_any(improved) && lane == 0) {
                if (lab_atomic_or(&pl.state[t], ST_ACTIVE) == 0u) {
                    lab_atomic_add(&p.ctl->pending, 1u);
                    queue_push(p, static_cast<uint32_t>(plane) * static_cast<uint32_t>(g.ntiles) + static_cast<uint32_t>(t)); // (as init_run_body queues a source)
                    lab_atomic_add(activated, 1u);
                }
            }
        }
    }
}

// One thread: make the control block ready for another launch of the search kernel.
LAB_DEV void band_rearm_body(Params p, unsigned long long budget_ns) {
    p.ctl->head = 0;
    p.ctl->tail = 0;
    p.ctl->done = 0;
    p.ctl->deadline_ns = lab_clock_ns() + budget_ns;
}

} // namespace lab
