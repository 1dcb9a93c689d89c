// room.cuh -- the open room: a maze whose passable squares are exactly one full rectangle.
//
// That is what the reference's arena builder makes (builders/arena.cc:17-26: every interior square a
// path), and the one family of mazes on which the FIFO search of painters/distance.cc:129-153 and
// solvers/bfs_threads.cc:35-85 has nothing to find out: inside an obstacle-free rectangle every
// shortest 4-neighbour path is a staircase, so the BFS level of a square is its Manhattan distance from
// the start, whatever order the reference's queue visits the squares in.  The tile wave of engine.cuh
// gets the same numbers after ~(rows + cols) / 64 dependent tile hand-overs; here they are written in
// ONE streaming pass at HBM speed.
//
// Checked, not assumed: tree_count_body (tree.cuh) counts the passable squares V and takes their
// bounding box; a box that holds exactly V squares contains nothing else (tree_init_body sets
// Control::room).  Anything that is not such a rectangle -- a single wall square inside an arena is
// enough -- goes to the tile wave (or the tree pipeline).
//
//   distance   room_distance_body: Square | measure bit, level            8 B per square (= SURVEY 8(d))
//   solvers    room_paint_body:    Square | paint / cache bits            4 B per square, no level field
//
// Both walk the grid as a FLAT array, 8 squares (16 bytes of Squares, 32 bytes of levels) per thread and
// step, because the rows of an odd-sized maze are not 16-byte aligned but the arrays are.
#pragma once
#include "engine.cuh"

namespace lab {

// (signed: in a row band the box is the part of the room inside the band -- empty if r0 > r1 -- and the source may
// lie in another band, above row 0 or below the last row)
struct Room {
    int r0, r1, c0, c1, sr, sc;
};

LAB_DEV Room room_load(Control const *ctl) { return {ctl->room_r0, ctl->room_r1, ctl->room_c0, ctl->room_c1, ctl->room_sr, ctl->room_sc}; }

LAB_DEV uint32_t room_absdiff(int a, int b) { return static_cast<uint32_t>(a > b ? a - b : b - a); }

// plane 0's level of square (r, c): Manhattan distance inside the box, INF outside (nothing is passable there)
LAB_DEV uint32_t room_level(Room const &m, uint32_t r, uint32_t c) {
    int const ri = static_cast<int>(r), ci = static_cast<int>(c);
    bool const inside = ri >= m.r0 && ri <= m.r1 && ci >= m.c0 && ci <= m.c1;
    return inside ? room_absdiff(ri, m.sr) + room_absdiff(ci, m.sc) : INF;
}

// What a paint rule set gives a square at level lv (Paint_rule of post.cuh, flattened: below[k], bits[k]).
struct Room_rules {
    uint32_t n;
    uint32_t below[4], bits[4];
};
LAB_DEV uint32_t room_rule_bits(Room_rules const &q, uint32_t lv) {
    uint32_t b = 0;
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
        b |= (k < q.n && lv < q.below[k]) ? q.bits[k] : 0u;
    }
    return b;
}

// The shared skeleton: every thread takes chunks of 8 consecutive squares of the flat grid.
//   MODE 0 (distance): levels out, measure bit on the squares of the room
//   MODE 1 (solvers):  the rules' bits on the squares whose level is below the rules' thresholds
// Returns how many squares this thread marked.
template <int MODE>
LAB_DEV uint32_t room_stream(Params const &p, Room const &m, Room_rules const &q, uint16_t *grid, long long gthread, long long nthreads) {
    uint32_t const cols = static_cast<uint32_t>(p.g.cols);
    unsigned long long const n = static_cast<unsigned long long>(p.g.rows) * cols; // < 2^32 (make_geom)
    uint32_t *dist = p.plane[0].dist;
    bool const aligned = (reinterpret_cast<uintptr_t>(grid) & 15u) == 0 && (MODE == 1 || (reinterpret_cast<uintptr_t>(dist) & 15u) == 0);
    unsigned long long const nvec = aligned ? n / 8 : 0;
    uint32_t marked = 0;
    // U chunks per step: the loads of all of them are in flight before the first is needed
    constexpr int U = MODE == 0 ? 2 : 4;
    unsigned long long const stride = static_cast<unsigned long long>(nthreads);
    for (unsigned long long k0 = static_cast<unsigned long long>(gthread); k0 < nvec; k0 += stride * U) {
        lab_u32x4 in[U];
#pragma unroll
        for (int j = 0; j < U; j++) {
            unsigned long long const k = k0 + j * stride;
            in[j] = k < nvec ? lab_ld_stream_v4(grid + k * 8) : lab_u32x4{0u, 0u, 0u, 0u};
        }
#pragma unroll
        for (int j = 0; j < U; j++) {
            unsigned long long const k = k0 + j * stride;
            if (k >= nvec) {
                break;
            }
            uint32_t const i0 = static_cast<uint32_t>(k * 8);
            uint32_t r = i0 / cols, c = i0 - r * cols;
            uint32_t w[4] = {in[j].x, in[j].y, in[j].z, in[j].w};
            uint32_t lv[8];
            uint32_t any = 0;
            // Nearly every chunk lies in one row and wholly inside the room: its levels are dr + |d0 + u|, and
            // (solvers) usually all of them on the same side of every rule's threshold.
            if (c + 8u <= cols && static_cast<int>(r) >= m.r0 && static_cast<int>(r) <= m.r1 && static_cast<int>(c) >= m.c0 &&
                static_cast<int>(c) + 7 <= m.c1) {
                uint32_t const dr = room_absdiff(static_cast<int>(r), m.sr);
                int const d0 = static_cast<int>(c) - m.sc;
                if (MODE == 0) {
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        int const d = d0 + u;
                        lv[u] = dr + static_cast<uint32_t>(d < 0 ? -d : d);
                    }
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        w[h] |= static_cast<uint32_t>(SQ_MEASURE) * 0x10001u;
                    }
                    any = 1;
                    marked += 8;
                } else {
                    uint32_t const e0 = static_cast<uint32_t>(d0 < 0 ? -d0 : d0), e7 = static_cast<uint32_t>(d0 + 7 < 0 ? -(d0 + 7) : d0 + 7);
                    uint32_t const hi = dr + lab_umax(e0, e7), lo = (d0 <= 0 && d0 + 7 >= 0) ? dr : dr + lab_umin(e0, e7);
                    uint32_t all = 0;
                    bool mixed = false;
#pragma unroll
                    for (uint32_t rk = 0; rk < 4; rk++) {
                        if (rk < q.n) {
                            all |= hi < q.below[rk] ? q.bits[rk] : 0u;
                            mixed |= hi >= q.below[rk] && lo < q.below[rk];
                        }
                    }
                    if (!mixed) {
#pragma unroll
                        for (int h = 0; h < 4; h++) {
                            w[h] |= all * 0x10001u;
                        }
                        any = all;
                        marked += all ? 8u : 0u;
                    } else {
#pragma unroll
                        for (int u = 0; u < 8; u++) {
                            int const d = d0 + u;
                            uint32_t const b = room_rule_bits(q, dr + static_cast<uint32_t>(d < 0 ? -d : d));
                            w[u >> 1] |= b << (16 * (u & 1));
                            any |= b;
                            marked += b != 0;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    lv[u] = room_level(m, r, c);
                    uint32_t const b = MODE == 0 ? (lv[u] != INF ? static_cast<uint32_t>(SQ_MEASURE) : 0u) : room_rule_bits(q, lv[u]);
                    w[u >> 1] |= b << (16 * (u & 1));
                    any |= b;
                    marked += b != 0;
                    if (++c == cols) {
                        c = 0;
                        r++;
                    }
                }
            }
            if (any) {
                lab_st_stream_v4(grid + i0, {w[0], w[1], w[2], w[3]});
            }
            if (MODE == 0) {
                lab_st_stream_v4(dist + i0, {lv[0], lv[1], lv[2], lv[3]});
                lab_st_stream_v4(dist + i0 + 4, {lv[4], lv[5], lv[6], lv[7]});
            }
        }
    }
    // the tail (and everything, if a caller's array is not 16-byte aligned): square by square
    for (unsigned long long i = nvec * 8 + static_cast<unsigned long long>(gthread); i < n; i += static_cast<unsigned long long>(nthreads)) {
        uint32_t const r = static_cast<uint32_t>(i / cols), c = static_cast<uint32_t>(i - static_cast<unsigned long long>(r) * cols);
        uint32_t const lv = room_level(m, r, c);
        uint32_t const b = MODE == 0 ? (lv != INF ? static_cast<uint32_t>(SQ_MEASURE) : 0u) : room_rule_bits(q, lv);
        if (b) {
            grid[i] = static_cast<uint16_t>(grid[i] | b);
            marked++;
        }
        if (MODE == 0) {
            dist[i] = lv;
        }
    }
    return marked;
}

// distance map of an open room (returns at once anywhere else): levels, measure bits, map size.
LAB_DEV void room_distance_body(Params const &p, uint16_t *grid, unsigned long long *map_size, long long gthread, long long nthreads) {
    if (!p.ctl->room) {
        return;
    }
    Room const m = room_load(p.ctl);
    Room_rules const none{};
    uint32_t marked = room_stream<0>(p, m, none, grid, gthread, nthreads);
    marked = lab_reduce_add(marked);
    if ((gthread & 31) == 0 && marked) {
        lab_atomic_add(reinterpret_cast<uint64_t *>(map_size), static_cast<uint64_t>(marked));
    }
}

// solvers' paint pass in an open room; `painted` counts the squares that received a bit.
LAB_DEV void room_paint_body(Params const &p, Room_rules const &q, uint16_t *grid, unsigned long long *painted, long long gthread, long long nthreads) {
    Room const m = room_load(p.ctl);
    uint32_t marked = room_stream<1>(p, m, q, grid, gthread, nthreads);
    marked = lab_reduce_add(marked);
    if ((gthread & 31) == 0 && marked) {
        lab_atomic_add(reinterpret_cast<uint64_t *>(painted), static_cast<uint64_t>(marked));
    }
}

} // namespace lab
