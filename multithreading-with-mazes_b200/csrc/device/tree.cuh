// tree.cuh -- spanning-tree mazes without a wave: exact BFS levels in O(log) dependent steps.
//
// The reference's builders (kruskal, wilson, rdfs, ...) make PERFECT mazes: the passable squares
// form a tree, whose diameter is 10^4 .. 10^7 levels at the sizes of BASELINE.json.  Any wave
// algorithm -- the reference's FIFO queue (painters/distance.cc:129-153, solvers/bfs_threads.cc:35-85)
// or the tile wave of engine.cuh -- has that many dependent steps on its critical path.  On a tree
// the BFS level of a square is its depth below the root, and depths can be had without following
// the wave:
//
//   1 walk    per 64x64 tile, one thread per border crossing follows the wall (right-hand rule) from
//             the crossing it enters through to the crossing it leaves through.  These segments,
//             chained through the crossings, ARE the Euler tour of the tree, cut at the tile borders.
//   2 rank    pointer jumping over the segments (a list of ~64 entries per tile) orders the crossings
//             along the tour that starts at the root.
//   3 orient  of the two directions of a crossing the one met first leads AWAY from the root.  Every
//             tile-local component of the tree now knows the one crossing it is entered through.
//   4 flood   every tile's components are traversed from their entry crossings, all tiles at once, no
//             tile waiting for another: every square gets component << 16 | local level (its parent's
//             value + 1; junctions hand their other branches to idle lanes).
//   5 prefix  the components form a tree themselves (parent = the component on the other side of the
//             entry crossing, edge weight = local level of that crossing + 1); pointer jumping sums
//             the weights down from the root: the exact level of every entry crossing.
//   6 fix-up  one streaming pass over the field: level of the component's entry + local level.
//
// Whether the maze IS a tree is checked, not assumed: passable squares V and adjacent passable pairs
// E are counted (a tree has E == V-1), every crossing must lie on the root's tour and the local
// floods must reach all V squares (=> connected).  Anything else (arena, grid, -m cross/x, squares
// masked by measure bits) takes the tile wave of engine.cuh instead; both give identical levels.
#pragma once
#include "engine.cuh"
#include "room.cuh"

namespace lab {

constexpr uint32_t SLOTS = 4 * TS; // crossing slots of a tile: [side][pos] = the arc ENTERING the tile there

struct Tree_ctl {
    unsigned long long V, E; // passable squares / adjacent passable pairs
    unsigned long long local_settled;
    uint32_t ok;         // cleared as soon as the maze turns out not to be a tree
    uint32_t unreached;  // crossings that are not on the root's tour
    uint32_t bad_weight; // an entry crossing whose far side got no local level
    uint32_t walk_error; // a walker ran longer than any tile allows
    uint32_t ticket[3];  // dynamic tile counters: walk, flood
    uint32_t rank_buf, prefix_buf; // which ping-pong buffer holds the finished ranking / prefix
    uint32_t rank_more[40], prefix_more[40]; // round r left work for round r+1
    uint32_t n_arcs;     // crossings in total = entries of Tree::alist
    uint32_t used;       // the tree pipeline produced the field (set by gate)
    uint32_t barrier[2]; // grid barriers of the two multi-round kernels (rank, prefix)
    uint32_t dirty;      // the flood has started writing its packed field into the dense level field
    // bounding box of the passable squares (0x7FFFFFFF - min as a running max, so that the all-zero reset
    // means "nothing yet"): a box that holds exactly V squares is an OPEN ROOM (device/room.cuh)
    uint32_t box_r0_inv, box_r1, box_c0_inv, box_c1;
    unsigned long long marked; // squares the fix-up has given Tree::mark_bits
    uint32_t crossings;  // border crossings of the tiles counted (= the entries the walk will add to alist)
    uint32_t alist_base; // row bands: where this band's piece of the global crossing list starts
    // the lattice road (device/lattice.cuh), tried first on mazes carved on the lattice of odd squares
    uint32_t lat_err;        // why it gave up (1 not a lattice, 2 labels, 3 tour, 4 floods, 5 weights); nothing was written then
    uint32_t lat_on;         // tree_init_body: the counts allow a tree and the host found odd sizes and an (odd, odd) start
    uint32_t lat_root_label[2]; // the root components' labels in their tiles (two roots: a start on a passage square)
    uint32_t lat_head_b;     // 1 + the arc the second root's tour begins with (0: it has none)
    uint32_t lat_flood_bad;
    uint32_t lat_ticket[4];  // dynamic tile counters: link, local ranking, flood, levels
    unsigned long long lat_settled;
    unsigned long long lat_settled_band; // row bands: this band's share (lat_settled is the whole maze's after the exchange)
    // paths from the tour's ranks (lat_path_*): rank of the entry arc of the finish's component (0: a root's
    // component), the finish's level, segments found
    uint32_t lat_path_rank[4], lat_path_level[4], lat_path_nseg[4], lat_path_on; // (per path: hunt has one, gather four; _on: a bit per path)
};

struct Tree {
    uint32_t *succ0;     // [nslots+2] successor segment, as walked (kept: it also chains a component's crossings)
    uint32_t *sr[2];     // [2 * (nslots+2)] ping-pong, (pointer, value) PAIRS: [2a] successor / parent entry, [2a+1] segments
                         //                  to the end of the tour / level of the entry (one 8-byte access: lab_ld_cg_x2)
    uint64_t *inmask;    // [ntiles][4]  crossings through which a tile's components are ENTERED
    uint32_t *loc;       // [nslots]     local level of border squares ([side][pos], like edges)
    uint32_t *alist;     // [n_arcs]     the slots that ARE crossings (the pointer-jumping rounds only visit these)
    Tree_ctl *tc;
    uint32_t nslots;     // ntiles * SLOTS; ROOT = nslots, END = nslots + 1
    // Row bands (one band of a taller maze per GPU): the arrays above are indexed by GLOBAL
    // slot / tile (every rank holds all of them; the ranks exchange their slices between the phases), nslots
    // counts the whole maze's slots, and Params describes this band only: its tile t is global tile
    // tile_base + t.  Single GPU: tile_base = 0, banded = 0.
    uint32_t tile_base;
    uint32_t banded;
    // Pointer jumping in a WINDOW (row bands): a pointer is only followed while it stays inside slots
    // [jump_lo, jump_hi) -- the band's own -- so a band contracts its part of the list without looking at
    // anybody else's; what is left is a list over the bands' border tile rows, short enough for every rank
    // to finish on its own after ONE exchange of those rows (tree_border_*_body), and a last local pass adds
    // the two (tree_finish_jump_body).  jump_hi == 0: no window.  jump_skip: bit 0 = no parent pass before
    // the prefix rounds, bit 1 = no orient pass after the ranking rounds (they run as launches of their own).
    uint32_t jump_lo, jump_hi, jump_skip;
    // Square bits the fix-up ORs into every square it gives a level: Rgb::measure for the distance map
    // (painters/distance.cc:138,149 -- saves measure_body's pass over the grid), 0 for the solvers.
    uint32_t mark_bits;
    // the host found odd sizes and an (odd, odd) start: the lattice road (device/lattice.cuh) is tried first
    uint32_t lat_try;
};

LAB_DEV bool tree_follow(Tree const &tr, uint32_t slot) { // may a round follow a pointer to `slot`?
    return slot < tr.nslots && (tr.jump_hi == 0u || slot - tr.jump_lo < tr.jump_hi - tr.jump_lo);
}

// Row bands: the ranks found (from their counts and bounding boxes) that the whole maze is an open room; this band's
// part of it and the source, relative to the band, come from the host.  One thread.
LAB_DEV void room_set_body(Control *ctl, int r0, int r1, int c0, int c1, int sr, int sc) {
    ctl->room_r0 = r0;
    ctl->room_r1 = r1;
    ctl->room_c0 = c0;
    ctl->room_c1 = c1;
    ctl->room_sr = sr;
    ctl->room_sc = sc;
    ctl->room = 1u;
}

// The lattice road's kernels all run before the general pipeline's: once they are through, lat_on without an error
// means the level field is made.
LAB_DEV bool lat_made(Tree_ctl const *tc) { return tc->lat_on != 0 && tc->lat_err == 0; }
LAB_DEV bool tree_live(Tree_ctl const *tc) {
    return tc->ok != 0 && !lat_made(tc) && tc->unreached == 0 && tc->bad_weight == 0 && tc->walk_error == 0;
}

// the arc entering the neighbouring tile through the same crossing
LAB_DEV uint32_t tree_twin(Geom const &g, uint32_t a) {
    uint32_t const t = a / SLOTS, s = a % SLOTS, side = s / TS, pos = s % TS;
    switch (side) {
    case SIDE_N: return (t - static_cast<uint32_t>(g.tiles_x)) * SLOTS + SIDE_S * TS + pos;
    case SIDE_S: return (t + static_cast<uint32_t>(g.tiles_x)) * SLOTS + SIDE_N * TS + pos;
    case SIDE_W: return (t - 1u) * SLOTS + SIDE_E * TS + pos;
    default: return (t + 1u) * SLOTS + SIDE_W * TS + pos;
    }
}

// ---------------------------------------------------------------------------------------------
// V and E.  Warp per tile, grid stride; lane i holds rows 2i, 2i+1.
LAB_DEV void tree_count_body(Params const &p, Tree const &tr, long long gwarp, long long nwarps, int lane) {
    unsigned long long v = 0, e = 0;
    uint32_t r0i = 0, r1 = 0, c0i = 0, c1 = 0; // the warp's share of the bounding box (see Tree_ctl)
    uint32_t nx = 0;
    for (long long t = gwarp; t < p.g.ntiles; t += nwarps) {
        uint64_t const P0 = lab_ld_ro(p.path + t * TS + 2 * lane), P1 = lab_ld_ro(p.path + t * TS + 2 * lane + 1);
        if (P0 | P1) {
            uint32_t const row0 = static_cast<uint32_t>(t / p.g.tiles_x) * TS, col0 = static_cast<uint32_t>(t % p.g.tiles_x) * TS;
            uint32_t const ra = row0 + static_cast<uint32_t>(2 * lane + (P0 ? 0 : 1)), rb = row0 + static_cast<uint32_t>(2 * lane + (P1 ? 1 : 0));
            uint32_t const ca = col0 + static_cast<uint32_t>(lab_ffs64(P0 | P1) - 1), cb = col0 + static_cast<uint32_t>(63 - lab_clz64(P0 | P1));
            r0i = lab_umax(r0i, 0x7FFFFFFFu - ra);
            r1 = lab_umax(r1, rb);
            c0i = lab_umax(c0i, 0x7FFFFFFFu - ca);
            c1 = lab_umax(c1, cb);
        }
        uint64_t below = lab_shfl_down64(P0, 1); // row 2*lane+2
        if (lane == 31) {
            below = 0;
        }
        v += static_cast<unsigned>(lab_popc64(P0) + lab_popc64(P1));
        e += static_cast<unsigned>(lab_popc64(P0 & (P0 >> 1)) + lab_popc64(P1 & (P1 >> 1)) + lab_popc64(P0 & P1) + lab_popc64(P1 & below));
        if (lane < 4) {
            uint32_t const x = static_cast<uint32_t>(lab_popc64(lab_ld_ro(p.cross + t * 4 + lane)));
            nx += x;
            if (lane == SIDE_S || lane == SIDE_E) { // (as adjacent pairs they are counted once: this tile's S and E borders)
                e += x;
            }
        }
    }
    uint32_t const vs = lab_reduce_add(static_cast<uint32_t>(v)), es = lab_reduce_add(static_cast<uint32_t>(e));
    uint32_t const vh = lab_reduce_add(static_cast<uint32_t>(v >> 32)), eh = lab_reduce_add(static_cast<uint32_t>(e >> 32));
    r0i = lab_reduce_max(r0i);
    r1 = lab_reduce_max(r1);
    c0i = lab_reduce_max(c0i);
    c1 = lab_reduce_max(c1);
    nx = lab_reduce_add(nx);
    if (lane == 0) {
        if (nx) {
            lab_atomic_add(&tr.tc->crossings, nx);
        }
        lab_atomic_add(reinterpret_cast<uint64_t *>(&tr.tc->V), (static_cast<uint64_t>(vh) << 32) + vs);
        lab_atomic_add(reinterpret_cast<uint64_t *>(&tr.tc->E), (static_cast<uint64_t>(eh) << 32) + es);
        if (r0i) { // (some square of this warp's tiles is passable)
            lab_atomic_max(&tr.tc->box_r0_inv, r0i);
            lab_atomic_max(&tr.tc->box_r1, r1);
            lab_atomic_max(&tr.tc->box_c0_inv, c0i);
            lab_atomic_max(&tr.tc->box_c1, c1);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// A tile as a byte per square, in shared memory: bits 0-3 = the neighbour to the N, E, S, W is
// passable and inside the tile, bits 4-7 = ... is passable and across the tile's border (an open
// crossing).  One look-up per step for the walkers below.  Lane i brings rows 2i, 2i+1.
LAB_DEV uint32_t spread4(uint32_t bits) { // bits 0..3 -> bit 0 of bytes 0..3
    return ((bits & 0xFu) * 0x00204081u) & 0x01010101u;
}
LAB_DEV void tile_open_row(uint8_t *O, int r, uint64_t n, uint64_t e, uint64_t s, uint64_t w) {
    uint32_t *out = reinterpret_cast<uint32_t *>(O + r * TS);
#pragma unroll
    for (int half = 0; half < 2; half++) {
        uint32_t const n32 = static_cast<uint32_t>(n >> (32 * half)), e32 = static_cast<uint32_t>(e >> (32 * half));
        uint32_t const s32 = static_cast<uint32_t>(s >> (32 * half)), w32 = static_cast<uint32_t>(w >> (32 * half));
#pragma unroll
        for (int k = 0; k < 8; k++) {
            out[half * 8 + k] = spread4(n32 >> (4 * k)) | (spread4(e32 >> (4 * k)) << 1) | (spread4(s32 >> (4 * k)) << 2) | (spread4(w32 >> (4 * k)) << 3);
        }
    }
}
LAB_DEV void tile_open_table(uint8_t *O, uint64_t P0, uint64_t P1, uint64_t const (&X)[4], int lane) {
    int const r0 = 2 * lane, r1 = 2 * lane + 1;
    uint64_t const up = lab_shfl_up64(P1, 1), down = lab_shfl_down64(P0, 1);
    tile_open_row(O, r0, lane == 0 ? 0ull : (P0 & up), P0 & (P0 >> 1), P0 & P1, P0 & (P0 << 1));
    tile_open_row(O, r1, P1 & P0, P1 & (P1 >> 1), lane == 31 ? 0ull : (P1 & down), P1 & (P1 << 1));
    lab_syncwarp();
    // the open crossings: lane looks after columns / rows lane and lane + 32 of each side
#pragma unroll
    for (int half = 0; half < 2; half++) {
        int const q = lane + 32 * half;
        if ((X[SIDE_N] >> q) & 1ull) O[q] |= 0x10;
        if ((X[SIDE_S] >> q) & 1ull) O[(TS - 1) * TS + q] |= 0x40;
    }
    lab_syncwarp(); // (a corner square belongs to two sides)
#pragma unroll
    for (int half = 0; half < 2; half++) {
        int const q = lane + 32 * half;
        if ((X[SIDE_E] >> q) & 1ull) O[q * TS + TS - 1] |= 0x20;
        if ((X[SIDE_W] >> q) & 1ull) O[q * TS] |= 0x80;
    }
    lab_syncwarp();
}

// ---------------------------------------------------------------------------------------------
// 1 walk.  Directions clockwise: N=0, E=1, S=2, W=3.  A walker's state is (square, heading it
// arrived with); it takes the first open direction among right, straight, left, back.  On a tree
// that visits every arc exactly once, i.e. it is the Euler tour.
//
// A segment between two crossings of a tile can be thousands of steps long (the tour of a whole
// subtree), and one thread would walk it alone.  So the walk is cut once more, at the borders of
// the tile's sixteen 16x16 BLOCKS: phase A starts a walker at every open border between two blocks
// as well (in both directions) and every walker stops at the first block border it crosses -- a few
// dozen steps; phase B then hops from piece to piece (one shared-memory look-up each) from every
// tile crossing to the crossing its segment ends at.
//
// walker ids: [0,256) tile crossings [side][pos]; 256 + h*192 + (b-1)*64 + pos: heading h over inner
// border b (1..3, at row/column 16*b), pos = row (h = E, W) or column (h = N, S); 1024: the root.
// piece results (uint16): < 1024 the next piece; 1024 + s: leaves the tile into the neighbour's slot s;
// 0xFFFF: the end of the tour.
//
// shared memory per warp (uint64 units): byte table[4096] | res16[1032] | list16[1032] (later res32[256]) | masks[16] | counter
constexpr int WALK_IDS = 1025;
constexpr int WALK_TAB = TS * TS / 8;
constexpr int WALK_LIST = 640; // walkers of one tile (a 64x64 piece of a perfect maze has at most 513)
constexpr int WALK_U64 = WALK_TAB + 260 + WALK_LIST / 4 + 2; // 7.4 KB: seven 4-warp blocks per SM
constexpr uint32_t WALK_END = 0xFFFFu;

LAB_DEV void walk_tile(Params const &p, Tree const &tr, uint32_t t, int lane, uint64_t *sm) {
    Geom const &g = p.g;
    uint8_t *O = reinterpret_cast<uint8_t *>(sm);
    uint16_t *res16 = reinterpret_cast<uint16_t *>(sm + WALK_TAB);
    uint16_t *list = reinterpret_cast<uint16_t *>(sm + WALK_TAB + 260);
    uint32_t *res32 = reinterpret_cast<uint32_t *>(sm + WALK_TAB + 260); // after phase A the list is dead
    uint64_t *masks = sm + WALK_TAB + 260 + WALK_LIST / 4 - 16; // (read before the list grows that far)
    uint32_t *cnt = reinterpret_cast<uint32_t *>(sm + WALK_TAB + 260 + WALK_LIST / 4);
    uint32_t const ROOT = tr.nslots, END = tr.nslots + 1u;
    uint32_t const gt = t + tr.tile_base; // the tile's id in the whole maze (row bands)
    int const r0 = 2 * lane, r1 = 2 * lane + 1;

    uint64_t const P0 = lab_ld_ro(p.path + static_cast<long long>(t) * TS + r0), P1 = lab_ld_ro(p.path + static_cast<long long>(t) * TS + r1);
    uint64_t X[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        X[s] = lab_ld_ro(p.cross + static_cast<long long>(t) * 4 + s);
    }
    tile_open_table(O, P0, P1, X, lane);

    // which walkers exist: 16 masks of 64 positions (4 tile sides, then heading-major inner borders)
    if (lane < 4) {
        masks[lane] = X[lane];
    }
#pragma unroll
    for (int b = 1; b <= 3; b++) {
        // vertical inner border b: squares (pos, 16b-1) and (pos, 16b) both passable
        uint64_t const v = lab_ballot64((O[lane * TS + 16 * b - 1] >> 1) & 1, (O[(lane + 32) * TS + 16 * b - 1] >> 1) & 1);
        // horizontal inner border b: squares (16b-1, pos) and (16b, pos)
        uint64_t const hmask = lab_ballot64((O[(16 * b - 1) * TS + lane] >> 2) & 1, (O[(16 * b - 1) * TS + lane + 32] >> 2) & 1);
        if (lane == 0) {
            masks[4 + 0 * 3 + (b - 1)] = hmask; // heading N
            masks[4 + 1 * 3 + (b - 1)] = v;     // heading E
            masks[4 + 2 * 3 + (b - 1)] = hmask; // heading S
            masks[4 + 3 * 3 + (b - 1)] = v;     // heading W
        }
    }
    lab_syncwarp();
    uint32_t const my_bits = static_cast<uint32_t>(masks[lane >> 1] >> (32 * (lane & 1)));
    uint32_t const mine = static_cast<uint32_t>(lab_popc64(my_bits));
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t const v = lab_shfl_up(incl, d);
        if (lane >= d) {
            incl += v;
        }
    }
    uint32_t total = lab_shfl(incl, 31);
    lab_syncwarp(); // every lane has read its mask word: the list may overwrite them
    if (total > static_cast<uint32_t>(WALK_LIST)) { // more walkers than any maze tile has: leave it to the wave
        lab_atomic_max(&tr.tc->walk_error, 4u);
        total = 0;
    } else {
        uint32_t q = incl - mine, m = my_bits;
        while (m) {
            int const b = lab_ffs64(m) - 1;
            m &= m - 1;
            list[q++] = static_cast<uint16_t>(32 * lane + b);
        }
    }
    if (lane == 0) {
        *cnt = 0;
    }
    lab_syncwarp();

    // the root: the tour starts in the state "arrived at the root from one of its open neighbours"
    bool const root_tile = p.ctl->src_tile[0] == t;
    int rr = -1, rc = -1, rh = 0;
    bool root_open = false;
    if (root_tile) {
        uint32_t const rcv = p.ctl->src_rc[0];
        rr = static_cast<int>(rcv >> 8);
        rc = static_cast<int>(rcv & 0xFF);
        uint32_t const b = O[rr * TS + rc];
        for (int d = 3; d >= 0; d--) {
            if (((b | (b >> 4)) >> d) & 1u) {
                rh = (d + 2) & 3;
                root_open = true;
            }
        }
    }
    // ---- phase A: every walker to the first block border ----------------------------------------
    // One flat loop, one step per iteration: a lane whose walker has arrived fetches its next one in
    // the same iteration, so the lanes of the warp stay converged whatever the lengths of their pieces.
    // A walker's state is one word: square (row << 6 | column) << 2 | heading.
    uint32_t const nitems = total + (root_tile ? 1u : 0u);
    uint32_t const root_state = root_tile && root_open ? ((static_cast<uint32_t>(rr) << 6 | static_cast<uint32_t>(rc)) << 2 | static_cast<uint32_t>(rh))
                                                       : 0xFFFFFFFFu;
    bool have = false, skip = false;
    uint32_t st = 0, id = 0;
    for (;;) {
        if (!have) {
            uint32_t const idx = lab_atomic_add(cnt, 1u);
            if (idx >= nitems) {
                break;
            }
            have = true;
            skip = idx == total; // the root's own walker leaves the root state instead of ending there
            if (skip) {
                id = WALK_IDS - 1;
                st = (static_cast<uint32_t>(rr) << 6 | static_cast<uint32_t>(rc)) << 2 | static_cast<uint32_t>(rh);
            } else {
                id = list[idx];
                uint32_t const pos = id & 63u;
                uint32_t r, c, h;
                if (id < SLOTS) { // a tile crossing: the first square inside, heading inwards
                    uint32_t const side = id >> 6;
                    bool const ns = side < 2u;
                    r = ns ? side * (TS - 1u) : pos;
                    c = ns ? pos : (side & 1u) * (TS - 1u);
                    h = (0xD2u >> (2u * side)) & 3u; // N -> S, S -> N, W -> E, E -> W
                } else { // over an inner border at row / column `line`
                    uint32_t const k = (id - SLOTS) >> 6; // heading * 3 + (border - 1)
                    h = (k * 11u) >> 5;
                    uint32_t const line = 16u * (k - 3u * h + 1u);
                    uint32_t const at = line - ((h == 0u || h == 3u) ? 1u : 0u); // heading N / W: the square before the line
                    r = (h & 1u) ? pos : at;
                    c = (h & 1u) ? at : pos;
                }
                st = (r << 6 | c) << 2 | h;
            }
        }
        uint32_t result = WALK_END;
        bool arrived = st == root_state && !skip; // the tour ends where it began
        skip = false;
        uint32_t const x = st >> 2;
        uint32_t h = st & 3u;
        uint32_t const b = O[x];
        uint32_t const nib = (b | (b >> 4)) & 15u;
        arrived |= nib == 0; // a square on its own (only the root can be one)
        if (!arrived) {
            uint32_t const rot = ((nib >> h) | (nib << (4u - h))) & 15u; // bit j: direction heading + j
            h = (h + ((0x53535252u >> (2u * rot)) & 3u)) & 3u;           // right, else straight, else left, else back
            uint32_t const nx = x + ((0x3F804100u >> (8u * h)) & 0xFFu) - 64u; // -64, +1, +64, -1
            if ((b >> (4u + h)) & 1u) { // through an open crossing: into the neighbour's slot
                uint32_t const r = x >> 6, c = x & 63u;
                // heading N enters the neighbour's S side, E its W side, S its N side, W its E side
                result = 1024u + ((0xC9u >> (2u * h)) & 3u) * TS + ((h & 1u) ? r : c);
                arrived = true;
            } else if ((x ^ nx) & 0xC30u) { // crossed a block border: the next piece starts here
                uint32_t const r = x >> 6, c = x & 63u, nr = nx >> 6, nc = nx & 63u;
                uint32_t const line = (h == 1u ? nc : (h == 3u ? c : (h == 2u ? nr : r))) >> 4; // 1..3
                result = SLOTS + (h * 3u + line - 1u) * 64u + ((h & 1u) ? r : c);
                arrived = true;
            }
            st = nx << 2 | h;
        }
        if (arrived) {
            res16[id] = static_cast<uint16_t>(result);
            have = false;
        }
    }
    lab_syncwarp();
    // ---- phase B: from every tile crossing (and the root) piece by piece to the end of its segment ----
    for (int k = 0; k < 9; k++) {
        int const s = k < 8 ? lane + 32 * k : WALK_IDS - 1;
        bool const active = k < 8 ? ((X[s >> 6] >> (s & 63)) & 1ull) != 0 : (root_tile && lane == 0);
        uint32_t out = END;
        if (active) {
            uint32_t cur = static_cast<uint32_t>(s);
            uint32_t v = WALK_END;
            int hops = 0;
            for (; hops <= WALK_IDS; hops++) {
                v = res16[cur];
                if (v >= 1024u) {
                    break;
                }
                cur = v;
            }
            if (hops > WALK_IDS) {
                lab_atomic_max(&tr.tc->walk_error, 2u);
                v = WALK_END;
            }
            if (v != WALK_END) {
                uint32_t const ns = v - 1024u, side = ns >> 6;
                uint32_t const nt = side == SIDE_S ? gt - static_cast<uint32_t>(g.tiles_x)
                                                   : (side == SIDE_N ? gt + static_cast<uint32_t>(g.tiles_x) : (side == SIDE_W ? gt + 1u : gt - 1u));
                out = nt * SLOTS + ns;
            }
        }
        if (k < 8) {
            res32[s] = out; // (the list it overwrites was last read before the barrier above)
        } else if (root_tile && lane == 0) {
            tr.succ0[ROOT] = out;
            lab_st_x2(tr.sr[0], ROOT, out, 1u);
        }
    }
    lab_syncwarp();
    long long const o = static_cast<long long>(gt) * SLOTS;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int const s = lane + 32 * k;
        uint32_t const v = res32[s];
        bool const active = (X[s >> 6] >> (s & 63)) & 1ull;
        tr.succ0[o + s] = v;
        lab_st_x2(tr.sr[0], static_cast<uint32_t>(o + s), v, active ? 1u : 0u);
    }
    // the tile's crossings join the global list of arcs
    uint32_t const xbits = static_cast<uint32_t>((X[lane >> 3] >> ((lane & 7) * 8)) & 0xFFull); // slots 8*lane .. 8*lane+7
    uint32_t const xmine = static_cast<uint32_t>(lab_popc64(xbits));
    uint32_t xincl = xmine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t const v = lab_shfl_up(xincl, d);
        if (lane >= d) {
            xincl += v;
        }
    }
    uint32_t const xtotal = lab_shfl(xincl, 31);
    uint32_t xbase = 0;
    if (lane == 0 && xtotal) {
        xbase = lab_atomic_add(&tr.tc->n_arcs, xtotal) + tr.tc->alist_base;
    }
    xbase = lab_shfl(xbase, 0);
    {
        uint32_t q = xbase + xincl - xmine, m = xbits;
        while (m) {
            int const b = lab_ffs64(m) - 1;
            m &= m - 1;
            tr.alist[q++] = static_cast<uint32_t>(o) + static_cast<uint32_t>(8 * lane + b);
        }
    }
}

LAB_DEV void tree_walk_body(Params const &p, Tree const &tr, int lane, uint64_t *sm) {
    if (!tree_live(tr.tc)) {
        return;
    }
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) {
            t = lab_atomic_add(&tr.tc->ticket[0], 1u);
        }
        t = lab_shfl(t, 0);
        if (t >= static_cast<uint32_t>(p.g.ntiles)) {
            break;
        }
        walk_tile(p, tr, t, lane, sm);
        lab_syncwarp();
    }
}

// One thread, after the count and before the walk: can it be a tree at all?  And the two list ends.
LAB_DEV void tree_init_body(Params const &p, Tree const &tr) {
    uint32_t const ROOT = tr.nslots, END = tr.nslots + 1u;
    Tree_ctl *tc = tr.tc;
    tc->ok = (tc->V >= 2 && tc->E + 1 == tc->V) ? 1u : 0u; // a tree has squares - 1 adjacent pairs
    tc->lat_on = (tc->ok && tr.lat_try && !tr.banded) ? 1u : 0u;
    // Not a tree, but the bounding box of the passable squares holds nothing else: an open room (the
    // reference's arena builder).  Levels are a closed form there (device/room.cuh); the source is one of
    // the passable squares (force_sources_body), so it lies inside.
    if (!tc->ok && tc->V >= 1 && !tr.banded) {
        uint32_t const r0 = 0x7FFFFFFFu - tc->box_r0_inv, c0 = 0x7FFFFFFFu - tc->box_c0_inv;
        unsigned long long const area = static_cast<unsigned long long>(tc->box_r1 - r0 + 1u) * (tc->box_c1 - c0 + 1u);
        if (area == tc->V && p.ctl->src_tile[0] != 0xFFFFFFFFu) {
            Control *ctl = p.ctl;
            ctl->room_r0 = static_cast<int32_t>(r0);
            ctl->room_r1 = static_cast<int32_t>(tc->box_r1);
            ctl->room_c0 = static_cast<int32_t>(c0);
            ctl->room_c1 = static_cast<int32_t>(tc->box_c1);
            ctl->room_sr = static_cast<int32_t>((ctl->src_tile[0] / static_cast<uint32_t>(p.g.tiles_x)) * TS + (ctl->src_rc[0] >> 8));
            ctl->room_sc = static_cast<int32_t>((ctl->src_tile[0] % static_cast<uint32_t>(p.g.tiles_x)) * TS + (ctl->src_rc[0] & 0xFFu));
            ctl->room = 1u;
        }
    }
    tr.succ0[ROOT] = END;
    tr.succ0[END] = END;
    for (int b = 0; b < 2; b++) {
        lab_st_x2(tr.sr[b], ROOT, END, 1u);
        lab_st_x2(tr.sr[b], END, END, 0u);
    }
}

// ---------------------------------------------------------------------------------------------
// 2 rank.  One round of pointer jumping: r = segments from here to the end of the tour.
// Round `round` reads buffer round&1 and writes the other one; a round that finds the previous one
// left nothing to do returns at once (the launches are fixed in number, the work is not).
// Only crossings (Tree::alist) and the root are visited.  The same body runs as one launch per round
// (tests/simt) or inside one persistent launch with a grid barrier between rounds (csrc/lab.cu), hence
// the loads that bypass L1: the other buffer was written by other SMs a moment ago.
LAB_DEV void tree_rank_round_body(Tree const &tr, uint32_t round, long long gthread, long long nthreads) {
    Tree_ctl *tc = tr.tc;
    if (!tree_live(tc) || (round > 0 && lab_ld_volatile(&tc->rank_more[round - 1]) == 0)) {
        return;
    }
    uint32_t const ROOT = tr.nslots;
    bool const odd = round & 1u;
    uint32_t const *in = odd ? tr.sr[1] : tr.sr[0];
    uint32_t *out = odd ? tr.sr[0] : tr.sr[1];
    long long const n = static_cast<long long>(tc->n_arcs) + 1;
    bool more = false;
    for (long long i = gthread; i < n; i += nthreads) {
        uint32_t const a = i + 1 < n ? tr.alist[i] : ROOT;
        lab_u32x2 const me = lab_ld_cg_x2(in, a);
        if (!tree_follow(tr, me.x)) { // the end of the tour, or (row bands) a segment of another band
            lab_st_x2(out, a, me.x, me.y);
        } else {
            lab_u32x2 const nx = lab_ld_cg_x2(in, me.x);
            lab_st_x2(out, a, nx.x, me.y + nx.y);
            more |= tree_follow(tr, nx.x);
        }
    }
    // (thousands of warps storing to -- or reading past L1 from -- one address would serialise in L2:
    // look before writing, through L1; a stale zero only costs a redundant store)
    if (lab_any(more) && (gthread & 31) == 0 && lab_ld_ro(&tc->rank_more[round]) == 0u) {
        lab_st_volatile(&tc->rank_more[round], 1u);
    }
    if (gthread == 0) {
        lab_st_volatile(&tc->rank_buf, (round & 1u) ^ 1u);
    }
}

// ---------------------------------------------------------------------------------------------
// 3 orient.  Warp per tile: which crossings are entered through (met earlier on the tour than their
// twin).  Also resets the tile's local border levels.
LAB_DEV void tree_orient_body(Params const &p, Tree const &tr, long long gwarp, long long nwarps, int lane) {
    Tree_ctl *tc = tr.tc;
    if (!tree_live(tc)) {
        return;
    }
    uint32_t const END = tr.nslots + 1u;
    // (the rounds' result has been brought back to buffer 0: tree_jump_settle_body)
    uint32_t const *fin = tr.sr[0];
    uint32_t bad = 0;
    for (long long t = gwarp; t < p.g.ntiles; t += nwarps) {
        long long const o = (t + tr.tile_base) * SLOTS;
        uint64_t masks[4];
#pragma unroll
        for (int side = 0; side < 4; side++) {
            uint64_t const X = lab_ld_ro(p.cross + t * 4 + side);
            bool in[2];
#pragma unroll
            for (int half = 0; half < 2; half++) {
                int const pos = lane + 32 * half;
                uint32_t const a = static_cast<uint32_t>(o) + static_cast<uint32_t>(side * TS + pos);
                in[half] = false;
                if ((X >> pos) & 1ull) {
                    uint32_t const tw = tree_twin(p.g, a);
                    lab_u32x2 const me = lab_ld_cg_x2(fin, a);
                    in[half] = me.y > lab_ld_cg_x2(fin, tw).y;
                    bad += me.x != END;
                }
                tr.loc[a] = INF;
            }
            masks[side] = lab_ballot64(in[0], in[1]);
        }
        if (lane < 4) {
            tr.inmask[(t + tr.tile_base) * 4 + lane] = masks[lane];
        }
    }
    if (lab_any(bad != 0) && lane == 0) {
        lab_atomic_add(&tc->unreached, 1u);
    }
}

// ---------------------------------------------------------------------------------------------
// 4 flood.  Every component of every tile gets  component << 16 | local level  on each of its
// squares (component = the slot of its entry crossing, 256 = the root's), all tiles at once, no tile
// waiting for another.  On a tree that needs no wave either: a square's value is its parent's + 1,
// so a component is a set of independent TASKS "square x holds v, came from direction b": a lane
// follows its branch square by square (one look-up of the path rows, one store per square), and at
// every junction pushes the other branches on the warp's shared-memory stack for idle lanes to pop.
// No level synchronisation, no frontier, every lane busy as long as the stack holds work; a warp keeps
// FLOOD_SLOTS tiles open at a time so that the tail of one tile overlaps the bulk of the next.
// Once the entry crossings' exact levels are known (5 prefix) one streaming pass turns the packed
// field into global levels (6 fix-up).  The local levels of border squares go to tr.loc.
constexpr int FLOOD_SLOTS = 1;
constexpr int FLOOD_STACK = 256; // tasks
// per warp, uint64 units: byte tables of the open tiles | task stack | control words
constexpr int FLOOD_TAB = TS * TS / 8; // one byte table (tile_open_table) per open tile
constexpr int FLOOD_STEPS = 4;         // squares a lane advances between two looks at the stack
constexpr int FLOOD_U64 = FLOOD_SLOTS * FLOOD_TAB + FLOOD_STACK + 2 + 2 * FLOOD_SLOTS;
constexpr uint32_t ROOT_COMP = SLOTS;
constexpr uint32_t NO_TILE = 0xFFFFFFFFu;

LAB_DEV void flood_note_border(Tree const &tr, uint32_t t, int r, int c, uint32_t ld) {
    uint32_t *loc = tr.loc + static_cast<long long>(t + tr.tile_base) * SLOTS;
    if (r == 0) loc[SIDE_N * TS + c] = ld;
    if (r == TS - 1) loc[SIDE_S * TS + c] = ld;
    if (c == 0) loc[SIDE_W * TS + r] = ld;
    if (c == TS - 1) loc[SIDE_E * TS + r] = ld;
}

struct Flood_sm {
    uint8_t *O;           // [FLOOD_SLOTS][4096] tile_open_table
    uint64_t *stack;      // [FLOOD_STACK]  v | (x | back << 12 | slot << 15) << 32
    uint32_t *top;        // tasks on the stack
    uint32_t *error;
    uint32_t *live;       // [FLOOD_SLOTS] tasks of the slot's tile not finished yet (0 = slot free)
    uint32_t *tile;       // [FLOOD_SLOTS]
    uint64_t *origin;     // [FLOOD_SLOTS] index of the tile's first square in the dense field
};

LAB_DEV void flood_push(Flood_sm const &f, uint32_t v, uint32_t x, uint32_t back, uint32_t slot) {
    uint32_t const pos = lab_atomic_add(f.top, 1u);
    if (pos < static_cast<uint32_t>(FLOOD_STACK)) {
        f.stack[pos] = static_cast<uint64_t>(v) | (static_cast<uint64_t>(x | (back << 12) | (slot << 15)) << 32);
    } else {
        *f.error = 1u;
    }
}

// Take the next tile into `slot`: its path rows, and one task per entry crossing (and the root).
// Returns false when there are no tiles left.  Whole warp.
LAB_DEV bool flood_open(Params const &p, Tree const &tr, Flood_sm const &f, int slot, int lane, Warp_stats &ws) {
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) {
            t = lab_atomic_add(&tr.tc->ticket[1], 1u);
        }
        t = lab_shfl(t, 0);
        if (t >= static_cast<uint32_t>(p.g.ntiles)) {
            return false;
        }
        ws.passes++;
        uint64_t const in_side = lane < 4 ? tr.inmask[static_cast<long long>(t + tr.tile_base) * 4 + lane] : 0ull;
        bool const root_here = p.ctl->src_tile[0] == t;
        if (!lab_any(in_side != 0) && !root_here) {
            ws.empty++;
            continue; // nothing enters this tile
        }
        {
            uint64_t X[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                X[k] = lab_ld_ro(p.cross + static_cast<long long>(t) * 4 + k);
            }
            tile_open_table(f.O + slot * TS * TS, lab_ld_ro(p.path + static_cast<long long>(t) * TS + 2 * lane),
                            lab_ld_ro(p.path + static_cast<long long>(t) * TS + 2 * lane + 1), X, lane);
        }
        // lane looks after slots 8*lane .. 8*lane+7 (side = lane / 8)
        uint64_t const m_side = lab_shfl64(in_side, lane >> 3);
        uint32_t m = static_cast<uint32_t>((m_side >> ((lane & 7) * 8)) & 0xFFull);
        uint32_t n = static_cast<uint32_t>(lab_popc64(m));
        int const side = lane >> 3;
        while (m) {
            int const b = lab_ffs64(m) - 1;
            m &= m - 1;
            uint32_t const pos = static_cast<uint32_t>((lane & 7) * 8 + b);
            uint32_t const r = side == SIDE_N ? 0u : (side == SIDE_S ? TS - 1u : pos);
            uint32_t const c = side == SIDE_W ? 0u : (side == SIDE_E ? TS - 1u : pos);
            flood_push(f, (static_cast<uint32_t>(side) * TS + pos) << 16, (r << 6) | c, 4u, static_cast<uint32_t>(slot));
        }
        if (root_here && lane == 0) {
            uint32_t const rc = p.ctl->src_rc[0];
            flood_push(f, ROOT_COMP << 16, ((rc >> 8) << 6) | (rc & 0xFFu), 4u, static_cast<uint32_t>(slot));
            n++;
        }
        n = lab_reduce_add(n);
        if (lane == 0) {
            int const ty = static_cast<int>(t / static_cast<uint32_t>(p.g.tiles_x)), tx = static_cast<int>(t % static_cast<uint32_t>(p.g.tiles_x));
            f.live[slot] = n;
            f.tile[slot] = t;
            f.origin[slot] = static_cast<uint64_t>(ty) * TS * static_cast<uint64_t>(p.g.cols) + static_cast<uint64_t>(tx) * TS;
        }
        lab_syncwarp();
        return true;
    }
}

LAB_DEV void tree_flood_body(Params const &p, Tree const &tr, int lane, uint64_t *sm) {
    if (!tree_live(tr.tc)) {
        return;
    }
    if (lane == 0 && lab_ld_ro(&tr.tc->dirty) == 0u) {
        lab_st_volatile(&tr.tc->dirty, 1u);
    }
    Flood_sm f;
    f.O = reinterpret_cast<uint8_t *>(sm);
    f.stack = sm + FLOOD_SLOTS * FLOOD_TAB;
    f.top = reinterpret_cast<uint32_t *>(sm + FLOOD_SLOTS * FLOOD_TAB + FLOOD_STACK);
    f.error = f.top + 1;
    f.live = f.top + 2;
    f.tile = f.live + FLOOD_SLOTS;
    f.origin = sm + FLOOD_SLOTS * FLOOD_TAB + FLOOD_STACK + 2 + FLOOD_SLOTS;
    if (lane == 0) {
        *f.top = 0;
        *f.error = 0;
        for (int s = 0; s < FLOOD_SLOTS; s++) {
            f.live[s] = 0;
            f.tile[s] = NO_TILE;
        }
    }
    lab_syncwarp();
    Warp_stats ws;
    uint32_t *dist = p.plane[0].dist;
    long long const cols = p.g.cols;
    bool exhausted = false, have = false;
    uint32_t xs = 0, back = 4, v = 0, settled = 0; // xs = slot << 12 | row << 6 | column
    uint32_t di = 0;                                // the square's index in the dense field (< 2^32: make_geom)
    int const icols = static_cast<int>(cols);
    for (;;) {
        if (*f.error) { // (nobody writes it between the barrier that ended the last round and here)
            break;      // not a tree (a cycle keeps spawning tasks): the result is void anyway, stop
        }
        // keep the stack fed: a free slot takes the next tile
        uint32_t top = *f.top;
        if (top < 32u && !exhausted) {
            for (int s = 0; s < FLOOD_SLOTS && !exhausted; s++) {
                if (f.live[s] == 0) {
                    exhausted = !flood_open(p, tr, f, s, lane, ws);
                }
            }
            top = *f.top;
        }
        if (top > static_cast<uint32_t>(FLOOD_STACK)) {
            top = FLOOD_STACK; // pushes beyond the stack were dropped (and flagged): not a tree
        }
        // idle lanes pop (lane 0 settles the new top afterwards)
        uint32_t const need = lab_ballot(!have);
        if (need) {
            uint32_t const rank = static_cast<uint32_t>(lab_popc64(need & ((1u << lane) - 1u)));
            if (!have && rank < top) {
                uint64_t const task = f.stack[top - 1u - rank];
                v = static_cast<uint32_t>(task);
                uint32_t const hi = static_cast<uint32_t>(task >> 32);
                back = (hi >> 12) & 7u;
                xs = (hi & 0xFFFu) | ((hi >> 15) << 12);
                di = static_cast<uint32_t>(f.origin[hi >> 15]) + ((hi >> 6) & 63u) * static_cast<uint32_t>(icols) + (hi & 63u);
                have = true;
            }
            uint32_t const want = static_cast<uint32_t>(lab_popc64(need));
            lab_syncwarp();
            if (lane == 0) {
                *f.top = top - (want < top ? want : top);
            }
        } else if (lane == 0 && *f.top != top) {
            *f.top = top;
        }
        lab_syncwarp();
        if (!lab_any(have)) {
            if (exhausted) {
                break; // no tasks, no tiles
            }
            continue;
        }
#pragma unroll 1
        for (int k = 0; k < FLOOD_STEPS; k++) {
            if (have) {
                uint32_t const b = f.O[xs];
                uint32_t const nib = b & 15u & ~(1u << back);
                dist[di] = v;
                if (b >> 4) { // a square with an open crossing: its local level is an edge weight of 5 prefix
                    flood_note_border(tr, f.tile[xs >> 12], static_cast<int>((xs >> 6) & 63u), static_cast<int>(xs & 63u), v & 0xFFFFu);
                }
                settled++;
                if (nib == 0 || (v & 0xFFFFu) > static_cast<uint32_t>(TS * TS)) {
                    // the branch ends here (or runs in circles: not a tree, the checks after the flood say so)
                    if (nib != 0) {
                        *f.error = 1u;
                    }
                    have = false;
                    lab_atomic_sub(&f.live[xs >> 12], 1u);
                } else {
                    int const d = lab_ffs32(nib) - 1;
                    uint32_t rest = nib & (nib - 1u);
                    while (rest) { // junction: the other branches are somebody else's
                        int const d2 = lab_ffs32(rest) - 1;
                        rest &= rest - 1u;
                        lab_atomic_add(&f.live[xs >> 12], 1u);
                        flood_push(f, v + 1u, static_cast<uint32_t>(static_cast<int>(xs & 0xFFFu) + ((d2 & 1) ? 2 - d2 : (d2 - 1) * TS)),
                                   static_cast<uint32_t>((d2 + 2) & 3), xs >> 12);
                    }
                    xs = static_cast<uint32_t>(static_cast<int>(xs) + ((d & 1) ? 2 - d : (d - 1) * TS));
                    di = static_cast<uint32_t>(static_cast<int>(di) + ((d & 1) ? 2 - d : (d - 1) * icols));
                    back = static_cast<uint32_t>((d + 2) & 3);
                    v++;
                }
            }
        }
        lab_syncwarp();
    }
    if (*f.error) {
        lab_atomic_max(&tr.tc->walk_error, 3u);
    }
    settled = lab_reduce_add(settled);
    if (lane == 0) {
        if (settled) {
            lab_atomic_add(reinterpret_cast<uint64_t *>(&tr.tc->local_settled), static_cast<uint64_t>(settled));
        }
        lab_atomic_add(reinterpret_cast<uint64_t *>(&p.ctl->n_activations), static_cast<uint64_t>(ws.passes));
        lab_atomic_add(reinterpret_cast<uint64_t *>(&p.ctl->n_empty), static_cast<uint64_t>(ws.empty));
    }
}

// ---------------------------------------------------------------------------------------------
// 5 prefix.  parent: an entry crossing e (into tile T from square q of tile T') hangs below the entry
// crossing of q's component of T'.  The crossings of one component are chained by the walk: the
// segment that starts at crossing a ends at crossing succ0[a], whose twin starts the next one.
// Thread per slot (buffer 0 of the ping-pong pair; the ranking is finished with them).
LAB_DEV void tree_parent_body(Params const &p, Tree const &tr, long long gthread, long long nthreads) {
    Tree_ctl *tc = tr.tc;
    if (!tree_live(tc)) {
        return;
    }
    if (tc->local_settled != tc->V) { // some square cannot be reached: not one tree
        if (gthread == 0) {
            tc->ok = 0;
        }
        return;
    }
    uint32_t const ROOT = tr.nslots, END = tr.nslots + 1u;
    uint32_t *parval = tr.sr[0];
    long long const n = static_cast<long long>(tc->n_arcs) + 1;
    for (long long i = gthread; i < n; i += nthreads) {
        uint32_t pa = ROOT, v = INF;
        uint32_t const e = i + 1 < n ? tr.alist[i] : ROOT;
        if (e == ROOT) {
            v = 0;
        } else {
            uint32_t const t = e / SLOTS, s = e % SLOTS;
            if ((tr.inmask[static_cast<long long>(t) * 4 + s / TS] >> (s % TS)) & 1ull) {
                uint32_t const a = tree_twin(p.g, e);
                uint32_t const lq = tr.loc[a];
                if (lq == INF) {
                    lab_atomic_max(&tc->bad_weight, 1u);
                } else {
                    v = lq + 1u;
                }
                uint32_t cur = a;
                int guard = 0;
                for (; guard <= static_cast<int>(SLOTS); guard++) {
                    uint32_t const ct = cur / SLOTS, cs = cur % SLOTS;
                    if ((tr.inmask[static_cast<long long>(ct) * 4 + cs / TS] >> (cs % TS)) & 1ull) {
                        pa = cur;
                        break;
                    }
                    uint32_t const nx = tr.succ0[cur];
                    if (nx == END) { // the component that holds the root
                        pa = ROOT;
                        break;
                    }
                    cur = tree_twin(p.g, nx);
                }
                if (guard > static_cast<int>(SLOTS)) {
                    lab_atomic_max(&tc->bad_weight, 2u);
                }
            }
        }
        lab_st_x2(parval, e, pa, v);
    }
}

LAB_DEV void tree_prefix_round_body(Tree const &tr, uint32_t round, long long gthread, long long nthreads) {
    Tree_ctl *tc = tr.tc;
    if (!tree_live(tc) || (round > 0 && lab_ld_volatile(&tc->prefix_more[round - 1]) == 0)) {
        return;
    }
    uint32_t const ROOT = tr.nslots;
    bool const odd = round & 1u;
    uint32_t const *in = odd ? tr.sr[1] : tr.sr[0];
    uint32_t *out = odd ? tr.sr[0] : tr.sr[1];
    long long const n = static_cast<long long>(tc->n_arcs) + 1;
    bool more = false;
    for (long long i = gthread; i < n; i += nthreads) {
        uint32_t const a = i + 1 < n ? tr.alist[i] : ROOT;
        lab_u32x2 const me = lab_ld_cg_x2(in, a);
        if (!tree_follow(tr, me.x)) { // the root, or (row bands) an entry of another band
            lab_st_x2(out, a, me.x, me.y);
        } else {
            lab_u32x2 const up = lab_ld_cg_x2(in, me.x);
            lab_st_x2(out, a, up.x, me.y + up.y);
            more |= tree_follow(tr, up.x);
        }
    }
    if (lab_any(more) && (gthread & 31) == 0 && lab_ld_ro(&tc->prefix_more[round]) == 0u) {
        lab_st_volatile(&tc->prefix_more[round], 1u);
    }
    if (gthread == 0) {
        lab_st_volatile(&tc->prefix_buf, (round & 1u) ^ 1u);
    }
}

// After the rounds of a ranking (kind 0) / prefix (kind 1): bring the entries of the list back to ping-pong
// buffer 0 if the last round left them in buffer 1, so that everything downstream (orient, fix-up, the band
// exchange) reads ONE place.  `buf` = the buffer the rounds ended in (read by the caller before anybody
// could change it).
LAB_DEV void tree_jump_settle_body(Tree const &tr, uint32_t buf, long long gthread, long long nthreads) {
    if (buf == 0u) {
        return;
    }
    uint32_t const ROOT = tr.nslots;
    long long const n = static_cast<long long>(tr.tc->n_arcs) + 1;
    for (long long i = gthread; i < n; i += nthreads) {
        uint32_t const a = i + 1 < n ? tr.alist[i] : ROOT;
        lab_u32x2 const v = lab_ld_cg_x2(tr.sr[1], a);
        lab_st_x2(tr.sr[0], a, v.x, v.y);
    }
}

// ---------------------------------------------------------------------------------------------
// Row bands: what the ranks tell each other between the phases.  Every rank contributes one CHUNK of
// uint32: a header (its error flags and counts) and the slots of its first and last TILE ROW -- the only
// tiles another band's pointers can lead into -- of up to three of the global arrays; after the all-gather
// every rank copies the other ranks' rows into its own copy of those arrays and folds the headers into its
// control block, so that all ranks come to the same verdict without a round trip through the host.
//   stage 1 (after the windowed ranking)  sr[0] (successor, segments) pairs, succ0   + the tour's start (ROOT) from its rank
//   stage 2 (after orient + flood)        loc, inmask
//   stage 3 (after the windowed prefix)   sr[0] (parent, level so far) pairs
constexpr uint32_t BORDER_HDR = 8; // unreached, walk_error, bad_weight, !ok, settled lo, settled hi, ROOT pointer, ROOT value
struct Band_layout {          // of every rank, in rank order (device memory, owned by the caller)
    uint32_t tile_base, ntiles; // the band's first tile (global id) and number of tiles
};
LAB_DEV uint32_t border_row_words(Params const &p) { return static_cast<uint32_t>(p.g.tiles_x) * SLOTS; }
LAB_DEV uint32_t border_chunk_words(Params const &p, int stage) {
    uint32_t const R = border_row_words(p);
    return BORDER_HDR + (stage == 1 ? 6u * R : (stage == 2 ? 2u * R + 2u * (R / 32u) : 4u * R));
}
// the arrays of a stage, as (pointer to uint32 view, words per tile)
struct Border_arr {
    uint32_t *base;
    uint32_t per_tile;
};
LAB_DEV int border_arrays(Tree const &tr, int stage, Border_arr (&arr)[3]) {
    if (stage == 1) {
        arr[0] = {tr.sr[0], 2u * SLOTS};
        arr[1] = {tr.succ0, SLOTS};
        return 2;
    }
    if (stage == 2) {
        arr[0] = {tr.loc, SLOTS};
        arr[1] = {reinterpret_cast<uint32_t *>(tr.inmask), 8u}; // 4 x uint64 per tile
        return 2;
    }
    arr[0] = {tr.sr[0], 2u * SLOTS};
    return 1;
}

LAB_DEV void tree_border_export_body(Params const &p, Tree const &tr, int stage, uint32_t *chunk, long long gthread, long long nthreads) {
    Tree_ctl *tc = tr.tc;
    uint32_t const tiles_x = static_cast<uint32_t>(p.g.tiles_x), ntiles = static_cast<uint32_t>(p.g.ntiles);
    if (gthread == 0) {
        chunk[0] = tc->unreached;
        chunk[1] = tc->walk_error;
        chunk[2] = tc->bad_weight;
        chunk[3] = tc->ok ? 0u : 1u;
        chunk[4] = static_cast<uint32_t>(tc->local_settled);
        chunk[5] = static_cast<uint32_t>(tc->local_settled >> 32);
        lab_u32x2 const root = lab_ld_cg_x2(tr.sr[0], tr.nslots);
        chunk[6] = root.x;
        chunk[7] = root.y;
    }
    Border_arr arr[3];
    int const na = border_arrays(tr, stage, arr);
    uint32_t off = BORDER_HDR;
    for (int k = 0; k < na; k++) {
        uint32_t const row = tiles_x * arr[k].per_tile; // words of one tile row of this array
        for (int which = 0; which < 2; which++) {        // first, last tile row of the band
            uint32_t const t0 = tr.tile_base + (which == 0 ? 0u : ntiles - tiles_x);
            uint32_t const *src = arr[k].base + static_cast<size_t>(t0) * arr[k].per_tile;
            for (long long i = gthread; i < row; i += nthreads) {
                chunk[off + i] = src[i];
            }
            off += row;
        }
    }
}

// `all` = the gathered chunks (world x chunk words).  root_rank = the rank whose band holds the start.
LAB_DEV void tree_border_import_body(Params const &p, Tree const &tr, int stage, uint32_t const *all, Band_layout const *lay, int world, int me,
                                     int root_rank, long long gthread, long long nthreads) {
    Tree_ctl *tc = tr.tc;
    uint32_t const tiles_x = static_cast<uint32_t>(p.g.tiles_x);
    uint32_t const words = border_chunk_words(p, stage);
    if (gthread == 0) {
        unsigned long long settled = 0;
        uint32_t bad = 0;
        for (int r = 0; r < world; r++) {
            uint32_t const *h = all + static_cast<size_t>(r) * words;
            bad |= h[0] | h[1] | h[2] | h[3];
            settled += static_cast<unsigned long long>(h[4]) | (static_cast<unsigned long long>(h[5]) << 32);
        }
        if (bad) { // some rank found out that this is not one tree: every rank reads the same flags and stops here
            tc->ok = 0;
        }
        if (stage == 2) {
            tc->local_settled = settled; // (tree_parent_body compares it with V: both are the whole maze's now)
        }
        if (stage == 1 && me != root_rank) {
            uint32_t const *h = all + static_cast<size_t>(root_rank) * words;
            lab_st_x2(tr.sr[0], tr.nslots, h[6], h[7]);
        }
    }
    Border_arr arr[3];
    int const na = border_arrays(tr, stage, arr);
    for (int r = 0; r < world; r++) {
        if (r == me) {
            continue;
        }
        uint32_t const *c = all + static_cast<size_t>(r) * words;
        uint32_t off = BORDER_HDR;
        for (int k = 0; k < na; k++) {
            uint32_t const row = tiles_x * arr[k].per_tile;
            for (int which = 0; which < 2; which++) {
                uint32_t const t0 = lay[r].tile_base + (which == 0 ? 0u : lay[r].ntiles - tiles_x);
                uint32_t *dst = arr[k].base + static_cast<size_t>(t0) * arr[k].per_tile;
                for (long long i = gthread; i < row; i += nthreads) {
                    dst[i] = c[off + i];
                }
                off += row;
            }
        }
    }
}

// The last pass of a windowed ranking (kind 0) / prefix (kind 1): an entry of the band's own list that stopped
// at a pointer into another band adds what that band's border row -- finished by the short list's rounds --
// holds.  Entries of the band's own border rows were on the short list themselves and are final already.
LAB_DEV void tree_finish_jump_body(Params const &p, Tree const &tr, long long gthread, long long nthreads) {
    Tree_ctl *tc = tr.tc;
    if (!tree_live(tc)) {
        return;
    }
    uint32_t const ROOT = tr.nslots;
    uint32_t const tiles_x = static_cast<uint32_t>(p.g.tiles_x), ntiles = static_cast<uint32_t>(p.g.ntiles);
    long long const n = static_cast<long long>(tc->n_arcs) + 1;
    for (long long i = gthread; i < n; i += nthreads) {
        uint32_t const a = i + 1 < n ? tr.alist[i] : ROOT;
        if (a != ROOT) {
            uint32_t const t = a / SLOTS - tr.tile_base;
            if (t < tiles_x || t >= ntiles - tiles_x) {
                continue; // a border row's entry
            }
        }
        lab_u32x2 const me = lab_ld_cg_x2(tr.sr[0], a);
        if (me.x < tr.nslots) { // stopped at the window's edge: me.x is an entry of another band's border row
            lab_u32x2 const far = lab_ld_cg_x2(tr.sr[0], me.x);
            lab_st_x2(tr.sr[0], a, far.x, me.y + far.y);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 6 fix-up.  component << 16 | local level  ->  level of the component's entry + local level.
// One warp-item = 64 squares of one row inside one tile (so the entry levels it needs are one 1 KB
// table); lane reads squares lane and lane + 32.
LAB_DEV void tree_fixup_body(Params const &p, Tree const &tr, uint16_t *grid, long long gwarp, long long nwarps, int lane) {
    Tree_ctl *tc = tr.tc;
    if (!tree_live(tc)) {
        return;
    }
    Geom const &g = p.g;
    uint32_t const *lvl = tr.sr[0]; // (tree_jump_settle_body) pairs: the level is the second word
    uint32_t *dist = p.plane[0].dist;
    uint32_t mx = 0, marked = 0;
    uint16_t const mark = static_cast<uint16_t>(tr.mark_bits);
    // one warp-item = four tiles' worth of one row: eight independent loads per lane in flight
    constexpr int U = 4;
    long long const chunks = (g.tiles_x + U - 1) / U;
    long long const items = static_cast<long long>(g.rows) * chunks;
    for (long long it = gwarp; it < items; it += nwarps) {
        int const r = static_cast<int>(it / chunks);
        int const j0 = static_cast<int>(it % chunks) * U;
        uint32_t *row = dist + static_cast<long long>(r) * g.cols;
        uint16_t *srow = grid + static_cast<long long>(r) * g.cols;
        long long const trow = static_cast<long long>(r / TS) * g.tiles_x;
        // The field was NOT reset before the flood: what is passable comes from the tile masks (all
        // passable squares were reached: one tree), everything else becomes INF here.
        uint64_t w[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            w[u] = j0 + u < g.tiles_x ? lab_ld_ro(p.path + (trow + j0 + u) * TS + (r % TS)) : 0ull;
        }
        uint32_t v[2 * U];
        uint16_t sq[2 * U];
#pragma unroll
        for (int u = 0; u < 2 * U; u++) {
            int const c = j0 * TS + lane + 32 * u;
            bool const pass = (w[u / 2] >> (lane + 32 * (u & 1))) & 1ull;
            v[u] = pass ? row[c] : INF;
            sq[u] = (pass && mark) ? srow[c] : static_cast<uint16_t>(0);
        }
#pragma unroll
        for (int u = 0; u < 2 * U; u++) {
            int const c = j0 * TS + lane + 32 * u;
            if (v[u] != INF) {
                uint32_t const comp = v[u] >> 16;
                uint32_t const *b = lvl + (trow + j0 + u / 2 + tr.tile_base) * (2 * SLOTS);
                uint32_t const out = (comp == ROOT_COMP ? 0u : b[2 * comp + 1]) + (v[u] & 0xFFFFu);
                row[c] = out;
                mx = out > mx ? out : mx;
                if (mark) {
                    srow[c] = static_cast<uint16_t>(sq[u] | mark);
                    marked++;
                }
            } else if (c < g.cols) {
                row[c] = INF;
            }
        }
    }
    mx = lab_reduce_max(mx);
    marked = lab_reduce_add(marked);
    if (lane == 0 && mx) {
        lab_atomic_max(&p.ctl->max_level, mx);
    }
    if (lane == 0 && marked) {
        lab_atomic_add(reinterpret_cast<uint64_t *>(&tc->marked), static_cast<uint64_t>(marked));
    }
}

// After the fix-up.  If the tree pipeline made the field (thread 0): take the targets' levels (the
// solvers' bounds and paint rules come from them) and call the wave off.  If it gave up after the
// flood had already written its packed field: wipe field and visited masks, and the tile wave finds
// everything as init_run_body armed it.
LAB_DEV void tree_gate_body(Params const &p, Tree const &tr, long long gthread, long long nthreads) {
    Tree_ctl *tc = tr.tc;
    Control *ctl = p.ctl;
    if (ctl->room) {
        // An open room: nothing to search.  The levels are a closed form (room.cuh): the distance map's
        // fused pass has written them (and the measure bits: `visited` stays empty for measure_body), the
        // solvers paint straight from the formula and never materialise the field.
        if (gthread != 0) {
            return;
        }
        Room const m = room_load(ctl);
        for (uint32_t j = 0; j < ctl->n_targets; j++) {
            unsigned long long const cell = ctl->target_cell[j];
            ctl->target_val[j] = room_level(m, static_cast<uint32_t>(cell / static_cast<unsigned long long>(p.g.cols)),
                                            static_cast<uint32_t>(cell % static_cast<unsigned long long>(p.g.cols)));
        }
        if (m.r0 <= m.r1 && m.c0 <= m.c1) { // the farthest corner of the box (row bands: of the band's part of it, if any)
            ctl->max_level = lab_umax(room_absdiff(m.r0, m.sr), room_absdiff(m.r1, m.sr)) + lab_umax(room_absdiff(m.c0, m.sc), room_absdiff(m.c1, m.sc));
        }
        for (unsigned long long q = 0; q < ctl->tail; q++) {
            p.queue[q & p.qmask] = Q_EMPTY;
        }
        if (ctl->src_tile[0] != 0xFFFFFFFFu) {
            p.plane[0].state[ctl->src_tile[0]] = 0;
        }
        ctl->tail = 0;
        ctl->pending = 0;
        ctl->done = 1;
        ctl->n_settled = tc->V;
        tc->used = 2;
        return;
    }
    bool const lattice = tc->ok != 0 && lat_made(tc); // the lattice road made the field (it marks nothing itself)
    if (!tree_live(tc) && !lattice) {
        // the wave takes over: it needs a field of INF and empty visited masks (search_prepare left the
        // reset of the field to whoever writes it: the fix-up above, or this)
        long long const cells = static_cast<long long>(p.g.rows) * p.g.cols;
        for (long long i = gthread; i < cells; i += nthreads) {
            p.plane[0].dist[i] = INF;
        }
        for (long long i = gthread; i < static_cast<long long>(p.g.ntiles) * TS; i += nthreads) {
            p.plane[0].visited[i] = 0;
        }
        return;
    }
    bool const spec_final = lattice && ctl->spec_mark != 0; // pack_body's speculative marks are exactly the squares reached
    if ((!tr.mark_bits || lattice) && !spec_final) { // (else the squares are marked already and measure_body has nothing to look at)
        for (long long i = gthread; i < static_cast<long long>(p.g.ntiles) * TS; i += nthreads) {
            p.plane[0].visited[i] = p.path[i]; // one tree: every passable square was reached
        }
    }
    if (gthread != 0) {
        return;
    }
    if (tr.mark_bits && !lattice) {
        ctl->measured = 1u;
        ctl->n_marked = tc->marked;
    }
    if (spec_final) {
        ctl->measured = 1u;
        ctl->n_marked = tr.banded ? tc->lat_settled_band : tc->lat_settled;
    }
    for (uint32_t j = 0; j < ctl->n_targets && !ctl->painted; j++) { // (painted: lat_targets_body has taken them; the field may not exist)
        ctl->target_val[j] = p.plane[0].dist[ctl->target_cell[j]];
    }
    for (unsigned long long q = 0; q < ctl->tail; q++) {
        p.queue[q & p.qmask] = Q_EMPTY;
    }
    if (ctl->src_tile[0] != 0xFFFFFFFFu) {
        p.plane[0].state[ctl->src_tile[0]] = 0;
    }
    ctl->tail = 0;
    ctl->pending = 0;
    ctl->done = 1;
    ctl->n_settled = lattice ? tc->lat_settled : tc->local_settled;
    tc->used = lattice ? 3u : 1u;
}

} // namespace lab
