// lattice.cuh -- the reference builders' perfect mazes as what they are: a spanning tree of a CELL LATTICE.
//
// Every builder of the reference without a -m modification (kruskal, wilson, rdfs, ...; builders/*.cc) carves a
// perfect maze on the lattice of its odd squares: cells at (odd, odd), a passage square between two
// neighbouring cells at (odd, even) / (even, odd), (even, even) squares always wall.  The
// spanning-tree pipeline of tree.cuh takes any tree of squares and pays for that with divergent, square-by-square
// walks (12 of 32 lanes busy, profiles/r01b_*).  On the lattice a 64x64-square tile is a 32x32 CELL board that
// fits one 32-bit word per lane, and everything a tile has to do becomes a few warp-wide bit operations:
//
//   1 link    components of a tile: the vertical runs of a cell column are bit operations on the column's word, the
//             W passages join them through a union-find in shared memory (one union per passage).  The tree is
//             planar, so the Euler tour that enters a component through one border crossing leaves it through the
//             NEXT crossing of the same component clockwise around the tile: no wall following at all, the tour's
//             successor pointers come from the labels of the <= 128 crossing slots (match_any).
//   2 rank    hops to the end of the tour for every crossing arc, hierarchically: inside a super-tile of
//             LST x LST tiles by in-place pointer jumping in shared memory (one warp, single-word (pointer,
//             hops) entries, so any interleaving of reads and writes is valid and no barrier is needed), then
//             over the super-tiles' border arcs only (1/LST of all arcs, L2 resident) by the same barrier-free
//             in-place jumping in global memory, then one gather per arc.
//   3 flood   of the two arcs of a crossing the one met first on the tour leads away from the root: every
//             component knows its entry crossing.  One multi-source bit-board BFS per tile from all entry cells
//             gives every cell its depth below its component's entry, recorded bit-sliced (10 planes).  The
//             depth of a crossing cell + 1 is the weight of the tour's arc through it: +w going down, -w coming
//             back, scattered to the arcs' tour positions.
//   4 scan    an ordinary prefix sum over the tour positions: the depth of every component's entry.  (The second
//             pointer jumping of tree.cuh is gone: the ranks of phase 2 sort the list.)
//   5 levels  level of a cell = 2 * (entry depth + local depth); of a passage = the deeper of its two cells - 1.
//             Written once, whole rows of the tile, from registers.
//
// Checked, not assumed: the lattice shape (pillars are walls, a passage has both its cells), E == V - 1
// (tree_count_body), every arc reaches the end of the tour, the floods reach all V squares.  If anything fails
// nothing has been written to the level field and the general pipeline (tree.cuh), then the tile wave
// (engine.cuh), take over.  Results are bit-identical on all three roads (tests/test_lattice.py).
#pragma once
#include "post.cuh"
#include "tree.cuh"

namespace lab {

constexpr int LC = 32;             // cells per tile side
constexpr uint32_t LSLOTS = 128;   // crossing slots of a tile: side * 32 + index, sides CLOCKWISE
enum Lside : uint32_t { L_N = 0, L_E = 1, L_S = 2, L_W = 3 }; // index = cell column (N, S) / cell row (E, W)
constexpr uint32_t L_END = 0xFFFFFFFEu, L_NONE = 0xFFFFFFFFu;
constexpr uint32_t L_LINK = 0xFFFFFFFDu; // the first root's tour is over: the list goes on with the second root's (Tree_ctl::lat_head_b)
constexpr uint32_t LO_NONE = 0xFFu, LO_END = 0xFEu, LO_LINK = 0xFDu; // Lattice::outslot's markers (slots are 0 .. 127)
constexpr int LST = 4;                                  // super-tile side, tiles
constexpr uint32_t LB = 4u * LST * LC;                  // border slots of a super-tile
constexpr uint32_t LSUP_SLOTS = LST * LST * LSLOTS;     // slots of a super-tile (8192: 13-bit local index)
constexpr int L_DEPTH_PLANES = 10;                          // a cell's depth below its component's entry, bit-sliced
constexpr uint32_t L_LINK_SM = 1536;                    // words of shared memory per warp: lat_link_tile, lat_levels_tile
constexpr uint32_t L_CHUNK = 1024;                      // tour positions per scan chunk

struct Lattice {
    uint32_t *mask;   // [ntiles][3][32]  cell boards in COLUMN layout (lane = cell column, bit = cell row): cells, W passage open, N passage open
    uint32_t *xmask;  // [ntiles][4]      open crossings per side (bit = index)
    uint8_t *outslot; // [ntiles][128]    the slot the tour LEAVES the tile through after entering through this one; LO_END, LO_LINK, LO_NONE
    uint16_t *label;  // [ntiles][128]    component of the slot's cell (= the component's lowest piece id)
    uint16_t *cmap;   // [ntiles][1024]   component of the vertical run that begins at cell (row, column), at [row * 32 + column]
    uint64_t *xd;     // [ntiles][128]    after the local ranking: (dense border slot the chain leaves the super-tile to | hops to there << 32)
    uint64_t *gb;     // [nsuper][LB]     the border arcs' list: (next dense border slot | hops << 32), jumped in place
    uint32_t *R;      // [ntiles][128]    hops to the end of the tour (>= 1; larger = earlier)
    uint32_t *dplane; // [ntiles][10][32] local depth of every cell, bit-sliced
    uint32_t *emask;  // [ntiles][4]      slots that are their component's entry
    uint32_t *wsum;   // [nslots + 2]     arc weights by R, then their inclusive prefix sums per chunk
    uint32_t *chunk;  // [nslots / L_CHUNK + 2] exclusive prefix of the chunk totals
    Tree_ctl *tc;
    int sup_y, sup_x, nsuper;
    uint32_t nslots; // ntiles * 128
    // Row bands (one band of a taller maze per GPU).  Bands are whole super-tile rows.  Arc ids, `R`, the
    // dense border slots of `gb` and the tour positions of `wsum` are GLOBAL (the whole maze's); the per-tile arrays
    // hold the band's tiles only: local tile t is global tile tile_base + t, local super-tile S is global sup_base + S.
    // gb and wsum are whole-maze arrays on every rank: a rank fills in its own part -- into every peer's copy, by
    // stores over NVLink from inside lat_rank_local_super / lat_flood_tile (peers > 1), or into its own copy alone,
    // the copies then being joined by an all-gather / all-reduce (peers == 1) -- and reads the whole.
    // Single GPU: tile_base = sup_base = 0, the *_global fields equal the local ones, peers = 1.
    uint32_t tile_base, sup_base;
    int nsuper_global;
    uint32_t nslots_global;
    int peers;                 // copies of gb / wsum this rank writes (its own is [0])
    uint64_t *gb_peer[8];
    uint32_t *wsum_peer[8];
};

LAB_DEV bool lat_live(Tree_ctl const *tc) { return tc->ok != 0 && tc->lat_on != 0 && tc->lat_err == 0; }

// bits 0, 2, 4, ... of a 64-bit word, packed
LAB_DEV uint32_t lat_even16(uint32_t x) {
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}
LAB_DEV uint32_t lat_even(uint64_t x) { return lat_even16(static_cast<uint32_t>(x)) | (lat_even16(static_cast<uint32_t>(x >> 32)) << 16); }
LAB_DEV uint32_t lat_odd(uint64_t x) { return lat_even(x >> 1); }

// 32 x 32 bit matrix, one row per lane -> one column per lane
LAB_DEV uint32_t lat_transpose(uint32_t x, int lane) {
    uint32_t const m[5] = {0x0000FFFFu, 0x00FF00FFu, 0x0F0F0F0Fu, 0x33333333u, 0x55555555u};
#pragma unroll
    for (int k = 0; k < 5; k++) {
        int const s = 16 >> k;
        uint32_t const y = lab_shfl(x, lane ^ s);
        x = (lane & s) ? ((x & ~m[k]) | ((y & ~m[k]) >> s)) : ((x & m[k]) | ((y & m[k]) << s));
    }
    return x;
}

// the arc entering the neighbouring tile through the same crossing
LAB_DEV uint32_t lat_twin(Geom const &g, uint32_t t, uint32_t slot) {
    uint32_t const side = slot >> 5, idx = slot & 31u;
    switch (side) {
    case L_N: return (t - static_cast<uint32_t>(g.tiles_x)) * LSLOTS + L_S * LC + idx;
    case L_S: return (t + static_cast<uint32_t>(g.tiles_x)) * LSLOTS + L_N * LC + idx;
    case L_W: return (t - 1u) * LSLOTS + L_E * LC + idx;
    default: return (t + 1u) * LSLOTS + L_W * LC + idx;
    }
}

// The root.  A start on a cell (odd, odd) is the tree's root.  A start on a PASSAGE square (the solvers' starts are
// any path square: solve_utilities.cc:275-285) is taken out of the lattice: what is left are two trees, rooted at the
// passage's two cells, both at level 1; their tours are chained into one list (L_LINK) and every level is
// 2 * depth + 1.  The passage itself gets level 0 from the tile that holds it.
struct Lat_roots {
    int n;               // root cells
    uint32_t tile[2];
    int row[2], col[2];  // cell coordinates inside the tile
    bool passage, horizontal;
    uint32_t ptile;      // the passage: W passage (horizontal) / N passage of cell (pi, pj) of tile ptile
    int pi, pj;
};
LAB_DEV Lat_roots lat_roots(Params const &p) {
    Lat_roots r{};
    uint32_t const t0 = p.ctl->src_tile[0];
    if (t0 == 0xFFFFFFFFu) { // row bands: the root is in another band (and on a cell: lab_band_lat_link)
        r.ptile = t0;
        return r;
    }
    int const sr = static_cast<int>(p.ctl->src_rc[0] >> 8), sc = static_cast<int>(p.ctl->src_rc[0] & 0xFFu);
    int const i = sr >> 1, j = sc >> 1;
    r.ptile = t0;
    r.pi = i;
    r.pj = j;
    if ((sr & 1) && (sc & 1)) {
        r.n = 1;
        r.tile[0] = t0;
        r.row[0] = i;
        r.col[0] = j;
        return r;
    }
    r.n = 2;
    r.passage = true;
    r.horizontal = (sr & 1) != 0;
    r.tile[1] = t0; // the cell the passage belongs to
    r.row[1] = i;
    r.col[1] = j;
    if (r.horizontal) { // the cell to the left
        r.tile[0] = j ? t0 : t0 - 1u;
        r.row[0] = i;
        r.col[0] = j ? j - 1 : LC - 1;
    } else { // the cell above
        r.tile[0] = i ? t0 : t0 - static_cast<uint32_t>(p.g.tiles_x);
        r.row[0] = i ? i - 1 : LC - 1;
        r.col[0] = j;
    }
    return r;
}

// A tile's boards in column layout and the step masks of the bit-board BFS.
struct Lat_tile {
    uint32_t C, Wo, No;         // cells / W passage of the cell open / N passage of the cell open
    uint32_t XN, XE, XS, XW;    // open crossings (uniform)
    uint32_t mS, mN, mE, mW;    // a frontier bit may move south / north / east / west INTO these cells
};
LAB_DEV void lat_step_masks(Lat_tile &T, int lane) {
    T.mS = T.No & ~1u;                                   // into row i from row i-1: the N passage of (i, j); bit 0 is a crossing
    T.mN = T.No >> 1;                                    // into row i-1 from row i
    T.mE = lane ? T.Wo : 0u;                             // into column j from column j-1: the W passage of (i, j); lane 0: crossings
    uint32_t const right = lab_shfl_down(T.Wo, 1);
    T.mW = lane < 31 ? right : 0u;                       // into column j from column j+1
}
LAB_DEV Lat_tile lat_load(Lattice const &lt, long long t, int lane) {
    Lat_tile T;
    uint32_t const *m = lt.mask + t * 96;
    T.C = lab_ld_ro(m + lane);
    T.Wo = lab_ld_ro(m + 32 + lane);
    T.No = lab_ld_ro(m + 64 + lane);
    uint32_t const *x = lt.xmask + t * 4;
    T.XN = lab_ld_ro(x + L_N);
    T.XE = lab_ld_ro(x + L_E);
    T.XS = lab_ld_ro(x + L_S);
    T.XW = lab_ld_ro(x + L_W);
    lat_step_masks(T, lane);
    return T;
}
// one level of the bit-board BFS: F = the cells reached now, V = every cell reached so far
LAB_DEV void lat_step(Lat_tile const &T, uint32_t &F, uint32_t &V) {
    uint32_t const west = lab_shfl_up(F, 1), east = lab_shfl_down(F, 1);
    uint32_t a = (F << 1) & T.mS;
    a |= (F >> 1) & T.mN;
    a |= west & T.mE;
    a |= east & T.mW;
    F = a & ~V;
    V |= F;
}

// ---------------------------------------------------------------------------------------------
// 1 link.  Warp per tile.  sm: L_LINK_SM words of shared memory per warp.
LAB_DEV void lat_link_tile(Params const &p, Lattice const &lt, uint32_t t, int lane, uint32_t *sm, uint32_t &arcs) {
    Geom const &g = p.g;
    Tree_ctl *tc = lt.tc;
    Lat_roots const roots = lat_roots(p);
    uint64_t const P0 = lab_ld_ro(p.path + static_cast<long long>(t) * TS + 2 * lane), P1 = lab_ld_ro(p.path + static_cast<long long>(t) * TS + 2 * lane + 1);
    Lat_tile T;
    T.XN = lat_odd(lab_ld_ro(p.cross + static_cast<long long>(t) * 4 + SIDE_N));
    T.XE = lat_odd(lab_ld_ro(p.cross + static_cast<long long>(t) * 4 + SIDE_E));
    T.XS = lat_odd(lab_ld_ro(p.cross + static_cast<long long>(t) * 4 + SIDE_S));
    T.XW = lat_odd(lab_ld_ro(p.cross + static_cast<long long>(t) * 4 + SIDE_W));
    {
        // row layout first (lane = cell row i): the lattice's shape
        uint32_t const rc = lat_odd(P1), rw = lat_even(P1), rn = lat_odd(P0), pillars = lat_even(P0);
        uint32_t const above = lab_shfl_up(rc, 1);
        bool bad = pillars != 0;
        bad |= (rw & ~rc) != 0 || (rw & ~((rc << 1) | ((T.XW >> lane) & 1u))) != 0;          // a W passage has its cell and the cell to the left
        bad |= (rn & ~rc) != 0 || (rn & ~(lane ? above : T.XN)) != 0;                          // an N passage has its cell and the cell above
        if (lab_any(bad)) {
            if (lane == 0 && lab_ld_ro(&tc->lat_err) == 0u) {
                lab_st_volatile(&tc->lat_err, 1u);
            }
            return;
        }
        // a root on a passage square: that passage is closed (lat_roots)
        uint32_t rw2 = rw, rn2 = rn;
        if (roots.passage) {
            if (roots.ptile == t && lane == roots.pi) {
                if (roots.horizontal) rw2 &= ~(1u << roots.pj);
                else rn2 &= ~(1u << roots.pj);
            }
            if (roots.horizontal && roots.pj == 0) {
                if (roots.ptile == t) T.XW &= ~(1u << roots.pi);
                if (roots.ptile == t + 1u) T.XE &= ~(1u << roots.pi);
            }
            if (!roots.horizontal && roots.pi == 0) {
                if (roots.ptile == t) T.XN &= ~(1u << roots.pj);
                if (roots.ptile == t + static_cast<uint32_t>(g.tiles_x)) T.XS &= ~(1u << roots.pj);
            }
        }
        T.C = lat_transpose(rc, lane);
        T.Wo = lat_transpose(rw2, lane);
        T.No = lat_transpose(rn2, lane);
    }
    lat_step_masks(T, lane);
    {
        uint32_t *m = lt.mask + static_cast<long long>(t) * 96;
        m[lane] = T.C;
        m[32 + lane] = T.Wo;
        m[64 + lane] = T.No;
        if (lane < 4) {
            lt.xmask[static_cast<long long>(t) * 4 + lane] = lane == L_N ? T.XN : (lane == L_E ? T.XE : (lane == L_S ? T.XS : T.XW));
        }
    }
    arcs += lane == 0 ? static_cast<uint32_t>(lab_popc64(T.XN) + lab_popc64(T.XE) + lab_popc64(T.XS) + lab_popc64(T.XW)) : 0u;

    // ---- A: components.  A column's cells fall into vertical RUNS (cells joined through open N passages); bit i of
    // `st`: cell (i, lane) begins one.  A run is named after its first cell, id = row * 32 + column, so the runs cost
    // nothing; the open W passages join runs of neighbouring columns: a union-find over the ids in shared memory (the
    // larger root is hooked under the smaller: pointers only ever lead to smaller ids, no cycle can form, and a
    // component's name is its lowest run id).  One union per W passage, ~16 per lane. ----
    uint16_t *parent = reinterpret_cast<uint16_t *>(sm);                 // [1024]
    uint8_t *first = reinterpret_cast<uint8_t *>(sm + 512);              // [4][1024] see C; all 0xFF between tiles (lat_link_body)
    uint32_t const st = T.C & ~(T.No & ~1u);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        int const w = k * 32 + lane; // ids 2w, 2w + 1
        sm[w] = static_cast<uint32_t>(2 * w) | (static_cast<uint32_t>(2 * w + 1) << 16);
    }
    lab_syncwarp();
    auto run_of = [](uint32_t starts, int i) { return 31 - lab_clz32(starts & (0xFFFFFFFFu >> (31 - i))); }; // first row of cell i's run
    auto find = [&](uint32_t a) {
        // path halving (a stale grandparent is still an ancestor: roots only ever get hooked under smaller ids)
        for (uint32_t pa; (pa = lab_ld_volatile(parent + a)) != a;) {
            uint32_t const ga = lab_ld_volatile(parent + pa);
            if (ga != pa) lab_st_volatile(parent + a, static_cast<uint16_t>(ga));
            a = ga;
        }
        return a;
    };
    {
        // Warp-synchronous rounds instead of atomics (a shared-memory CAS costs the LSU 2-4 cycles per LANE): every lane
        // finds the two roots of its next passage; after a barrier all lanes hook at once with a plain store -- roots are
        // only changed by hooks, so every lane still holds true roots; lanes hooking the SAME root collide, one store
        // wins, the others see it after the next barrier and go round again.
        uint32_t const st_left = lab_shfl_up(st, 1);
        uint32_t w = lane ? T.Wo : 0u; // (lane 0's W passages are crossings)
        uint32_t a = 0, b = 0;
        bool busy = false;
        for (int round = 0; round < 2048; round++) {
            if (!busy && w) {
                int const i = lab_ffs32(w) - 1;
                w &= w - 1u;
                a = static_cast<uint32_t>(run_of(st, i) * 32 + lane);
                b = static_cast<uint32_t>(run_of(st_left, i) * 32 + lane - 1);
                busy = true;
            }
            if (!lab_any(busy)) {
                break;
            }
            uint32_t hi = 0, lo = 0;
            if (busy) {
                a = find(a);
                b = find(b);
                hi = a > b ? a : b;
                lo = a > b ? b : a;
                busy = a != b; // (equal: a cycle, not a tree; the counts or the tour will say so)
            }
            lab_syncwarp();
            if (busy) {
                lab_st_volatile(parent + hi, static_cast<uint16_t>(lo));
            }
            lab_syncwarp();
            if (busy && lab_ld_volatile(parent + hi) == lo) {
                busy = false;
            }
        }
    }
    lab_syncwarp();
    // every id's component, flat (parent[id] = root), by pointer doubling over all 1024 ids in step: 32 independent
    // two-hop look-ups per lane and pass, no divergence (a walk to the root per run is a long tail of a few lanes).
    // cmap: the flat array as it is, for lat_levels_tile (ids that begin no run hold themselves).
    for (int pass = 0; pass < 12; pass++) {
        bool changed = false;
#pragma unroll
        for (int k = 0; k < 32; k++) {
            uint32_t const pa = parent[k * 32 + lane], ga = parent[pa];
            if (ga != pa) {
                parent[k * 32 + lane] = static_cast<uint16_t>(ga);
                changed = true;
            }
        }
        lab_syncwarp();
        if (!lab_any(changed)) {
            break;
        }
    }
    {
        uint32_t *cm = reinterpret_cast<uint32_t *>(lt.cmap + static_cast<long long>(t) * 1024);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            cm[k * 32 + lane] = sm[k * 32 + lane];
        }
    }
    // ---- C: the tour: from a slot to the next slot of the same component, clockwise (N ->, E down, S <-, W up) ----
    uint32_t labs[4] = {0xFFFFu, 0xFFFFu, 0xFFFFu, 0xFFFFu}; // component of slots (side, lane)
    {
        uint32_t const st_w = lab_shfl(st, 0), st_e = lab_shfl(st, 31); // W / E slots: lane = cell row, columns 0 / 31
        if ((T.XN >> lane) & 1u) labs[L_N] = parent[lane];
        if ((T.XS >> lane) & 1u) labs[L_S] = parent[run_of(st, 31) * 32 + lane];
        if ((T.XW >> lane) & 1u) labs[L_W] = parent[run_of(st_w, lane) * 32];
        if ((T.XE >> lane) & 1u) labs[L_E] = parent[run_of(st_e, lane) * 32 + 31];
    }
    uint32_t root_label[2] = {0xFFFFu, 0xFFFFu};
#pragma unroll
    for (int q = 0; q < 2; q++) {
        if (roots.n > q && roots.tile[q] == t) { // (uniform)
            uint32_t const st_r = lab_shfl(st, roots.col[q]);
            root_label[q] = parent[run_of(st_r, roots.row[q]) * 32 + roots.col[q]];
            if (lane == 0) lab_st_volatile(&tc->lat_root_label[q], root_label[q]);
        }
    }
    // first[side][component]: the component's first slot on that side in clockwise order (0xFF: none)
    uint32_t peers[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        peers[s] = lab_match_any(labs[s]);
        if (labs[s] != 0xFFFFu) {
            bool const ascending = s == L_N || s == L_E;
            int const lead = ascending ? lab_ffs32(peers[s]) - 1 : 31 - lab_clz32(peers[s]);
            if (lead == lane) {
                first[s * 1024 + labs[s]] = static_cast<uint8_t>(lane);
            }
        }
    }
    lab_syncwarp();
    // a root component's first slot clockwise: its tour is cut there (the last root's: the end of the list; with two
    // roots the first one's tour is followed by the second's)
    uint32_t root_first[2] = {L_NONE, L_NONE};
#pragma unroll
    for (int q = 0; q < 2; q++) {
        if (root_label[q] != 0xFFFFu) {
            for (int s = 3; s >= 0; s--) {
                uint32_t const f = first[s * 1024 + root_label[q]];
                if (f != 0xFFu) root_first[q] = static_cast<uint32_t>(s) * LC + f;
            }
        }
    }
    uint32_t const gt = t + lt.tile_base; // the tile's id in the whole maze (row bands)
#pragma unroll
    for (int s = 0; s < 4; s++) {
        uint32_t out = LO_NONE;
        uint32_t const lab = labs[s];
        if (lab != 0xFFFFu) {
            bool const ascending = s == L_N || s == L_E;
            uint32_t const rest = ascending ? (peers[s] & ~((2u << lane) - 1u)) : (peers[s] & ((1u << lane) - 1u));
            uint32_t nslot;
            if (rest) {
                nslot = static_cast<uint32_t>(s) * LC + static_cast<uint32_t>(ascending ? lab_ffs32(rest) - 1 : 31 - lab_clz32(rest));
            } else {
                nslot = L_NONE;
#pragma unroll
                for (int d = 1; d <= 4; d++) {
                    int const s2 = (s + d) & 3;
                    uint32_t const f = first[s2 * 1024 + lab];
                    if (nslot == L_NONE && f != 0xFFu) {
                        nslot = static_cast<uint32_t>(s2) * LC + f;
                    }
                }
            }
            uint32_t const me = static_cast<uint32_t>(s) * LC + static_cast<uint32_t>(lane);
            out = nslot;
            if (roots.n > 0 && me == root_first[roots.n - 1]) { // the last root's tour ends here ...
                if (roots.n == 2) {
                    lab_st_volatile(&tc->lat_head_b, lat_twin(g, gt, nslot) + 1u); // ... and would have begun with this arc
                }
                out = LO_END;
            } else if (roots.n == 2 && me == root_first[0]) {
                out = LO_LINK;
            }
        }
        lt.outslot[static_cast<long long>(t) * LSLOTS + s * LC + lane] = static_cast<uint8_t>(out);
        lt.label[static_cast<long long>(t) * LSLOTS + s * LC + lane] = static_cast<uint16_t>(lab);
    }
    lab_syncwarp();
#pragma unroll
    for (int s = 0; s < 4; s++) { // `first` back to all 0xFF for the warp's next tile
        if (labs[s] != 0xFFFFu) {
            first[s * 1024 + labs[s]] = 0xFFu;
        }
    }
    lab_syncwarp();
}

LAB_DEV void lat_link_body(Params const &p, Lattice const &lt, int lane, uint32_t *sm) {
    if (!lat_live(lt.tc)) {
        return;
    }
    uint32_t arcs = 0;
#pragma unroll
    for (int k = 0; k < 32; k++) {
        sm[512 + k * 32 + lane] = 0xFFFFFFFFu; // lat_link_tile's `first`
    }
    lab_syncwarp();
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) {
            t = lab_atomic_add(&lt.tc->lat_ticket[0], 1u);
        }
        t = lab_shfl(t, 0);
        if (t >= static_cast<uint32_t>(p.g.ntiles)) {
            break;
        }
        lat_link_tile(p, lt, t, lane, sm, arcs);
    }
    if (lane == 0 && arcs) {
        lab_atomic_add(&lt.tc->n_arcs, arcs);
    }
}

// ---------------------------------------------------------------------------------------------
// 2a local ranking.  Warp per super-tile; sm: LSUP_SLOTS words.  Entry: bit 31 = the chain has left the
// super-tile (then bits 0-12 = the LAST slot inside), else bits 0-12 = a slot further along; bits 13-30 = hops.
// One word per entry, read and written whole: whatever mix of old and new entries a lane sees, "slot a
// reaches slot x in d hops" stays true, so the jumping runs in place with no barrier between rounds.
constexpr uint32_t LX_OUT = 0x80000000u, LX_IDX = 0x1FFFu;
LAB_DEV uint32_t lat_dense_border(Lattice const &lt, Geom const &g, uint32_t arc) { // arc: an inbound border slot of its super-tile
    uint32_t const t = arc / LSLOTS, slot = arc % LSLOTS, side = slot >> 5, idx = slot & 31u;
    uint32_t const ty = t / static_cast<uint32_t>(g.tiles_x), tx = t % static_cast<uint32_t>(g.tiles_x);
    uint32_t const S = (ty / LST) * static_cast<uint32_t>(lt.sup_x) + tx / LST;
    uint32_t const along = (side == L_N || side == L_S) ? tx % LST : ty % LST;
    return S * LB + side * (LST * LC) + along * LC + idx;
}
// Where the tour goes after entering super-tile-local tile (lt_y, lt_x) through a slot whose out-slot is `o` (< 128):
// the neighbouring tile on o's side, through the mirrored slot.  (dy, dx, slot there)
LAB_DEV void lat_step_out(uint32_t o, int &dy, int &dx, uint32_t &slot) {
    uint32_t const side = o >> 5, idx = o & 31u;
    dy = side == L_N ? -1 : (side == L_S ? 1 : 0);
    dx = side == L_W ? -1 : (side == L_E ? 1 : 0);
    slot = ((side + 2u) & 3u) * LC + idx; // N <-> S, E <-> W
}
LAB_DEV void lat_rank_local_super(Params const &p, Lattice const &lt, uint32_t S, int lane, uint32_t *sm) {
    Geom const &g = p.g;
    // (sy0, sx0): the super-tile's first tile in the WHOLE maze; gy0: the band's first tile row there
    uint32_t const SG = S + lt.sup_base, gy0 = lt.tile_base / static_cast<uint32_t>(g.tiles_x);
    uint32_t const sup_x = static_cast<uint32_t>(lt.sup_x);
    uint32_t const sy0 = (SG / sup_x) * LST, sx0 = (SG % sup_x) * LST;
    // (the second root's first arc, if the list is two tours chained: the one entry that is not a neighbour's slot)
    uint32_t const head_b = lab_ld_cg(&lt.tc->lat_head_b);
    auto local_arc = [&](uint32_t lt_y, uint32_t lt_x, uint32_t slot) { // a tile of this super-tile (and of this band)
        return (static_cast<long long>(sy0 + lt_y - gy0) * g.tiles_x + sx0 + lt_x) * LSLOTS + slot;
    };
    // where the chain goes from entry i: -> (global tile row, column, slot), or none (the end of the list)
    auto successor = [&](uint32_t lt_y, uint32_t lt_x, uint32_t o, uint32_t &ny, uint32_t &nx, uint32_t &nslot) {
        if (o == LO_LINK) {
            if (!head_b) return false;
            uint32_t const a = head_b - 1u, t = a / LSLOTS;
            ny = t / static_cast<uint32_t>(g.tiles_x);
            nx = t % static_cast<uint32_t>(g.tiles_x);
            nslot = a % LSLOTS;
            return true;
        }
        if (o >= LO_LINK) return false; // END (or no crossing)
        int dy, dx;
        lat_step_out(o, dy, dx, nslot);
        ny = sy0 + lt_y + static_cast<uint32_t>(dy);
        nx = sx0 + lt_x + static_cast<uint32_t>(dx);
        return true;
    };
    for (uint32_t i = static_cast<uint32_t>(lane); i < LSUP_SLOTS; i += 32u) {
        uint32_t const lt_y = i / (LST * LSLOTS), lt_x = (i / LSLOTS) % LST, slot = i % LSLOTS;
        uint32_t e = LX_OUT | i;
        if (sy0 + lt_y - gy0 < static_cast<uint32_t>(g.tiles_y) && sx0 + lt_x < static_cast<uint32_t>(g.tiles_x)) {
            uint32_t const o = lab_ld_cg(lt.outslot + local_arc(lt_y, lt_x, slot));
            uint32_t ny, nx, nslot;
            if (successor(lt_y, lt_x, o, ny, nx, nslot) && ny - sy0 < static_cast<uint32_t>(LST) && nx - sx0 < static_cast<uint32_t>(LST)) {
                e = (((ny - sy0) * LST + (nx - sx0)) * LSLOTS + nslot) | (1u << 13);
            }
        }
        sm[i] = e;
    }
    lab_syncwarp();
    bool pending = true;
    for (int pass = 0; pass < 40 && pending; pass++) {
        bool mine = false;
        for (uint32_t i = static_cast<uint32_t>(lane); i < LSUP_SLOTS; i += 32u) {
            uint32_t const e = sm[i];
            if (!(e & LX_OUT)) {
                uint32_t const e2 = sm[e & LX_IDX];
                uint32_t const ne = (e2 & (LX_OUT | LX_IDX)) | ((e & ~(LX_OUT | LX_IDX)) + (e2 & ~(LX_OUT | LX_IDX)));
                sm[i] = ne;
                mine |= !(ne & LX_OUT);
            }
        }
        pending = lab_any(mine);
    }
    if (pending) { // a chain that never leaves: a cycle, not a tree
        if (lane == 0) lab_st_volatile(&lt.tc->lat_err, 3u);
    }
    lab_syncwarp();
    for (uint32_t i = static_cast<uint32_t>(lane); i < LSUP_SLOTS; i += 32u) {
        uint32_t const lt_y = i / (LST * LSLOTS), lt_x = (i / LSLOTS) % LST, slot = i % LSLOTS, side = slot >> 5;
        bool const present = sy0 + lt_y - gy0 < static_cast<uint32_t>(g.tiles_y) && sx0 + lt_x < static_cast<uint32_t>(g.tiles_x);
        uint64_t pair = 0xFFFFFFFFull; // (no arc: the end, 0 hops)
        bool open = false;
        if (present) {
            long long const a = local_arc(lt_y, lt_x, slot);
            open = lab_ld_cg(lt.outslot + a) != LO_NONE;
            if (open) {
                uint32_t const e = sm[i];
                uint32_t const last = e & LX_IDX, hops = ((e & ~LX_OUT) >> 13) + 1u;
                uint32_t const ly = last / (LST * LSLOTS), lx = (last / LSLOTS) % LST;
                uint32_t const o = lab_ld_cg(lt.outslot + local_arc(ly, lx, last % LSLOTS));
                uint32_t ny, nx, nslot, dense = 0xFFFFFFFFu;
                if ((e & LX_OUT) && successor(ly, lx, o, ny, nx, nslot)) { // an inbound border slot of another super-tile
                    uint32_t const nside = nslot >> 5;
                    uint32_t const along = (nside == L_N || nside == L_S) ? nx % LST : ny % LST;
                    dense = ((ny / LST) * sup_x + nx / LST) * LB + nside * (LST * LC) + along * LC + (nslot & 31u);
                }
                pair = static_cast<uint64_t>(dense) | (static_cast<uint64_t>(hops) << 32);
                lab_st_cg(lt.xd + a, pair);
            }
        }
        bool const border = (side == L_N && lt_y == 0) || (side == L_S && lt_y == LST - 1) || (side == L_W && lt_x == 0) || (side == L_E && lt_x == LST - 1);
        if (border) {
            uint32_t const along = (side == L_N || side == L_S) ? lt_x : lt_y;
            long long const at = static_cast<long long>(SG) * LB + side * (LST * LC) + along * LC + (slot & 31u);
            for (int q = 0; q < lt.peers; q++) { // every rank's copy of the border list (stores over NVLink)
                lab_st_cg(lt.gb_peer[q] + at, open ? pair : 0xFFFFFFFFull);
            }
        }
    }
    lab_syncwarp();
}
LAB_DEV void lat_rank_local_body(Params const &p, Lattice const &lt, int lane, uint32_t *sm) {
    if (!lat_live(lt.tc)) {
        return;
    }
    for (;;) {
        uint32_t S = 0;
        if (lane == 0) {
            S = lab_atomic_add(&lt.tc->lat_ticket[1], 1u);
        }
        S = lab_shfl(S, 0);
        if (S >= static_cast<uint32_t>(lt.nsuper)) {
            break;
        }
        lat_rank_local_super(p, lt, S, lane, sm);
    }
}

// 2b the border arcs.  In place, no barrier: an entry is one 8-byte word, so a thread that reads another thread's
// entry sees SOME true statement "slot reaches x in d hops", old or new, and its own stays true.
LAB_DEV void lat_rank_global_body(Lattice const &lt, long long gthread, long long nthreads) {
    Tree_ctl *tc = lt.tc;
    if (!lat_live(tc)) {
        return;
    }
    long long const n = static_cast<long long>(lt.nsuper_global) * LB;
    uint32_t const limit = lt.nslots_global + 2u;
    for (int pass = 0; pass < (1 << 20); pass++) { // (a cycle ends through `limit`, not through the pass count)
        bool pending = false;
        for (long long i = gthread; i < n; i += nthreads) {
            uint64_t e = lab_ld_cg(lt.gb + i);
            uint32_t x = static_cast<uint32_t>(e), d = static_cast<uint32_t>(e >> 32);
            if (x == 0xFFFFFFFFu) {
                continue;
            }
            for (int hop = 0; hop < 6 && x != 0xFFFFFFFFu; hop++) {
                uint64_t const e2 = lab_ld_cg(lt.gb + x);
                x = static_cast<uint32_t>(e2);
                d += static_cast<uint32_t>(e2 >> 32);
            }
            if (d > limit) { // longer than the tour can be: a cycle
                lab_st_volatile(&tc->lat_err, 3u);
                return;
            }
            lab_st_cg(lt.gb + i, static_cast<uint64_t>(x) | (static_cast<uint64_t>(d) << 32));
            pending |= x != 0xFFFFFFFFu;
        }
        if (!pending) {
            return;
        }
        if (lab_ld_volatile(&tc->lat_err)) {
            return;
        }
    }
    lab_st_volatile(&tc->lat_err, 3u);
}

// 2c every arc: its own hops to the border + the border arc's hops to the end
LAB_DEV void lat_rank_expand_body(Lattice const &lt, long long gthread, long long nthreads) {
    Tree_ctl *tc = lt.tc;
    if (!lat_live(tc)) {
        return;
    }
    bool bad = false;
    for (long long a = gthread; a < static_cast<long long>(lt.nslots); a += nthreads) {
        uint32_t r = 0;
        if (lab_ld_cg(lt.outslot + a) != LO_NONE) {
            uint64_t const e = lab_ld_cg(lt.xd + a);
            uint32_t const x = static_cast<uint32_t>(e);
            r = static_cast<uint32_t>(e >> 32);
            if (x != 0xFFFFFFFFu) {
                uint64_t const b = lab_ld_cg(lt.gb + x);
                bad |= static_cast<uint32_t>(b) != 0xFFFFFFFFu;
                r += static_cast<uint32_t>(b >> 32);
            }
        }
        lt.R[a] = r;
    }
    if (bad) {
        lab_st_volatile(&tc->lat_err, 3u);
    }
}

// ---------------------------------------------------------------------------------------------
// 3 flood.  Warp per tile.  Entry slots (the arc into the tile is met before its twin), one BFS from all entry
// cells, local depths bit-sliced, the weights of the arcs through the tile's other crossings.
LAB_DEV uint32_t lat_plane_value(uint32_t const (&D)[L_DEPTH_PLANES], int bit) {
    uint32_t v = 0;
#pragma unroll
    for (int q = 0; q < L_DEPTH_PLANES; q++) {
        v |= ((D[q] >> bit) & 1u) << q;
    }
    return v;
}
template <int U> LAB_DEV void lat_depth_step(Lat_tile const &T, uint32_t &F, uint32_t &V, uint32_t &G, uint32_t (&D)[L_DEPTH_PLANES]) {
    lat_step(T, F, V);
    G |= F;
    if (U & 1) D[0] |= F;
    if (U & 2) D[1] |= F;
    if (U & 4) D[2] |= F;
}
LAB_DEV void lat_flood_tile(Params const &p, Lattice const &lt, uint32_t t, int lane, unsigned long long &settled, uint32_t &bad) {
    Geom const &g = p.g;
    Lat_tile const T = lat_load(lt, t, lane);
    uint32_t const X[4] = {T.XN, T.XE, T.XS, T.XW};
    long long const base = static_cast<long long>(t) * LSLOTS;
    uint32_t Ro[4], Rt[4];
    bool open[4], entry[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        open[s] = (X[s] >> lane) & 1u;
        Ro[s] = open[s] ? lab_ld_cg(lt.R + base + s * LC + lane) : 0u;
        Rt[s] = 0u;
        if (open[s]) {
            uint32_t const tw = lat_twin(g, t + lt.tile_base, static_cast<uint32_t>(s * LC + lane)); // (global arc id)
            uint32_t const twl = tw - lt.tile_base * LSLOTS;
            // a twin in another band is a border arc of its super-tile: its hops are in the (whole-maze) border list
            Rt[s] = twl < lt.nslots ? lab_ld_cg(lt.R + twl) : static_cast<uint32_t>(lab_ld_cg(lt.gb + lat_dense_border(lt, g, tw)) >> 32);
        }
        entry[s] = open[s] && Ro[s] > Rt[s];
        bad |= (open[s] && (Ro[s] == Rt[s] || Ro[s] == 0u || Rt[s] == 0u)) ? 1u : 0u;
    }
    uint32_t const eW = lab_ballot(entry[L_W]), eE = lab_ballot(entry[L_E]), eN = lab_ballot(entry[L_N]), eS = lab_ballot(entry[L_S]);
    if (lane < 4) {
        lt.emask[static_cast<long long>(t) * 4 + lane] = lane == L_N ? eN : (lane == L_E ? eE : (lane == L_S ? eS : eW));
    }
    uint32_t F = (entry[L_N] ? 1u : 0u) | (entry[L_S] ? 0x80000000u : 0u) | (lane == 0 ? eW : 0u) | (lane == 31 ? eE : 0u);
    Lat_roots const roots = lat_roots(p);
#pragma unroll
    for (int q = 0; q < 2; q++) {
        if (roots.n > q && roots.tile[q] == t && lane == roots.col[q]) F |= 1u << roots.row[q];
    }
    uint32_t V = F, G = 0;
    uint32_t D[L_DEPTH_PLANES];
#pragma unroll
    for (int q = 0; q < L_DEPTH_PLANES; q++) {
        D[q] = 0;
    }
    // depth = 8 * group + u: the seeds are (0, 0); the low three planes come from the unrolled position
    lat_depth_step<1>(T, F, V, G, D);
    lat_depth_step<2>(T, F, V, G, D);
    lat_depth_step<3>(T, F, V, G, D);
    lat_depth_step<4>(T, F, V, G, D);
    lat_depth_step<5>(T, F, V, G, D);
    lat_depth_step<6>(T, F, V, G, D);
    lat_depth_step<7>(T, F, V, G, D);
    for (uint32_t grp = 1; grp < 128u && lab_any(F != 0); grp++) {
        G = 0;
        lat_depth_step<0>(T, F, V, G, D);
        lat_depth_step<1>(T, F, V, G, D);
        lat_depth_step<2>(T, F, V, G, D);
        lat_depth_step<3>(T, F, V, G, D);
        lat_depth_step<4>(T, F, V, G, D);
        lat_depth_step<5>(T, F, V, G, D);
        lat_depth_step<6>(T, F, V, G, D);
        lat_depth_step<7>(T, F, V, G, D);
#pragma unroll
        for (int q = 3; q < L_DEPTH_PLANES; q++) {
            if ((grp >> (q - 3)) & 1u) {
                D[q] |= G;
            }
        }
    }
    {
        uint32_t *pl = lt.dplane + static_cast<long long>(t) * (L_DEPTH_PLANES * 32);
#pragma unroll
        for (int q = 0; q < L_DEPTH_PLANES; q++) {
            pl[q * 32 + lane] = D[q];
        }
    }
    // every cell reached <=> all the tile's passable squares have a level coming
    bad |= (V != T.C) ? 2u : 0u;
    settled += static_cast<unsigned>(lab_popc64(V) + lab_popc64(T.Wo) + lab_popc64(T.No));
    if (roots.passage && roots.ptile == t && lane == 0) {
        settled++; // the root passage (closed in the boards)
    }
    // weights: through a crossing that is not an entry the tour goes DOWN into the neighbour (+w on the twin arc)
    // and comes back UP (-w on this slot's arc); w = depth of the cell at the crossing + 1
    uint32_t DW[L_DEPTH_PLANES], DE[L_DEPTH_PLANES];
#pragma unroll
    for (int q = 0; q < L_DEPTH_PLANES; q++) {
        DW[q] = lab_shfl(D[q], 0);
        DE[q] = lab_shfl(D[q], 31);
    }
    uint32_t const ld[4] = {lat_plane_value(D, 0), lat_plane_value(DE, lane), lat_plane_value(D, 31), lat_plane_value(DW, lane)};
#pragma unroll
    for (int s = 0; s < 4; s++) {
        if (open[s] && !entry[s]) {
            uint32_t const w = ld[s] + 1u;
            for (int q = 0; q < lt.peers; q++) { // every rank's copy of the weights (stores over NVLink)
                lt.wsum_peer[q][Rt[s]] = w;
                lt.wsum_peer[q][Ro[s]] = 0u - w;
            }
        }
    }
}
LAB_DEV void lat_flood_body(Params const &p, Lattice const &lt, int lane) {
    Tree_ctl *tc = lt.tc;
    if (!lat_live(tc)) {
        return;
    }
    unsigned long long settled = 0;
    uint32_t bad = 0;
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) {
            t = lab_atomic_add(&tc->lat_ticket[2], 1u);
        }
        t = lab_shfl(t, 0);
        if (t >= static_cast<uint32_t>(p.g.ntiles)) {
            break;
        }
        lat_flood_tile(p, lt, t, lane, settled, bad);
    }
    uint32_t const lo = lab_reduce_add(static_cast<uint32_t>(settled)), hi = lab_reduce_add(static_cast<uint32_t>(settled >> 32));
    bad = lab_reduce_or(bad);
    if (lane == 0) {
        lab_atomic_add(reinterpret_cast<uint64_t *>(&tc->lat_settled), (static_cast<uint64_t>(hi) << 32) + lo);
        if (bad) {
            lab_atomic_or(&tc->lat_flood_bad, bad);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 4 scan.  wsum[0 .. n_arcs] (index = R; [0] is nobody's) -> inclusive prefix sums inside chunks of L_CHUNK,
// then the chunk totals' exclusive prefix.  Warp per chunk; the second pass is one warp.
LAB_DEV void lat_scan_chunks_body(Lattice const &lt, long long gwarp, long long nwarps, int lane) {
    Tree_ctl *tc = lt.tc;
    if (!lat_live(tc)) {
        return;
    }
    if (tc->lat_flood_bad || tc->lat_settled != tc->V) { // the floods did not reach every square: not one tree
        if (gwarp == 0 && lane == 0) lab_st_volatile(&tc->lat_err, 4u);
        return;
    }
    long long const n = static_cast<long long>(tc->n_arcs) + 1;
    long long const chunks = (n + L_CHUNK - 1) / L_CHUNK;
    for (long long c = gwarp; c < chunks; c += nwarps) {
        uint32_t carry = 0;
        uint32_t *w = lt.wsum + c * L_CHUNK;
        long long const left = n - c * L_CHUNK;
        uint32_t v[L_CHUNK / 32];
#pragma unroll
        for (int k = 0; k < static_cast<int>(L_CHUNK / 32); k++) {
            long long const i = 32 * k + lane;
            v[k] = (i < left && (c | i) != 0) ? w[i] : 0u; // ([0] was never written)
        }
#pragma unroll
        for (int k = 0; k < static_cast<int>(L_CHUNK / 32); k++) {
            uint32_t x = v[k];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t const y = lab_shfl_up(x, d);
                if (lane >= d) x += y;
            }
            x += carry;
            carry = lab_shfl(x, 31);
            if (32 * k + lane < left) {
                w[32 * k + lane] = x;
            }
        }
        if (lane == 0) {
            lt.chunk[c] = carry;
        }
    }
}
LAB_DEV void lat_scan_tops_body(Lattice const &lt, int lane) {
    Tree_ctl *tc = lt.tc;
    if (!lat_live(tc)) {
        return;
    }
    long long const n = static_cast<long long>(tc->n_arcs) + 1;
    long long const chunks = (n + L_CHUNK - 1) / L_CHUNK;
    uint32_t carry = 0;
    for (long long c0 = 0; c0 < chunks; c0 += 32) {
        long long const c = c0 + lane;
        uint32_t const own = c < chunks ? lt.chunk[c] : 0u;
        uint32_t x = own;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t const y = lab_shfl_up(x, d);
            if (lane >= d) x += y;
        }
        if (c < chunks) {
            lt.chunk[c] = carry + x - own; // exclusive
        }
        carry += lab_shfl(x, 31);
    }
    if (lane == 0 && carry != 0u) { // every +w has its -w: a total that is not zero means a broken tour
        lab_st_volatile(&tc->lat_err, 5u);
    }
}
// sum of the weights of all arcs with R' < r
LAB_DEV uint32_t lat_prefix_below(Lattice const &lt, uint32_t r) {
    uint32_t const i = r - 1u;
    return lab_ld_cg(lt.wsum + i) + lab_ld_cg(lt.chunk + i / L_CHUNK);
}

// ---------------------------------------------------------------------------------------------
// 5 levels.  Warp per tile; sm: L_LINK_SM words.
// What lat_levels_tile does with a tile's levels: LAT_LEVELS writes them; LAT_PAINT also applies the solvers' paint rules
// to the Squares on the way (the field is in registers: k_paint's second pass over it, 4 B read per square, is saved);
// LAT_PROBE writes nothing and only takes the level of one square (the solvers' finishes: the paint rules come from them).
enum Lat_mode : int { LAT_LEVELS = 0, LAT_PAINT = 1, LAT_PROBE = 2, LAT_PAINT_ONLY = 3 }; // (PAINT_ONLY: the rules applied, the field itself NOT written -- see lat_targets_body)
struct Lat_paint { // resolve_body's rules (post.cuh), all on field 0
    uint32_t n;
    uint32_t below[4], bits[4];
    uint16_t *grid;
};
// rows of squares: row 2i = [pillar, N passage of cell (i, j)], row 2i+1 = [W passage of cell (i, j), cell (i, j)];
// lane j has columns 2j, 2j+1.  A passage lies one below the deeper of its two cells; across the tile's border the
// deeper one is this cell iff the crossing is its component's entry.  FULL: the tile lies inside the maze (no bounds
// to check); the pair (pillar, N passage) of an even row starts on an 8-byte boundary (cols is odd): one 8-byte store.
// UNIFORM (paint modes): no threshold of the rules falls among this tile's levels -- every passable square of the tile gets
// the same bits `ubits` (rules whose threshold lies above all of them), so a square's bits are one test instead of four.
template <int MODE, bool FULL, bool UNIFORM = false>
LAB_DEV void lat_emit_rows(Geom const &g, Lat_tile const &T, uint32_t const (&lv)[LC], uint32_t emW, bool entryN, int lane, int ty, int tx, uint32_t *dist,
                           Lat_paint const &pr, uint32_t &painted, int probe_r, int probe_c, uint32_t &probe_out, uint32_t ubits = 0) {
    constexpr bool PAINTS = MODE == LAT_PAINT || MODE == LAT_PAINT_ONLY, STORES = MODE == LAT_LEVELS || MODE == LAT_PAINT;
    long long const c0 = static_cast<long long>(tx) * TS + 2 * lane, r_base = static_cast<long long>(ty) * TS;
    bool const in0 = FULL || c0 < g.cols, in1 = FULL || c0 + 1 < g.cols;
    long long at = r_base * g.cols + c0; // square (row 2i, column c0)
    bool const pairs = FULL && (reinterpret_cast<uintptr_t>(dist) & 7u) == 0; // (a caller's field may start anywhere)
    auto bits_of = [&](uint32_t v) {
        if (UNIFORM) {
            return v != INF ? ubits : 0u;
        }
        uint32_t b = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            b |= (static_cast<uint32_t>(k) < pr.n && v < pr.below[k]) ? pr.bits[k] : 0u;
        }
        return b;
    };
    // LAT_PAINT: four cell rows at a time -- their levels are stored, then ALL their Squares that get paint are loaded, then
    // all are stored (a load / store pair per square one after the other leaves one DRAM round trip per square on the
    // warp's critical path: the compiler cannot move a load above the store before it)
    constexpr int G = 4;
#pragma unroll
    for (int i0 = 0; i0 < LC; i0 += G) {
        uint32_t pb[3 * G];
        long long const at0 = at;
#pragma unroll
        for (int ii = 0; ii < G; ii++) {
            int const i = i0 + ii;
            uint32_t const me = lv[i];
            uint32_t const left_in = lab_shfl_up(me, 1);
            uint32_t const up = i ? lv[i ? i - 1 : 0] : (entryN ? me : me + 2u);
            uint32_t const left = lane ? left_in : (((emW >> i) & 1u) ? me : me + 2u);
            uint32_t const np = ((T.No >> i) & 1u) ? (me > up ? me : up) - 1u : INF;
            uint32_t const wp = ((T.Wo >> i) & 1u) ? (me > left ? me : left) - 1u : INF;
            if (MODE == LAT_PROBE) {
                if (lane == (probe_c >> 1)) {
                    if (probe_r == 2 * i) probe_out = (probe_c & 1) ? np : INF;
                    if (probe_r == 2 * i + 1) probe_out = (probe_c & 1) ? me : wp;
                }
                continue;
            }
            bool const ok0 = FULL || r_base + 2 * i < g.rows, ok1 = FULL || r_base + 2 * i + 1 < g.rows;
            if (STORES && ok0) {
                if (pairs) {
                    lab_st_pair(dist + at, INF, np);
                } else {
                    if (in0) dist[at] = INF;
                    if (in1) dist[at + 1] = np;
                }
            }
            if (STORES && ok1) {
                long long const at1 = at + g.cols;
                if (in0) dist[at1] = wp;
                if (in1) dist[at1 + 1] = me;
            }
            if (PAINTS) {
                pb[3 * ii] = ok0 && in1 ? bits_of(np) : 0u;
                pb[3 * ii + 1] = ok1 && in0 ? bits_of(wp) : 0u;
                pb[3 * ii + 2] = ok1 && in1 ? bits_of(me) : 0u;
            }
            at += 2 * static_cast<long long>(g.cols);
        }
        if (PAINTS) {
            uint16_t pv[3 * G];
#pragma unroll
            for (int ii = 0; ii < G; ii++) {
                long long const a0 = at0 + 2 * static_cast<long long>(g.cols) * ii, a1 = a0 + g.cols;
                pv[3 * ii] = pb[3 * ii] ? pr.grid[a0 + 1] : static_cast<uint16_t>(0);
                pv[3 * ii + 1] = pb[3 * ii + 1] ? pr.grid[a1] : static_cast<uint16_t>(0);
                pv[3 * ii + 2] = pb[3 * ii + 2] ? pr.grid[a1 + 1] : static_cast<uint16_t>(0);
            }
#pragma unroll
            for (int ii = 0; ii < G; ii++) {
                long long const a0 = at0 + 2 * static_cast<long long>(g.cols) * ii, a1 = a0 + g.cols;
                if (pb[3 * ii]) pr.grid[a0 + 1] = static_cast<uint16_t>(pv[3 * ii] | pb[3 * ii]);
                if (pb[3 * ii + 1]) pr.grid[a1] = static_cast<uint16_t>(pv[3 * ii + 1] | pb[3 * ii + 1]);
                if (pb[3 * ii + 2]) pr.grid[a1 + 1] = static_cast<uint16_t>(pv[3 * ii + 2] | pb[3 * ii + 2]);
                painted += (pb[3 * ii] ? 1u : 0u) + (pb[3 * ii + 1] ? 1u : 0u) + (pb[3 * ii + 2] ? 1u : 0u);
            }
        }
    }
}
template <int MODE>
LAB_DEV void lat_levels_tile(Params const &p, Lattice const &lt, uint32_t t, int lane, uint32_t *sm, uint32_t &mx, Lat_paint const &pr, uint32_t &painted,
                             int probe_r, int probe_c, uint32_t &probe_out) {
    Geom const &g = p.g;
    Lat_tile const T = lat_load(lt, t, lane);
    long long const base = static_cast<long long>(t) * LSLOTS;
    uint32_t em[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        em[s] = lab_ld_cg(lt.emask + static_cast<long long>(t) * 4 + s);
    }
    // shared memory: entry depth per component [1024] | component of every run [1024 x 16 bit] (lat_link_tile's cmap)
    uint32_t *ecomp = sm;
    uint16_t *comp_of = reinterpret_cast<uint16_t *>(sm + 1024);
    {
        uint32_t const *cm = reinterpret_cast<uint32_t const *>(lt.cmap + static_cast<long long>(t) * 1024);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            sm[1024 + k * 32 + lane] = lab_ld_cg(cm + k * 32 + lane);
        }
    }
    // entry depth of every component: minus the weights of the arcs later on the tour than its entry arc
#pragma unroll
    for (int s = 0; s < 4; s++) {
        if ((em[s] >> lane) & 1u) {
            uint32_t const lab = lab_ld_cg(lt.label + base + s * LC + lane);
            ecomp[lab & 1023u] = 0u - lat_prefix_below(lt, lab_ld_cg(lt.R + base + s * LC + lane));
        }
    }
    Lat_roots const roots = lat_roots(p);
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            if (roots.n > q && roots.tile[q] == t) ecomp[lt.tc->lat_root_label[q] & 1023u] = 0u;
        }
    }
    uint32_t const off = roots.passage ? 1u : 0u; // a root on a passage: its two cells are at level 1
    // Depth planes -> one word per cell: a 32 x 32 bit transpose in registers (rows 0-9 the depth)
    uint32_t A[32];
    {
        uint32_t const *pd = lt.dplane + static_cast<long long>(t) * (L_DEPTH_PLANES * 32);
#pragma unroll
        for (int q = 0; q < L_DEPTH_PLANES; q++) {
            A[q] = lab_ld_cg(pd + q * 32 + lane);
        }
#pragma unroll
        for (int q = L_DEPTH_PLANES; q < 32; q++) {
            A[q] = 0;
        }
        uint32_t m = 0x0000FFFFu;
#pragma unroll
        for (int jj = 16; jj != 0; jj >>= 1) {
#pragma unroll
            for (int k = 0; k < 32; k = (k + jj + 1) & ~jj) {
                uint32_t const x = ((A[k] >> jj) ^ A[k + jj]) & m;
                A[k] ^= x << jj;
                A[k + jj] ^= x;
            }
            m ^= m << (jj >> 1);
        }
    }
    lab_syncwarp();
    // lane = cell column j: its 32 cells' levels (INF: no cell).  The cells of a vertical run share their component:
    // its entry depth is looked up where a run begins (bit i of st, as in lat_link_tile)
    uint32_t lv[LC];
    uint32_t const st = T.C & ~(T.No & ~1u);
    uint32_t e = 0, tmin = INF, tmax = 0; // (paint modes: the tile's cells' levels lie in [tmin, tmax])
#pragma unroll
    for (int i = 0; i < LC; i++) {
        if ((st >> i) & 1u) {
            e = ecomp[comp_of[i * 32 + lane] & 1023u];
        }
        uint32_t const v = 2u * (e + (A[i] & 1023u)) + off;
        lv[i] = ((T.C >> i) & 1u) ? v : INF;
        mx = ((T.C >> i) & 1u) && v > mx ? v : mx;
        if (MODE == LAT_PAINT || MODE == LAT_PAINT_ONLY) {
            tmin = ((T.C >> i) & 1u) && v < tmin ? v : tmin;
            tmax = ((T.C >> i) & 1u) && v > tmax ? v : tmax;
        }
    }
    int const ty = static_cast<int>(t / static_cast<uint32_t>(g.tiles_x)), tx = static_cast<int>(t % static_cast<uint32_t>(g.tiles_x));
    uint32_t *dist = p.plane[0].dist;
    bool const entryN = (em[L_N] >> lane) & 1u;
    bool const full = static_cast<long long>(ty + 1) * TS <= g.rows && static_cast<long long>(tx + 1) * TS <= g.cols;
    // paint modes: a passage lies one below the deeper of its two cells, a crossing passage one below or above its cell -- every
    // square of the tile has its level in [tmin - 1, tmax + 1].  A rule whose threshold lies above that paints all of them, one
    // whose threshold lies at or below it none; if every rule is one or the other the tile's paint is uniform (the common case:
    // four thresholds among 10^5-10^6 levels), and a tile nobody paints is not read at all.
    bool uniform = false;
    uint32_t ubits = 0;
    if (MODE == LAT_PAINT || MODE == LAT_PAINT_ONLY) {
        tmin = lab_reduce_min(tmin);
        tmax = lab_reduce_max(tmax);
        uniform = tmin != INF && tmax < INF - 2u;
        for (uint32_t k = 0; k < pr.n; k++) {
            bool const all = tmax + 1u < pr.below[k], none = pr.below[k] == 0u || (tmin > 0u && tmin - 1u >= pr.below[k]);
            uniform = uniform && (all || none);
            ubits |= all ? pr.bits[k] : 0u;
        }
    }
    if (MODE == LAT_PAINT_ONLY && uniform && ubits == 0u) {
        // nothing to write in this tile
    } else if (MODE == LAT_PAINT && uniform && ubits == 0u) {
        if (full) {
            lat_emit_rows<LAT_LEVELS, true>(g, T, lv, em[L_W], entryN, lane, ty, tx, dist, pr, painted, probe_r, probe_c, probe_out);
        } else {
            lat_emit_rows<LAT_LEVELS, false>(g, T, lv, em[L_W], entryN, lane, ty, tx, dist, pr, painted, probe_r, probe_c, probe_out);
        }
    } else if ((MODE == LAT_PAINT || MODE == LAT_PAINT_ONLY) && uniform) {
        if (full) {
            lat_emit_rows<MODE, true, true>(g, T, lv, em[L_W], entryN, lane, ty, tx, dist, pr, painted, probe_r, probe_c, probe_out, ubits);
        } else {
            lat_emit_rows<MODE, false, true>(g, T, lv, em[L_W], entryN, lane, ty, tx, dist, pr, painted, probe_r, probe_c, probe_out, ubits);
        }
    } else if (full) {
        lat_emit_rows<MODE, true>(g, T, lv, em[L_W], entryN, lane, ty, tx, dist, pr, painted, probe_r, probe_c, probe_out);
    } else {
        lat_emit_rows<MODE, false>(g, T, lv, em[L_W], entryN, lane, ty, tx, dist, pr, painted, probe_r, probe_c, probe_out);
    }
    lab_syncwarp(); // (the rows' INF at the root passage was another lane's store)
    if (roots.passage && roots.ptile == t && lane == roots.pj) { // the root passage itself: level 0
        int const lr = 2 * roots.pi + (roots.horizontal ? 1 : 0), lc = 2 * roots.pj + (roots.horizontal ? 0 : 1);
        long long const sq = (static_cast<long long>(ty) * TS + lr) * g.cols + static_cast<long long>(tx) * TS + lc;
        if (MODE == LAT_PROBE) {
            if (probe_r == lr && probe_c == lc) probe_out = 0u;
        } else {
            if (MODE != LAT_PAINT_ONLY) dist[sq] = 0u;
            if (MODE == LAT_PAINT || MODE == LAT_PAINT_ONLY) {
                uint32_t b = 0;
                for (uint32_t k = 0; k < pr.n; k++) {
                    b |= 0u < pr.below[k] ? pr.bits[k] : 0u;
                }
                if (b) {
                    pr.grid[sq] = static_cast<uint16_t>(pr.grid[sq] | b);
                    painted++;
                }
            }
        }
    }
    lab_syncwarp();
}
template <int MODE> LAB_DEV void lat_levels_body(Params const &p, Lattice const &lt, int lane, uint32_t *sm, Result *res, uint16_t *grid) {
    Tree_ctl *tc = lt.tc;
    if (!lat_live(tc)) {
        return;
    }
    Lat_paint pr{};
    if (MODE == LAT_PAINT || MODE == LAT_PAINT_ONLY) {
        pr.n = res->n_rules;
        for (int k = 0; k < 4; k++) {
            pr.below[k] = res->rule[k].below;
            pr.bits[k] = res->rule[k].bits;
        }
        pr.grid = grid;
    }
    uint32_t mx = 0, painted = 0, unused = 0;
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) {
            t = lab_atomic_add(&tc->lat_ticket[3], 1u);
        }
        t = lab_shfl(t, 0);
        if (t >= static_cast<uint32_t>(p.g.ntiles)) {
            break;
        }
        lat_levels_tile<MODE>(p, lt, t, lane, sm, mx, pr, painted, 0, 0, unused);
    }
    mx = lab_reduce_max(mx);
    if (lane == 0 && mx) {
        lab_atomic_max(&p.ctl->max_level, mx);
    }
    if (MODE == LAT_PAINT || MODE == LAT_PAINT_ONLY) {
        painted = lab_reduce_add(painted);
        if (lane == 0 && painted) {
            lab_atomic_add(reinterpret_cast<uint64_t *>(&res->settled), static_cast<uint64_t>(painted));
        }
    }
}
// The solvers' finishes: warp j takes finish j's level out of its tile (nothing is written to the field), then one
// thread runs resolve_body on them: the paint rules are known BEFORE the levels are written, so lat_levels_body<LAT_PAINT>
// can apply them on the way (Control::painted tells paint_body that it has been done).
LAB_DEV void lat_targets_body(Params const &p, Lattice const &lt, int warp, int lane, uint32_t *sm) {
    Control *ctl = p.ctl;
    if (!lat_live(lt.tc) || static_cast<uint32_t>(warp) >= ctl->n_targets) {
        return;
    }
    unsigned long long const cell = ctl->target_cell[warp];
    int const r = static_cast<int>(cell / static_cast<unsigned long long>(p.g.cols)), c = static_cast<int>(cell % static_cast<unsigned long long>(p.g.cols));
    uint32_t const t = static_cast<uint32_t>((r / TS) * p.g.tiles_x + c / TS);
    uint32_t mx = 0, painted = 0, out = INF;
    Lat_paint const none{};
    lat_levels_tile<LAT_PROBE>(p, lt, t, lane, sm, mx, none, painted, r % TS, c % TS, out);
    if (lane == ((c % TS) >> 1)) {
        ctl->target_val[warp] = out;
    }
}

// gather without paths, the field not wanted (LAT_PAINT_ONLY): the one thing besides the paint that reads levels is
// path.front() of every colour (bfs_threads.cc:355-364) -- the first of the finish's N, E, S, W neighbours one level lower.
// Warp k probes the four neighbours of the finish colour k claims (after resolve_body); front_rc[2k], [2k+1] as k_walk leaves them.
LAB_DEV void lat_fronts_body(Params const &p, Lattice const &lt, Result const *res, int32_t const *finish_rc, int32_t *front_rc, int warp, int lane, uint32_t *sm) {
    if (!lat_live(lt.tc) || warp >= 4) {
        return;
    }
    int const fj = res->finish_of[warp];
    uint32_t const D = res->game_level[warp];
    if (fj < 0 || D == INF || D == 0u) {
        return; // (front_rc stays -1)
    }
    int const fr = finish_rc[2 * fj], fc = finish_rc[2 * fj + 1];
    int const dr[4] = {-1, 0, 1, 0}, dc[4] = {0, 1, 0, -1};
    bool found = false;
    for (int k = 0; k < 4; k++) {
        int const r = fr + dr[k], c = fc + dc[k];
        if (found || r < 0 || r >= p.g.rows || c < 0 || c >= p.g.cols) {
            continue; // (uniform)
        }
        uint32_t const t = static_cast<uint32_t>((r / TS) * p.g.tiles_x + c / TS);
        uint32_t mx = 0, painted = 0, out = INF;
        Lat_paint const none{};
        lat_levels_tile<LAT_PROBE>(p, lt, t, lane, sm, mx, none, painted, r % TS, c % TS, out);
        out = lab_shfl(out, (c % TS) >> 1);
        if (out == D - 1u) {
            found = true;
            if (lane == 0) {
                front_rc[2 * warp] = r;
                front_rc[2 * warp + 1] = c;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// hunt: the winner's path WITHOUT the walk.  The canonical path (from the finish to the first neighbour one level lower,
// post.cuh walk_body) is one dependent chain, 10^5-10^6 squares long on the reference's mazes at the headline sizes -- tens
// of milliseconds for one warp, ten times the search.  On a tree it is the finish's ancestors, and the tour's ranks name
// them: a component of a tile lies on the path iff the subtree behind one of its crossings holds the finish's component,
// i.e. iff the rank R_F of that component's entry arc lies between the ranks of the crossing's two arcs (down: the
// twin's, earlier on the tour, larger R; up: this slot's).  So every tile finds its own SEGMENTS of the path (first
// squares: the crossing passage if it lies in the tile -- N and W sides -- else the cell at the crossing), and every
// segment is walked on its own, inside its tile (walk_body, confine): a few hundred steps each, all at once.  Squares are
// written at their index along the path (finish level - 1 - own level).
// Row bands: the ranks are the whole maze's, so every band finds and walks the segments in its own rows with no hand-over
// at all; the band that holds the finish tells the others R_F and the finish's level (lat_path_rank_body), every band
// then runs lat_path_setup_body + find + walk (lab_band_lat_path_*).
// rank of the entry arc of the component that holds square (fr, fc) of this band (0: a root's component).  One warp.
LAB_DEV uint32_t lat_path_rank_body(Params const &p, Lattice const &lt, int fr, int fc, int lane) {
    Geom const &g = p.g;
    // the component of the cell that owns the square (a passage is its cell's N or W passage)
    uint32_t const t = static_cast<uint32_t>((fr / TS) * g.tiles_x + fc / TS);
    int const i = (fr % TS) >> 1, j = (fc % TS) >> 1;
    uint32_t const *m = lt.mask + static_cast<long long>(t) * 96;
    uint32_t const st = lab_ld_cg(m + j) & ~(lab_ld_cg(m + 64 + j) & ~1u);
    uint32_t const run = static_cast<uint32_t>(31 - lab_clz32(st & (0xFFFFFFFFu >> (31 - i))));
    uint32_t const comp = lab_ld_cg(lt.cmap + static_cast<long long>(t) * 1024 + run * 32u + static_cast<uint32_t>(j));
    uint32_t rank = 0;
#pragma unroll
    for (int s = 0; s < 4; s++) {
        long long const a = static_cast<long long>(t) * LSLOTS + s * LC + lane;
        if (((lab_ld_cg(lt.emask + static_cast<long long>(t) * 4 + s) >> lane) & 1u) && lab_ld_cg(lt.label + a) == comp) {
            rank = lab_ld_cg(lt.R + a);
        }
    }
    return lab_reduce_max(rank);
}
// One thread.  (fr, fc) / (sr, sc): the finish / the start in this grid's rows, fr < 0 / sr < 0: in another band.
// q: which path (hunt: 0; gather: colour q's); seg: path q's region of the segment list; out: path q's buffer
LAB_DEV void lat_path_setup_body(Params const &p, Lattice const &lt, int q, uint32_t rank, uint32_t D, int fr, int fc, int sr, int sc, uint32_t repaint,
                                 uint16_t *grid, int32_t *out, unsigned long long max_out, uint32_t *seg) {
    Tree_ctl *tc = lt.tc;
    Geom const &g = p.g;
    uint32_t const *dist = p.plane[0].dist;
    uint32_t nseg = 0;
    // the first square of the path, if it lies in the finish's tile: that tile's own segment begins there
    int const dr[4] = {-1, 0, 1, 0}, dc[4] = {0, 1, 0, -1};
    for (int k = 0; k < 4 && D > 0 && fr >= 0; k++) {
        int const r = fr + dr[k], c = fc + dc[k];
        if (r >= 0 && r < g.rows && c >= 0 && c < g.cols && dist[static_cast<long long>(r) * g.cols + c] == D - 1u) {
            if (r / TS == fr / TS && c / TS == fc / TS) {
                seg[nseg++] = static_cast<uint32_t>(r) * static_cast<uint32_t>(g.cols) + static_cast<uint32_t>(c);
            }
            break;
        }
    }
    // the last square is the start (a start on a passage square is in no component)
    if (D > 0 && sr >= 0) {
        long long const at = static_cast<long long>(sr) * g.cols + sc;
        if (repaint) {
            grid[at] = static_cast<uint16_t>((grid[at] & ~SQ_PAINT_MASK) | repaint);
        }
        if (out && D - 1u < max_out) {
            out[2ull * (D - 1u)] = sr;
            out[2ull * (D - 1u) + 1] = sc;
        }
    }
    tc->lat_path_rank[q] = rank;
    tc->lat_path_level[q] = D;
    tc->lat_path_nseg[q] = nseg;
    lab_atomic_or(&tc->lat_path_on, 1u << q);
}
// lat_path_begin_body: one GPU; warp q of four (after lat_path_on and the counts were cleared).  hunt: warp 0 alone, the winner's
// path, repainted in its colour; gather (paths wanted): warp q = colour q's path to the finish it claims, nothing repainted.
// seg: room for the segments' first squares (flat indices), a quarter per path; out: the paths' buffers, out_stride pairs apart.
LAB_DEV void lat_path_begin_body(Params const &p, Lattice const &lt, int game, Result *res, int32_t const *finish_rc, uint16_t *grid, int32_t *out,
                                 unsigned long long out_stride, unsigned long long max_out, uint32_t *seg, uint32_t seg_cap, int q, int lane) {
    Tree_ctl *tc = lt.tc;
    Control *ctl = p.ctl;
    Geom const &g = p.g;
    if (!lat_live(tc) || !ctl->painted || ctl->fronts_done) { // (fronts_done: the field was not written)
        return;
    }
    int fj = 0;
    if (game == GAME_HUNT) {
        if (q != 0 || res->winner < 0) return;
    } else {
        fj = res->finish_of[q];
        if (fj < 0) { // (as k_walk leaves an unclaimed colour)
            if (lane == 0) {
                res->path_len[q] = 0;
                if (out) out[static_cast<unsigned long long>(q) * out_stride * 2] = -1;
            }
            return;
        }
    }
    int const fr = finish_rc[2 * fj], fc = finish_rc[2 * fj + 1];
    uint32_t const D = p.plane[0].dist[static_cast<long long>(fr) * g.cols + fc];
    if (D == INF || ((fr | fc) & 1) == 0) {
        if (lane == 0) ctl->walk_fallback = 1u; // (the one-warp walk says what there is to say)
        return;
    }
    uint32_t const rank = lat_path_rank_body(p, lt, fr, fc, lane);
    if (lane != 0) {
        return;
    }
    uint32_t const t0 = ctl->src_tile[0];
    int const sr = static_cast<int>(t0 / static_cast<uint32_t>(g.tiles_x)) * TS + static_cast<int>(ctl->src_rc[0] >> 8);
    int const sc = static_cast<int>(t0 % static_cast<uint32_t>(g.tiles_x)) * TS + static_cast<int>(ctl->src_rc[0] & 0xFFu);
    uint32_t const repaint = game == GAME_HUNT ? (1u << res->winner) << SQ_PAINT_SHIFT : 0u;
    lat_path_setup_body(p, lt, q, rank, D, fr, fc, sr, sc, repaint, grid, out ? out + static_cast<unsigned long long>(q) * out_stride * 2 : nullptr, max_out,
                        seg + static_cast<size_t>(q) * (seg_cap / 4));
    res->walk_r[q] = D > 0 ? sr : fr;
    res->walk_c[q] = D > 0 ? sc : fc;
    res->path_len[q] = D;
    res->walk_left[q] = 0;
}
// warp per tile: the crossings the path comes up through
LAB_DEV void lat_path_find_body(Params const &p, Lattice const &lt, uint32_t *seg, uint32_t seg_cap, long long gwarp, long long nwarps, int lane) {
    Tree_ctl *tc = lt.tc;
    Geom const &g = p.g;
    uint32_t const on = tc->lat_path_on;
    uint32_t RF[4];
    bool any = false;
    for (int q = 0; q < 4; q++) {
        RF[q] = ((on >> q) & 1u) ? tc->lat_path_rank[q] : 0u; // (0: no such path, or its finish lies in a root's component: nobody's descendant)
        any = any || RF[q] != 0u;
    }
    if (!any) {
        return;
    }
    for (long long t = gwarp; t < g.ntiles; t += nwarps) {
        long long const base = t * LSLOTS;
        uint32_t const ty = static_cast<uint32_t>(t / g.tiles_x), tx = static_cast<uint32_t>(t % g.tiles_x);
#pragma unroll
        for (int s = 0; s < 4; s++) {
            uint32_t const x = lab_ld_cg(lt.xmask + t * 4 + s), e = lab_ld_cg(lt.emask + t * 4 + s);
            if (((x & ~e) >> lane) & 1u) {
                uint32_t const up = lab_ld_cg(lt.R + base + s * LC + lane);
                uint32_t const tw = lat_twin(g, static_cast<uint32_t>(t) + lt.tile_base, static_cast<uint32_t>(s * LC + lane)); // (global arc id)
                uint32_t const twl = tw - lt.tile_base * LSLOTS;
                // (a twin in another band is a border arc of its super-tile: its hops are in the whole-maze border list)
                uint32_t const down = twl < lt.nslots ? lab_ld_cg(lt.R + twl) : static_cast<uint32_t>(lab_ld_cg(lt.gb + lat_dense_border(lt, g, tw)) >> 32);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (RF[q] != 0u && up < RF[q] && RF[q] <= down) {
                        uint32_t const r = ty * TS + (s == L_N ? 0u : (s == L_S ? TS - 1u : 2u * lane + 1u));
                        uint32_t const c = tx * TS + (s == L_W ? 0u : (s == L_E ? TS - 1u : 2u * lane + 1u));
                        uint32_t const at = lab_atomic_add(&tc->lat_path_nseg[q], 1u);
                        if (at < seg_cap / 4) {
                            seg[static_cast<size_t>(q) * (seg_cap / 4) + at] = r * static_cast<uint32_t>(g.cols) + c;
                        }
                    }
                }
            }
        }
    }
}
// warp per segment (wid of nw warps; sm: WALK_SM_U32 words)
LAB_DEV void lat_path_walk_body(Params const &p, Lattice const &lt, Result *res, uint32_t repaint, uint16_t *grid, int32_t *out, unsigned long long out_stride,
                                unsigned long long max_out, uint32_t const *seg, uint32_t seg_cap, long long wid, long long nw, int lane, uint32_t *sm) {
    Tree_ctl const *tc = lt.tc;
    uint32_t const on = tc->lat_path_on;
    for (int q = 0; q < 4; q++) {
        if (!((on >> q) & 1u)) {
            continue;
        }
        uint32_t const n = tc->lat_path_nseg[q] < seg_cap / 4 ? tc->lat_path_nseg[q] : seg_cap / 4, D = tc->lat_path_level[q];
        uint32_t const *list = seg + static_cast<size_t>(q) * (seg_cap / 4);
        int32_t *out_q = out ? out + static_cast<unsigned long long>(q) * out_stride * 2 : nullptr;
        for (long long i = wid; i < n; i += nw) {
            uint32_t const sq = lab_ld_cg(list + i);
            int const r = static_cast<int>(sq / static_cast<uint32_t>(p.g.cols)), c = static_cast<int>(sq % static_cast<uint32_t>(p.g.cols));
            uint32_t const d = p.plane[0].dist[sq];
            walk_body(p, 0, r, c, grid, out_q, max_out, repaint, 0, res, 0, lane, sm, 1, D - 1u - d, 1);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Row bands: the glue between the phases.  A band runs count + link + local ranking on its own tiles,
// the ranks put their border-list slices and one header each together (lat_band_header_body -> all-gather, or peer
// stores), every rank checks the WHOLE maze's counts (lat_band_check_body), finishes the ranking over the whole
// border list, floods its tiles and scatters its arcs' weights; the weights are joined (all-reduce or peer stores,
// with each band's count of reached squares in the tail), and every rank scans them and writes its band's levels.
// Two exchanges per map, whatever its diameter.  All single-thread bodies.
constexpr uint32_t LAT_HDR = 8;   // words of a band's header
constexpr uint32_t LAT_TAIL = 8;  // words after wsum[nslots_global + 2): reached squares (lo, hi), error flags
LAB_DEV void lat_band_open_body(Lattice const &lt) { // before the link: whether the whole maze has a tree's counts is not known yet
    lt.tc->ok = 1u;
    lt.tc->lat_on = 1u;
}
LAB_DEV void lat_band_header_body(Lattice const &lt, uint32_t *hdr) {
    Tree_ctl const *tc = lt.tc;
    hdr[0] = static_cast<uint32_t>(tc->V);
    hdr[1] = static_cast<uint32_t>(tc->V >> 32);
    hdr[2] = static_cast<uint32_t>(tc->E);
    hdr[3] = static_cast<uint32_t>(tc->E >> 32);
    hdr[4] = tc->n_arcs;
    hdr[5] = tc->lat_err;
    hdr[6] = hdr[7] = 0u;
}
LAB_DEV void lat_band_check_body(Lattice const &lt, uint32_t const *all_hdr, int world) {
    Tree_ctl *tc = lt.tc;
    unsigned long long V = 0, E = 0, arcs = 0;
    uint32_t err = 0;
    for (int r = 0; r < world; r++) {
        uint32_t const *h = all_hdr + static_cast<size_t>(r) * LAT_HDR;
        V += static_cast<unsigned long long>(h[0]) | (static_cast<unsigned long long>(h[1]) << 32);
        E += static_cast<unsigned long long>(h[2]) | (static_cast<unsigned long long>(h[3]) << 32);
        arcs += h[4];
        err = err ? err : h[5];
    }
    tc->V = V;
    tc->E = E;
    tc->n_arcs = static_cast<uint32_t>(arcs);
    tc->ok = (V >= 2 && E + 1 == V && arcs <= lt.nslots_global) ? 1u : 0u; // every rank reads the same headers: they all agree
    if (err && !tc->lat_err) {
        tc->lat_err = err;
    }
}
// after the flood: this band's reached squares and flags go where the all-reduce (or the peers' stores) join them
LAB_DEV void lat_band_tail_body(Lattice const &lt) {
    Tree_ctl const *tc = lt.tc;
    uint32_t const at = lt.nslots_global + 2u;
    for (int q = 0; q < lt.peers; q++) {
        uint32_t *w = lt.wsum_peer[q] + at;
        if (lt.peers == 1) { // (to be summed)
            w[0] = static_cast<uint32_t>(tc->lat_settled); // (a maze has fewer than 2^32 squares: make_geom)
            w[1] = 0u;
            w[2] = (tc->lat_flood_bad || tc->lat_err) ? 1u : 0u;
        } else { // (peer stores: everybody adds into everybody's tail)
            lab_atomic_add(w, static_cast<uint32_t>(tc->lat_settled));
            lab_atomic_add(w + 2, (tc->lat_flood_bad || tc->lat_err) ? 1u : 0u);
        }
    }
}
// before the gate: this control block belongs to the lattice road alone -- if it gave up, no pipeline made the field
LAB_DEV void lat_band_close_body(Lattice const &lt) {
    Tree_ctl *tc = lt.tc;
    if (!(tc->ok != 0 && lat_made(tc))) {
        tc->ok = 0u;
    }
}
LAB_DEV void lat_band_settled_body(Lattice const &lt) {
    Tree_ctl *tc = lt.tc;
    uint32_t const *w = lt.wsum + lt.nslots_global + 2u;
    tc->lat_settled_band = tc->lat_settled;
    tc->lat_settled = w[0];
    tc->lat_flood_bad = w[2];
    if (w[2] && !tc->lat_err) {
        tc->lat_err = 4u;
    }
}

} // namespace lab
