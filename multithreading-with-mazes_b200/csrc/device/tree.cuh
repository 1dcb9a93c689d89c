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
//   4 local   the bit-sliced tile engine (engine.cuh step_group) floods every tile from its entry
//             crossings, all tiles at once, no tile waiting for another: local levels of the border
//             squares.
//   5 prefix  the components form a tree themselves (parent = the component on the other side of the
//             entry crossing, edge weight = local level of that crossing + 1); pointer jumping sums
//             the weights down from the root: the exact level of every entry crossing.
//   6 final   the tile engine floods every tile once more from its entry crossings, now at their exact
//             levels, and writes the dense level field.  Again all tiles are independent.
//
// Whether the maze IS a tree is checked, not assumed: passable squares V and adjacent passable pairs
// E are counted (a tree has E == V-1), every crossing must lie on the root's tour and the local
// floods must reach all V squares (=> connected).  Anything else (arena, grid, -m cross/x, squares
// masked by measure bits) takes the tile wave of engine.cuh instead; both give identical levels.
#pragma once
#include "engine.cuh"

namespace lab {

constexpr uint32_t SLOTS = 4 * TS; // crossing slots of a tile: [side][pos] = the arc ENTERING the tile there

struct Tree_ctl {
    unsigned long long V, E; // passable squares / adjacent passable pairs
    unsigned long long local_settled;
    uint32_t ok;         // cleared as soon as the maze turns out not to be a tree
    uint32_t unreached;  // crossings that are not on the root's tour
    uint32_t bad_weight; // an entry crossing whose far side got no local level
    uint32_t walk_error; // a walker ran longer than any tile allows
    uint32_t ticket[3];  // dynamic tile counters: walk, local, final
    uint32_t rank_buf, prefix_buf; // which ping-pong buffer holds the finished ranking / prefix
    uint32_t rank_more[40], prefix_more[40]; // round r left work for round r+1
    uint32_t n_arcs;     // crossings in total (diagnostics)
    uint32_t used;       // the tree pipeline produced the field (set by gate)
};

struct Tree {
    uint32_t *succ0;     // [nslots+2] successor segment, as walked (kept: it also chains a component's crossings)
    uint32_t *s[2];      // [nslots+2] ping-pong: successor / parent entry
    uint32_t *r[2];      // [nslots+2] ping-pong: segments to the end of the tour / level of the entry
    uint64_t *inmask;    // [ntiles][4]  crossings through which a tile's components are ENTERED
    uint32_t *loc;       // [nslots]     local level of border squares ([side][pos], like edges)
    Tree_ctl *tc;
    uint32_t nslots;     // ntiles * SLOTS; ROOT = nslots, END = nslots + 1
};

LAB_DEV bool tree_live(Tree_ctl const *tc) { return tc->ok != 0 && tc->unreached == 0 && tc->bad_weight == 0 && tc->walk_error == 0; }

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
    for (long long t = gwarp; t < p.g.ntiles; t += nwarps) {
        uint64_t const P0 = lab_ld_ro(p.path + t * TS + 2 * lane), P1 = lab_ld_ro(p.path + t * TS + 2 * lane + 1);
        uint64_t below = lab_shfl_down64(P0, 1); // row 2*lane+2
        if (lane == 31) {
            below = 0;
        }
        v += static_cast<unsigned>(lab_popc64(P0) + lab_popc64(P1));
        e += static_cast<unsigned>(lab_popc64(P0 & (P0 >> 1)) + lab_popc64(P1 & (P1 >> 1)) + lab_popc64(P0 & P1) + lab_popc64(P1 & below));
        if (lane < 2) { // crossings counted once: this tile's S and E borders
            e += static_cast<unsigned>(lab_popc64(lab_ld_ro(p.cross + t * 4 + (lane == 0 ? SIDE_S : SIDE_E))));
        }
    }
    uint32_t const vs = lab_reduce_add(static_cast<uint32_t>(v)), es = lab_reduce_add(static_cast<uint32_t>(e));
    uint32_t const vh = lab_reduce_add(static_cast<uint32_t>(v >> 32)), eh = lab_reduce_add(static_cast<uint32_t>(e >> 32));
    if (lane == 0) {
        lab_atomic_add(reinterpret_cast<uint64_t *>(&tr.tc->V), (static_cast<uint64_t>(vh) << 32) + vs);
        lab_atomic_add(reinterpret_cast<uint64_t *>(&tr.tc->E), (static_cast<uint64_t>(eh) << 32) + es);
    }
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
// shared memory per warp (uint64 units): open[64][4] | res16[1032] | list16[1032] (later res32[256]) | masks[16] | counter
constexpr int WALK_IDS = 1025;
constexpr int WALK_U64 = 4 * TS + 260 + 260 + 16 + 2;
constexpr uint32_t WALK_END = 0xFFFFu;

LAB_DEV void walk_tile(Params const &p, Tree const &tr, uint32_t t, int lane, uint64_t *sm) {
    Geom const &g = p.g;
    uint64_t *open = sm;
    uint16_t *res16 = reinterpret_cast<uint16_t *>(sm + 4 * TS);
    uint16_t *list = reinterpret_cast<uint16_t *>(sm + 4 * TS + 260);
    uint32_t *res32 = reinterpret_cast<uint32_t *>(sm + 4 * TS + 260); // after phase A the list is dead
    uint64_t *masks = sm + 4 * TS + 520;
    uint32_t *cnt = reinterpret_cast<uint32_t *>(sm + 4 * TS + 536);
    uint32_t const ROOT = tr.nslots, END = tr.nslots + 1u;
    int const r0 = 2 * lane, r1 = 2 * lane + 1;

    uint64_t const P0 = lab_ld_ro(p.path + static_cast<long long>(t) * TS + r0), P1 = lab_ld_ro(p.path + static_cast<long long>(t) * TS + r1);
    uint64_t X[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        X[s] = lab_ld_ro(p.cross + static_cast<long long>(t) * 4 + s);
    }
    uint64_t const up = lab_shfl_up64(P1, 1), down = lab_shfl_down64(P0, 1);
    open[r0 * 4 + 0] = lane == 0 ? X[SIDE_N] : (P0 & up);
    open[r1 * 4 + 0] = P1 & P0;
    open[r0 * 4 + 2] = P0 & P1;
    open[r1 * 4 + 2] = lane == 31 ? X[SIDE_S] : (P1 & down);
    open[r0 * 4 + 1] = (P0 & (P0 >> 1)) | (((X[SIDE_E] >> r0) & 1ull) << 63);
    open[r1 * 4 + 1] = (P1 & (P1 >> 1)) | (((X[SIDE_E] >> r1) & 1ull) << 63);
    open[r0 * 4 + 3] = (P0 & (P0 << 1)) | ((X[SIDE_W] >> r0) & 1ull);
    open[r1 * 4 + 3] = (P1 & (P1 << 1)) | ((X[SIDE_W] >> r1) & 1ull);
    lab_syncwarp();

    // which walkers exist: 16 masks of 64 positions (4 tile sides, then heading-major inner borders)
    if (lane < 4) {
        masks[lane] = X[lane];
    }
#pragma unroll
    for (int b = 1; b <= 3; b++) {
        // vertical inner border b: squares (pos, 16b-1) and (pos, 16b) both passable
        uint64_t const v = lab_ballot64((open[lane * 4 + 1] >> (16 * b - 1)) & 1ull, (open[(lane + 32) * 4 + 1] >> (16 * b - 1)) & 1ull);
        if (lane == 0) {
            uint64_t const hmask = open[(16 * b - 1) * 4 + 2]; // horizontal inner border b
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
    uint32_t const total = lab_shfl(incl, 31);
    {
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
        for (int d = 3; d >= 0; d--) {
            if ((open[rr * 4 + d] >> rc) & 1ull) {
                rh = (d + 2) & 3;
                root_open = true;
            }
        }
    }
    // ---- phase A: every walker to the first block border ----------------------------------------
    // One flat loop, one step per iteration: a lane whose walker has arrived fetches its next one in
    // the same iteration, so the lanes of the warp stay converged whatever the lengths of their pieces.
    uint32_t const nitems = total + (root_tile ? 1u : 0u);
    bool have = false, skip = false;
    int r = 0, c = 0, h = 0;
    uint32_t id = 0, steps = 0;
    for (;;) {
        if (!have) {
            uint32_t const idx = lab_atomic_add(cnt, 1u);
            if (idx >= nitems) {
                break;
            }
            have = true;
            steps = 0;
            skip = idx == total; // the root's own walker leaves the root state instead of ending there
            if (skip) {
                id = WALK_IDS - 1;
                r = rr;
                c = rc;
                h = rh;
            } else {
                id = list[idx];
                int const pos = static_cast<int>(id & 63u);
                if (id < SLOTS) {
                    int const side = static_cast<int>(id >> 6);
                    r = side == SIDE_N ? 0 : (side == SIDE_S ? TS - 1 : pos);
                    c = side == SIDE_W ? 0 : (side == SIDE_E ? TS - 1 : pos);
                    h = side == SIDE_N ? 2 : (side == SIDE_S ? 0 : (side == SIDE_W ? 1 : 3));
                } else {
                    uint32_t const k = (id - SLOTS) >> 6; // h*3 + (b-1)
                    h = static_cast<int>(k / 3u);
                    int const line = 16 * (static_cast<int>(k % 3u) + 1);
                    r = h == 2 ? line : (h == 0 ? line - 1 : pos);
                    c = h == 1 ? line : (h == 3 ? line - 1 : pos);
                }
            }
        }
        uint32_t result = WALK_END;
        bool arrived = root_tile && !skip && r == rr && c == rc && h == rh && root_open; // the tour ends where it began
        skip = false;
        uint64_t const *o = open + r * 4;
        uint32_t const nib = static_cast<uint32_t>((o[0] >> c) & 1ull) | (static_cast<uint32_t>((o[1] >> c) & 1ull) << 1)
                             | (static_cast<uint32_t>((o[2] >> c) & 1ull) << 2) | (static_cast<uint32_t>((o[3] >> c) & 1ull) << 3);
        arrived |= nib == 0; // a square on its own (only the root can be one)
        if (!arrived) {
            uint32_t const rot = ((nib >> h) | (nib << (4 - h))) & 15u; // bit j: direction h + j
            int const j = (rot & 2u) ? 1 : ((rot & 1u) ? 0 : ((rot & 8u) ? 3 : 2));
            h = (h + j) & 3;
            int const nr = r + (h == 2) - (h == 0), nc = c + (h == 1) - (h == 3);
            if ((nr | nc) & ~(TS - 1)) { // left the tile: into the neighbour's slot
                uint32_t const ns = h == 0 ? SIDE_S * TS + static_cast<uint32_t>(c)
                                           : (h == 2 ? SIDE_N * TS + static_cast<uint32_t>(c)
                                                     : (h == 1 ? SIDE_W * TS + static_cast<uint32_t>(r) : SIDE_E * TS + static_cast<uint32_t>(r)));
                result = 1024u + ns;
                arrived = true;
            } else if (((r ^ nr) | (c ^ nc)) & 0x30) { // crossed a block border: the next piece starts here
                int const line = (h == 1 ? nc : (h == 3 ? c : (h == 2 ? nr : r))) >> 4; // 1..3
                result = SLOTS + static_cast<uint32_t>(h * 3 + line - 1) * 64u + static_cast<uint32_t>((h & 1) ? r : c);
                arrived = true;
            } else if (++steps > 4u * 16u * 16u + 8u) {
                lab_atomic_max(&tr.tc->walk_error, 1u);
                arrived = true;
            }
            r = nr;
            c = nc;
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
                uint32_t const nt = side == SIDE_S ? t - static_cast<uint32_t>(g.tiles_x)
                                                   : (side == SIDE_N ? t + static_cast<uint32_t>(g.tiles_x) : (side == SIDE_W ? t + 1u : t - 1u));
                out = nt * SLOTS + ns;
            }
        }
        if (k < 8) {
            res32[s] = out; // (the list it overwrites was last read before the barrier above)
        } else if (root_tile && lane == 0) {
            tr.succ0[ROOT] = out;
            tr.s[0][ROOT] = out;
            tr.r[0][ROOT] = 1u;
        }
    }
    lab_syncwarp();
    long long const o = static_cast<long long>(t) * SLOTS;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int const s = lane + 32 * k;
        uint32_t const v = res32[s];
        bool const active = (X[s >> 6] >> (s & 63)) & 1ull;
        tr.succ0[o + s] = v;
        tr.s[0][o + s] = v;
        tr.r[0][o + s] = active ? 1u : 0u;
    }
    if (lane == 0 && total) {
        lab_atomic_add(&tr.tc->n_arcs, static_cast<uint32_t>(lab_popc64(X[0]) + lab_popc64(X[1]) + lab_popc64(X[2]) + lab_popc64(X[3])));
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

// One thread, before the walk: the two list ends.
LAB_DEV void tree_init_body(Tree const &tr) {
    uint32_t const ROOT = tr.nslots, END = tr.nslots + 1u;
    tr.succ0[ROOT] = END;
    tr.succ0[END] = END;
    for (int b = 0; b < 2; b++) {
        tr.s[b][ROOT] = END;
        tr.s[b][END] = END;
        tr.r[b][ROOT] = 1u;
        tr.r[b][END] = 0u;
    }
}

// ---------------------------------------------------------------------------------------------
// 2 rank.  One round of pointer jumping: r = segments from here to the end of the tour.
// Round `round` reads buffer round&1 and writes the other one; a round that finds the previous one
// left nothing to do returns at once (the launches are fixed in number, the work is not).
LAB_DEV void tree_rank_round_body(Tree const &tr, uint32_t round, long long gthread, long long nthreads) {
    Tree_ctl *tc = tr.tc;
    if (!tree_live(tc) || (round > 0 && tc->rank_more[round - 1] == 0)) {
        return;
    }
    uint32_t const END = tr.nslots + 1u;
    bool const odd = round & 1u;
    uint32_t const *sin = odd ? tr.s[1] : tr.s[0], *rin = odd ? tr.r[1] : tr.r[0];
    uint32_t *sout = odd ? tr.s[0] : tr.s[1], *rout = odd ? tr.r[0] : tr.r[1];
    long long const n = static_cast<long long>(tr.nslots) + 2;
    bool more = false;
    for (long long i = gthread; i < n; i += nthreads) {
        uint32_t const s = sin[i];
        uint32_t const r = rin[i];
        if (s == END) {
            sout[i] = END;
            rout[i] = r;
        } else {
            uint32_t const s2 = sin[s];
            sout[i] = s2;
            rout[i] = r + rin[s];
            more |= s2 != END;
        }
    }
    // (thousands of warps storing to -- or reading past L1 from -- one address would serialise in L2:
    // look before writing, through L1; a stale zero only costs a redundant store)
    if (lab_any(more) && (gthread & 31) == 0 && lab_ld_ro(&tc->rank_more[round]) == 0u) {
        lab_st_volatile(&tc->rank_more[round], 1u);
    }
    if (gthread == 0) {
        tc->rank_buf = (round & 1u) ^ 1u;
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
    uint32_t const *sfin = tc->rank_buf ? tr.s[1] : tr.s[0], *rfin = tc->rank_buf ? tr.r[1] : tr.r[0];
    uint32_t bad = 0;
    for (long long t = gwarp; t < p.g.ntiles; t += nwarps) {
        long long const o = t * SLOTS;
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
                    in[half] = rfin[a] > rfin[tw];
                    bad += sfin[a] != END;
                }
                tr.loc[a] = INF;
            }
            masks[side] = lab_ballot64(in[0], in[1]);
        }
        if (lane < 4) {
            tr.inmask[t * 4 + lane] = masks[lane];
        }
    }
    if (lab_any(bad != 0) && lane == 0) {
        lab_atomic_add(&tc->unreached, 1u);
    }
}

// ---------------------------------------------------------------------------------------------
// emit_borders with the facing levels "never worth waking"
LAB_DEV uint32_t tree_emit_borders(Tile_ctx const &tc, uint64_t *stage, Wave const &w, uint64_t S0, uint64_t S1, uint32_t base,
                                   int lane, Emit_acc &acc, bool edges) {
    if (edges) {
        return emit_borders<true>(tc, stage, w, S0, S1, 0ull, 0ull, base, lane, 0u, 0u, 0u, 0u, 0u, 0u, acc);
    }
    return emit_borders<false>(tc, stage, w, S0, S1, 0ull, 0ull, base, lane, 0u, 0u, 0u, 0u, 0u, 0u, acc);
}

// 4 local.  Every tile flooded from its entry crossings (and the root) at local level 0; the local
// levels of its border squares go to tr.loc.  stage: 7*64 uint64 per warp.
constexpr int LOCAL_STAGE_U64 = 7 * TS;

LAB_DEV void local_fill_tile(Params const &p, Tree const &tr, uint32_t t, int lane, uint64_t *stage, uint32_t &settled) {
    int const r0 = 2 * lane, r1 = 2 * lane + 1;
    uint64_t const inN = tr.inmask[static_cast<long long>(t) * 4 + SIDE_N], inS = tr.inmask[static_cast<long long>(t) * 4 + SIDE_S];
    uint64_t const inW = tr.inmask[static_cast<long long>(t) * 4 + SIDE_W], inE = tr.inmask[static_cast<long long>(t) * 4 + SIDE_E];
    Seeds sd;
    sd.src0 = 0;
    sd.src1 = 0;
    if (p.ctl->src_tile[0] == t) {
        uint32_t const rc = p.ctl->src_rc[0];
        int const sr = static_cast<int>(rc >> 8), sc = static_cast<int>(rc & 0xFF);
        if (sr == r0) sd.src0 = 1ull << sc;
        if (sr == r1) sd.src1 = 1ull << sc;
    }
    bool const has_src = lab_any((sd.src0 | sd.src1) != 0);
    if (!(inN | inS | inW | inE) && !has_src) {
        return; // nothing enters this tile
    }
    Wave w;
    w.A0 = lab_ld_ro(p.path + static_cast<long long>(t) * TS + r0);
    w.A1 = lab_ld_ro(p.path + static_cast<long long>(t) * TS + r1);
    w.F0 = 0;
    w.F1 = 0;
    sd.rW0 = ((inW >> r0) & 1ull) ? 0u : INF;
    sd.rW1 = ((inW >> r1) & 1ull) ? 0u : INF;
    sd.rE0 = ((inE >> r0) & 1ull) ? 0u : INF;
    sd.rE1 = ((inE >> r1) & 1ull) ? 0u : INF;
    sd.rSrc = has_src ? 0u : INF;
    sd.nsv = lane == 0 ? inN : (lane == 31 ? inS : 0ull);
#pragma unroll
    for (int b = 0; b < 6; b++) {
        sd.nsp[b] = 0;
    }
    Tile_ctx tc;
    tc.cols = 0;
    tc.dist0 = nullptr;
    tc.edges = tr.loc + static_cast<long long>(t) * SLOTS;
    Emit_acc acc{0u, 0u, 0u, 0u};
    uint32_t base = 0;
    bool first = true;
    for (;;) {
#pragma unroll
        for (int b = 0; b < 6; b++) {
            w.B0[b] = 0;
            w.B1[b] = 0;
        }
        uint64_t const A0s = w.A0, A1s = w.A1;
        bool moving = true;
        for (int gi = 0; gi < WINDOW / 8 && moving; gi++) {
            if (first && gi == 0) {
                step_group<true, true>(w, sd, 0, lane);
            } else {
                step_group<false, false>(w, sd, gi, lane);
            }
            moving = lab_any((w.F0 | w.F1) != 0);
        }
        first = false;
        tree_emit_borders(tc, stage, w, A0s & ~w.A0, A1s & ~w.A1, base, lane, acc, true);
        lab_syncwarp();
        if (!moving) {
            break;
        }
        base += WINDOW;
    }
    settled += acc.settled;
}

LAB_DEV void tree_local_body(Params const &p, Tree const &tr, int lane, uint64_t *stage) {
    if (!tree_live(tr.tc)) {
        return;
    }
    uint32_t settled = 0;
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) {
            t = lab_atomic_add(&tr.tc->ticket[1], 1u);
        }
        t = lab_shfl(t, 0);
        if (t >= static_cast<uint32_t>(p.g.ntiles)) {
            break;
        }
        local_fill_tile(p, tr, t, lane, stage, settled);
    }
    settled = lab_reduce_add(settled);
    if (lane == 0 && settled) {
        lab_atomic_add(reinterpret_cast<uint64_t *>(&tr.tc->local_settled), static_cast<uint64_t>(settled));
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
    uint32_t *par = tr.s[0], *val = tr.r[0];
    for (long long i = gthread; i < static_cast<long long>(tr.nslots) + 2; i += nthreads) {
        uint32_t pa = ROOT, v = INF;
        if (i == ROOT) {
            v = 0;
        } else if (i < tr.nslots) {
            uint32_t const e = static_cast<uint32_t>(i);
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
                for (int guard = 0; guard <= static_cast<int>(SLOTS); guard++) {
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
                    if (guard == static_cast<int>(SLOTS)) {
                        lab_atomic_max(&tc->bad_weight, 2u);
                    }
                }
            }
        }
        par[i] = pa;
        val[i] = v;
    }
}

LAB_DEV void tree_prefix_round_body(Tree const &tr, uint32_t round, long long gthread, long long nthreads) {
    Tree_ctl *tc = tr.tc;
    if (!tree_live(tc) || (round > 0 && tc->prefix_more[round - 1] == 0)) {
        return;
    }
    uint32_t const ROOT = tr.nslots;
    bool const odd = round & 1u;
    uint32_t const *pin = odd ? tr.s[1] : tr.s[0], *vin = odd ? tr.r[1] : tr.r[0];
    uint32_t *pout = odd ? tr.s[0] : tr.s[1], *vout = odd ? tr.r[0] : tr.r[1];
    long long const n = static_cast<long long>(tr.nslots) + 2;
    bool more = false;
    for (long long i = gthread; i < n; i += nthreads) {
        uint32_t const pa = pin[i];
        uint32_t const v = vin[i];
        if (pa == ROOT) {
            pout[i] = ROOT;
            vout[i] = v;
        } else {
            uint32_t const p2 = pin[pa];
            pout[i] = p2;
            vout[i] = v + vin[pa];
            more |= p2 != ROOT;
        }
    }
    if (lab_any(more) && (gthread & 31) == 0 && lab_ld_ro(&tc->prefix_more[round]) == 0u) {
        lab_st_volatile(&tc->prefix_more[round], 1u);
    }
    if (gthread == 0) {
        tc->prefix_buf = (round & 1u) ^ 1u;
    }
}

// ---------------------------------------------------------------------------------------------
// 6 final.  Every tile flooded from its entry crossings at their exact levels; dense field written.
LAB_DEV void final_fill_tile(Params const &p, Tree const &tr, uint32_t const *lvl, uint32_t t, int lane, uint64_t *stage, Warp_stats &ws) {
    Geom const &g = p.g;
    Plane const &pl = p.plane[0];
    int const ty = static_cast<int>(t / static_cast<uint32_t>(g.tiles_x)), tx = static_cast<int>(t % static_cast<uint32_t>(g.tiles_x));
    int const r0 = 2 * lane, r1 = 2 * lane + 1;
    uint32_t const *b = lvl + static_cast<long long>(t) * SLOTS;
    uint32_t sN0 = b[SIDE_N * TS + lane], sN1 = b[SIDE_N * TS + lane + 32];
    uint32_t sS0 = b[SIDE_S * TS + lane], sS1 = b[SIDE_S * TS + lane + 32];
    uint32_t sW0 = b[SIDE_W * TS + r0], sW1 = b[SIDE_W * TS + r1];
    uint32_t sE0 = b[SIDE_E * TS + r0], sE1 = b[SIDE_E * TS + r1];
    Seeds sd;
    sd.src0 = 0;
    sd.src1 = 0;
    if (p.ctl->src_tile[0] == t) {
        uint32_t const rc = p.ctl->src_rc[0];
        int const sr = static_cast<int>(rc >> 8), sc = static_cast<int>(rc & 0xFF);
        if (sr == r0) sd.src0 = 1ull << sc;
        if (sr == r1) sd.src1 = 1ull << sc;
    }
    uint32_t sSrc = (sd.src0 | sd.src1) ? 0u : INF;
    uint32_t pending_min = lab_umin(lab_umin(lab_min3(sN0, sN1, sS0), lab_min3(sS1, sW0, sW1)), lab_min3(sE0, sE1, sSrc));
    pending_min = lab_reduce_min(pending_min);
    ws.passes++;
    if (pending_min == INF) {
        ws.empty++;
        return;
    }
    Tile_ctx tc;
    tc.cols = g.cols;
    tc.edges = nullptr;
    tc.dist0 = pl.dist + (static_cast<long long>(ty) * TS) * g.cols + static_cast<long long>(tx) * TS;
    bool const any_ns = lab_any((sN0 & sN1 & sS0 & sS1) != INF);
    Wave w;
    w.A0 = lab_ld_ro(p.path + static_cast<long long>(t) * TS + r0);
    w.A1 = lab_ld_ro(p.path + static_cast<long long>(t) * TS + r1);
    uint64_t const P0 = w.A0, P1 = w.A1;
    w.F0 = 0;
    w.F1 = 0;
    Emit_acc acc{0u, 0u, 0u, 0u};
    uint32_t base = pending_min;
    for (;;) { // one batch of up to WINDOW levels (cf. relax_tile)
#pragma unroll
        for (int k = 0; k < 6; k++) {
            w.B0[k] = 0;
            w.B1[k] = 0;
        }
        uint64_t const A0s = w.A0, A1s = w.A1;
        sd.rW0 = sW0 - base;
        sd.rW1 = sW1 - base;
        sd.rE0 = sE0 - base;
        sd.rE1 = sE1 - base;
        sd.rSrc = sSrc - base;
        uint32_t grp = seed_group_bit(sW0, base) | seed_group_bit(sW1, base) | seed_group_bit(sE0, base) | seed_group_bit(sE1, base)
                       | seed_group_bit(sSrc, base);
        sd.nsv = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            sd.nsp[k] = 0;
        }
        bool ns_window = false;
        if (any_ns) {
            uint32_t const rN0 = sN0 - base, rN1 = sN1 - base, rS0 = sS0 - base, rS1 = sS1 - base;
            bool const vN0 = rN0 < WINDOW, vN1 = rN1 < WINDOW, vS0 = rS0 < WINDOW, vS1 = rS1 < WINDOW;
            uint64_t const wn = lab_ballot64(vN0, vN1), wsb = lab_ballot64(vS0, vS1);
            sd.nsv = lane == 0 ? wn : (lane == 31 ? wsb : 0ull);
            if (wn | wsb) {
                ns_window = true;
                grp |= seed_group_bit(sN0, base) | seed_group_bit(sN1, base) | seed_group_bit(sS0, base) | seed_group_bit(sS1, base);
#pragma unroll
                for (int k = 0; k < 6; k++) {
                    uint64_t const pn = lab_ballot64(vN0 && ((rN0 >> k) & 1), vN1 && ((rN1 >> k) & 1));
                    uint64_t const ps = lab_ballot64(vS0 && ((rS0 >> k) & 1), vS1 && ((rS1 >> k) & 1));
                    sd.nsp[k] = lane == 0 ? pn : (lane == 31 ? ps : 0ull);
                }
            }
        }
        grp = lab_reduce_or(grp);
        int gi = 0;
        for (; gi < WINDOW / 8; gi++) {
            if (!((grp >> gi) & 1u)) {
                step_group<false, false>(w, sd, gi, lane);
            } else if (ns_window) {
                step_group<true, true>(w, sd, gi, lane);
            } else {
                step_group<true, false>(w, sd, gi, lane);
            }
            ws.levels += 8;
            if (!lab_any((w.F0 | w.F1) != 0)) {
                uint32_t const later = grp & ~((2u << gi) - 1u);
                if (later == 0) {
                    gi++;
                    break;
                }
                gi = lab_ffs64(later) - 2;
            }
        }
        uint64_t const S0 = A0s & ~w.A0, S1 = A1s & ~w.A1;
        uint32_t const n_batch = tree_emit_borders(tc, stage, w, S0, S1, base, lane, acc, false);
        emit_interior(tc, stage, S0, S1, base, lane, n_batch);
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
        if (lab_any((w.F0 | w.F1) != 0)) {
            base = done_lvl + 1;
        } else {
            pending_min = lab_umin(lab_umin(lab_min3(sN0, sN1, sS0), lab_min3(sS1, sW0, sW1)), lab_min3(sE0, sE1, sSrc));
            pending_min = lab_reduce_min(pending_min);
            if (pending_min == INF) {
                break;
            }
            base = pending_min;
        }
    }
    lab_st_cg(pl.visited + static_cast<long long>(t) * TS + r0, P0 & ~w.A0);
    lab_st_cg(pl.visited + static_cast<long long>(t) * TS + r1, P1 & ~w.A1);
    ws.settled += acc.settled;
    ws.max_lvl = acc.max_lvl > ws.max_lvl ? acc.max_lvl : ws.max_lvl;
}

LAB_DEV void tree_final_body(Params const &p, Tree const &tr, int lane, uint64_t *stage) {
    Tree_ctl *tc = tr.tc;
    if (!tree_live(tc)) {
        return;
    }
    uint32_t const *lvl = tc->prefix_buf ? tr.r[1] : tr.r[0];
    Warp_stats ws;
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) {
            t = lab_atomic_add(&tc->ticket[2], 1u);
        }
        t = lab_shfl(t, 0);
        if (t >= static_cast<uint32_t>(p.g.ntiles)) {
            break;
        }
        final_fill_tile(p, tr, lvl, t, lane, stage, ws);
    }
    flush_stats(p.ctl, ws, lane);
}

// One thread, after the final flood: if the tree pipeline made the field, take the targets' levels
// (the solvers' bounds and paint rules come from them) and call the wave off; otherwise leave the
// control block as init_run_body armed it and the tile wave runs as if nothing had happened.
LAB_DEV void tree_gate_body(Params const &p, Tree const &tr) {
    Tree_ctl *tc = tr.tc;
    Control *ctl = p.ctl;
    if (!tree_live(tc)) {
        return;
    }
    for (uint32_t j = 0; j < ctl->n_targets; j++) {
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
    tc->used = 1;
}

} // namespace lab
