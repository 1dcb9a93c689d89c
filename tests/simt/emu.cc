// emu.cc -- TEST INFRASTRUCTURE: runs the product's kernel BODIES (csrc/device/engine.cuh,
// post.cuh) on the CPU warp emulator and exposes the same operations as the C ABI, so pytest can
// compare them with the oracle without a GPU.  The orchestration mirrors csrc/lab.cu.
#define LAB_SIMT_HOST 1
#include "simt_host.hh"

#include "../../multithreading-with-mazes_b200/csrc/device/engine.cuh"
#include "../../multithreading-with-mazes_b200/csrc/device/post.cuh"
#include "../../multithreading-with-mazes_b200/csrc/device/tree.cuh"
#include "../../multithreading-with-mazes_b200/csrc/device/lattice.cuh"

#include <cstring>
#include <vector>

using namespace lab;

namespace {

struct Workspace {
    Geom g;
    uint16_t *grid_ptr = nullptr;
    std::vector<uint64_t> path, cross;
    std::vector<uint64_t> visited[MAX_PLANES];
    std::vector<uint32_t> edges[MAX_PLANES], state[MAX_PLANES], dist[MAX_PLANES];
    Control ctl;
    std::vector<uint32_t> queue;
    Result res;
};

int g_tree_mode = 1;      // 0: never, 1: when the counts allow it (as csrc/lab.cu does)
int g_lattice_mode = 1;   // 0: the general pipeline only, 1: the lattice road first (as csrc/lab.cu does)
Tree_ctl g_tree_last{};
int g_paint_fused = 0; // the last search applied the solvers' paint rules inside lat_levels_body
int g_path_segments = -1; // ... and marked hunt's winner path from the tour's ranks: segments walked (-1: it did not)
int g_opt_solver_levels = 1; // lab_grid_set_option(LAB_OPT_SOLVER_LEVELS)
int g_skip_field = 0;        // this solver call may leave its field unwritten (gather without paths, the option off)
int32_t *g_front_rc = nullptr;
int32_t *g_hunt_out = nullptr; // where hunt's path goes (set by emu_solver before the search, as solver_common does)
unsigned long long g_hunt_out_cap = 0;

// The spanning-tree pipeline, orchestrated as run_tree() in csrc/lab.cu does.
void run_tree(Workspace &w, Params &p, Run_desc const &d, int nwarps, unsigned sched_seed) {
    g_tree_last = Tree_ctl{};
    g_paint_fused = 0;
    g_path_segments = -1;
    if (d.nplanes != 1 || d.src_r[0] < 0 || !g_tree_mode) return;
    size_t const nt = static_cast<size_t>(w.g.ntiles);
    size_t const n = nt * SLOTS + 2;
    std::vector<uint32_t> succ0(n), sr0(2 * n), sr1(2 * n), loc(nt * SLOTS), alist(nt * SLOTS);
    std::vector<uint64_t> inmask(nt * 4);
    Tree_ctl tc{};
    Tree tr{};
    tr.succ0 = succ0.data();
    tr.sr[0] = sr0.data();
    tr.sr[1] = sr1.data();
    tr.inmask = inmask.data();
    tr.loc = loc.data();
    tr.alist = alist.data();
    tr.tc = &tc;
    tr.nslots = static_cast<uint32_t>(nt * SLOTS);
    tr.mark_bits = d.game == GAME_DISTANCE ? SQ_MEASURE : 0u;
    simt::launch(nwarps, [&](int wid, int lane) { tree_count_body(p, tr, wid, nwarps, lane); });
    bool const lat_try = g_lattice_mode && (w.g.rows & 1) && (w.g.cols & 1) && ((d.src_r[0] | d.src_c[0]) & 1);
    tr.lat_try = lat_try ? 1u : 0u;
    tree_init_body(p, tr);
    if (lat_try) { // the lattice road, orchestrated as run_lattice() in csrc/lab.cu does
        Lattice lt{};
        lt.sup_y = (w.g.tiles_y + LST - 1) / LST;
        lt.sup_x = (w.g.tiles_x + LST - 1) / LST;
        lt.nsuper = lt.sup_y * lt.sup_x;
        lt.nslots = static_cast<uint32_t>(nt * LSLOTS);
        std::vector<uint32_t> mask(nt * 96, 0xDEADBEEFu), xmask(nt * 4), R(nt * LSLOTS, 0xDEADBEEFu),
            dplane(nt * L_DEPTH_PLANES * 32), emask(nt * 4), wsum(nt * LSLOTS + 2 + L_CHUNK, 0xDEADBEEFu), chunk(nt * LSLOTS / L_CHUNK + 2);
        std::vector<uint16_t> label(nt * LSLOTS), cmap(nt * 1024, 0xDEAD);
        std::vector<uint8_t> outslot(nt * LSLOTS, 0xAB);
        std::vector<uint64_t> xd(nt * LSLOTS, 0xDEADBEEFDEADBEEFull), gb(static_cast<size_t>(lt.nsuper) * LB, 0xDEADBEEFDEADBEEFull);
        lt.mask = mask.data(); lt.xmask = xmask.data(); lt.outslot = outslot.data(); lt.label = label.data(); lt.cmap = cmap.data();
        lt.xd = xd.data(); lt.gb = gb.data(); lt.R = R.data(); lt.dplane = dplane.data(); lt.emask = emask.data(); lt.wsum = wsum.data();
        lt.chunk = chunk.data();
        lt.tc = &tc;
        lt.nsuper_global = lt.nsuper;
        lt.nslots_global = lt.nslots;
        lt.peers = 1;
        lt.gb_peer[0] = lt.gb;
        lt.wsum_peer[0] = lt.wsum;
        std::vector<uint32_t> lsm(static_cast<size_t>(nwarps) * LSUP_SLOTS);
        long long const nth = static_cast<long long>(nwarps) * 32;
        simt::launch(nwarps, [&](int wid, int lane) { lat_link_body(p, lt, lane, lsm.data() + static_cast<size_t>(wid) * LSUP_SLOTS); }, sched_seed);
        simt::launch(nwarps, [&](int wid, int lane) { lat_rank_local_body(p, lt, lane, lsm.data() + static_cast<size_t>(wid) * LSUP_SLOTS); }, sched_seed);
        simt::launch(nwarps, [&](int wid, int lane) { lat_rank_global_body(lt, static_cast<long long>(wid) * 32 + lane, nth); }, sched_seed);
        simt::launch(nwarps, [&](int wid, int lane) { lat_rank_expand_body(lt, static_cast<long long>(wid) * 32 + lane, nth); });
        simt::launch(nwarps, [&](int wid, int lane) { lat_flood_body(p, lt, lane); }, sched_seed);
        simt::launch(nwarps, [&](int wid, int lane) { lat_scan_chunks_body(lt, wid, nwarps, lane); });
        simt::launch(1, [&](int, int lane) { lat_scan_tops_body(lt, lane); });
        if (d.game == GAME_HUNT || d.game == GAME_GATHER) { // as run_lattice(): k_lat_targets, then the levels with the paint fused
            int32_t frc[8] = {0};
            for (int j = 0; j < d.n_targets; j++) {
                frc[2 * j] = d.tgt_r[j];
                frc[2 * j + 1] = d.tgt_c[j];
            }
            std::vector<uint32_t> tsm(4 * static_cast<size_t>(L_LINK_SM));
            simt::launch(4, [&](int wid, int lane) { lat_targets_body(p, lt, wid, lane, tsm.data() + static_cast<size_t>(wid) * L_LINK_SM); });
            bool const skip_field = g_skip_field && d.game == GAME_GATHER;
            if (lat_live(lt.tc)) {
                resolve_body(p, d, &w.res, frc);
                p.ctl->painted = 1u;
                if (skip_field) p.ctl->fronts_done = 1u;
                g_paint_fused = skip_field ? 2 : 1;
            }
            if (skip_field) {
                simt::launch(4, [&](int wid, int lane) { lat_fronts_body(p, lt, &w.res, frc, g_front_rc, wid, lane, tsm.data() + static_cast<size_t>(wid) * L_LINK_SM); });
                simt::launch(nwarps, [&](int wid, int lane) { lat_levels_body<LAT_PAINT_ONLY>(p, lt, lane, lsm.data() + static_cast<size_t>(wid) * LSUP_SLOTS, &w.res, w.grid_ptr); }, sched_seed);
            } else
            simt::launch(nwarps, [&](int wid, int lane) { lat_levels_body<LAT_PAINT>(p, lt, lane, lsm.data() + static_cast<size_t>(wid) * LSUP_SLOTS, &w.res, w.grid_ptr); }, sched_seed);
            if ((d.game == GAME_HUNT || (d.game == GAME_GATHER && g_hunt_out)) && !skip_field) { // paths from the tour's ranks, as run_lattice() launches them
                uint32_t *seg = reinterpret_cast<uint32_t *>(xd.data());
                uint32_t const seg_cap = lt.nslots;
                tc.lat_path_on = 0u;
                for (int q = 0; q < 4; q++) tc.lat_path_nseg[q] = 0u;
                simt::launch(4, [&](int wid, int lane) {
                    lat_path_begin_body(p, lt, d.game, &w.res, frc, w.grid_ptr, g_hunt_out, g_hunt_out_cap, g_hunt_out_cap, seg, seg_cap, wid, lane);
                });
                if (lat_live(lt.tc) && p.ctl->painted && !p.ctl->fronts_done && !p.ctl->walk_fallback && (d.game != GAME_HUNT || w.res.winner >= 0)) {
                    p.ctl->walked = 1u;
                }
                simt::launch(nwarps, [&](int wid, int lane) { lat_path_find_body(p, lt, seg, seg_cap, wid, nwarps, lane); });
                std::vector<uint32_t> wsm(static_cast<size_t>(nwarps) * WALK_SM_U32);
                uint32_t const repaint = d.game == GAME_HUNT && w.res.winner >= 0 ? (1u << w.res.winner) << SQ_PAINT_SHIFT : 0u;
                simt::launch(nwarps, [&](int wid, int lane) {
                    if (tc.lat_path_on) lat_path_walk_body(p, lt, &w.res, repaint, w.grid_ptr, g_hunt_out, g_hunt_out_cap, g_hunt_out_cap, seg, seg_cap, wid, nwarps, lane, wsm.data() + static_cast<size_t>(wid) * WALK_SM_U32);
                });
                if (tc.lat_path_on) {
                    g_path_segments = 0;
                    for (int q = 0; q < 4; q++) g_path_segments += static_cast<int>(tc.lat_path_nseg[q]);
                }
            }
        } else {
            simt::launch(nwarps, [&](int wid, int lane) { lat_levels_body<LAT_LEVELS>(p, lt, lane, lsm.data() + static_cast<size_t>(wid) * LSUP_SLOTS, nullptr, nullptr); }, sched_seed);
        }
    }
    size_t const SM = FLOOD_U64 > WALK_U64 ? FLOOD_U64 : WALK_U64;
    std::vector<uint64_t> sm(static_cast<size_t>(nwarps) * SM);
    simt::launch(nwarps, [&](int wid, int lane) { tree_walk_body(p, tr, lane, sm.data() + static_cast<size_t>(wid) * SM); }, sched_seed);
    long long const nthreads = static_cast<long long>(nwarps) * 32;
    int rounds = 0;
    while ((1ull << rounds) < n) rounds++;
    for (int r = 0; r < rounds; r++) {
        simt::launch(nwarps, [&](int wid, int lane) { tree_rank_round_body(tr, static_cast<uint32_t>(r), static_cast<long long>(wid) * 32 + lane, nthreads); });
    }
    simt::launch(nwarps, [&](int wid, int lane) { tree_jump_settle_body(tr, tc.rank_buf, static_cast<long long>(wid) * 32 + lane, nthreads); });
    simt::launch(nwarps, [&](int wid, int lane) { tree_orient_body(p, tr, wid, nwarps, lane); });
    simt::launch(nwarps, [&](int wid, int lane) { tree_flood_body(p, tr, lane, sm.data() + static_cast<size_t>(wid) * SM); }, sched_seed);
    simt::launch(nwarps, [&](int wid, int lane) { tree_parent_body(p, tr, static_cast<long long>(wid) * 32 + lane, nthreads); });
    for (int r = 0; r < rounds; r++) {
        simt::launch(nwarps, [&](int wid, int lane) { tree_prefix_round_body(tr, static_cast<uint32_t>(r), static_cast<long long>(wid) * 32 + lane, nthreads); });
    }
    simt::launch(nwarps, [&](int wid, int lane) { tree_jump_settle_body(tr, tc.prefix_buf, static_cast<long long>(wid) * 32 + lane, nthreads); });
    simt::launch(nwarps, [&](int wid, int lane) { tree_fixup_body(p, tr, w.grid_ptr, wid, nwarps, lane); });
    if (d.game == GAME_DISTANCE) {
        simt::launch(nwarps, [&](int wid, int lane) { room_distance_body(p, w.grid_ptr, &w.res.settled, static_cast<long long>(wid) * 32 + lane, nthreads); });
    }
    simt::launch(nwarps, [&](int wid, int lane) { tree_gate_body(p, tr, static_cast<long long>(wid) * 32 + lane, nthreads); });
    g_tree_last = tc;
}

Params make_params(Workspace &w, int nplanes) {
    Params p{};
    p.g = w.g;
    p.path = w.path.data();
    p.cross = w.cross.data();
    for (int k = 0; k < MAX_PLANES; k++) {
        int const kk = k < nplanes ? k : 0;
        p.plane[k] = {w.visited[kk].data(), w.edges[kk].data(), w.state[kk].data(), w.dist[kk].data()};
    }
    p.nplanes = nplanes;
    p.ctl = &w.ctl;
    p.queue = w.queue.data();
    p.qmask = static_cast<uint32_t>(w.queue.size() - 1);
    return p;
}

int search(Workspace &w, uint16_t *grid, int rows, int cols, Run_desc const &d, uint16_t mask, uint16_t value,
           int nwarps, unsigned sched_seed) {
    Geom &g = w.g;
    w.grid_ptr = grid;
    g.rows = rows;
    g.cols = cols;
    g.tiles_y = (rows + TS - 1) / TS;
    g.tiles_x = (cols + TS - 1) / TS;
    g.ntiles = g.tiles_y * g.tiles_x;
    size_t const nt = static_cast<size_t>(g.ntiles), cells = static_cast<size_t>(rows) * cols;
    w.path.assign(nt * TS, 0);
    w.cross.assign(nt * 4, 0);
    for (int k = 0; k < d.nplanes; k++) {
        w.visited[k].assign(nt * TS, 0);
        w.edges[k].assign((nt + 2 * static_cast<size_t>(g.tiles_x)) * 4 * TS, INF);
        w.state[k].assign(nt, 0);
        // (as csrc/lab.cu: when the tree pipeline may run, the field is NOT reset here -- garbage in, so
        // that the tests see the fix-up / the gate write every square)
        bool const tree_may = d.nplanes == 1 && d.src_r[0] >= 0 && g_tree_mode;
        w.dist[k].assign(cells, tree_may ? 0xABABABABu : INF);
    }
    std::memset(&w.ctl, 0, sizeof w.ctl);
    std::memset(&w.res, 0, sizeof w.res);
    size_t cap = 1024;
    while (cap < 2 * (static_cast<size_t>(d.nplanes) * nt + static_cast<size_t>(nwarps) + 64)) cap <<= 1;
    w.queue.assign(cap, Q_EMPTY);
    Params p = make_params(w, d.nplanes);

    simt::launch(nwarps, [&](int wid, int lane) { pack_body(g, grid, w.path.data(), mask, value, static_cast<uint16_t>(d.spec_mark), wid, nwarps, lane); });
    force_sources_body(g, w.path.data(), d, grid);
    simt::launch(nwarps, [&](int wid, int lane) { cross_body(g, w.path.data(), w.cross.data(), nullptr, nullptr, wid, nwarps, lane); });
    init_run_body(p, d, 60ull * 1000000000ull);
    run_tree(w, p, d, nwarps, sched_seed);
    std::vector<uint64_t> stages(static_cast<size_t>(nwarps) * STAGE_U64);
    simt::launch(nwarps, [&](int wid, int lane) { bfs_worker_body(p, lane, stages.data() + static_cast<size_t>(wid) * STAGE_U64); }, sched_seed);
    if (w.ctl.error) return -100 - static_cast<int>(w.ctl.error);
    if (w.ctl.pending != 0) return -200;
    return 0;
}

void fill_desc(Run_desc &d) {
    std::memset(&d, 0, sizeof d);
    for (int k = 0; k < MAX_PLANES; k++) d.src_r[k] = d.src_c[k] = -1;
}

void stats_out(Workspace const &w, uint64_t *st) {
    if (!st) return;
    st[0] = w.ctl.n_activations;
    st[1] = w.ctl.n_empty;
    st[2] = w.ctl.n_levels;
    st[3] = w.ctl.n_settled;
    st[4] = w.ctl.n_rechecks;
    st[5] = w.ctl.n_resettles;
    st[6] = w.ctl.n_pushes;
    st[7] = w.ctl.n_direct;
}

} // namespace

extern "C" {

// Runs::paint_runs' run lengths from a level field (post.cuh runs_body), as lab_runs_map launches it
int emu_runs(int rows, int cols, uint32_t const *dist, uint32_t *runs, uint32_t *max_out, int nwarps) {
    Geom g{};
    g.rows = rows;
    g.cols = cols;
    *max_out = 0;
    long long const nthreads = static_cast<long long>(nwarps) * 32;
    simt::launch(nwarps, [&](int wid, int lane) { runs_body(g, dist, runs, max_out, static_cast<long long>(wid) * 32 + lane, nthreads); });
    return 0;
}
void emu_set_tree(int mode) { g_tree_mode = mode; }
void emu_set_option_solver_levels(int v) { g_opt_solver_levels = v; }
void emu_set_lattice(int mode) { g_lattice_mode = mode; }
int emu_last_lattice_error(void) { return static_cast<int>(g_tree_last.lat_err); }
// 0: the tile wave made the field; 1: the tree pipeline did
int emu_last_tree_used(void) { return static_cast<int>(g_tree_last.used); }
int emu_last_paint_fused(void) { return g_paint_fused; }
int emu_last_path_segments(void) { return g_path_segments; }

int emu_distance(int rows, int cols, uint16_t *grid, int sr, int sc, uint32_t *dist_out, uint64_t *max_out,
                 uint64_t *settled_out, int nwarps, unsigned sched_seed, uint64_t *stats) {
    Workspace w;
    Run_desc d;
    fill_desc(d);
    d.game = GAME_DISTANCE;
    d.nplanes = 1;
    d.src_r[0] = sr;
    d.src_c[0] = sc;
    d.bound_mode = BOUND_NONE;
    // (as lab_distance_map: where the lattice road is tried, pack sets the measure bits in advance)
    d.spec_mark = (g_tree_mode && g_lattice_mode && (rows & 1) && (cols & 1) && ((sr | sc) & 1)) ? SQ_MEASURE : 0u;
    int rc = search(w, grid, rows, cols, d, SQ_PATH | SQ_MEASURE, SQ_PATH, nwarps, sched_seed);
    if (rc) return rc;
    Params p = make_params(w, 1);
    int32_t frc[8] = {0};
    resolve_body(p, d, &w.res, frc);
    long long const nthreads = static_cast<long long>(nwarps) * 32;
    simt::launch(nwarps, [&](int wid, int lane) {
        max_level_body(w.g, w.dist[0].data(), &w.ctl, &w.res, static_cast<long long>(wid) * 32 + lane, nthreads);
    });
    simt::launch(nwarps, [&](int wid, int lane) { measure_body(w.g, grid, w.visited[0].data(), w.path.data(), &w.ctl, &w.res, wid, nwarps, lane); });
    std::memcpy(dist_out, w.dist[0].data(), w.dist[0].size() * sizeof(uint32_t));
    *max_out = w.res.max_level;
    *settled_out = w.res.settled;
    stats_out(w, stats);
    return 0;
}

// The heat map of a level field (post.cuh heatmap_body), as lab_distance_paint runs it.
int emu_heatmap(int rows, int cols, uint32_t const *dist, uint64_t max, int channel, uint32_t *pixels, int nwarps) {
    Geom g{};
    g.rows = rows;
    g.cols = cols;
    Result res{};
    res.max_level = max;
    long long const nthreads = static_cast<long long>(nwarps) * 32;
    simt::launch(nwarps, [&](int wid, int lane) { heatmap_body(g, dist, &res, channel, pixels, static_cast<long long>(wid) * 32 + lane, nthreads); });
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Row bands: the lab_band_* / lab_band_tree_* entry points of csrc/lab.cu, same names, same protocol, with
// host memory standing in for device memory (so that the real driver, multithreading-with-mazes_b200/bands.py,
// can run under gloo on CPU tensors with these kernel bodies underneath).
struct Emu_band {
    Workspace w;
    Run_desc d;
    std::vector<uint64_t> ghost_n, ghost_s;
    std::vector<uint64_t> stages, sm;
    int nwarps = 4;
    unsigned sched = 0;
    Tree tr{};
    Tree_ctl tc{};
    uint32_t *alist = nullptr;
    uint32_t const *border_list = nullptr;
    uint32_t border_count = 0, crossings = 0;
    Band_layout const *layout = nullptr;
    int world = 1, me = 0;
    Tree_ctl tree_last{};
    // the lattice road across bands (emu_band_lat_*)
    Lattice lat{};
    Tree_ctl lat_tc{};
    std::vector<uint8_t> l_outslot;
    std::vector<uint32_t> l_mask, l_xmask, l_R, l_dplane, l_emask, l_sm;
    std::vector<uint16_t> l_label, l_cmap;
    std::vector<uint64_t> l_xd;
};

void *emu_band_create(int rows, int cols, uint16_t *grid, int nwarps, unsigned sched_seed) {
    auto *e = new Emu_band;
    Geom &g = e->w.g;
    g.rows = rows;
    g.cols = cols;
    g.tiles_y = (rows + TS - 1) / TS;
    g.tiles_x = (cols + TS - 1) / TS;
    g.ntiles = g.tiles_y * g.tiles_x;
    e->w.grid_ptr = grid;
    e->nwarps = nwarps;
    e->sched = sched_seed;
    return e;
}
void emu_band_destroy(void *h) { delete static_cast<Emu_band *>(h); }
int emu_band_cols_padded(void *h) { return static_cast<Emu_band *>(h)->w.g.tiles_x * TS; }
uint32_t *emu_band_levels(void *h) { return static_cast<Emu_band *>(h)->w.dist[0].data(); }
uint32_t *emu_band_levels_plane(void *h, int k) { return static_cast<Emu_band *>(h)->w.dist[k].data(); }
int emu_band_tree_used(void *h) { return static_cast<int>(static_cast<Emu_band *>(h)->tree_last.used); }

int emu_band_begin_lattice(void *h, int game, int src_r, int src_c, int lattice_next);
int emu_band_begin_game(void *h, int game, int src_r, int src_c) { return emu_band_begin_lattice(h, game, src_r, src_c, 0); }
int emu_band_begin(void *h, int src_r, int src_c) { return emu_band_begin_game(h, GAME_DISTANCE, src_r, src_c); }

static int emu_band_begin_common(Emu_band *e, int lattice_next) {
    Workspace &w = e->w;
    Geom const &g = w.g;
    int const game = e->d.game;
    size_t const nt = static_cast<size_t>(g.ntiles), cells = static_cast<size_t>(g.rows) * g.cols;
    w.path.assign(nt * TS, 0);
    w.cross.assign(nt * 4, 0);
    for (int k = 0; k < e->d.nplanes; k++) {
        w.visited[k].assign(nt * TS, 0);
        w.edges[k].assign((nt + 2 * static_cast<size_t>(g.tiles_x)) * 4 * TS, INF);
        w.state[k].assign(nt, 0);
        w.dist[k].assign(cells, lattice_next ? 0xABABABABu : INF); // (as lab_band_begin_lattice: the lattice road writes or wipes the field)
    }
    std::memset(&w.ctl, 0, sizeof w.ctl);
    std::memset(&w.res, 0, sizeof w.res);
    size_t cap = 1024;
    while (cap < 2 * (static_cast<size_t>(e->d.nplanes) * nt + static_cast<size_t>(e->nwarps) + 64)) cap <<= 1;
    w.queue.assign(cap, Q_EMPTY);
    e->ghost_n.assign(g.tiles_x, 0);
    e->ghost_s.assign(g.tiles_x, 0);
    e->tree_last = Tree_ctl{};
    int const nw = e->nwarps;
    uint16_t const mask = game == GAME_DISTANCE ? (SQ_PATH | SQ_MEASURE) : SQ_PATH;
    simt::launch(nw, [&](int wid, int lane) { pack_body(g, w.grid_ptr, w.path.data(), mask, SQ_PATH, static_cast<uint16_t>(e->d.spec_mark), wid, nw, lane); });
    force_sources_body(g, w.path.data(), e->d, w.grid_ptr);
    return 0;
}
int emu_band_begin_lattice(void *h, int game, int src_r, int src_c, int lattice_next) {
    auto *e = static_cast<Emu_band *>(h);
    fill_desc(e->d);
    e->d.game = game;
    e->d.nplanes = 1;
    if (src_r >= 0) {
        e->d.src_r[0] = src_r;
        e->d.src_c[0] = src_c;
    }
    e->d.bound_mode = BOUND_NONE;
    e->d.spec_mark = (lattice_next && game == GAME_DISTANCE) ? SQ_MEASURE : 0u;
    return emu_band_begin_common(e, lattice_next);
}
int emu_band_begin_corners(void *h, int32_t const *src_rc) { // mirrors lab_band_begin_corners
    auto *e = static_cast<Emu_band *>(h);
    fill_desc(e->d);
    e->d.game = GAME_CORNERS;
    e->d.nplanes = 4;
    for (int k = 0; k < 4; k++) {
        if (src_rc[2 * k] >= 0) {
            e->d.src_r[k] = src_rc[2 * k];
            e->d.src_c[k] = src_rc[2 * k + 1];
        }
    }
    e->d.bound_mode = BOUND_NONE;
    return emu_band_begin_common(e, 0);
}
int emu_band_planes(void *h) { return static_cast<Emu_band *>(h)->d.nplanes; }

int emu_band_level_at_plane(void *h, int plane, int r, int c, uint32_t *level) {
    auto *e = static_cast<Emu_band *>(h);
    *level = e->w.dist[plane][static_cast<size_t>(r) * e->w.g.cols + c];
    return 0;
}
int emu_band_level_at(void *h, int r, int c, uint32_t *level) { return emu_band_level_at_plane(h, 0, r, c, level); }

int emu_band_paint_planes(void *h, int n_rules, uint32_t const *plane, uint32_t const *below, uint32_t const *bits, uint64_t *painted);
int emu_band_paint(void *h, int n_rules, uint32_t const *below, uint32_t const *bits, uint64_t *painted) {
    uint32_t const plane0[4] = {0u, 0u, 0u, 0u};
    return emu_band_paint_planes(h, n_rules, plane0, below, bits, painted);
}
int emu_band_paint_planes(void *h, int n_rules, uint32_t const *plane, uint32_t const *below, uint32_t const *bits, uint64_t *painted) {
    auto *e = static_cast<Emu_band *>(h);
    Workspace &w = e->w;
    Params p = make_params(w, e->d.nplanes);
    int const nw = e->nwarps;
    long long const nthreads = static_cast<long long>(nw) * 32;
    w.res = Result{};
    w.res.winner = -1;
    w.res.n_rules = static_cast<uint32_t>(n_rules);
    for (int k = 0; k < n_rules; k++) w.res.rule[k] = {plane[k], below[k], bits[k], 0u};
    simt::launch(nw, [&](int wid, int lane) { paint_body(p, w.grid_ptr, &w.res, static_cast<long long>(wid) * 32 + lane, nthreads); });
    *painted = w.res.settled;
    return 0;
}

int emu_band_walk(void *h, int r, int c, int include_first, int first_only, uint32_t repaint_bits, int32_t *path, uint64_t path_cap,
                  int64_t *state) {
    auto *e = static_cast<Emu_band *>(h);
    Workspace &w = e->w;
    Params p = make_params(w, 1);
    std::vector<uint32_t> wsm(WALK_SM_U32);
    simt::launch(1, [&](int, int lane) {
        walk_body(p, 0, r, c, w.grid_ptr, path, path_cap, repaint_bits, first_only, &w.res, 0, lane, wsm.data(), include_first);
    });
    state[0] = static_cast<int64_t>(w.res.path_len[0]);
    state[1] = w.res.walk_r[0];
    state[2] = w.res.walk_c[0];
    state[3] = w.res.walk_left[0] == INF ? -1 : static_cast<int64_t>(w.res.walk_left[0]);
    return 0;
}

int emu_band_recolour(void *h, int r, int c, uint16_t paint_bits) {
    auto *e = static_cast<Emu_band *>(h);
    uint16_t &sq = e->w.grid_ptr[static_cast<size_t>(r) * e->w.g.cols + c];
    sq = static_cast<uint16_t>((sq & ~SQ_PAINT_MASK) | paint_bits);
    return 0;
}

int emu_band_finish_game(void *h) {
    auto *e = static_cast<Emu_band *>(h);
    return e->w.ctl.pending != 0 ? -200 : 0;
}

int emu_band_path_rows(void *h, uint8_t *top, uint8_t *bottom) {
    auto *e = static_cast<Emu_band *>(h);
    for (long long i = 0; i < static_cast<long long>(e->w.g.tiles_x) * TS; i++) {
        band_path_rows_body(e->w.g, e->w.path.data(), top, bottom, i);
    }
    return 0;
}

int emu_band_set_neighbour_paths(void *h, uint8_t const *above_bottom, uint8_t const *below_top) {
    auto *e = static_cast<Emu_band *>(h);
    Workspace &w = e->w;
    int const nw = e->nwarps;
    simt::launch(nw, [&](int wid, int lane) { band_ghost_words_body(w.g, above_bottom, below_top, e->ghost_n.data(), e->ghost_s.data(), wid, nw, lane); });
    simt::launch(nw, [&](int wid, int lane) { cross_body(w.g, w.path.data(), w.cross.data(), e->ghost_n.data(), e->ghost_s.data(), wid, nw, lane); });
    Params p = make_params(w, e->d.nplanes);
    init_run_body(p, e->d, 60ull * 1000000000ull);
    return 0;
}

int emu_band_relax(void *h) {
    auto *e = static_cast<Emu_band *>(h);
    Params p = make_params(e->w, e->d.nplanes);
    int const nw = e->nwarps;
    e->stages.assign(static_cast<size_t>(nw) * STAGE_U64, 0);
    simt::launch(nw, [&](int wid, int lane) { bfs_worker_body(p, lane, e->stages.data() + static_cast<size_t>(wid) * STAGE_U64); }, e->sched);
    if (e->w.ctl.error) return -100 - static_cast<int>(e->w.ctl.error);
    return 0;
}

int emu_band_export_edges(void *h, uint32_t *top, uint32_t *bottom) {
    auto *e = static_cast<Emu_band *>(h);
    long long const n = static_cast<long long>(e->w.g.tiles_x) * TS;
    for (int k = 0; k < e->d.nplanes; k++) {
        for (long long i = 0; i < n; i++) {
            band_export_edges_body(e->w.g, e->w.edges[k].data(), top + k * n, bottom + k * n, i);
        }
    }
    return 0;
}

int emu_band_import_edges(void *h, uint32_t const *above_bottom, uint32_t const *below_top, uint32_t *activated) {
    auto *e = static_cast<Emu_band *>(h);
    Params p = make_params(e->w, e->d.nplanes);
    int const nw = e->nwarps;
    long long const n = static_cast<long long>(e->w.g.tiles_x) * TS;
    *activated = 0;
    band_rearm_body(p, 60ull * 1000000000ull);
    for (int k = 0; k < e->d.nplanes; k++) {
        simt::launch(nw, [&](int wid, int lane) {
            band_import_edges_body(p, k, above_bottom ? above_bottom + k * n : nullptr, below_top ? below_top + k * n : nullptr, activated, wid, nw, lane);
        });
    }
    return 0;
}

int emu_band_finish(void *h, uint64_t *local_max, uint64_t *local_settled) {
    auto *e = static_cast<Emu_band *>(h);
    Workspace &w = e->w;
    Params p = make_params(w, 1);
    int const nw = e->nwarps;
    int32_t frc[8] = {0};
    resolve_body(p, e->d, &w.res, frc);
    long long const nthreads = static_cast<long long>(nw) * 32;
    simt::launch(nw, [&](int wid, int lane) { max_level_body(w.g, w.dist[0].data(), &w.ctl, &w.res, static_cast<long long>(wid) * 32 + lane, nthreads); });
    simt::launch(nw, [&](int wid, int lane) { measure_body(w.g, w.grid_ptr, w.visited[0].data(), w.path.data(), &w.ctl, &w.res, wid, nw, lane); });
    if (w.ctl.pending != 0) return -200;
    *local_max = w.res.max_level;
    *local_settled = w.res.settled;
    return 0;
}

int emu_band_tree_bind(void *h, uint32_t tile_base, uint32_t ntiles_global, uint32_t *succ0, uint32_t *sr0, uint32_t *sr1,
                       uint32_t *loc, uint64_t *inmask, uint32_t *alist, uint32_t const *border_list, uint32_t border_count,
                       Band_layout const *layout, int world, int me) {
    auto *e = static_cast<Emu_band *>(h);
    Tree &tr = e->tr;
    tr.succ0 = succ0;
    tr.sr[0] = sr0;
    tr.sr[1] = sr1;
    tr.loc = loc;
    tr.inmask = inmask;
    tr.alist = alist;
    tr.tc = &e->tc;
    tr.nslots = ntiles_global * SLOTS;
    tr.tile_base = tile_base;
    tr.banded = 1;
    e->alist = alist;
    e->border_list = border_list;
    e->border_count = border_count;
    e->layout = layout;
    e->world = world;
    e->me = me;
    size_t const SM = FLOOD_U64 > WALK_U64 ? FLOOD_U64 : WALK_U64;
    e->sm.assign(static_cast<size_t>(e->nwarps) * SM, 0);
    return 0;
}

static Tree emu_band_local(Emu_band *e) {
    Tree tr = e->tr;
    tr.mark_bits = e->d.game == GAME_DISTANCE ? SQ_MEASURE : 0u;
    tr.alist = e->alist;
    tr.jump_lo = tr.tile_base * SLOTS;
    tr.jump_hi = tr.jump_lo + static_cast<uint32_t>(e->w.g.ntiles) * SLOTS;
    tr.jump_skip = 3u;
    return tr;
}
static Tree emu_band_border(Emu_band *e) {
    Tree tr = e->tr;
    tr.alist = const_cast<uint32_t *>(e->border_list);
    tr.jump_lo = tr.jump_hi = 0;
    tr.jump_skip = 3u;
    return tr;
}
// the rounds of one pointer-jumping launch of csrc/lab.cu (k_tree_jump), one emulated launch per round
static void emu_band_jump(Emu_band *e, Tree const &tr, int kind, uint32_t count) {
    int const nw = e->nwarps;
    long long const nthreads = static_cast<long long>(nw) * 32;
    if (!tree_live(&e->tc)) return;
    e->tc.n_arcs = count;
    std::memset(kind == 0 ? e->tc.rank_more : e->tc.prefix_more, 0, sizeof e->tc.rank_more);
    int rounds = 0;
    while ((1ull << rounds) < static_cast<unsigned long long>(tr.nslots) + 2) rounds++;
    (kind == 0 ? e->tc.rank_buf : e->tc.prefix_buf) = 0;
    for (int r = 0; r < rounds; r++) {
        if (kind == 0) {
            simt::launch(nw, [&](int wid, int lane) { tree_rank_round_body(tr, static_cast<uint32_t>(r), static_cast<long long>(wid) * 32 + lane, nthreads); });
        } else {
            simt::launch(nw, [&](int wid, int lane) { tree_prefix_round_body(tr, static_cast<uint32_t>(r), static_cast<long long>(wid) * 32 + lane, nthreads); });
        }
    }
    uint32_t const buf = kind == 0 ? e->tc.rank_buf : e->tc.prefix_buf;
    simt::launch(nw, [&](int wid, int lane) { tree_jump_settle_body(tr, buf, static_cast<long long>(wid) * 32 + lane, nthreads); });
}

int emu_band_tree_chunk_words(void *h, int stage) {
    auto *e = static_cast<Emu_band *>(h);
    Params p = make_params(e->w, 1);
    return static_cast<int>(border_chunk_words(p, stage));
}

int emu_band_tree_count(void *h, uint64_t *squares, uint64_t *pairs, uint32_t *crossings, int32_t *box) {
    auto *e = static_cast<Emu_band *>(h);
    Params p = make_params(e->w, 1);
    int const nw = e->nwarps;
    e->tc = Tree_ctl{};
    Tree const tr = e->tr;
    simt::launch(nw, [&](int wid, int lane) { tree_count_body(p, tr, wid, nw, lane); });
    e->crossings = e->tc.crossings;
    *squares = e->tc.V;
    *pairs = e->tc.E;
    *crossings = e->tc.crossings;
    box[0] = e->tc.V ? static_cast<int32_t>(0x7FFFFFFFu - e->tc.box_r0_inv) : 1;
    box[1] = e->tc.V ? static_cast<int32_t>(e->tc.box_r1) : 0;
    box[2] = e->tc.V ? static_cast<int32_t>(0x7FFFFFFFu - e->tc.box_c0_inv) : 1;
    box[3] = e->tc.V ? static_cast<int32_t>(e->tc.box_c1) : 0;
    return 0;
}

int emu_band_room(void *h, int r0, int r1, int c0, int c1, int src_r, int src_c) {
    auto *e = static_cast<Emu_band *>(h);
    Params p = make_params(e->w, 1);
    int const nw = e->nwarps;
    long long const nthreads = static_cast<long long>(nw) * 32;
    Tree const tr = e->tr;
    room_set_body(&e->w.ctl, r0, r1, c0, c1, src_r, src_c);
    if (e->d.game == GAME_DISTANCE) {
        simt::launch(nw, [&](int wid, int lane) { room_distance_body(p, e->w.grid_ptr, &e->w.res.settled, static_cast<long long>(wid) * 32 + lane, nthreads); });
    }
    simt::launch(nw, [&](int wid, int lane) { tree_gate_body(p, tr, static_cast<long long>(wid) * 32 + lane, nthreads); });
    e->tree_last = e->tc;
    return 0;
}

int emu_band_tree_rank_local(void *h, uint64_t squares_global, uint64_t pairs_global, uint32_t *chunk1) {
    auto *e = static_cast<Emu_band *>(h);
    Params p = make_params(e->w, 1);
    int const nw = e->nwarps;
    long long const nthreads = static_cast<long long>(nw) * 32;
    Tree const tr = emu_band_local(e);
    e->tc.V = squares_global;
    e->tc.E = pairs_global;
    tree_init_body(p, tr);
    size_t const SM = FLOOD_U64 > WALK_U64 ? FLOOD_U64 : WALK_U64;
    simt::launch(nw, [&](int wid, int lane) { tree_walk_body(p, tr, lane, e->sm.data() + static_cast<size_t>(wid) * SM); }, e->sched);
    emu_band_jump(e, tr, 0, e->crossings);
    simt::launch(nw, [&](int wid, int lane) { tree_border_export_body(p, tr, 1, chunk1, static_cast<long long>(wid) * 32 + lane, nthreads); });
    return 0;
}

int emu_band_tree_flood(void *h, uint32_t const *all1, int root_rank, uint32_t *chunk2) {
    auto *e = static_cast<Emu_band *>(h);
    Params p = make_params(e->w, 1);
    int const nw = e->nwarps;
    long long const nthreads = static_cast<long long>(nw) * 32;
    Tree const local = emu_band_local(e), border = emu_band_border(e);
    simt::launch(nw, [&](int wid, int lane) { tree_border_import_body(p, local, 1, all1, e->layout, e->world, e->me, root_rank, static_cast<long long>(wid) * 32 + lane, nthreads); });
    emu_band_jump(e, border, 0, e->border_count);
    e->tc.n_arcs = e->crossings;
    simt::launch(nw, [&](int wid, int lane) { tree_finish_jump_body(p, local, static_cast<long long>(wid) * 32 + lane, nthreads); });
    simt::launch(nw, [&](int wid, int lane) { tree_orient_body(p, local, wid, nw, lane); });
    size_t const SM = FLOOD_U64 > WALK_U64 ? FLOOD_U64 : WALK_U64;
    simt::launch(nw, [&](int wid, int lane) { tree_flood_body(p, local, lane, e->sm.data() + static_cast<size_t>(wid) * SM); }, e->sched);
    simt::launch(nw, [&](int wid, int lane) { tree_border_export_body(p, local, 2, chunk2, static_cast<long long>(wid) * 32 + lane, nthreads); });
    return 0;
}

int emu_band_tree_prefix_local(void *h, uint32_t const *all2, uint32_t *chunk3) {
    auto *e = static_cast<Emu_band *>(h);
    Params p = make_params(e->w, 1);
    int const nw = e->nwarps;
    long long const nthreads = static_cast<long long>(nw) * 32;
    Tree const local = emu_band_local(e);
    simt::launch(nw, [&](int wid, int lane) { tree_border_import_body(p, local, 2, all2, e->layout, e->world, e->me, 0, static_cast<long long>(wid) * 32 + lane, nthreads); });
    e->tc.n_arcs = e->crossings;
    simt::launch(nw, [&](int wid, int lane) { tree_parent_body(p, local, static_cast<long long>(wid) * 32 + lane, nthreads); });
    emu_band_jump(e, local, 1, e->crossings);
    simt::launch(nw, [&](int wid, int lane) { tree_border_export_body(p, local, 3, chunk3, static_cast<long long>(wid) * 32 + lane, nthreads); });
    return 0;
}

int emu_band_tree_levels(void *h, uint32_t const *all3, int32_t *used) {
    auto *e = static_cast<Emu_band *>(h);
    Params p = make_params(e->w, 1);
    int const nw = e->nwarps;
    long long const nthreads = static_cast<long long>(nw) * 32;
    Tree const local = emu_band_local(e), border = emu_band_border(e);
    simt::launch(nw, [&](int wid, int lane) { tree_border_import_body(p, local, 3, all3, e->layout, e->world, e->me, 0, static_cast<long long>(wid) * 32 + lane, nthreads); });
    emu_band_jump(e, border, 1, e->border_count);
    e->tc.n_arcs = e->crossings;
    simt::launch(nw, [&](int wid, int lane) { tree_finish_jump_body(p, local, static_cast<long long>(wid) * 32 + lane, nthreads); });
    simt::launch(nw, [&](int wid, int lane) { tree_fixup_body(p, local, e->w.grid_ptr, wid, nw, lane); });
    simt::launch(nw, [&](int wid, int lane) { tree_gate_body(p, local, static_cast<long long>(wid) * 32 + lane, nthreads); });
    e->tree_last = e->tc;
    *used = static_cast<int32_t>(e->tree_last.used);
    return 0;
}

// ---- the lattice road across bands: emu_band_lat_* mirror lab_band_lat_* (csrc/lab.cu) ----
int emu_band_lat_super_rows(void) { return LST; }
int emu_band_lat_words(int tiles_x, int super_rows_global, uint64_t *border_entries, uint64_t *weight_words, uint64_t *chunk_words) {
    unsigned long long const sup_x = (static_cast<unsigned long long>(tiles_x) + LST - 1) / LST;
    unsigned long long const nslots = static_cast<unsigned long long>(super_rows_global) * LST * tiles_x * LSLOTS;
    *border_entries = static_cast<unsigned long long>(super_rows_global) * sup_x * LB;
    *weight_words = nslots + 2 + LAT_TAIL + L_CHUNK;
    *chunk_words = nslots / L_CHUNK + 2;
    return 0;
}
int emu_band_lat_bind(void *h, int super_row_base, int super_rows_band, int super_rows_global, uint64_t *border_list, uint32_t *weights, uint32_t *chunks) {
    auto *e = static_cast<Emu_band *>(h);
    Geom const &g = e->w.g;
    if (g.tiles_y > super_rows_band * LST) return -1;
    size_t const nt = static_cast<size_t>(g.ntiles);
    e->l_mask.assign(nt * 96, 0xDEADBEEFu);
    e->l_xmask.assign(nt * 4, 0);
    e->l_outslot.assign(nt * LSLOTS, 0xAB);
    e->l_label.assign(nt * LSLOTS, 0);
    e->l_cmap.assign(nt * 1024, 0xDEAD);
    e->l_xd.assign(nt * LSLOTS, 0xDEADBEEFDEADBEEFull);
    e->l_R.assign(nt * LSLOTS, 0xDEADBEEFu);
    e->l_dplane.assign(nt * L_DEPTH_PLANES * 32, 0);
    e->l_emask.assign(nt * 4, 0);
    e->l_sm.assign(static_cast<size_t>(e->nwarps) * LSUP_SLOTS, 0);
    Lattice &lt = e->lat;
    lt = Lattice{};
    lt.mask = e->l_mask.data(); lt.xmask = e->l_xmask.data(); lt.outslot = e->l_outslot.data(); lt.label = e->l_label.data(); lt.cmap = e->l_cmap.data();
    lt.xd = e->l_xd.data(); lt.R = e->l_R.data(); lt.dplane = e->l_dplane.data();
    lt.emask = e->l_emask.data();
    lt.tc = &e->lat_tc;
    lt.sup_x = (g.tiles_x + LST - 1) / LST;
    lt.sup_y = super_rows_band;
    lt.nsuper = lt.sup_y * lt.sup_x;
    lt.nslots = static_cast<uint32_t>(nt * LSLOTS);
    lt.tile_base = static_cast<uint32_t>(super_row_base) * LST * static_cast<uint32_t>(g.tiles_x);
    lt.sup_base = static_cast<uint32_t>(super_row_base) * static_cast<uint32_t>(lt.sup_x);
    lt.nsuper_global = super_rows_global * lt.sup_x;
    lt.nslots_global = static_cast<uint32_t>(static_cast<unsigned long long>(super_rows_global) * LST * g.tiles_x * LSLOTS);
    lt.gb = border_list;
    lt.wsum = weights;
    lt.chunk = chunks;
    lt.peers = 1;
    lt.gb_peer[0] = lt.gb;
    lt.wsum_peer[0] = lt.wsum;
    return 0;
}
int emu_band_lat_set_peers(void *h, int n, void *const *border_lists, void *const *weights) {
    auto *e = static_cast<Emu_band *>(h);
    e->lat.peers = n;
    for (int q = 0; q < n; q++) {
        e->lat.gb_peer[q] = static_cast<uint64_t *>(border_lists[q]);
        e->lat.wsum_peer[q] = static_cast<uint32_t *>(weights[q]);
    }
    return 0;
}
int emu_band_lat_link(void *h, uint32_t *header) {
    auto *e = static_cast<Emu_band *>(h);
    Params p = make_params(e->w, 1);
    int const nw = e->nwarps;
    Lattice const lt = e->lat;
    e->lat_tc = Tree_ctl{};
    Tree tr{};
    tr.tc = &e->lat_tc;
    tr.banded = 1;
    simt::launch(nw, [&](int wid, int lane) { tree_count_body(p, tr, wid, nw, lane); });
    lat_band_open_body(lt);
    simt::launch(nw, [&](int wid, int lane) { lat_link_body(p, lt, lane, e->l_sm.data() + static_cast<size_t>(wid) * LSUP_SLOTS); }, e->sched);
    simt::launch(nw, [&](int wid, int lane) { lat_rank_local_body(p, lt, lane, e->l_sm.data() + static_cast<size_t>(wid) * LSUP_SLOTS); }, e->sched);
    lat_band_header_body(lt, header);
    return 0;
}
int emu_band_lat_flood(void *h, uint32_t const *all_headers, int world, int zero_weights) {
    auto *e = static_cast<Emu_band *>(h);
    Params p = make_params(e->w, 1);
    int const nw = e->nwarps;
    long long const nth = static_cast<long long>(nw) * 32;
    Lattice const lt = e->lat;
    if (zero_weights) std::memset(lt.wsum, 0, (static_cast<size_t>(lt.nslots_global) + 2 + LAT_TAIL) * sizeof(uint32_t));
    lat_band_check_body(lt, all_headers, world);
    simt::launch(nw, [&](int wid, int lane) { lat_rank_global_body(lt, static_cast<long long>(wid) * 32 + lane, nth); }, e->sched);
    simt::launch(nw, [&](int wid, int lane) { lat_rank_expand_body(lt, static_cast<long long>(wid) * 32 + lane, nth); });
    simt::launch(nw, [&](int wid, int lane) { lat_flood_body(p, lt, lane); }, e->sched);
    lat_band_tail_body(lt);
    return 0;
}
int emu_band_lat_levels(void *h, int32_t *used) {
    auto *e = static_cast<Emu_band *>(h);
    Params p = make_params(e->w, 1);
    int const nw = e->nwarps;
    long long const nth = static_cast<long long>(nw) * 32;
    Lattice const lt = e->lat;
    lat_band_settled_body(lt);
    simt::launch(nw, [&](int wid, int lane) { lat_scan_chunks_body(lt, wid, nw, lane); });
    simt::launch(1, [&](int, int lane) { lat_scan_tops_body(lt, lane); });
    simt::launch(nw, [&](int wid, int lane) { lat_levels_body<LAT_LEVELS>(p, lt, lane, e->l_sm.data() + static_cast<size_t>(wid) * LSUP_SLOTS, nullptr, nullptr); }, e->sched);
    lat_band_close_body(lt);
    Tree tr{};
    tr.tc = &e->lat_tc;
    tr.banded = 1;
    simt::launch(nw, [&](int wid, int lane) { tree_gate_body(p, tr, static_cast<long long>(wid) * 32 + lane, nth); });
    e->tree_last = e->lat_tc;
    *used = e->lat_tc.used == 3u ? 1 : 0;
    return 0;
}

// mirrors lab_band_lat_path_rank / _mark (csrc/lab.cu)
int emu_band_lat_path_rank(void *h, int r, int c, uint32_t *rank, uint32_t *level) {
    auto *e = static_cast<Emu_band *>(h);
    Params p = make_params(e->w, 1);
    Lattice const lt = e->lat;
    uint32_t out = 0;
    simt::launch(1, [&](int, int lane) {
        uint32_t const v = lat_path_rank_body(p, lt, r, c, lane);
        if (lane == 0) out = v;
    });
    *rank = out;
    *level = e->w.dist[0][static_cast<size_t>(r) * e->w.g.cols + c];
    return 0;
}
int emu_band_lat_path_mark(void *h, uint32_t rank, uint32_t level, int fin_r, int fin_c, int start_r, int start_c, uint32_t repaint_bits, int32_t *path,
                           uint64_t path_cap, uint64_t *segments) {
    auto *e = static_cast<Emu_band *>(h);
    Params p = make_params(e->w, 1);
    Lattice const lt = e->lat;
    int const nw = e->nwarps;
    uint32_t *seg = reinterpret_cast<uint32_t *>(e->l_xd.data());
    uint32_t const seg_cap = lt.nslots;
    e->lat_tc.lat_path_on = 0u;
    for (int q = 0; q < 4; q++) e->lat_tc.lat_path_nseg[q] = 0u;
    lat_path_setup_body(p, lt, 0, rank, level, fin_r, fin_c, start_r, start_c, repaint_bits, e->w.grid_ptr, path, path_cap, seg);
    simt::launch(nw, [&](int wid, int lane) { lat_path_find_body(p, lt, seg, seg_cap, wid, nw, lane); });
    std::vector<uint32_t> wsm(static_cast<size_t>(nw) * WALK_SM_U32);
    simt::launch(nw, [&](int wid, int lane) {
        lat_path_walk_body(p, lt, &e->w.res, repaint_bits, e->w.grid_ptr, path, 0, path_cap, seg, seg_cap, wid, nw, lane, wsm.data() + static_cast<size_t>(wid) * WALK_SM_U32);
    });
    if (segments) *segments = e->lat_tc.lat_path_nseg[0];
    return 0;
}

// game: 1 hunt, 2 gather, 3 corners.  starts: 4 (r,c) (hunt/gather use the first).  finishes: 4 (r,c)
// (hunt/corners use the first).  paths_out: 4 * max_path (r,c).
int emu_solver(int game, int rows, int cols, uint16_t *grid, int32_t const *starts, int32_t const *finishes,
               int32_t *winner, int32_t *claimed_by, int32_t *paths_out, long long max_path, long long *path_len,
               uint64_t *painted, int nwarps, unsigned sched_seed, uint64_t *stats) {
    Workspace w;
    Run_desc d;
    fill_desc(d);
    d.game = game;
    int32_t frc[8] = {0};
    if (game == GAME_HUNT) {
        d.nplanes = 1;
        d.src_r[0] = starts[0];
        d.src_c[0] = starts[1];
        d.n_targets = 1;
        d.tgt_r[0] = finishes[0];
        d.tgt_c[0] = finishes[1];
        d.bound_mode = BOUND_MIN_ANY;
        frc[0] = finishes[0];
        frc[1] = finishes[1];
    } else if (game == GAME_GATHER) {
        d.nplanes = 1;
        d.src_r[0] = starts[0];
        d.src_c[0] = starts[1];
        d.n_targets = 4;
        for (int j = 0; j < 4; j++) {
            d.tgt_r[j] = finishes[2 * j];
            d.tgt_c[j] = finishes[2 * j + 1];
            frc[2 * j] = finishes[2 * j];
            frc[2 * j + 1] = finishes[2 * j + 1];
        }
        d.bound_mode = BOUND_MAX_ALL;
    } else {
        d.nplanes = 4;
        d.n_targets = 4;
        for (int k = 0; k < 4; k++) {
            d.src_r[k] = starts[2 * k];
            d.src_c[k] = starts[2 * k + 1];
            d.tgt_plane[k] = k;
            d.tgt_r[k] = finishes[0];
            d.tgt_c[k] = finishes[1];
        }
        d.bound_mode = BOUND_MIN_ANY;
        frc[0] = finishes[0];
        frc[1] = finishes[1];
    }
    g_hunt_out = ((game == GAME_HUNT || game == GAME_GATHER) && max_path > 0) ? paths_out : nullptr; // (gather: four buffers, max_path pairs apart)
    g_hunt_out_cap = g_hunt_out ? static_cast<unsigned long long>(max_path) : 0ull;
    int32_t front[8];
    for (auto &x : front) x = -1;
    g_front_rc = front;
    g_skip_field = game == GAME_GATHER && max_path == 0 && !g_opt_solver_levels; // (as solver_common: gather, no paths, the option off)
    int rc = search(w, grid, rows, cols, d, SQ_PATH, SQ_PATH, nwarps, sched_seed);
    g_hunt_out = nullptr;
    g_front_rc = nullptr;
    g_skip_field = 0;
    if (rc) return rc;
    Params p = make_params(w, d.nplanes);
    resolve_body(p, d, &w.res, frc);
    long long const nthreads = static_cast<long long>(nwarps) * 32;
    simt::launch(nwarps, [&](int wid, int lane) { paint_body(p, grid, &w.res, static_cast<long long>(wid) * 32 + lane, nthreads); });
    auto walk = [&](int slot, int32_t *out, unsigned long long max_out, int first_only) {
        int plane = 0, fj = 0;
        uint32_t repaint = 0;
        if (game == GAME_HUNT && w.ctl.walked) { // (k_walk: lat_path_* has marked and written the path)
            w.res.path_len[slot] = w.res.game_level[0];
            return;
        }
        if (game == GAME_HUNT) {
            if (w.res.winner < 0) return;
            repaint = (1u << w.res.winner) << SQ_PAINT_SHIFT;
        } else if (game == GAME_CORNERS) {
            if (w.res.winner < 0) return;
            plane = w.res.winner;
        } else {
            if (first_only && w.ctl.fronts_done) return; // (k_walk: lat_fronts_body has found path.front())
            if (!first_only && w.ctl.walked) {           // (k_walk: lat_path_* has written every colour's path)
                w.res.path_len[slot] = w.res.finish_of[slot] >= 0 ? w.res.game_level[slot] : 0;
                return;
            }
            fj = w.res.finish_of[slot];
            if (fj < 0) return;
        }
        std::vector<uint32_t> wsm(WALK_SM_U32);
        simt::launch(1, [&](int, int lane) {
            walk_body(p, plane, frc[2 * fj], frc[2 * fj + 1], grid, out, max_out, repaint, first_only, &w.res, slot, lane, wsm.data());
        });
    };
    if (game == GAME_GATHER) {
        for (int k = 0; k < 4; k++) walk(k, front + 2 * k, 1, 1);
        gather_fixup_body(w.g, grid, &w.res, frc, front);
        for (int k = 0; k < 4 && max_path > 0; k++) walk(k, paths_out + static_cast<size_t>(k) * max_path * 2, static_cast<unsigned long long>(max_path), 0);
    } else {
        walk(0, paths_out, static_cast<unsigned long long>(max_path), 0);
    }
    *winner = w.res.winner;
    for (int k = 0; k < 4; k++) {
        claimed_by[k] = w.res.claimed_by[k];
        path_len[k] = static_cast<long long>(w.res.path_len[k]);
    }
    *painted = w.res.settled;
    stats_out(w, stats);
    return 0;
}

} // extern "C"
