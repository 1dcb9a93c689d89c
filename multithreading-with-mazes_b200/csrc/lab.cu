// lab.cu -- __global__ wrappers of the kernel bodies and the C ABI of include/labyrinth_b200.h.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a (see ../Makefile).  No torch, no CPU
// fallback: every compute entry point needs a CUDA device and says so loudly otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/labyrinth_b200.h"
#include "device/engine.cuh"
#include "device/post.cuh"
#include "device/tree.cuh"
#include "device/lattice.cuh"
#include "host/builders.hh"
#include "host/mazefile.hh"
#include "host/points.hh"

using namespace lab;

// ------------------------------------------------------------------------------------- kernels
namespace {

__device__ __forceinline__ long long gwarp_id() {
    return (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
}
__device__ __forceinline__ long long nwarps_total() { return (static_cast<long long>(gridDim.x) * blockDim.x) >> 5; }
__device__ __forceinline__ int lane_id() { return static_cast<int>(threadIdx.x & 31); }

__global__ void __launch_bounds__(256) k_pack(Geom g, uint16_t *grid, uint64_t *path, uint16_t mask, uint16_t value, uint16_t mark) {
    pack_body(g, grid, path, mask, value, mark, gwarp_id(), nwarps_total(), lane_id());
}

__global__ void __launch_bounds__(256) k_cross(Geom g, uint64_t const *path, uint64_t *cross, uint64_t const *ghost_n,
                                               uint64_t const *ghost_s) {
    cross_body(g, path, cross, ghost_n, ghost_s, gwarp_id(), nwarps_total(), lane_id());
}

__global__ void k_force_sources(Geom g, uint64_t *path, Run_desc d, uint16_t *grid) { force_sources_body(g, path, d, grid); }

__global__ void k_init_run(Params p, Run_desc d, unsigned long long budget_ns) { init_run_body(p, d, budget_ns); }

// The persistent search kernel: every warp is a worker until the search is done.
__global__ void __launch_bounds__(128, 3) k_bfs_tiles(Params p) {
    __shared__ __align__(16) uint64_t stage[4][STAGE_U64];
    bfs_worker_body(p, lane_id(), stage[threadIdx.x >> 5]);
}

// ---- spanning-tree pipeline (device/tree.cuh) ----
__global__ void __launch_bounds__(256) k_tree_count(Params p, Tree tr) { tree_count_body(p, tr, gwarp_id(), nwarps_total(), lane_id()); }
__global__ void k_tree_init(Params p, Tree tr) { tree_init_body(p, tr); }
__global__ void __launch_bounds__(128) k_tree_walk(Params p, Tree tr) {
    __shared__ __align__(16) uint64_t sm[4][WALK_U64];
    tree_walk_body(p, tr, lane_id(), sm[threadIdx.x >> 5]);
}
// All blocks of a cooperative launch meet here (they are co-resident: cudaLaunchCooperativeKernel).
// `target` counts the arrivals expected so far; the counter only ever grows.
__device__ __forceinline__ void grid_barrier(uint32_t *counter, uint32_t &target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        while (*reinterpret_cast<uint32_t volatile *>(counter) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}
// Pointer jumping, all rounds in one persistent launch: a round costs a grid barrier, not a launch.
// The rounds are tree_rank_round_body / tree_prefix_round_body (device/tree.cuh, which tests/simt runs
// one launch per round) with one refinement: a thread keeps its first JUMP_K list entries (slot,
// pointer, value) in registers from round to round, so a round is ONE dependent pair of L2 loads.
// KIND 0: rank  (pointer = successor, stop at END; then 3 orient, which only needs the ranks)
// KIND 1: prefix (after 5's parent pass; pointer = parent entry, stop at ROOT)
constexpr int JUMP_K = 4;
template <int KIND> __global__ void __launch_bounds__(512) k_tree_jump(Params p, Tree tr, uint32_t max_rounds) {
    Tree_ctl *tc = tr.tc;
    if (!tree_live(tc)) {
        return; // every block sees the same flags: nobody waits at a barrier
    }
    long long const gthread = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    long long const nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
    uint32_t target = 0;
    uint32_t *const barrier = &tc->barrier[KIND];
    if (KIND == 1 && !(tr.jump_skip & 1u)) {
        tree_parent_body(p, tr, gthread, nthreads);
        grid_barrier(barrier, target);
        if (*reinterpret_cast<uint32_t volatile *>(&tc->ok) == 0u || *reinterpret_cast<uint32_t volatile *>(&tc->bad_weight) != 0u) {
            return;
        }
    }
    uint32_t const ROOT = tr.nslots, STOP = KIND == 0 ? tr.nslots + 1u : tr.nslots;
    uint32_t *const more = KIND == 0 ? tc->rank_more : tc->prefix_more;
    long long const n = static_cast<long long>(tc->n_arcs) + 1;
    uint32_t a[JUMP_K], s[JUMP_K], r[JUMP_K];
    bool own[JUMP_K];
#pragma unroll
    for (int k = 0; k < JUMP_K; k++) {
        long long const i = gthread + k * nthreads;
        own[k] = i < n;
        a[k] = own[k] ? (i + 1 < n ? tr.alist[i] : ROOT) : 0u;
    }
#pragma unroll
    for (int k = 0; k < JUMP_K; k++) { // (pointer, value) pairs: one 8-byte load each
        uint2 const v = own[k] ? __ldcg(reinterpret_cast<uint2 const *>(tr.sr[0]) + a[k]) : make_uint2(STOP, 0u);
        s[k] = v.x;
        r[k] = v.y;
    }
    bool const spill = n > JUMP_K * nthreads; // more entries than registers: the rest goes the plain way
    uint32_t last = 0; // the buffer the rounds end in
    for (uint32_t round = 0; round < max_rounds; round++) {
        bool const odd = round & 1u;
        uint2 const *in = reinterpret_cast<uint2 const *>(odd ? tr.sr[1] : tr.sr[0]);
        uint2 *out = reinterpret_cast<uint2 *>(odd ? tr.sr[0] : tr.sr[1]);
        // three hops per round through the buffer of the last round: a pointer's span grows fourfold.
        // (tree_follow: not past the end of the list, and -- row bands -- not out of the band's window)
        uint32_t s2[JUMP_K], r2[JUMP_K];
#pragma unroll
        for (int k = 0; k < JUMP_K; k++) {
            s2[k] = s[k];
            r2[k] = 0u;
        }
#pragma unroll
        for (int hop = 0; hop < 3; hop++) {
            uint32_t sn[JUMP_K], rn[JUMP_K];
#pragma unroll
            for (int k = 0; k < JUMP_K; k++) {
                bool const go = tree_follow(tr, s2[k]);
                uint2 const v = go ? __ldcg(in + s2[k]) : make_uint2(s2[k], 0u);
                sn[k] = v.x;
                rn[k] = v.y;
            }
#pragma unroll
            for (int k = 0; k < JUMP_K; k++) {
                s2[k] = sn[k];
                r2[k] += rn[k];
            }
        }
        bool m = false;
#pragma unroll
        for (int k = 0; k < JUMP_K; k++) {
            if (own[k]) {
                if (tree_follow(tr, s[k])) {
                    r[k] += r2[k];
                    s[k] = s2[k];
                    m |= tree_follow(tr, s[k]);
                }
                out[a[k]] = make_uint2(s[k], r[k]);
            }
        }
        if (spill) {
            for (long long i = gthread + JUMP_K * nthreads; i < n; i += nthreads) {
                uint32_t const ai = i + 1 < n ? tr.alist[i] : ROOT;
                uint2 const me = __ldcg(in + ai);
                if (!tree_follow(tr, me.x)) {
                    out[ai] = me;
                } else {
                    uint2 const nx = __ldcg(in + me.x);
                    out[ai] = make_uint2(nx.x, me.y + nx.y);
                    m |= tree_follow(tr, nx.x);
                }
            }
        }
        if (__syncthreads_or(m) && threadIdx.x == 0) {
            *reinterpret_cast<uint32_t volatile *>(&more[round]) = 1u;
        }
        last = odd ? 0u : 1u;
        grid_barrier(barrier, target);
        if (*reinterpret_cast<uint32_t volatile *>(&more[round]) == 0u) {
            break;
        }
    }
    // Everything downstream reads buffer 0 (tree_jump_settle_body): the registers still hold this thread's entries.
    if (last == 1u) {
#pragma unroll
        for (int k = 0; k < JUMP_K; k++) {
            if (own[k]) {
                reinterpret_cast<uint2 *>(tr.sr[0])[a[k]] = make_uint2(s[k], r[k]);
            }
        }
        if (spill) {
            for (long long i = gthread + JUMP_K * nthreads; i < n; i += nthreads) {
                uint32_t const ai = i + 1 < n ? tr.alist[i] : ROOT;
                reinterpret_cast<uint2 *>(tr.sr[0])[ai] = __ldcg(reinterpret_cast<uint2 const *>(tr.sr[1]) + ai);
            }
        }
    }
    if (KIND == 0 && !(tr.jump_skip & 2u)) {
        grid_barrier(barrier, target); // (orient looks at other threads' entries -- twins -- in buffer 0)
        tree_orient_body(p, tr, gthread >> 5, nthreads >> 5, static_cast<int>(threadIdx.x & 31));
    }
}
__global__ void __launch_bounds__(256) k_tree_border_export(Params p, Tree tr, int stage, uint32_t *chunk) {
    tree_border_export_body(p, tr, stage, chunk, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x, static_cast<long long>(gridDim.x) * blockDim.x);
}
__global__ void __launch_bounds__(256) k_tree_border_import(Params p, Tree tr, int stage, uint32_t const *all, Band_layout const *lay, int world, int me,
                                                            int root_rank) {
    tree_border_import_body(p, tr, stage, all, lay, world, me, root_rank, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x,
                            static_cast<long long>(gridDim.x) * blockDim.x);
}
__global__ void __launch_bounds__(256) k_tree_finish_jump(Params p, Tree tr) {
    tree_finish_jump_body(p, tr, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x, static_cast<long long>(gridDim.x) * blockDim.x);
}
// ---- the lattice road (device/lattice.cuh) ----
__global__ void __launch_bounds__(128, 8) k_lat_link(Params p, Lattice lt) {
    __shared__ uint32_t sm[4][L_LINK_SM];
    lat_link_body(p, lt, lane_id(), sm[threadIdx.x >> 5]);
}
constexpr int LAT_RANK_WARPS = 4; // warps per block of the local ranking (LSUP_SLOTS words of shared memory each)
__global__ void __launch_bounds__(32 * LAT_RANK_WARPS) k_lat_rank_local(Params p, Lattice lt) {
    extern __shared__ uint32_t lat_dsm[];
    lat_rank_local_body(p, lt, lane_id(), lat_dsm + (threadIdx.x >> 5) * LSUP_SLOTS);
}
__global__ void __launch_bounds__(256) k_lat_rank_global(Lattice lt) {
    lat_rank_global_body(lt, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x, static_cast<long long>(gridDim.x) * blockDim.x);
}
__global__ void __launch_bounds__(256) k_lat_rank_expand(Lattice lt) {
    lat_rank_expand_body(lt, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x, static_cast<long long>(gridDim.x) * blockDim.x);
}
__global__ void __launch_bounds__(128) k_lat_flood(Params p, Lattice lt) { lat_flood_body(p, lt, lane_id()); }
__global__ void __launch_bounds__(256) k_lat_scan_chunks(Lattice lt) { lat_scan_chunks_body(lt, gwarp_id(), nwarps_total(), lane_id()); }
__global__ void k_lat_scan_tops(Lattice lt) { lat_scan_tops_body(lt, lane_id()); }
template <int MODE> __global__ void __launch_bounds__(128, (MODE == LAT_PAINT || MODE == LAT_PAINT_ONLY) ? 5 : 8) k_lat_levels(Params p, Lattice lt, Result *res, uint16_t *grid) {
    __shared__ uint32_t sm[4][L_LINK_SM];
    lat_levels_body<MODE>(p, lt, lane_id(), sm[threadIdx.x >> 5], res, grid);
}
// hunt (the winner's path) and gather with paths (every colour's) on the lattice road: paths from the tour's ranks (lat_path_*),
// every segment walked on its own
__global__ void __launch_bounds__(128) k_lat_path_begin(Params p, Lattice lt, int game, Result *res, int32_t const *finish_rc, uint16_t *grid, int32_t *out,
                                                        unsigned long long out_stride, unsigned long long max_out, uint32_t *seg, uint32_t seg_cap) {
    Tree_ctl *tc = lt.tc;
    if (threadIdx.x == 0) {
        tc->lat_path_on = 0u;
        for (int q = 0; q < 4; q++) tc->lat_path_nseg[q] = 0u;
    }
    __syncthreads();
    lat_path_begin_body(p, lt, game, res, finish_rc, grid, out, out_stride, max_out, seg, seg_cap, static_cast<int>(threadIdx.x >> 5), lane_id());
    __syncthreads();
    if (threadIdx.x == 0 && lat_live(tc) && p.ctl->painted && !p.ctl->fronts_done && !p.ctl->walk_fallback && (game != GAME_HUNT || res->winner >= 0)) {
        p.ctl->walked = 1u; // (k_walk: nothing left to do)
    }
}
__global__ void __launch_bounds__(256) k_lat_path_find(Params p, Lattice lt, uint32_t *seg, uint32_t seg_cap) {
    lat_path_find_body(p, lt, seg, seg_cap, gwarp_id(), nwarps_total(), lane_id());
}
// repaint_bits == 0xFFFFFFFF: the winner's colour (Result::winner, known on the device only)
__global__ void k_lat_path_walk(Params p, Lattice lt, Result *res, uint32_t repaint_bits, uint16_t *grid, int32_t *out, unsigned long long out_stride,
                                unsigned long long max_out, uint32_t const *seg, uint32_t seg_cap) {
    __shared__ uint32_t sm[WALK_SM_U32]; // one warp per block
    if (!lt.tc->lat_path_on) {
        return;
    }
    uint32_t const repaint = repaint_bits == 0xFFFFFFFFu ? (1u << res->winner) << SQ_PAINT_SHIFT : repaint_bits;
    lat_path_walk_body(p, lt, res, repaint, grid, out, out_stride, max_out, seg, seg_cap, blockIdx.x, gridDim.x, lane_id(), sm);
}
// row bands (lab_band_lat_path_*)
__global__ void k_lat_path_rank(Params p, Lattice lt, int r, int c, uint32_t *out2) {
    uint32_t const rank = lat_path_rank_body(p, lt, r, c, lane_id());
    if (lane_id() == 0) {
        out2[0] = rank;
        out2[1] = p.plane[0].dist[static_cast<long long>(r) * p.g.cols + c];
    }
}
__global__ void k_lat_path_setup(Params p, Lattice lt, uint32_t rank, uint32_t level, int fr, int fc, int sr, int sc, uint32_t repaint, uint16_t *grid,
                                 int32_t *out, unsigned long long max_out, uint32_t *seg) {
    lt.tc->lat_path_on = 0u;
    for (int q = 0; q < 4; q++) lt.tc->lat_path_nseg[q] = 0u;
    lat_path_setup_body(p, lt, 0, rank, level, fr, fc, sr, sc, repaint, grid, out, max_out, seg);
}
// hunt / gather on the lattice road: the finishes' levels and the paint rules BEFORE the field is written (lat_targets_body)
// front_rc != null (gather, no paths, the field not wanted): every colour's path.front() by probing too (lat_fronts_body)
__global__ void __launch_bounds__(128) k_lat_targets(Params p, Lattice lt, Run_desc d, Result *res, int32_t const *finish_rc, int32_t *front_rc) {
    __shared__ uint32_t sm[4][L_LINK_SM];
    lat_targets_body(p, lt, static_cast<int>(threadIdx.x >> 5), lane_id(), sm[threadIdx.x >> 5]);
    __syncthreads();
    if (threadIdx.x == 0 && lat_live(lt.tc)) {
        resolve_body(p, d, res, finish_rc);
        p.ctl->painted = 1u;
        if (front_rc) p.ctl->fronts_done = 1u;
    }
    __syncthreads();
    if (front_rc) {
        lat_fronts_body(p, lt, res, finish_rc, front_rc, static_cast<int>(threadIdx.x >> 5), lane_id(), sm[threadIdx.x >> 5]);
    }
}
__global__ void k_lat_band_open(Lattice lt) { lat_band_open_body(lt); }
__global__ void k_lat_band_header(Lattice lt, uint32_t *hdr) { lat_band_header_body(lt, hdr); }
__global__ void k_lat_band_check(Lattice lt, uint32_t const *all_hdr, int world) { lat_band_check_body(lt, all_hdr, world); }
__global__ void k_lat_band_tail(Lattice lt) { lat_band_tail_body(lt); }
__global__ void k_lat_band_settled(Lattice lt) { lat_band_settled_body(lt); }
__global__ void k_lat_band_close(Lattice lt) { lat_band_close_body(lt); }
__global__ void __launch_bounds__(256) k_tree_orient(Params p, Tree tr) { tree_orient_body(p, tr, gwarp_id(), nwarps_total(), lane_id()); }
__global__ void __launch_bounds__(128) k_tree_flood(Params p, Tree tr) {
    extern __shared__ __align__(16) uint64_t flood_sm[];
    tree_flood_body(p, tr, lane_id(), flood_sm + static_cast<size_t>(threadIdx.x >> 5) * FLOOD_U64);
}
__global__ void __launch_bounds__(256) k_tree_fixup(Params p, Tree tr, uint16_t *grid) { tree_fixup_body(p, tr, grid, gwarp_id(), nwarps_total(), lane_id()); }
// the open room (device/room.cuh): the whole distance map in one streaming pass
__global__ void __launch_bounds__(256) k_room_distance(Params p, uint16_t *grid, Result *res) {
    room_distance_body(p, grid, &res->settled, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x,
                       static_cast<long long>(gridDim.x) * blockDim.x);
}
__global__ void k_room_set(Control *ctl, int r0, int r1, int c0, int c1, int sr, int sc) { room_set_body(ctl, r0, r1, c0, c1, sr, sc); }
__global__ void __launch_bounds__(256) k_tree_gate(Params p, Tree tr) {
    tree_gate_body(p, tr, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x, static_cast<long long>(gridDim.x) * blockDim.x);
}

__global__ void __launch_bounds__(256) k_measure(Geom g, uint16_t *grid, uint64_t const *visited, uint64_t const *path, Control const *ctl, Result *res) {
    measure_body(g, grid, visited, path, ctl, res, gwarp_id(), nwarps_total(), lane_id());
}

__global__ void __launch_bounds__(256) k_heatmap(Geom g, uint32_t const *dist, Result const *res, int channel, uint32_t *pixels) {
    heatmap_body(g, dist, res, channel, pixels, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x,
                 static_cast<long long>(gridDim.x) * blockDim.x);
}

__global__ void __launch_bounds__(256) k_runs(Geom g, uint32_t const *dist, uint32_t *runs, Result *res) {
    // (max_level is a 64-bit field holding a value below 2^32: its low word is what atomicMax sees)
    runs_body(g, dist, runs, reinterpret_cast<uint32_t *>(&res->max_level), static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x,
              static_cast<long long>(gridDim.x) * blockDim.x);
}
__global__ void k_resolve(Params p, Run_desc d, Result *res, int32_t const *finish_rc) {
    resolve_body(p, d, res, finish_rc);
}

__global__ void __launch_bounds__(256) k_max_level(Geom g, uint32_t const *dist, Control const *ctl, Result *res) {
    max_level_body(g, dist, ctl, res, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x,
                   static_cast<long long>(gridDim.x) * blockDim.x);
}

__global__ void __launch_bounds__(256) k_paint(Params p, uint16_t *grid, Result *res) {
    paint_body(p, grid, res, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x,
               static_cast<long long>(gridDim.x) * blockDim.x);
}

// One warp per path (blockIdx.x = colour / slot).  The finish and the field are looked up from
// the resolved result so that no host round trip sits between the search and the walk.
__global__ void k_walk(Params p, int game, uint16_t *grid, Result *res, int32_t const *finish_rc, int32_t *out,
                       unsigned long long out_stride, unsigned long long max_out, int first_only) {
    __shared__ uint32_t sm[WALK_SM_U32]; // one warp per block
    int const slot = static_cast<int>(blockIdx.x);
    int plane = 0, fj = 0;
    uint32_t repaint = 0;
    if (game == GAME_HUNT && p.ctl->walked) { // lat_path_*: the path is marked and written already; its length is the finish's level
        if (lane_id() == 0) res->path_len[slot] = res->game_level[0];
        return;
    }
    if (game == GAME_GATHER && !first_only && p.ctl->walked) { // lat_path_*: every colour's path is written already
        if (lane_id() == 0) res->path_len[slot] = res->finish_of[slot] >= 0 ? res->game_level[slot] : 0;
        return;
    }
    if (game == GAME_GATHER && first_only && p.ctl->fronts_done) { // lat_fronts_body: path.front() is known (and no field exists to step on)
        if (lane_id() == 0) res->path_len[slot] = res->finish_of[slot] >= 0 ? res->game_level[slot] : 0;
        return;
    }
    if (game == GAME_HUNT) {
        if (res->winner < 0) {
            if (lane_id() == 0) res->path_len[slot] = 0;
            return;
        }
        repaint = (1u << res->winner) << SQ_PAINT_SHIFT;
    } else if (game == GAME_CORNERS) {
        if (res->winner < 0) {
            if (lane_id() == 0) res->path_len[slot] = 0;
            return;
        }
        plane = res->winner;
    } else { // gather: colour `slot`
        fj = res->finish_of[slot];
        if (fj < 0) {
            if (lane_id() == 0) {
                res->path_len[slot] = 0;
                if (out) {
                    out[slot * out_stride * 2] = -1;
                }
            }
            return;
        }
    }
    walk_body(p, plane, finish_rc[2 * fj], finish_rc[2 * fj + 1], grid, out ? out + slot * out_stride * 2 : nullptr,
              max_out, repaint, first_only, res, slot, lane_id(), sm);
}

__global__ void k_band_walk(Params p, int r, int c, uint16_t *grid, int32_t *out, unsigned long long max_out, uint32_t repaint, int first_only,
                            int include_first, Result *res) {
    __shared__ uint32_t sm[WALK_SM_U32];
    walk_body(p, 0, r, c, grid, out, max_out, repaint, first_only, res, 0, lane_id(), sm, include_first);
}
__global__ void k_recolour(uint16_t *square, uint16_t paint_bits) { *square = static_cast<uint16_t>((*square & ~SQ_PAINT_MASK) | paint_bits); }

__global__ void k_gather_fixup(Geom g, uint16_t *grid, Result const *res, int32_t const *finish_rc,
                               int32_t const *front_rc) {
    gather_fixup_body(g, grid, res, finish_rc, front_rc);
}

__global__ void k_or_bits(uint16_t *grid, long long cols, int32_t const *cells_rc, uint16_t const *bits, int n) {
    for (int i = static_cast<int>(threadIdx.x); i < n; i += static_cast<int>(blockDim.x)) {
        grid[static_cast<long long>(cells_rc[2 * i]) * cols + cells_rc[2 * i + 1]] |= bits[i];
    }
}

__global__ void __launch_bounds__(256) k_band_path_rows(Geom g, uint64_t const *path, uint8_t *top, uint8_t *bottom) {
    band_path_rows_body(g, path, top, bottom, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(256) k_band_ghost_words(Geom g, uint8_t const *above_bottom, uint8_t const *below_top,
                                                          uint64_t *ghost_n, uint64_t *ghost_s) {
    band_ghost_words_body(g, above_bottom, below_top, ghost_n, ghost_s, gwarp_id(), nwarps_total(), lane_id());
}
__global__ void __launch_bounds__(256) k_band_export_edges(Geom g, uint32_t const *edges, uint32_t *top, uint32_t *bottom) {
    band_export_edges_body(g, edges, top, bottom, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x);
}
__global__ void __launch_bounds__(256) k_band_import_edges(Params p, int plane, uint32_t const *above_bottom, uint32_t const *below_top,
                                                           uint32_t *activated) {
    band_import_edges_body(p, plane, above_bottom, below_top, activated, gwarp_id(), nwarps_total(), lane_id());
}
__global__ void k_band_rearm(Params p, unsigned long long budget_ns) { band_rearm_body(p, budget_ns); }

__global__ void __launch_bounds__(256) k_and_bits(uint16_t *grid, long long n, uint16_t mask) {
    long long const stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        grid[i] &= mask;
    }
}

} // namespace

// ------------------------------------------------------------------------------------ host side
namespace {

thread_local std::string g_error;

int fail(int code, char const *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t const e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                         \
            return fail(LAB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                                \
    } while (0)

// LAB_TRACE=1: synchronise after every launch and say which one it was (debugging hangs/faults)
bool trace_on() {
    static int const on = std::getenv("LAB_TRACE") ? 1 : 0;
    return on != 0;
}
#define TRACE(name, stream)                                                                              \
    do {                                                                                                 \
        if (trace_on()) {                                                                                \
            fprintf(stderr, "[lab] %s ...", name);                                                       \
            fflush(stderr);                                                                              \
            cudaError_t const e_ = cudaStreamSynchronize(stream);                                        \
            fprintf(stderr, " %s\n", cudaGetErrorString(e_));                                            \
        }                                                                                                \
    } while (0)

enum Ev { EV_SETUP0, EV_PACK0, EV_SEARCH0, EV_POST0, EV_END, EV_COPY0, EV_COPY1, EV_COUNT };

} // namespace

struct lab_grid {
    int device = 0;
    Geom g{};
    uint16_t *squares = nullptr;
    bool owns_squares = false;
    uint64_t *path = nullptr;
    uint64_t *cross = nullptr;
    Plane plane[MAX_PLANES]{};
    bool owns_dist[MAX_PLANES]{};
    int planes_alloc = 0;
    Control *ctl = nullptr;
    uint32_t *queue = nullptr;
    uint32_t qcap = 0;
    Result *res = nullptr;
    int32_t *d_finish_rc = nullptr; // 4 (r,c)
    int32_t *d_front_rc = nullptr;  // 4 (r,c)
    int32_t *d_paths = nullptr;     // path_slots * path_cap (r,c)
    unsigned long long path_cap = 0;
    int path_slots = 0;
    uint32_t *pixels = nullptr; // heat map of the last distance map (lab_distance_paint), rows*cols
    bool owns_pixels = false;
    bool have_map = false;      // plane 0 holds a finished distance map
    uint32_t *runs = nullptr;   // Runs::paint_runs: run lengths of the last lab_runs_map, rows*cols
    Result *runs_res = nullptr; // ... and their maximum (max_level), where k_heatmap looks for it
    bool have_runs = false;
    int32_t *d_cells = nullptr; // scratch for or_bits
    uint16_t *d_bits = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[EV_COUNT]{};
    int sm_count = 0;
    int bfs_blocks = 0;
    int bfs_threads = 128;
    lab_timing timing{};
    lab_stats stats{};
    Result last{};
    float h2d_ms = 0, d2h_ms = 0;
    // row-band state
    uint64_t *ghost_n = nullptr, *ghost_s = nullptr; // neighbours' border path rows, one word per tile column
    uint32_t *d_activated = nullptr;
    Run_desc band_desc{};
    bool band_open = false;
    bool band_mode = false; // the search being prepared is one band of a sharded one
    bool band_skip_field_reset = false; // ... whose level field the lattice road across bands will write or wipe
    // spanning-tree pipeline workspace (allocated on first use)
    Tree tree{};
    bool tree_alloc = false;
    Tree_ctl tree_last{};
    bool tree_tried = false;
    bool opt_solver_levels = true; // lab_grid_set_option(LAB_OPT_SOLVER_LEVELS)
    bool skip_field_now = false;   // this solver call may leave its level field unwritten (gather, no paths, the option off)
    bool field_skipped = false;    // ... and the road that ran did (lab_levels_download says so)
    int32_t *hunt_out = nullptr; // where hunt's winner path goes (solver_common sets both before the search; null: not wanted)
    unsigned long long hunt_out_cap = 0;
    Lattice lat{}; // the lattice road's workspace (allocated on first use)
    bool lat_alloc = false;
    Lattice band_lat{}; // ... across row bands (lab_band_lat_*): whole-maze geometry, the caller's whole-maze arrays
    bool band_lat_bound = false;
    Tree_ctl *band_lat_tc = nullptr;
    uint64_t launches = 0; // kernels launched by the last search (bench.py reports them)
    bool tree_cfg = false;
    // the pipeline across row bands (lab_band_tree_*): the caller's global arrays, this band's control block
    Tree band_tree{};
    uint32_t band_ntiles_global = 0;
    uint32_t *band_alist = nullptr;            // the band's own crossings
    uint32_t const *band_border_list = nullptr; // every band's border-row slots (the short list)
    uint32_t band_border_count = 0, band_crossings = 0;
    Band_layout const *band_layout = nullptr;
    int band_world = 1, band_me = 0;
    bool band_tree_bound = false;
    int tree_coop_blocks = 0; // co-resident blocks of the multi-round kernels
    int tree_flood_per_sm = 1, tree_walk_per_sm = 1;
    cudaEvent_t tev[10]{}; // LAB_TREE_TIMING=1: per-phase events
    bool tev_on = false;
    cudaEvent_t lev[10]{}; // ... and the lattice road's
    bool lev_on = false;
};

namespace {

long long env_ll(char const *name, long long dflt) {
    char const *v = std::getenv(name);
    return v && *v ? std::atoll(v) : dflt;
}

int make_geom(int rows, int cols, Geom &g) {
    if (rows < 3 || cols < 3 || static_cast<long long>(rows) * cols >= (1ll << 32)) {
        return fail(LAB_ERR_ARG, "bad maze size %dx%d", rows, cols);
    }
    g.rows = rows;
    g.cols = cols;
    g.tiles_y = (rows + TS - 1) / TS;
    g.tiles_x = (cols + TS - 1) / TS;
    g.ntiles = g.tiles_y * g.tiles_x;
    return LAB_OK;
}

int grid_common_init(lab_grid *h) {
    CU(cudaGetDevice(&h->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, h->device));
    h->sm_count = prop.multiProcessorCount;
    size_t const nt = static_cast<size_t>(h->g.ntiles);
    CU(cudaMalloc(&h->path, nt * TS * sizeof(uint64_t)));
    CU(cudaMalloc(&h->cross, nt * 4 * sizeof(uint64_t)));
    CU(cudaMalloc(&h->ctl, sizeof(Control)));
    CU(cudaMalloc(&h->res, sizeof(Result)));
    CU(cudaMalloc(&h->d_finish_rc, 8 * sizeof(int32_t)));
    CU(cudaMalloc(&h->d_front_rc, 8 * sizeof(int32_t)));
    CU(cudaMalloc(&h->d_cells, 64 * sizeof(int32_t)));
    CU(cudaMalloc(&h->d_bits, 32 * sizeof(uint16_t)));
    for (auto &e : h->ev) {
        CU(cudaEventCreate(&e));
    }
    // persistent launch: as many 4-warp blocks as stay resident
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_bfs_tiles, h->bfs_threads, 0));
    long long const want_warps_per_sm = env_ll("LAB_BFS_WARPS_PER_SM", 0);
    if (want_warps_per_sm > 0) {
        per_sm = std::max<int>(1, std::min<long long>(per_sm, want_warps_per_sm / (h->bfs_threads / 32)));
    }
    h->bfs_blocks = std::max(1, per_sm) * h->sm_count;
    return LAB_OK;
}

int ensure_planes(lab_grid *h, int n) {
    size_t const nt = static_cast<size_t>(h->g.ntiles);
    size_t const cells = static_cast<size_t>(h->g.rows) * h->g.cols;
    for (int k = h->planes_alloc; k < n; k++) {
        Plane &p = h->plane[k];
        CU(cudaMalloc(&p.visited, nt * TS * sizeof(uint64_t)));
        CU(cudaMalloc(&p.edges, (nt + 2 * static_cast<size_t>(h->g.tiles_x)) * 4 * TS * sizeof(uint32_t)));
        CU(cudaMalloc(&p.state, nt * sizeof(uint32_t)));
        if (!p.dist) {
            CU(cudaMalloc(&p.dist, cells * sizeof(uint32_t)));
            h->owns_dist[k] = true;
        }
        h->planes_alloc = k + 1;
    }
    // queue: every (plane, tile) can be queued at most once at a time, plus one ticket per waiting warp
    unsigned long long need = static_cast<unsigned long long>(n) * nt + static_cast<unsigned long long>(h->bfs_blocks) * (h->bfs_threads / 32) + 64;
    uint32_t cap = 1024;
    while (cap < need * 2) {
        cap <<= 1;
    }
    if (cap > h->qcap) {
        if (h->queue) {
            CU(cudaFree(h->queue));
        }
        CU(cudaMalloc(&h->queue, static_cast<size_t>(cap) * sizeof(uint32_t)));
        h->qcap = cap;
    }
    return LAB_OK;
}

Params make_params(lab_grid *h, int nplanes) {
    Params p{};
    p.g = h->g;
    p.path = h->path;
    p.cross = h->cross;
    for (int k = 0; k < MAX_PLANES; k++) {
        p.plane[k] = h->plane[k < nplanes ? k : 0];
    }
    p.nplanes = nplanes;
    p.ctl = h->ctl;
    p.queue = h->queue;
    p.qmask = h->qcap - 1;
    return p;
}

int blocks_for(lab_grid *h, long long warps_wanted, int threads) {
    long long const per_block = threads / 32;
    long long const cap = static_cast<long long>(h->sm_count) * (2048 / threads) * 4; // a few waves
    long long b = (warps_wanted + per_block - 1) / per_block;
    return static_cast<int>(std::max<long long>(1, std::min(b, cap)));
}

unsigned long long budget_ns() { return static_cast<unsigned long long>(env_ll("LAB_BFS_BUDGET_MS", 20000)) * 1000000ull; }

// The spanning-tree pipeline is tried for single-field searches that have their source in this grid.
// (not for row bands: lab_band_* runs the tile wave in every band)
bool tree_may_run(lab_grid const *h, Run_desc const &d) {
    return !h->band_mode && d.nplanes == 1 && d.src_r[0] >= 0 && env_ll("LAB_TREE", 1) != 0;
}

// The lattice road (device/lattice.cuh) is tried where the maze can be one: odd sizes, the start on a cell or a passage
// (not on an (even, even) square: those are the lattice's pillars).
bool lattice_may_run(lab_grid const *h, Run_desc const &d) {
    return tree_may_run(h, d) && env_ll("LAB_LATTICE", 1) != 0 && (h->g.rows & 1) && (h->g.cols & 1) && ((d.src_r[0] | d.src_c[0]) & 1);
}

// Reset the workspace and pack the grid into tile bit masks (sources forced passable).
int search_prepare(lab_grid *h, Run_desc const &d, uint16_t pass_mask, uint16_t pass_value) {
    int rc = ensure_planes(h, d.nplanes);
    if (rc) {
        return rc;
    }
    cudaStream_t s = h->stream;
    Geom const &g = h->g;
    size_t const nt = static_cast<size_t>(g.ntiles);
    size_t const cells = static_cast<size_t>(g.rows) * g.cols;
    h->launches = 0;
    CU(cudaEventRecord(h->ev[EV_SETUP0], s));
    CU(cudaMemsetAsync(h->ctl, 0, sizeof(Control), s));
    CU(cudaMemsetAsync(h->res, 0, sizeof(Result), s));
    CU(cudaMemsetAsync(h->queue, 0xFF, static_cast<size_t>(h->qcap) * sizeof(uint32_t), s));
    for (int k = 0; k < d.nplanes; k++) {
        Plane &pl = h->plane[k];
        CU(cudaMemsetAsync(pl.visited, 0, nt * TS * sizeof(uint64_t), s));
        CU(cudaMemsetAsync(pl.edges, 0xFF, (nt + 2 * static_cast<size_t>(g.tiles_x)) * 4 * TS * sizeof(uint32_t), s));
        CU(cudaMemsetAsync(pl.state, 0, nt * sizeof(uint32_t), s));
        if (!tree_may_run(h, d) && !(h->band_mode && h->band_skip_field_reset)) { // (else the pipeline's last kernel or the gate writes every square of the field)
            CU(cudaMemsetAsync(pl.dist, 0xFF, cells * sizeof(uint32_t), s));
        }
    }
    CU(cudaEventRecord(h->ev[EV_PACK0], s));
    k_pack<<<blocks_for(h, static_cast<long long>(g.tiles_y) * TS, 256), 256, 0, s>>>(g, h->squares, h->path, pass_mask, pass_value, static_cast<uint16_t>(d.spec_mark));
    TRACE("k_pack", s);
    k_force_sources<<<1, 1, 0, s>>>(g, h->path, d, h->squares);
    TRACE("k_force_sources", s);
    h->launches += 2;
    CU(cudaGetLastError());
    return LAB_OK;
}

// Border masks (optionally against the neighbouring bands' rows) and the seeded control block.
int search_arm(lab_grid *h, Run_desc const &d, uint64_t const *ghost_n, uint64_t const *ghost_s) {
    cudaStream_t s = h->stream;
    Geom const &g = h->g;
    Params p = make_params(h, d.nplanes);
    k_cross<<<blocks_for(h, g.ntiles, 256), 256, 0, s>>>(g, h->path, h->cross, ghost_n, ghost_s);
    TRACE("k_cross", s);
    k_init_run<<<1, 1, 0, s>>>(p, d, budget_ns());
    TRACE("k_init_run", s);
    h->launches += 2;
    CU(cudaGetLastError());
    return LAB_OK;
}

int search_launch(lab_grid *h, int nplanes) {
    cudaStream_t s = h->stream;
    Params p = make_params(h, nplanes);
    k_bfs_tiles<<<h->bfs_blocks, h->bfs_threads, 0, s>>>(p);
    TRACE("k_bfs_tiles", s);
    h->launches += 1;
    CU(cudaGetLastError());
    return LAB_OK;
}

// Launch shapes of the pipeline's kernels (occupancy of the co-resident ones), once per grid.
int tree_configure(lab_grid *h) {
    if (h->tree_cfg) {
        return LAB_OK;
    }
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tree_jump<0>, 512, 0));
    int per_sm2 = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, k_tree_jump<1>, 512, 0));
    h->tree_coop_blocks = std::max(1, std::min(per_sm, per_sm2)) * h->sm_count;
    size_t const flood_smem = 4 * FLOOD_U64 * sizeof(uint64_t);
    CU(cudaFuncSetAttribute(k_tree_flood, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(flood_smem)));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->tree_flood_per_sm, k_tree_flood, 128, flood_smem));
    h->tree_flood_per_sm = std::max(1, h->tree_flood_per_sm);
    CU(cudaFuncSetAttribute(k_tree_walk, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(k_tree_flood, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->tree_walk_per_sm, k_tree_walk, 128, 0));
    h->tree_walk_per_sm = std::max(1, h->tree_walk_per_sm);
    h->tree_cfg = true;
    return LAB_OK;
}

int ensure_tree(lab_grid *h) {
    if (h->tree_alloc) {
        return LAB_OK;
    }
    int const rc = tree_configure(h);
    if (rc) return rc;
    Tree &tr = h->tree;
    size_t const nt = static_cast<size_t>(h->g.ntiles);
    if (nt * SLOTS + 2 >= (1ull << 32)) {
        return fail(LAB_ERR_ARG, "maze too large for the tree pipeline");
    }
    tr.nslots = static_cast<uint32_t>(nt * SLOTS);
    size_t const n = (nt * SLOTS + 2) * sizeof(uint32_t);
    CU(cudaMalloc(&tr.succ0, n));
    CU(cudaMalloc(&tr.sr[0], 2 * n));
    CU(cudaMalloc(&tr.sr[1], 2 * n));
    CU(cudaMalloc(&tr.inmask, nt * 4 * sizeof(uint64_t)));
    CU(cudaMalloc(&tr.loc, nt * SLOTS * sizeof(uint32_t)));
    CU(cudaMalloc(&tr.alist, nt * SLOTS * sizeof(uint32_t)));
    CU(cudaMalloc(&tr.tc, sizeof(Tree_ctl)));
    h->tree_alloc = true;
    return LAB_OK;
}

int ensure_lattice(lab_grid *h) {
    if (h->lat_alloc) {
        return LAB_OK;
    }
    Lattice &lt = h->lat;
    size_t const nt = static_cast<size_t>(h->g.ntiles);
    lt.sup_y = (h->g.tiles_y + LST - 1) / LST;
    lt.sup_x = (h->g.tiles_x + LST - 1) / LST;
    lt.nsuper = lt.sup_y * lt.sup_x;
    lt.nslots = static_cast<uint32_t>(nt * LSLOTS);
    CU(cudaMalloc(&lt.mask, nt * 96 * sizeof(uint32_t)));
    CU(cudaMalloc(&lt.xmask, nt * 4 * sizeof(uint32_t)));
    CU(cudaMalloc(&lt.outslot, nt * LSLOTS));
    CU(cudaMalloc(&lt.label, nt * LSLOTS * sizeof(uint16_t)));
    CU(cudaMalloc(&lt.cmap, nt * 1024 * sizeof(uint16_t)));
    CU(cudaMalloc(&lt.xd, nt * LSLOTS * sizeof(uint64_t)));
    CU(cudaMalloc(&lt.gb, static_cast<size_t>(lt.nsuper) * LB * sizeof(uint64_t)));
    CU(cudaMalloc(&lt.R, nt * LSLOTS * sizeof(uint32_t)));
    CU(cudaMalloc(&lt.dplane, nt * L_DEPTH_PLANES * 32 * sizeof(uint32_t)));
    CU(cudaMalloc(&lt.emask, nt * 4 * sizeof(uint32_t)));
    CU(cudaMalloc(&lt.wsum, (nt * LSLOTS + 2 + L_CHUNK) * sizeof(uint32_t)));
    CU(cudaMalloc(&lt.chunk, (nt * LSLOTS / L_CHUNK + 2) * sizeof(uint32_t)));
    CU(cudaFuncSetAttribute(k_lat_rank_local, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(LAT_RANK_WARPS * LSUP_SLOTS * sizeof(uint32_t))));
    lt.tile_base = lt.sup_base = 0;
    lt.nsuper_global = lt.nsuper;
    lt.nslots_global = lt.nslots;
    lt.peers = 1;
    lt.gb_peer[0] = lt.gb;
    lt.wsum_peer[0] = lt.wsum;
    h->lat_alloc = true;
    return LAB_OK;
}

// The lattice road (device/lattice.cuh), tried before the general pipeline on mazes that can be lattices: odd
// sizes, the start on an (odd, odd) square.  Whether the maze IS one is decided on the device, kernel by kernel;
// after a failure the remaining launches return at once and nothing has been written.
int run_lattice(lab_grid *h, Params const &p, Run_desc const &d) {
    int const rc = ensure_lattice(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    Lattice lt = h->lat;
    lt.tc = h->tree.tc;
    int const nt = h->g.ntiles;
    int const tile_blocks = static_cast<int>(std::min<long long>((nt + 3) / 4, static_cast<long long>(h->sm_count) * 16));
    h->lev_on = env_ll("LAB_TREE_TIMING", 0) != 0;
    int lev_i = 0;
    auto mark = [&]() {
        if (h->lev_on) {
            if (!h->lev[lev_i]) cudaEventCreate(&h->lev[lev_i]);
            cudaEventRecord(h->lev[lev_i++], s);
        }
    };
    mark();
    k_lat_link<<<tile_blocks, 128, 0, s>>>(p, lt);
    TRACE("k_lat_link", s);
    mark();
    size_t const rank_smem = LAT_RANK_WARPS * LSUP_SLOTS * sizeof(uint32_t);
    int const rank_blocks = std::min((lt.nsuper + LAT_RANK_WARPS - 1) / LAT_RANK_WARPS, h->sm_count * 6);
    k_lat_rank_local<<<rank_blocks, 32 * LAT_RANK_WARPS, rank_smem, s>>>(p, lt);
    TRACE("k_lat_rank_local", s);
    mark();
    long long const nb = static_cast<long long>(lt.nsuper) * LB;
    k_lat_rank_global<<<static_cast<int>(std::min<long long>((nb + 255) / 256, static_cast<long long>(h->sm_count) * 8)), 256, 0, s>>>(lt);
    TRACE("k_lat_rank_global", s);
    mark();
    k_lat_rank_expand<<<static_cast<int>(std::min<long long>((static_cast<long long>(lt.nslots) + 255) / 256, static_cast<long long>(h->sm_count) * 16)), 256, 0, s>>>(lt);
    TRACE("k_lat_rank_expand", s);
    mark();
    k_lat_flood<<<tile_blocks, 128, 0, s>>>(p, lt);
    TRACE("k_lat_flood", s);
    mark();
    k_lat_scan_chunks<<<static_cast<int>(std::min<long long>((static_cast<long long>(lt.nslots) / L_CHUNK + 8) / 8, static_cast<long long>(h->sm_count) * 8)), 256, 0, s>>>(lt);
    TRACE("k_lat_scan_chunks", s);
    mark();
    k_lat_scan_tops<<<1, 32, 0, s>>>(lt);
    TRACE("k_lat_scan_tops", s);
    mark();
    // hunt / gather: the paint rules are applied while the levels are written (k_paint then returns at once)
    bool const fuse_paint = (d.game == GAME_HUNT || d.game == GAME_GATHER) && env_ll("LAB_LAT_FUSE_PAINT", 1) != 0;
    if (fuse_paint) {
        bool const skip_field = h->skip_field_now; // (gather without paths, LAB_OPT_SOLVER_LEVELS off: nobody reads the field)
        k_lat_targets<<<1, 128, 0, s>>>(p, lt, d, h->res, h->d_finish_rc, skip_field ? h->d_front_rc : nullptr);
        TRACE("k_lat_targets", s);
        if (skip_field) {
            k_lat_levels<LAT_PAINT_ONLY><<<tile_blocks, 128, 0, s>>>(p, lt, h->res, h->squares);
        } else {
            k_lat_levels<LAT_PAINT><<<tile_blocks, 128, 0, s>>>(p, lt, h->res, h->squares);
        }
        h->launches += 1;
        if ((d.game == GAME_HUNT || (d.game == GAME_GATHER && h->hunt_out)) && !skip_field && env_ll("LAB_LAT_PATH", 1) != 0) {
            // hunt: the winner's path, painted over the colours just applied; gather with paths: every colour's path written
            // out (k_walk then returns at once).  The segments' list borrows the local ranking's pair array, which nobody
            // reads after k_lat_rank_expand.
            uint32_t *seg = reinterpret_cast<uint32_t *>(lt.xd);
            uint32_t const seg_cap = lt.nslots;
            k_lat_path_begin<<<1, 128, 0, s>>>(p, lt, d.game, h->res, h->d_finish_rc, h->squares, h->hunt_out, h->path_cap, h->hunt_out_cap, seg, seg_cap);
            k_lat_path_find<<<blocks_for(h, nt, 256), 256, 0, s>>>(p, lt, seg, seg_cap);
            k_lat_path_walk<<<h->sm_count * 8, 32, 0, s>>>(p, lt, h->res, d.game == GAME_HUNT ? 0xFFFFFFFFu : 0u, h->squares, h->hunt_out, h->path_cap,
                                                           h->hunt_out_cap, seg, seg_cap);
            TRACE("k_lat_path_walk", s);
            h->launches += 3;
        }
    } else {
        k_lat_levels<LAT_LEVELS><<<tile_blocks, 128, 0, s>>>(p, lt, nullptr, nullptr);
    }
    TRACE("k_lat_levels", s);
    mark();
    h->launches += 8;
    CU(cudaGetLastError());
    return LAB_OK;
}

int ceil_log2(unsigned long long n) {
    int r = 0;
    while ((1ull << r) < n) {
        r++;
    }
    return r;
}

// The spanning-tree pipeline (device/tree.cuh), between search_arm and the wave kernel.  Everything is
// decided and gated on the device (no host round trip): if the counts or any later check say "not a
// tree", the remaining kernels return at once and the wave kernel finds its control block untouched.
int run_tree(lab_grid *h, Run_desc const &d) {
    h->tree_last = Tree_ctl{};
    h->tree_tried = false;
    if (!tree_may_run(h, d)) {
        return LAB_OK;
    }
    int rc = ensure_tree(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    Tree tr = h->tree;
    tr.mark_bits = d.game == GAME_DISTANCE ? SQ_MEASURE : 0u;
    Params p = make_params(h, 1);
    CU(cudaMemsetAsync(tr.tc, 0, sizeof(Tree_ctl), s));
    k_tree_count<<<std::min(blocks_for(h, h->g.ntiles, 256), h->sm_count * 8), 256, 0, s>>>(p, tr); // (few warps: seven same-address atomics each)
    TRACE("k_tree_count", s);
    h->launches += 1;
    // No host round trip: k_tree_init decides on the device whether the counts allow a tree, and every
    // later kernel of the pipeline returns at once if not (a handful of empty launches on an arena).
    h->tev_on = env_ll("LAB_TREE_TIMING", 0) != 0;
    int tev_i = 0;
    auto mark = [&]() {
        if (h->tev_on) {
            if (!h->tev[tev_i]) cudaEventCreate(&h->tev[tev_i]);
            cudaEventRecord(h->tev[tev_i++], s);
        }
    };
    // odd sizes and an (odd, odd) start: the maze may be carved on the lattice of odd squares (every reference builder's is)
    bool const lat_try = lattice_may_run(h, d);
    tr.lat_try = lat_try ? 1u : 0u;
    mark();
    k_tree_init<<<1, 1, 0, s>>>(p, tr);
    if (lat_try) {
        rc = run_lattice(h, p, d);
        if (rc) return rc;
    }
    int const walk_blocks = std::min<long long>((h->g.ntiles + 3) / 4, static_cast<long long>(h->sm_count) * h->tree_walk_per_sm);
    k_tree_walk<<<walk_blocks, 128, 0, s>>>(p, tr);
    TRACE("k_tree_walk", s);
    mark();
    long long const nslots = static_cast<long long>(tr.nslots) + 2;
    uint32_t rounds = static_cast<uint32_t>(ceil_log2(static_cast<unsigned long long>(nslots)));
    // few crossings: fewer blocks make a cheaper barrier (a 64x64 tile has ~64 crossings)
    int const coop_blocks = static_cast<int>(std::max<long long>(
        1, std::min<long long>(h->tree_coop_blocks, (static_cast<long long>(h->g.ntiles) * 64 + 512 * JUMP_K - 1) / (512 * JUMP_K))));
    {
        void *args[] = {&p, const_cast<Tree *>(&tr), &rounds};
        CU(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(k_tree_jump<0>), dim3(coop_blocks), dim3(512), args, 0, s));
    }
    TRACE("k_tree_jump<rank+orient>", s);
    mark();
    mark();
    int const fill_blocks = std::min<long long>((h->g.ntiles + 3) / 4, static_cast<long long>(h->sm_count) * h->tree_flood_per_sm);
    k_tree_flood<<<fill_blocks, 128, 4 * FLOOD_U64 * sizeof(uint64_t), s>>>(p, tr);
    TRACE("k_tree_flood", s);
    mark();
    mark();
    {
        void *args[] = {&p, const_cast<Tree *>(&tr), &rounds};
        CU(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(k_tree_jump<1>), dim3(coop_blocks), dim3(512), args, 0, s));
    }
    TRACE("k_tree_jump<parent+prefix>", s);
    mark();
    k_tree_fixup<<<blocks_for(h, static_cast<long long>(h->g.rows) * ((h->g.tiles_x + 3) / 4), 256), 256, 0, s>>>(p, tr, h->squares);
    TRACE("k_tree_fixup", s);
    mark();
    if (d.game == GAME_DISTANCE) { // (the solvers paint an open room straight from the closed form: k_paint)
        k_room_distance<<<h->sm_count * 8, 256, 0, s>>>(p, h->squares, h->res);
        TRACE("k_room_distance", s);
        h->launches += 1;
    }
    k_tree_gate<<<h->sm_count * 2, 256, 0, s>>>(p, tr);
    TRACE("k_tree_gate", s);
    mark();
    h->tree_tried = true; // finish_run reads the pipeline's control block back
    h->launches += 7;     // init, walk, jump<rank>, flood, jump<prefix>, fixup, gate
    CU(cudaGetLastError());
    return LAB_OK;
}

// Reset the workspace, pack the grid, run the persistent search.  Leaves EV_POST0 recorded.
int run_search(lab_grid *h, Run_desc const &d, uint16_t pass_mask, uint16_t pass_value) {
    h->band_mode = false;
    int rc = search_prepare(h, d, pass_mask, pass_value);
    if (rc) return rc;
    rc = search_arm(h, d, nullptr, nullptr);
    if (rc) return rc;
    CU(cudaEventRecord(h->ev[EV_SEARCH0], h->stream));
    rc = run_tree(h, d);
    if (rc) return rc;
    rc = search_launch(h, d.nplanes);
    if (rc) return rc;
    CU(cudaEventRecord(h->ev[EV_POST0], h->stream));
    return LAB_OK;
}

// After the post kernels: record the end, sync, read results/statistics/timings back.
int finish_run(lab_grid *h) {
    cudaStream_t s = h->stream;
    CU(cudaEventRecord(h->ev[EV_END], s));
    Control c{};
    CU(cudaMemcpyAsync(&h->last, h->res, sizeof(Result), cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(&c, h->ctl, sizeof(Control), cudaMemcpyDeviceToHost, s));
    if (h->tree_tried) {
        CU(cudaMemcpyAsync(&h->tree_last, h->tree.tc, sizeof(Tree_ctl), cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
    if (h->tree_tried && h->lev_on) {
        static char const *const names[] = {"link", "rank_local", "rank_global", "expand", "flood", "scan", "tops", "levels"};
        fprintf(stderr, "[lab lattice]");
        for (int i = 0; i < 8; i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, h->lev[i], h->lev[i + 1]);
            fprintf(stderr, " %s %.1fus", names[i], ms * 1000.f);
        }
        fprintf(stderr, " (err %u)\n", h->tree_last.lat_err);
        h->lev_on = false;
    }
    if (h->tree_tried && h->tev_on) {
        static char const *const names[] = {"walk", "rank", "orient", "flood", "parent", "prefix", "fixup", "gate"};
        fprintf(stderr, "[lab tree]");
        for (int i = 0; i < 8; i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, h->tev[i], h->tev[i + 1]);
            fprintf(stderr, " %s %.1fus", names[i], ms * 1000.f);
        }
        fprintf(stderr, "\n");
    }
    lab_timing &t = h->timing;
    CU(cudaEventElapsedTime(&t.setup_ms, h->ev[EV_SETUP0], h->ev[EV_PACK0]));
    CU(cudaEventElapsedTime(&t.pack_ms, h->ev[EV_PACK0], h->ev[EV_SEARCH0]));
    CU(cudaEventElapsedTime(&t.search_ms, h->ev[EV_SEARCH0], h->ev[EV_POST0]));
    CU(cudaEventElapsedTime(&t.post_ms, h->ev[EV_POST0], h->ev[EV_END]));
    CU(cudaEventElapsedTime(&t.total_ms, h->ev[EV_SETUP0], h->ev[EV_END]));
    t.h2d_ms = h->h2d_ms;
    lab_stats &st = h->stats;
    st.tiles = static_cast<uint64_t>(h->g.ntiles);
    st.activations = c.n_activations;
    st.empty = c.n_empty;
    st.levels = c.n_levels;
    st.settled = c.n_settled;
    st.rechecks = c.n_rechecks;
    st.resettles = c.n_resettles;
    st.queue_pushes = c.n_pushes;
    st.direct = c.n_direct;
    st.worker_warps = static_cast<uint32_t>(h->bfs_blocks * (h->bfs_threads / 32));
    st.error = c.error;
    st.cyc_load = c.cyc_load;
    st.cyc_loop = c.cyc_loop;
    st.cyc_publish = c.cyc_publish;
    st.cyc_idle = c.cyc_idle;
    st.cyc_step = c.cyc_step;
    st.cyc_emit = c.cyc_emit;
    st.batches = c.n_batches;
    st.tree_squares = h->tree_last.V;
    st.tree_pairs = h->tree_last.E;
    st.tree_crossings = h->tree_last.n_arcs;
    st.tree_used = h->tree_last.used;
    st.launches = h->launches;
    if (c.error == 1) {
        return fail(LAB_ERR_TIMEOUT, "search kernel passed its deadline (LAB_BFS_BUDGET_MS); %llu tiles still active",
                    static_cast<unsigned long long>(c.pending));
    }
    if (c.error) {
        return fail(LAB_ERR_INTERNAL, "search kernel reported error %u (state word %u)", c.error & 0xFF, c.error >> 8);
    }
    if (c.pending != 0) {
        return fail(LAB_ERR_INTERNAL, "search ended with %u tiles still active", c.pending);
    }
    return LAB_OK;
}

bool in_grid(lab_grid const *h, lab_point p) { return p.row >= 0 && p.row < h->g.rows && p.col >= 0 && p.col < h->g.cols; }

// room for `slots` paths of `cap` squares each (gather: four, hunt and corners: the winner's alone)
int ensure_paths(lab_grid *h, unsigned long long cap, int slots) {
    if (cap > h->path_cap || slots > h->path_slots) {
        if (h->d_paths) {
            CU(cudaFree(h->d_paths));
            h->d_paths = nullptr;
        }
        cap = std::max(cap, h->path_cap);
        slots = std::max(slots, h->path_slots);
        h->path_cap = 0;
        h->path_slots = 0;
        CU(cudaMalloc(&h->d_paths, static_cast<size_t>(cap) * slots * 2 * sizeof(int32_t)));
        h->path_cap = cap;
        h->path_slots = slots;
    }
    return LAB_OK;
}

int grid_blocks(lab_grid *h) { return h->sm_count * 8; }

} // namespace

// Every entry point that takes a grid runs on the grid's device whatever device is current in the calling thread
// (one process driving several GPUs, or torch.cuda.set_device in between), and puts the caller's device back.
struct Device_guard {
    int prev = -1;
    bool switched = false;
    explicit Device_guard(lab_grid const *h) {
        if (h && cudaGetDevice(&prev) == cudaSuccess && prev != h->device) {
            switched = cudaSetDevice(h->device) == cudaSuccess;
        }
    }
    ~Device_guard() {
        if (switched) {
            cudaSetDevice(prev);
        }
    }
    Device_guard(Device_guard const &) = delete;
    Device_guard &operator=(Device_guard const &) = delete;
};

// ----------------------------------------------------------------------------------------- ABI
extern "C" {

const char *lab_last_error(void) { return g_error.c_str(); }

int lab_device_info(char *buf, size_t buflen) {
    int n = 0;
    cudaError_t const e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        n = 0;
        (void)cudaGetLastError();
    }
    if (buf && buflen) {
        if (n > 0) {
            cudaDeviceProp prop;
            int dev = 0;
            cudaGetDevice(&dev);
            cudaGetDeviceProperties(&prop, dev);
            snprintf(buf, buflen, "%s sm_%d%d, %d SMs, %.1f GB; liblabyrinth_b200 built for sm_100a", prop.name, prop.major,
                     prop.minor, prop.multiProcessorCount, static_cast<double>(prop.totalGlobalMem) / 1e9);
        } else {
            snprintf(buf, buflen, "no CUDA device (%s); liblabyrinth_b200 has no CPU fallback", cudaGetErrorString(e));
        }
    }
    return n;
}

int lab_grid_create(int rows, int cols, const uint16_t *host_squares, lab_grid **out) {
    if (!out) {
        return fail(LAB_ERR_ARG, "out is null");
    }
    *out = nullptr;
    auto *h = new lab_grid();
    int rc = make_geom(rows, cols, h->g);
    if (rc) {
        delete h;
        return rc;
    }
    size_t const bytes = static_cast<size_t>(rows) * cols * sizeof(uint16_t);
    cudaError_t e = cudaMalloc(&h->squares, bytes);
    if (e != cudaSuccess) {
        delete h;
        return fail(LAB_ERR_CUDA, "cudaMalloc(%zu) failed: %s -- liblabyrinth_b200 needs a CUDA device, there is no CPU fallback",
                    bytes, cudaGetErrorString(e));
    }
    h->owns_squares = true;
    rc = grid_common_init(h);
    if (rc) {
        lab_grid_destroy(h);
        return rc;
    }
    if (host_squares) {
        rc = lab_grid_upload(h, host_squares);
    } else {
        e = cudaMemset(h->squares, 0, bytes);
        rc = e == cudaSuccess ? LAB_OK : fail(LAB_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
    }
    if (rc) {
        lab_grid_destroy(h);
        return rc;
    }
    *out = h;
    return LAB_OK;
}

int lab_grid_create_on_device(int rows, int cols, void *device_squares, lab_grid **out) {
    if (!out || !device_squares) {
        return fail(LAB_ERR_ARG, "null argument");
    }
    *out = nullptr;
    auto *h = new lab_grid();
    int rc = make_geom(rows, cols, h->g);
    if (rc) {
        delete h;
        return rc;
    }
    h->squares = static_cast<uint16_t *>(device_squares);
    h->owns_squares = false;
    // the grid lives on the device that owns the caller's Squares, whatever device is current
    cudaPointerAttributes attr{};
    int prev = -1;
    bool switched = false;
    if (cudaPointerGetAttributes(&attr, device_squares) == cudaSuccess && attr.type == cudaMemoryTypeDevice && cudaGetDevice(&prev) == cudaSuccess &&
        prev != attr.device) {
        switched = cudaSetDevice(attr.device) == cudaSuccess;
    }
    (void)cudaGetLastError();
    rc = grid_common_init(h);
    if (switched) {
        cudaSetDevice(prev);
    }
    if (rc) {
        lab_grid_destroy(h);
        return rc;
    }
    *out = h;
    return LAB_OK;
}

int lab_grid_destroy(lab_grid *h) {
    Device_guard const guard_(h);
    if (!h) {
        return LAB_OK;
    }
    if (h->owns_squares) cudaFree(h->squares);
    cudaFree(h->path);
    cudaFree(h->cross);
    for (int k = 0; k < MAX_PLANES; k++) {
        cudaFree(h->plane[k].visited);
        cudaFree(h->plane[k].edges);
        cudaFree(h->plane[k].state);
        if (h->owns_dist[k]) cudaFree(h->plane[k].dist);
    }
    cudaFree(h->ctl);
    cudaFree(h->queue);
    cudaFree(h->res);
    cudaFree(h->d_finish_rc);
    cudaFree(h->d_front_rc);
    cudaFree(h->d_paths);
    cudaFree(h->band_tree.tc);
    cudaFree(h->band_lat_tc);
    cudaFree(h->runs);
    cudaFree(h->runs_res);
    if (h->owns_pixels) {
        cudaFree(h->pixels);
    }
    cudaFree(h->d_cells);
    cudaFree(h->d_bits);
    cudaFree(h->ghost_n);
    cudaFree(h->ghost_s);
    cudaFree(h->d_activated);
    cudaFree(h->tree.succ0);
    cudaFree(h->tree.sr[0]);
    cudaFree(h->tree.sr[1]);
    cudaFree(h->tree.inmask);
    cudaFree(h->tree.loc);
    cudaFree(h->tree.alist);
    cudaFree(h->tree.tc);
    if (h->lat_alloc) {
        Lattice &lt = h->lat;
        void *const arrays[] = {lt.mask, lt.xmask, lt.outslot, lt.label, lt.cmap, lt.xd, lt.gb, lt.R, lt.dplane, lt.emask, lt.wsum, lt.chunk};
        for (void *a : arrays) {
            cudaFree(a);
        }
    }
    for (auto &e : h->ev) {
        if (e) cudaEventDestroy(e);
    }
    for (auto &e : h->tev) {
        if (e) cudaEventDestroy(e);
    }
    for (auto &e : h->lev) {
        if (e) cudaEventDestroy(e);
    }
    delete h;
    return LAB_OK;
}

int lab_grid_set_stream(lab_grid *h, void *cuda_stream) {
    Device_guard const guard_(h);
    if (!h) {
        return fail(LAB_ERR_ARG, "null grid");
    }
    h->stream = static_cast<cudaStream_t>(cuda_stream);
    return LAB_OK;
}

int lab_grid_set_option(lab_grid *h, int option, int value) {
    if (!h) {
        return fail(LAB_ERR_ARG, "null grid");
    }
    if (option == LAB_OPT_SOLVER_LEVELS) {
        h->opt_solver_levels = value != 0;
        return LAB_OK;
    }
    return fail(LAB_ERR_ARG, "lab_grid_set_option: unknown option %d", option);
}

void *lab_grid_device_squares(lab_grid *h) { return h ? h->squares : nullptr; }
int lab_grid_rows(const lab_grid *h) { return h ? h->g.rows : 0; }
int lab_grid_cols(const lab_grid *h) { return h ? h->g.cols : 0; }

int lab_grid_upload(lab_grid *h, const uint16_t *host_squares) {
    Device_guard const guard_(h);
    if (!h || !host_squares) {
        return fail(LAB_ERR_ARG, "null argument");
    }
    size_t const bytes = static_cast<size_t>(h->g.rows) * h->g.cols * sizeof(uint16_t);
    CU(cudaEventRecord(h->ev[EV_COPY0], h->stream));
    CU(cudaMemcpyAsync(h->squares, host_squares, bytes, cudaMemcpyHostToDevice, h->stream));
    CU(cudaEventRecord(h->ev[EV_COPY1], h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaEventElapsedTime(&h->h2d_ms, h->ev[EV_COPY0], h->ev[EV_COPY1]));
    h->timing.h2d_ms = h->h2d_ms;
    return LAB_OK;
}

int lab_grid_download(lab_grid *h, uint16_t *host_squares) {
    Device_guard const guard_(h);
    if (!h || !host_squares) {
        return fail(LAB_ERR_ARG, "null argument");
    }
    size_t const bytes = static_cast<size_t>(h->g.rows) * h->g.cols * sizeof(uint16_t);
    CU(cudaEventRecord(h->ev[EV_COPY0], h->stream));
    CU(cudaMemcpyAsync(host_squares, h->squares, bytes, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaEventRecord(h->ev[EV_COPY1], h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaEventElapsedTime(&h->d2h_ms, h->ev[EV_COPY0], h->ev[EV_COPY1]));
    h->timing.d2h_ms = h->d2h_ms;
    return LAB_OK;
}

int lab_grid_or_bits(lab_grid *h, const lab_point *cells, const uint16_t *bits, int n) {
    Device_guard const guard_(h);
    if (!h || !cells || !bits || n < 0 || n > 32) {
        return fail(LAB_ERR_ARG, "bad or_bits arguments");
    }
    int32_t rc[64];
    for (int i = 0; i < n; i++) {
        if (!in_grid(h, cells[i])) {
            return fail(LAB_ERR_ARG, "square (%d,%d) outside the maze", cells[i].row, cells[i].col);
        }
        rc[2 * i] = cells[i].row;
        rc[2 * i + 1] = cells[i].col;
    }
    if (n == 0) {
        return LAB_OK;
    }
    // stream-ordered staging copies from pageable memory complete before returning
    CU(cudaMemcpyAsync(h->d_cells, rc, sizeof(int32_t) * 2 * n, cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_bits, bits, sizeof(uint16_t) * n, cudaMemcpyHostToDevice, h->stream));
    k_or_bits<<<1, 32, 0, h->stream>>>(h->squares, h->g.cols, h->d_cells, h->d_bits, n);
    TRACE("k_or_bits", h->stream);
    CU(cudaGetLastError());
    return LAB_OK;
}

int lab_grid_and_bits(lab_grid *h, uint16_t mask) {
    Device_guard const guard_(h);
    if (!h) {
        return fail(LAB_ERR_ARG, "null grid");
    }
    k_and_bits<<<grid_blocks(h), 256, 0, h->stream>>>(h->squares, static_cast<long long>(h->g.rows) * h->g.cols, mask);
    TRACE("k_and_bits", h->stream);
    CU(cudaGetLastError());
    return LAB_OK;
}

lab_point lab_distance_start(int rows, int cols) {
    int const row_mid = rows / 2, col_mid = cols / 2;
    return {row_mid + 1 - (row_mid % 2), col_mid + 1 - (col_mid % 2)}; // painters/distance.cc:131-134
}

void *lab_levels_device(lab_grid *h, int plane) {
    Device_guard const guard_(h);
    return (h && plane >= 0 && plane < MAX_PLANES) ? h->plane[plane].dist : nullptr;
}

// The level field of the last search's plane, copied to the host (what the animated variants replay their frames from).
// An open room (csrc/device/room.cuh) never materialises its field: LAB_ERR_ARG, the levels are Manhattan distances.
int lab_levels_download(lab_grid *h, int plane, uint32_t *levels_host) {
    Device_guard const guard_(h);
    if (!h || plane < 0 || plane >= MAX_PLANES || !levels_host || !h->plane[plane].dist) {
        return fail(LAB_ERR_ARG, "lab_levels_download: bad argument, or no such field");
    }
    if (h->tree_last.used == 2u) {
        return fail(LAB_ERR_ARG, "lab_levels_download: the last search was an open room (levels are Manhattan distances from the start; no field was written)");
    }
    if (h->field_skipped) {
        return fail(LAB_ERR_ARG, "lab_levels_download: the last game left no field (bfs-gather without paths with LAB_OPT_SOLVER_LEVELS off)");
    }
    CU(cudaMemcpyAsync(levels_host, h->plane[plane].dist, static_cast<size_t>(h->g.rows) * h->g.cols * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LAB_OK;
}

int lab_levels_bind(lab_grid *h, int plane, void *device_levels) {
    Device_guard const guard_(h);
    if (!h || plane < 0 || plane >= MAX_PLANES || !device_levels) {
        return fail(LAB_ERR_ARG, "bad levels_bind arguments");
    }
    if (h->owns_dist[plane] && h->plane[plane].dist) {
        CU(cudaFree(h->plane[plane].dist));
    }
    h->plane[plane].dist = static_cast<uint32_t *>(device_levels);
    h->owns_dist[plane] = false;
    if (plane == 0) {
        h->have_map = false; // (plane 0 now points at memory no search of this grid has written)
    }
    return LAB_OK;
}

int lab_distance_map(lab_grid *h, int sr, int sc, uint32_t *dist_host, uint64_t *max_out, uint64_t *settled_out) {
    Device_guard const guard_(h);
    if (!h) {
        return fail(LAB_ERR_ARG, "null grid");
    }
    if (!in_grid(h, {sr, sc})) {
        return fail(LAB_ERR_ARG, "start (%d,%d) outside the maze", sr, sc);
    }
    h->have_map = false; // (until this search has finished: a failed one leaves no map to paint)
    Run_desc d{};
    d.game = GAME_DISTANCE;
    d.nplanes = 1;
    for (int k = 0; k < MAX_PLANES; k++) {
        d.src_r[k] = d.src_c[k] = -1;
    }
    d.src_r[0] = sr;
    d.src_c[0] = sc;
    d.n_targets = 0;
    d.bound_mode = BOUND_NONE;
    // the lattice road makes no pass over the Squares of its own: pack sets the measure bits in advance (measure_body)
    h->band_mode = false;
    d.spec_mark = lattice_may_run(h, d) ? SQ_MEASURE : 0u;
    // passable <=> path_bit set and measure bit clear (distance.cc:145-148)
    int rc = run_search(h, d, SQ_PATH | SQ_MEASURE, SQ_PATH);
    if (rc) {
        return rc;
    }
    cudaStream_t s = h->stream;
    Params p = make_params(h, 1);
    k_resolve<<<1, 1, 0, s>>>(p, d, h->res, h->d_finish_rc);
    TRACE("k_resolve", s);
    k_max_level<<<grid_blocks(h), 256, 0, s>>>(h->g, h->plane[0].dist, h->ctl, h->res);
    TRACE("k_max_level", s);
    k_measure<<<blocks_for(h, static_cast<long long>(h->g.rows) * ((h->g.tiles_x + 3) / 4), 256), 256, 0, s>>>(h->g, h->squares, h->plane[0].visited, h->path, h->ctl, h->res);
    TRACE("k_measure", s);
    h->launches += 3;
    CU(cudaGetLastError());
    rc = finish_run(h);
    if (rc) {
        return rc;
    }
    h->have_map = true;
    if (max_out) *max_out = h->last.max_level;
    if (settled_out) *settled_out = h->last.settled;
    if (dist_host) {
        size_t const bytes = static_cast<size_t>(h->g.rows) * h->g.cols * sizeof(uint32_t);
        CU(cudaEventRecord(h->ev[EV_COPY0], s));
        CU(cudaMemcpyAsync(dist_host, h->plane[0].dist, bytes, cudaMemcpyDeviceToHost, s));
        CU(cudaEventRecord(h->ev[EV_COPY1], s));
        CU(cudaStreamSynchronize(s));
        CU(cudaEventElapsedTime(&h->timing.d2h_ms, h->ev[EV_COPY0], h->ev[EV_COPY1]));
    }
    return LAB_OK;
}

int lab_distance_paint(lab_grid *h, int channel, uint32_t *pixels_host) {
    Device_guard const guard_(h);
    if (!h) {
        return fail(LAB_ERR_ARG, "null grid");
    }
    if (channel < 0 || channel > 2) {
        return fail(LAB_ERR_ARG, "colour channel %d (0 red, 1 green, 2 blue)", channel);
    }
    if (!h->have_map) {
        return fail(LAB_ERR_ARG, "no distance map to paint: call lab_distance_map first");
    }
    size_t const bytes = static_cast<size_t>(h->g.rows) * h->g.cols * sizeof(uint32_t);
    if (!h->pixels) {
        CU(cudaMalloc(&h->pixels, bytes));
        h->owns_pixels = true;
    }
    cudaStream_t s = h->stream;
    k_heatmap<<<grid_blocks(h), 256, 0, s>>>(h->g, h->plane[0].dist, h->res, channel, h->pixels);
    TRACE("k_heatmap", s);
    CU(cudaGetLastError());
    if (pixels_host) {
        CU(cudaMemcpyAsync(pixels_host, h->pixels, bytes, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
    return LAB_OK;
}

// Runs::paint_runs (painters/runs.cc:130-161): the search is the distance map's (same start, same measure bits: :139,152);
// the run lengths follow from its level field on one tree.
int lab_runs_map(lab_grid *h, int sr, int sc, uint32_t *runs_host, uint64_t *max_out, uint64_t *settled_out) {
    Device_guard const guard_(h);
    if (!h) {
        return fail(LAB_ERR_ARG, "null grid");
    }
    h->have_runs = false;
    uint64_t settled = 0;
    int rc = lab_distance_map(h, sr, sc, nullptr, nullptr, &settled);
    if (rc) {
        return rc;
    }
    if (h->tree_last.used != 1u && h->tree_last.used != 3u) {
        return fail(LAB_ERR_ARG, "lab_runs_map: the maze is not one tree (the run of a square needs its parent in the reference's FIFO search, "
                                 "which only a tree's level field names)");
    }
    size_t const cells = static_cast<size_t>(h->g.rows) * h->g.cols;
    if (!h->runs) {
        CU(cudaMalloc(&h->runs, cells * sizeof(uint32_t)));
        CU(cudaMalloc(&h->runs_res, sizeof(Result)));
    }
    cudaStream_t s = h->stream;
    CU(cudaMemsetAsync(h->runs_res, 0, sizeof(Result), s));
    k_runs<<<grid_blocks(h), 256, 0, s>>>(h->g, h->plane[0].dist, h->runs, h->runs_res);
    TRACE("k_runs", s);
    CU(cudaGetLastError());
    Result r{};
    CU(cudaMemcpyAsync(&r, h->runs_res, sizeof r, cudaMemcpyDeviceToHost, s));
    if (runs_host) {
        CU(cudaMemcpyAsync(runs_host, h->runs, cells * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
    h->have_runs = true;
    if (max_out) *max_out = r.max_level;
    if (settled_out) *settled_out = settled;
    return LAB_OK;
}

// The runs painter's colours (painters/runs.cc:45-70: the distance painter's arithmetic on run lengths).
int lab_runs_paint(lab_grid *h, int channel, uint32_t *pixels_host) {
    Device_guard const guard_(h);
    if (!h) {
        return fail(LAB_ERR_ARG, "null grid");
    }
    if (channel < 0 || channel > 2) {
        return fail(LAB_ERR_ARG, "colour channel %d (0 red, 1 green, 2 blue)", channel);
    }
    if (!h->have_runs) {
        return fail(LAB_ERR_ARG, "no run map to paint: call lab_runs_map first");
    }
    size_t const bytes = static_cast<size_t>(h->g.rows) * h->g.cols * sizeof(uint32_t);
    if (!h->pixels) {
        CU(cudaMalloc(&h->pixels, bytes));
        h->owns_pixels = true;
    }
    cudaStream_t s = h->stream;
    k_heatmap<<<grid_blocks(h), 256, 0, s>>>(h->g, h->runs, h->runs_res, channel, h->pixels);
    TRACE("k_heatmap", s);
    CU(cudaGetLastError());
    if (pixels_host) {
        CU(cudaMemcpyAsync(pixels_host, h->pixels, bytes, cudaMemcpyDeviceToHost, s));
    }
    CU(cudaStreamSynchronize(s));
    return LAB_OK;
}

void *lab_pixels_device(lab_grid *h) { return h ? h->pixels : nullptr; }

int lab_pixels_bind(lab_grid *h, void *device_pixels) {
    Device_guard const guard_(h);
    if (!h || !device_pixels) {
        return fail(LAB_ERR_ARG, "null grid or pixel buffer");
    }
    if (h->owns_pixels && h->pixels) {
        CU(cudaFree(h->pixels));
    }
    h->pixels = static_cast<uint32_t *>(device_pixels);
    h->owns_pixels = false;
    return LAB_OK;
}

static int solver_common(lab_grid *h, Run_desc const &d, lab_point const *finishes, int nfinish, bool want_paths,
                         uint64_t path_cap) {
    h->have_map = false; // (plane 0 is about to be reused, or -- in an open room -- not written at all)
    int32_t frc[8] = {0};
    for (int i = 0; i < nfinish; i++) {
        frc[2 * i] = finishes[i].row;
        frc[2 * i + 1] = finishes[i].col;
    }
    cudaStream_t s = h->stream;
    CU(cudaMemcpyAsync(h->d_finish_rc, frc, sizeof frc, cudaMemcpyHostToDevice, s));
    int rc = LAB_OK;
    h->hunt_out = nullptr;
    h->hunt_out_cap = 0;
    h->skip_field_now = d.game == GAME_GATHER && !want_paths && !h->opt_solver_levels;
    h->field_skipped = false;
    if (d.game == GAME_GATHER) {
        CU(cudaMemsetAsync(h->d_front_rc, 0xFF, 8 * sizeof(int32_t), s)); // (before the search: the lattice road may fill it in)
    }
    if ((d.game == GAME_HUNT || d.game == GAME_GATHER) && want_paths) { // (the lattice road writes the paths itself: it needs the buffer before the search)
        rc = ensure_paths(h, path_cap, d.game == GAME_GATHER ? 4 : 1);
        if (rc) {
            return rc;
        }
        h->hunt_out = h->d_paths;
        h->hunt_out_cap = path_cap;
    }
    rc = run_search(h, d, SQ_PATH, SQ_PATH); // passable <=> path_bit (bfs_threads.cc:71-72)
    if (rc) {
        return rc;
    }
    Params p = make_params(h, d.nplanes);
    k_resolve<<<1, 1, 0, s>>>(p, d, h->res, h->d_finish_rc);
    TRACE("k_resolve", s);
    k_paint<<<grid_blocks(h), 256, 0, s>>>(p, h->squares, h->res);
    TRACE("k_paint", s);
    h->launches += 2 + (d.game == GAME_GATHER ? (want_paths ? 3 : 2) : 1);
    if (d.game == GAME_GATHER) {
        k_walk<<<4, 32, 0, s>>>(p, d.game, h->squares, h->res, h->d_finish_rc, h->d_front_rc, 1, 1, 1);
        TRACE("k_walk", s);
        k_gather_fixup<<<1, 1, 0, s>>>(h->g, h->squares, h->res, h->d_finish_rc, h->d_front_rc);
        TRACE("k_gather_fixup", s);
        if (want_paths) {
            rc = ensure_paths(h, path_cap, 4);
            if (rc) {
                return rc;
            }
            k_walk<<<4, 32, 0, s>>>(p, d.game, h->squares, h->res, h->d_finish_rc, h->d_paths, h->path_cap, path_cap, 0);
            TRACE("k_walk", s);
        }
    } else {
        // hunt repaints the winner's path (bfs_threads.cc:263-274), so it always walks;
        // corners walks only when the caller wants the path (static corners never repaints)
        if (d.game == GAME_HUNT || want_paths) {
            if (want_paths) {
                rc = ensure_paths(h, path_cap, 1);
                if (rc) {
                    return rc;
                }
            }
            k_walk<<<1, 32, 0, s>>>(p, d.game, h->squares, h->res, h->d_finish_rc, want_paths ? h->d_paths : nullptr,
                                    h->path_cap, want_paths ? path_cap : 0, 0);
            TRACE("k_walk", s);
        } else {
            k_walk<<<1, 32, 0, s>>>(p, d.game, h->squares, h->res, h->d_finish_rc, nullptr, 0, 0, 1);
            TRACE("k_walk", s);
        }
    }
    CU(cudaGetLastError());
    rc = finish_run(h);
    h->field_skipped = rc == LAB_OK && h->skip_field_now && h->tree_last.used == 3u && env_ll("LAB_LAT_FUSE_PAINT", 1) != 0;
    return rc;
}

static int copy_path(lab_grid *h, int slot, lab_point *dst, uint64_t cap, uint64_t len) {
    if (!dst || len == 0) {
        return LAB_OK;
    }
    uint64_t const n = std::min<uint64_t>(cap, len);
    static_assert(sizeof(lab_point) == 2 * sizeof(int32_t), "lab_point layout");
    CU(cudaMemcpyAsync(dst, h->d_paths + static_cast<size_t>(slot) * h->path_cap * 2, n * sizeof(lab_point),
                       cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LAB_OK;
}

int lab_bfs_hunt(lab_grid *h, lab_point start, lab_point finish, int32_t *winner, lab_point *path, uint64_t path_cap,
                 uint64_t *path_len) {
    Device_guard const guard_(h);
    if (!h || !in_grid(h, start) || !in_grid(h, finish)) {
        return fail(LAB_ERR_ARG, "bad hunt arguments");
    }
    Run_desc d{};
    d.game = GAME_HUNT;
    d.nplanes = 1;
    for (int k = 0; k < MAX_PLANES; k++) {
        d.src_r[k] = d.src_c[k] = -1;
    }
    d.src_r[0] = start.row;
    d.src_c[0] = start.col;
    d.n_targets = 1;
    d.tgt_plane[0] = 0;
    d.tgt_r[0] = finish.row;
    d.tgt_c[0] = finish.col;
    d.bound_mode = BOUND_MIN_ANY;
    bool const want = path != nullptr && path_cap > 0;
    int rc = solver_common(h, d, &finish, 1, want, path_cap);
    if (rc) {
        return rc;
    }
    if (winner) *winner = h->last.winner;
    if (path_len) *path_len = h->last.path_len[0];
    return want ? copy_path(h, 0, path, path_cap, h->last.path_len[0]) : LAB_OK;
}

int lab_bfs_gather(lab_grid *h, lab_point start, const lab_point finishes[4], int32_t claimed_by[4],
                   lab_point *const paths[4], uint64_t path_cap, uint64_t path_len[4]) {
    Device_guard const guard_(h);
    if (!h || !finishes || !in_grid(h, start)) {
        return fail(LAB_ERR_ARG, "bad gather arguments");
    }
    Run_desc d{};
    d.game = GAME_GATHER;
    d.nplanes = 1;
    for (int k = 0; k < MAX_PLANES; k++) {
        d.src_r[k] = d.src_c[k] = -1;
    }
    d.src_r[0] = start.row;
    d.src_c[0] = start.col;
    d.n_targets = 4;
    for (int j = 0; j < 4; j++) {
        if (!in_grid(h, finishes[j])) {
            return fail(LAB_ERR_ARG, "finish %d outside the maze", j);
        }
        d.tgt_plane[j] = 0;
        d.tgt_r[j] = finishes[j].row;
        d.tgt_c[j] = finishes[j].col;
    }
    d.bound_mode = BOUND_MAX_ALL;
    bool const want = paths != nullptr && path_cap > 0;
    int rc = solver_common(h, d, finishes, 4, want, path_cap);
    if (rc) {
        return rc;
    }
    for (int k = 0; k < 4; k++) {
        if (claimed_by) claimed_by[k] = h->last.claimed_by[k];
        if (path_len) path_len[k] = h->last.path_len[k];
        if (want && paths[k]) {
            rc = copy_path(h, k, paths[k], path_cap, h->last.path_len[k]);
            if (rc) {
                return rc;
            }
        }
    }
    return LAB_OK;
}

int lab_bfs_corners(lab_grid *h, const lab_point starts[4], lab_point finish, int32_t *winner, lab_point *path,
                    uint64_t path_cap, uint64_t *path_len) {
    Device_guard const guard_(h);
    if (!h || !starts || !in_grid(h, finish)) {
        return fail(LAB_ERR_ARG, "bad corners arguments");
    }
    Run_desc d{};
    d.game = GAME_CORNERS;
    d.nplanes = 4;
    d.n_targets = 4;
    for (int k = 0; k < 4; k++) {
        if (!in_grid(h, starts[k])) {
            return fail(LAB_ERR_ARG, "start %d outside the maze", k);
        }
        d.src_r[k] = starts[k].row;
        d.src_c[k] = starts[k].col;
        d.tgt_plane[k] = k;
        d.tgt_r[k] = finish.row;
        d.tgt_c[k] = finish.col;
    }
    d.bound_mode = BOUND_MIN_ANY;
    bool const want = path != nullptr && path_cap > 0;
    int rc = solver_common(h, d, &finish, 1, want, path_cap);
    if (rc) {
        return rc;
    }
    if (winner) *winner = h->last.winner;
    if (path_len) *path_len = h->last.path_len[0];
    return want ? copy_path(h, 0, path, path_cap, h->last.path_len[0]) : LAB_OK;
}

// ---- row bands -------------------------------------------------------------------------------
int lab_band_cols_padded(const lab_grid *h) { return h ? h->g.tiles_x * TS : 0; }

int lab_band_begin(lab_grid *h, int src_r, int src_c) { return lab_band_begin_game(h, GAME_DISTANCE, src_r, src_c); }
int lab_band_begin_game(lab_grid *h, int game, int src_r, int src_c) { return lab_band_begin_lattice(h, game, src_r, src_c, 0); }

// lattice_next != 0: the caller will run the lattice road across bands (lab_band_lat_link .. lab_band_lat_levels) next.  Its
// last step writes every square of the band's level field, or wipes it if the maze is no lattice tree, so the field is
// not cleared here; and for the distance map pack sets the measure bits on the way (see measure_body).
static int band_begin_common(lab_grid *h, Run_desc const &d, int lattice_next) {
    h->have_map = false; // (a band's field is not a map lab_distance_paint could colour: its maximum is the band's)
    h->band_desc = d;
    h->band_skip_field_reset = lattice_next != 0;
    if (!h->ghost_n) {
        CU(cudaMalloc(&h->ghost_n, static_cast<size_t>(h->g.tiles_x) * sizeof(uint64_t)));
        CU(cudaMalloc(&h->ghost_s, static_cast<size_t>(h->g.tiles_x) * sizeof(uint64_t)));
        CU(cudaMalloc(&h->d_activated, sizeof(uint32_t)));
    }
    h->band_mode = true;
    h->tree_last = Tree_ctl{};
    h->tree_tried = false;
    // passable: the distance map skips measured squares (distance.cc:145-148), the solvers look at path_bit alone
    int const rc = d.game == GAME_DISTANCE ? search_prepare(h, d, SQ_PATH | SQ_MEASURE, SQ_PATH) : search_prepare(h, d, SQ_PATH, SQ_PATH);
    h->band_open = rc == LAB_OK;
    return rc;
}

int lab_band_begin_lattice(lab_grid *h, int game, int src_r, int src_c, int lattice_next) {
    Device_guard const guard_(h);
    if (!h) {
        return fail(LAB_ERR_ARG, "null grid");
    }
    if (game != GAME_DISTANCE && game != GAME_HUNT && game != GAME_GATHER) {
        return fail(LAB_ERR_ARG, "sharded game %d (0 distance, 1 hunt, 2 gather; corners: lab_band_begin_corners)", game);
    }
    if (src_r >= 0 && !in_grid(h, {src_r, src_c})) {
        return fail(LAB_ERR_ARG, "source (%d,%d) outside the band", src_r, src_c);
    }
    Run_desc d{};
    d.game = game;
    d.nplanes = 1;
    for (int k = 0; k < MAX_PLANES; k++) {
        d.src_r[k] = d.src_c[k] = -1;
    }
    if (src_r >= 0) {
        d.src_r[0] = src_r;
        d.src_c[0] = src_c;
    }
    d.bound_mode = BOUND_NONE;
    d.spec_mark = (lattice_next && game == GAME_DISTANCE) ? SQ_MEASURE : 0u;
    return band_begin_common(h, d, lattice_next);
}

// Bfs::corners across row bands (bfs_threads.cc:420-453): FOUR fields searched at once, colour k from its own corner.
// src_rc: (row, col) of colour k's start in band rows, (-1, -1) for a start in another band.  The fields are searched to
// their fixpoint (no bound): the canonical model paints {d_k < D}, D = the lowest level any colour reaches the finish at
// (resolve_body), which the caller takes from lab_band_level_at_plane once the bands have converged.
int lab_band_begin_corners(lab_grid *h, const int32_t *src_rc) {
    Device_guard const guard_(h);
    if (!h || !src_rc) {
        return fail(LAB_ERR_ARG, "lab_band_begin_corners: null argument");
    }
    Run_desc d{};
    d.game = GAME_CORNERS;
    d.nplanes = 4;
    for (int k = 0; k < MAX_PLANES; k++) {
        d.src_r[k] = d.src_c[k] = -1;
    }
    for (int k = 0; k < 4; k++) {
        if (src_rc[2 * k] >= 0) {
            if (!in_grid(h, {src_rc[2 * k], src_rc[2 * k + 1]})) {
                return fail(LAB_ERR_ARG, "corner start %d (%d,%d) outside the band", k, src_rc[2 * k], src_rc[2 * k + 1]);
            }
            d.src_r[k] = src_rc[2 * k];
            d.src_c[k] = src_rc[2 * k + 1];
        }
    }
    d.bound_mode = BOUND_NONE;
    return band_begin_common(h, d, 0);
}

int lab_band_planes(const lab_grid *h) { return (h && h->band_open) ? h->band_desc.nplanes : 0; }

int lab_band_path_rows(lab_grid *h, void *top_dev, void *bottom_dev) {
    Device_guard const guard_(h);
    if (!h || !h->band_open || !top_dev || !bottom_dev) {
        return fail(LAB_ERR_ARG, "lab_band_path_rows: bad arguments or no band search open");
    }
    int const n = h->g.tiles_x * TS;
    k_band_path_rows<<<(n + 255) / 256, 256, 0, h->stream>>>(h->g, h->path, static_cast<uint8_t *>(top_dev), static_cast<uint8_t *>(bottom_dev));
    TRACE("k_band_path_rows", h->stream);
    h->launches += 1;
    CU(cudaGetLastError());
    return LAB_OK;
}

int lab_band_set_neighbour_paths(lab_grid *h, const void *above_bottom_dev, const void *below_top_dev) {
    Device_guard const guard_(h);
    if (!h || !h->band_open) {
        return fail(LAB_ERR_ARG, "no band search open");
    }
    if (below_top_dev && h->g.rows % TS != 0) {
        return fail(LAB_ERR_ARG, "a band with a neighbour below must have a multiple of %d rows (has %d)", TS, h->g.rows);
    }
    k_band_ghost_words<<<blocks_for(h, h->g.tiles_x, 256), 256, 0, h->stream>>>(
        h->g, static_cast<uint8_t const *>(above_bottom_dev), static_cast<uint8_t const *>(below_top_dev), h->ghost_n, h->ghost_s);
    TRACE("k_band_ghost_words", h->stream);
    h->launches += 1;
    int const rc = search_arm(h, h->band_desc, h->ghost_n, h->ghost_s);
    if (rc) return rc;
    CU(cudaEventRecord(h->ev[EV_SEARCH0], h->stream));
    return LAB_OK;
}

int lab_band_relax(lab_grid *h) {
    Device_guard const guard_(h);
    if (!h || !h->band_open) {
        return fail(LAB_ERR_ARG, "no band search open");
    }
    return search_launch(h, h->band_desc.nplanes);
}

int lab_band_export_edges(lab_grid *h, void *top_levels_dev, void *bottom_levels_dev) {
    Device_guard const guard_(h);
    if (!h || !h->band_open || !top_levels_dev || !bottom_levels_dev) {
        return fail(LAB_ERR_ARG, "lab_band_export_edges: bad arguments");
    }
    int const n = h->g.tiles_x * TS;
    for (int k = 0; k < h->band_desc.nplanes; k++) { // (field k's rows at [k * n, (k + 1) * n))
        k_band_export_edges<<<(n + 255) / 256, 256, 0, h->stream>>>(h->g, h->plane[k].edges, static_cast<uint32_t *>(top_levels_dev) + static_cast<size_t>(k) * n,
                                                                   static_cast<uint32_t *>(bottom_levels_dev) + static_cast<size_t>(k) * n);
        TRACE("k_band_export_edges", h->stream);
        h->launches += 1;
    }
    CU(cudaGetLastError());
    return LAB_OK;
}

int lab_band_import_edges(lab_grid *h, const void *above_bottom_levels_dev, const void *below_top_levels_dev, uint32_t *activated) {
    Device_guard const guard_(h);
    if (!h || !h->band_open || !activated) {
        return fail(LAB_ERR_ARG, "lab_band_import_edges: bad arguments");
    }
    cudaStream_t s = h->stream;
    int const planes = h->band_desc.nplanes;
    size_t const n = static_cast<size_t>(h->g.tiles_x) * TS;
    Params p = make_params(h, planes);
    CU(cudaMemsetAsync(h->d_activated, 0, sizeof(uint32_t), s));
    k_band_rearm<<<1, 1, 0, s>>>(p, budget_ns());
    for (int k = 0; k < planes; k++) {
        uint32_t const *above = static_cast<uint32_t const *>(above_bottom_levels_dev), *below = static_cast<uint32_t const *>(below_top_levels_dev);
        k_band_import_edges<<<blocks_for(h, h->g.tiles_x, 256), 256, 0, s>>>(p, k, above ? above + k * n : nullptr, below ? below + k * n : nullptr, h->d_activated);
        TRACE("k_band_import_edges", s);
    }
    h->launches += 1 + planes;
    CU(cudaMemcpyAsync(activated, h->d_activated, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    return LAB_OK;
}

int lab_band_finish(lab_grid *h, uint64_t *local_max, uint64_t *local_settled) {
    Device_guard const guard_(h);
    if (!h || !h->band_open) {
        return fail(LAB_ERR_ARG, "no band search open");
    }
    cudaStream_t s = h->stream;
    CU(cudaEventRecord(h->ev[EV_POST0], s));
    Params p = make_params(h, 1);
    k_resolve<<<1, 1, 0, s>>>(p, h->band_desc, h->res, h->d_finish_rc);
    k_max_level<<<grid_blocks(h), 256, 0, s>>>(h->g, h->plane[0].dist, h->ctl, h->res);
    k_measure<<<blocks_for(h, static_cast<long long>(h->g.rows) * ((h->g.tiles_x + 3) / 4), 256), 256, 0, s>>>(h->g, h->squares, h->plane[0].visited, h->path, h->ctl, h->res);
    h->launches += 3;
    CU(cudaGetLastError());
    h->band_open = false;
    int const rc = finish_run(h);
    if (rc) return rc;
    if (local_max) *local_max = h->last.max_level;
    if (local_settled) *local_settled = h->last.settled;
    return LAB_OK;
}

// ---- the lattice road across row bands (device/lattice.cuh; protocol in include/labyrinth_b200.h) ----
int lab_band_lat_super_rows(void) { return LST; }
int lab_band_lat_words(int tiles_x, int super_rows_global, uint64_t *border_entries, uint64_t *weight_words, uint64_t *chunk_words) {
    if (tiles_x < 1 || super_rows_global < 1) {
        return fail(LAB_ERR_ARG, "lab_band_lat_words: bad argument");
    }
    unsigned long long const sup_x = (static_cast<unsigned long long>(tiles_x) + LST - 1) / LST;
    unsigned long long const nslots = static_cast<unsigned long long>(super_rows_global) * LST * tiles_x * LSLOTS;
    if (nslots + 2 + LAT_TAIL >= (1ull << 32)) {
        return fail(LAB_ERR_ARG, "lab_band_lat_words: maze too large");
    }
    if (border_entries) *border_entries = static_cast<unsigned long long>(super_rows_global) * sup_x * LB;
    if (weight_words) *weight_words = nslots + 2 + LAT_TAIL + L_CHUNK;
    if (chunk_words) *chunk_words = nslots / L_CHUNK + 2;
    return LAB_OK;
}

int lab_band_lat_bind(lab_grid *h, int super_row_base, int super_rows_band, int super_rows_global, void *border_list, void *weights, void *chunks) {
    Device_guard const guard_(h);
    if (!h || !border_list || !weights || !chunks || super_row_base < 0 || super_rows_band < 1 || super_row_base + super_rows_band > super_rows_global ||
        (reinterpret_cast<uintptr_t>(border_list) & 7u)) {
        return fail(LAB_ERR_ARG, "lab_band_lat_bind: bad argument");
    }
    if (h->g.tiles_y > super_rows_band * LST) {
        return fail(LAB_ERR_ARG, "lab_band_lat_bind: the band has %d tile rows, more than %d super-tile rows hold", h->g.tiles_y, super_rows_band);
    }
    int rc = ensure_lattice(h); // (the per-tile arrays; its own border list and weights stay unused here)
    if (rc) return rc;
    if (!h->band_lat_tc) {
        CU(cudaMalloc(&h->band_lat_tc, sizeof(Tree_ctl)));
    }
    Lattice lt = h->lat;
    lt.tc = h->band_lat_tc;
    lt.sup_y = super_rows_band; // (the last band may have fewer tile rows than that: its missing tiles count as empty)
    lt.nsuper = lt.sup_y * lt.sup_x;
    lt.tile_base = static_cast<uint32_t>(super_row_base) * LST * static_cast<uint32_t>(h->g.tiles_x);
    lt.sup_base = static_cast<uint32_t>(super_row_base) * static_cast<uint32_t>(lt.sup_x);
    lt.nsuper_global = super_rows_global * lt.sup_x;
    lt.nslots_global = static_cast<uint32_t>(static_cast<unsigned long long>(super_rows_global) * LST * h->g.tiles_x * LSLOTS);
    lt.gb = static_cast<uint64_t *>(border_list);
    lt.wsum = static_cast<uint32_t *>(weights);
    lt.chunk = static_cast<uint32_t *>(chunks);
    lt.peers = 1;
    lt.gb_peer[0] = lt.gb;
    lt.wsum_peer[0] = lt.wsum;
    h->band_lat = lt;
    h->band_lat_bound = true;
    return LAB_OK;
}

// Peer mode: the border-list slices and the weights are stored straight into every rank's copy (over NVLink) from inside
// the kernels that make them; [0] must be this rank's own arrays.
int lab_band_lat_set_peers(lab_grid *h, int n, void *const *border_lists, void *const *weights) {
    Device_guard const guard_(h);
    if (!h || !h->band_lat_bound || n < 1 || n > 8 || !border_lists || !weights || border_lists[0] != h->band_lat.gb || weights[0] != h->band_lat.wsum) {
        return fail(LAB_ERR_ARG, "lab_band_lat_set_peers: bad argument");
    }
    h->band_lat.peers = n;
    for (int q = 0; q < n; q++) {
        h->band_lat.gb_peer[q] = static_cast<uint64_t *>(border_lists[q]);
        h->band_lat.wsum_peer[q] = static_cast<uint32_t *>(weights[q]);
    }
    return LAB_OK;
}

static int band_lat_ready(lab_grid *h, char const *who) {
    if (!h || !h->band_open || !h->band_lat_bound) {
        return fail(LAB_ERR_ARG, "%s: no band search open, or lab_band_lat_bind not called", who);
    }
    return LAB_OK;
}

// Phase 1 (after lab_band_set_neighbour_paths): counts, link, local ranking -> this band's slice of the border list (in
// place, or in every peer's copy) and its header.  The root must be on a cell ((odd, odd) in the whole maze).
int lab_band_lat_link(lab_grid *h, void *header_dev) {
    Device_guard const guard_(h);
    int rc = band_lat_ready(h, "lab_band_lat_link");
    if (rc) return rc;
    if (!header_dev) {
        return fail(LAB_ERR_ARG, "lab_band_lat_link: null header");
    }
    cudaStream_t s = h->stream;
    Lattice const lt = h->band_lat;
    Params p = make_params(h, 1);
    Tree tr{};
    tr.tc = lt.tc;
    tr.banded = 1;
    CU(cudaMemsetAsync(lt.tc, 0, sizeof(Tree_ctl), s));
    k_tree_count<<<std::min(blocks_for(h, h->g.ntiles, 256), h->sm_count * 8), 256, 0, s>>>(p, tr);
    k_lat_band_open<<<1, 1, 0, s>>>(lt);
    int const tile_blocks = static_cast<int>(std::min<long long>((h->g.ntiles + 3) / 4, static_cast<long long>(h->sm_count) * 16));
    k_lat_link<<<tile_blocks, 128, 0, s>>>(p, lt);
    size_t const rank_smem = LAT_RANK_WARPS * LSUP_SLOTS * sizeof(uint32_t);
    int const rank_blocks = std::min((lt.nsuper + LAT_RANK_WARPS - 1) / LAT_RANK_WARPS, h->sm_count * 6);
    k_lat_rank_local<<<rank_blocks, 32 * LAT_RANK_WARPS, rank_smem, s>>>(p, lt);
    k_lat_band_header<<<1, 1, 0, s>>>(lt, static_cast<uint32_t *>(header_dev));
    h->launches += 5;
    CU(cudaGetLastError());
    return LAB_OK;
}

// Phase 2 (the border list and the headers of all ranks are in place): the whole maze's counts, the ranking over the whole
// border list, this band's floods; its arcs' weights and its count of reached squares go into the weights array.
// zero_weights: clear the array first (the copies are summed afterwards); 0 in peer mode, where every position is
// written by the rank that owns it (the caller clears the tail words once per map).
int lab_band_lat_flood(lab_grid *h, const void *all_headers_dev, int world, int zero_weights) {
    Device_guard const guard_(h);
    int rc = band_lat_ready(h, "lab_band_lat_flood");
    if (rc) return rc;
    if (!all_headers_dev || world < 1) {
        return fail(LAB_ERR_ARG, "lab_band_lat_flood: bad argument");
    }
    cudaStream_t s = h->stream;
    Lattice const lt = h->band_lat;
    Params p = make_params(h, 1);
    if (zero_weights) {
        CU(cudaMemsetAsync(lt.wsum, 0, (static_cast<size_t>(lt.nslots_global) + 2 + LAT_TAIL) * sizeof(uint32_t), s));
    }
    k_lat_band_check<<<1, 1, 0, s>>>(lt, static_cast<uint32_t const *>(all_headers_dev), world);
    long long const nb = static_cast<long long>(lt.nsuper_global) * LB;
    k_lat_rank_global<<<static_cast<int>(std::min<long long>((nb + 255) / 256, static_cast<long long>(h->sm_count) * 8)), 256, 0, s>>>(lt);
    k_lat_rank_expand<<<static_cast<int>(std::min<long long>((static_cast<long long>(lt.nslots) + 255) / 256, static_cast<long long>(h->sm_count) * 16)), 256, 0, s>>>(lt);
    int const tile_blocks = static_cast<int>(std::min<long long>((h->g.ntiles + 3) / 4, static_cast<long long>(h->sm_count) * 16));
    k_lat_flood<<<tile_blocks, 128, 0, s>>>(p, lt);
    k_lat_band_tail<<<1, 1, 0, s>>>(lt);
    h->launches += 5;
    CU(cudaGetLastError());
    return LAB_OK;
}

// Phase 3 (the weights of all ranks are in place): prefix sums over the whole tour, this band's levels.  *used = 1: the
// band's level field is final (lab_band_finish is left); 0: not one lattice tree, nothing was written.
int lab_band_lat_levels(lab_grid *h, int32_t *used) {
    Device_guard const guard_(h);
    int rc = band_lat_ready(h, "lab_band_lat_levels");
    if (rc) return rc;
    cudaStream_t s = h->stream;
    Lattice const lt = h->band_lat;
    Params p = make_params(h, 1);
    k_lat_band_settled<<<1, 1, 0, s>>>(lt);
    k_lat_scan_chunks<<<static_cast<int>(std::min<long long>((static_cast<long long>(lt.nslots_global) / L_CHUNK + 8) / 8, static_cast<long long>(h->sm_count) * 8)), 256, 0, s>>>(lt);
    k_lat_scan_tops<<<1, 32, 0, s>>>(lt);
    int const tile_blocks = static_cast<int>(std::min<long long>((h->g.ntiles + 3) / 4, static_cast<long long>(h->sm_count) * 16));
    k_lat_levels<LAT_LEVELS><<<tile_blocks, 128, 0, s>>>(p, lt, nullptr, nullptr);
    k_lat_band_close<<<1, 1, 0, s>>>(lt);
    Tree tr{};
    tr.tc = lt.tc;
    tr.banded = 1;
    k_tree_gate<<<h->sm_count * 2, 256, 0, s>>>(p, tr);
    h->launches += 6;
    CU(cudaGetLastError());
    Tree_ctl c{};
    CU(cudaMemcpyAsync(&c, lt.tc, sizeof c, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    h->tree_last = c;
    if (used) {
        *used = c.used == 3u ? 1 : 0;
    }
    return LAB_OK;
}

// hunt's winner path on the lattice road across bands, no hand-over (bands.sharded_hunt; lattice.cuh lat_path_*): the band
// that holds the finish reads the rank of the entry arc of the finish's component and the finish's level ...
int lab_band_lat_path_rank(lab_grid *h, int r, int c, uint32_t *rank, uint32_t *level) {
    Device_guard const guard_(h);
    int rc = band_lat_ready(h, "lab_band_lat_path_rank");
    if (rc) return rc;
    if (!rank || !level || !in_grid(h, {r, c}) || ((r | c) & 1) == 0) {
        return fail(LAB_ERR_ARG, "lab_band_lat_path_rank: bad argument");
    }
    cudaStream_t s = h->stream;
    Params p = make_params(h, 1);
    uint32_t *d2 = reinterpret_cast<uint32_t *>(h->band_lat.xd); // (the local ranking's pairs: nobody reads them any more)
    k_lat_path_rank<<<1, 32, 0, s>>>(p, h->band_lat, r, c, d2);
    h->launches += 1;
    uint32_t out2[2] = {0, 0};
    CU(cudaMemcpyAsync(out2, d2, sizeof out2, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    *rank = out2[0];
    *level = out2[1];
    return LAB_OK;
}
// ... every band then marks the path's squares in its own rows: (fin_r, fin_c) / (start_r, start_c) in band rows, -1: in another
// band.  path_dev: NULL or room for path_cap (row, col) pairs, written at the squares' indices along the WHOLE path (band
// rows; other bands' squares left alone).  repaint_bits != 0 replaces the paint nibble (the winner's colour).
int lab_band_lat_path_mark(lab_grid *h, uint32_t rank, uint32_t level, int fin_r, int fin_c, int start_r, int start_c, uint32_t repaint_bits, void *path_dev,
                           uint64_t path_cap, uint64_t *segments) {
    Device_guard const guard_(h);
    int rc = band_lat_ready(h, "lab_band_lat_path_mark");
    if (rc) return rc;
    if ((fin_r >= 0 && !in_grid(h, {fin_r, fin_c})) || (start_r >= 0 && !in_grid(h, {start_r, start_c})) || level == INF) {
        return fail(LAB_ERR_ARG, "lab_band_lat_path_mark: bad argument");
    }
    cudaStream_t s = h->stream;
    Params p = make_params(h, 1);
    Lattice const lt = h->band_lat;
    uint32_t *seg = reinterpret_cast<uint32_t *>(lt.xd);
    uint32_t const seg_cap = lt.nslots;
    int32_t *out = static_cast<int32_t *>(path_dev);
    k_lat_path_setup<<<1, 1, 0, s>>>(p, lt, rank, level, fin_r, fin_c, start_r, start_c, repaint_bits, h->squares, out, path_cap, seg);
    k_lat_path_find<<<blocks_for(h, h->g.ntiles, 256), 256, 0, s>>>(p, lt, seg, seg_cap);
    k_lat_path_walk<<<h->sm_count * 8, 32, 0, s>>>(p, lt, h->res, repaint_bits, h->squares, out, 0, path_cap, seg, seg_cap);
    TRACE("k_lat_path_walk", s);
    h->launches += 3;
    CU(cudaGetLastError());
    uint32_t n = 0;
    CU(cudaMemcpyAsync(&n, &lt.tc->lat_path_nseg[0], sizeof n, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (segments) *segments = n;
    return LAB_OK;
}

// ---- the spanning-tree pipeline across row bands (device/tree.cuh; protocol in include/labyrinth_b200.h) ----
int lab_band_tree_bind(lab_grid *h, uint32_t tile_base, uint32_t ntiles_global, void *succ0, void *sr0, void *sr1,
                       void *loc, void *inmask, void *alist, const void *border_list, uint32_t border_count, const void *layout,
                       int world, int me) {
    Device_guard const guard_(h);
    if (!h || !succ0 || !sr0 || !sr1 || (reinterpret_cast<uintptr_t>(sr0) & 7u) || (reinterpret_cast<uintptr_t>(sr1) & 7u) || !loc || !inmask || !alist || !border_list || !layout || world < 1 || me < 0 || me >= world) {
        return fail(LAB_ERR_ARG, "lab_band_tree_bind: bad argument");
    }
    if (static_cast<unsigned long long>(ntiles_global) * SLOTS + 2 >= (1ull << 32) || tile_base + static_cast<uint32_t>(h->g.ntiles) > ntiles_global) {
        return fail(LAB_ERR_ARG, "lab_band_tree_bind: band tiles [%u, %u) do not fit a maze of %u tiles (or the maze is too large)", tile_base,
                    tile_base + static_cast<uint32_t>(h->g.ntiles), ntiles_global);
    }
    int const rc = tree_configure(h);
    if (rc) return rc;
    Tree &tr = h->band_tree;
    if (!tr.tc) {
        CU(cudaMalloc(&tr.tc, sizeof(Tree_ctl)));
    }
    tr.succ0 = static_cast<uint32_t *>(succ0);
    tr.sr[0] = static_cast<uint32_t *>(sr0);
    tr.sr[1] = static_cast<uint32_t *>(sr1);
    tr.loc = static_cast<uint32_t *>(loc);
    tr.inmask = static_cast<uint64_t *>(inmask);
    tr.alist = static_cast<uint32_t *>(alist);
    tr.nslots = ntiles_global * SLOTS;
    tr.tile_base = tile_base;
    tr.banded = 1;
    h->band_alist = static_cast<uint32_t *>(alist);
    h->band_border_list = static_cast<uint32_t const *>(border_list);
    h->band_border_count = border_count;
    h->band_layout = static_cast<Band_layout const *>(layout);
    h->band_world = world;
    h->band_me = me;
    h->band_ntiles_global = ntiles_global;
    h->band_tree_bound = true;
    return LAB_OK;
}

static int band_tree_ready(lab_grid *h, char const *who) {
    if (!h || !h->band_open || !h->band_tree_bound) {
        return fail(LAB_ERR_ARG, "%s: no band search open, or lab_band_tree_bind not called", who);
    }
    return LAB_OK;
}
static int band_tree_read(lab_grid *h, Tree_ctl &out) {
    CU(cudaMemcpyAsync(&out, h->band_tree.tc, sizeof(Tree_ctl), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LAB_OK;
}
// The band's own list, followed only inside the band's slots ...
static Tree band_tree_local(lab_grid *h) {
    Tree tr = h->band_tree;
    tr.mark_bits = h->band_desc.game == GAME_DISTANCE ? SQ_MEASURE : 0u; // (only the distance map measures)
    tr.alist = h->band_alist;
    tr.jump_lo = tr.tile_base * SLOTS;
    tr.jump_hi = tr.jump_lo + static_cast<uint32_t>(h->g.ntiles) * SLOTS;
    tr.jump_skip = 3u;
    return tr;
}
// ... and the short list over every band's border tile rows.
static Tree band_tree_border(lab_grid *h) {
    Tree tr = h->band_tree;
    tr.alist = const_cast<uint32_t *>(h->band_border_list);
    tr.jump_lo = tr.jump_hi = 0;
    tr.jump_skip = 3u;
    return tr;
}
// One pointer-jumping launch (rank: KIND 0, prefix: KIND 1) over `count` list entries.
extern "C++" template <int KIND> int band_tree_jump(lab_grid *h, Tree tr, Params p, uint32_t count) {
    cudaStream_t s = h->stream;
    Tree_ctl *tc = tr.tc;
    CU(cudaMemcpyAsync(&tc->n_arcs, &count, sizeof count, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(KIND == 0 ? tc->rank_more : tc->prefix_more, 0, sizeof tc->rank_more, s));
    CU(cudaMemsetAsync(&tc->barrier[KIND], 0, sizeof(uint32_t), s));
    uint32_t rounds = static_cast<uint32_t>(ceil_log2(static_cast<unsigned long long>(tr.nslots) + 2));
    int const blocks = static_cast<int>(std::max<long long>(1, std::min<long long>(h->tree_coop_blocks, (static_cast<long long>(count) + 512 * JUMP_K) / (512 * JUMP_K))));
    void *args[] = {&p, &tr, &rounds};
    CU(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(k_tree_jump<KIND>), dim3(blocks), dim3(512), args, 0, s));
    TRACE(KIND == 0 ? "k_tree_jump<rank>" : "k_tree_jump<prefix>", s);
    h->launches += 1;
    return LAB_OK;
}

int lab_band_tree_chunk_words(const lab_grid *h, int stage) {
    Device_guard const guard_(h);
    if (!h || stage < 1 || stage > 3) {
        return 0;
    }
    uint32_t const R = static_cast<uint32_t>(h->g.tiles_x) * SLOTS;
    return static_cast<int>(BORDER_HDR + (stage == 1 ? 6u * R : (stage == 2 ? 2u * R + 2u * (R / 32u) : 4u * R)));
}

int lab_band_tree_count(lab_grid *h, uint64_t *squares, uint64_t *pairs, uint32_t *crossings, int32_t box[4]) {
    Device_guard const guard_(h);
    int rc = band_tree_ready(h, "lab_band_tree_count");
    if (rc) return rc;
    cudaStream_t s = h->stream;
    Tree const tr = h->band_tree;
    Params p = make_params(h, 1);
    CU(cudaMemsetAsync(tr.tc, 0, sizeof(Tree_ctl), s));
    k_tree_count<<<std::min(blocks_for(h, h->g.ntiles, 256), h->sm_count * 8), 256, 0, s>>>(p, tr); // (few warps: seven same-address atomics each)
    TRACE("k_tree_count", s);
    h->launches += 1;
    CU(cudaGetLastError());
    Tree_ctl tc{};
    rc = band_tree_read(h, tc);
    if (rc) return rc;
    h->band_crossings = tc.crossings;
    if (squares) *squares = tc.V;
    if (pairs) *pairs = tc.E;
    if (crossings) *crossings = tc.crossings;
    if (box) { // bounding box of the band's passable squares, rows relative to the band (r0 > r1: none)
        box[0] = tc.V ? static_cast<int32_t>(0x7FFFFFFFu - tc.box_r0_inv) : 1;
        box[1] = tc.V ? static_cast<int32_t>(tc.box_r1) : 0;
        box[2] = tc.V ? static_cast<int32_t>(0x7FFFFFFFu - tc.box_c0_inv) : 1;
        box[3] = tc.V ? static_cast<int32_t>(tc.box_c1) : 0;
    }
    return LAB_OK;
}

// The ranks' counts and boxes say that the whole maze is one open room (device/room.cuh): this band's part of it
// [r0, r1] x [c0, c1] (empty if r0 > r1) and the source, all relative to the band; levels in one streaming pass.
int lab_band_room(lab_grid *h, int r0, int r1, int c0, int c1, int src_r, int src_c) {
    Device_guard const guard_(h);
    int rc = band_tree_ready(h, "lab_band_room");
    if (rc) return rc;
    cudaStream_t s = h->stream;
    Tree const tr = h->band_tree;
    Params p = make_params(h, 1);
    k_room_set<<<1, 1, 0, s>>>(h->ctl, r0, r1, c0, c1, src_r, src_c);
    if (h->band_desc.game == GAME_DISTANCE) { // (the solvers paint and walk from the closed form: no field, no measure bits)
        k_room_distance<<<h->sm_count * 8, 256, 0, s>>>(p, h->squares, h->res);
        TRACE("k_room_distance", s);
    }
    k_tree_gate<<<h->sm_count * 2, 256, 0, s>>>(p, tr);
    TRACE("k_tree_gate", s);
    h->launches += 3;
    CU(cudaGetLastError());
    return band_tree_read(h, h->tree_last);
}

// walk + the band's own part of the ranking -> chunk 1
int lab_band_tree_rank_local(lab_grid *h, uint64_t squares_global, uint64_t pairs_global, void *chunk1) {
    Device_guard const guard_(h);
    int rc = band_tree_ready(h, "lab_band_tree_rank_local");
    if (rc) return rc;
    cudaStream_t s = h->stream;
    Tree const tr = band_tree_local(h);
    Params p = make_params(h, 1);
    unsigned long long const ve[2] = {squares_global, pairs_global};
    CU(cudaMemcpyAsync(&tr.tc->V, ve, sizeof ve, cudaMemcpyHostToDevice, s)); // (V and E are the struct's first two members)
    k_tree_init<<<1, 1, 0, s>>>(p, tr);
    int const walk_blocks = std::min<long long>((h->g.ntiles + 3) / 4, static_cast<long long>(h->sm_count) * h->tree_walk_per_sm);
    k_tree_walk<<<walk_blocks, 128, 0, s>>>(p, tr);
    TRACE("k_tree_walk", s);
    h->launches += 2;
    rc = band_tree_jump<0>(h, tr, p, h->band_crossings);
    if (rc) return rc;
    k_tree_border_export<<<h->sm_count, 256, 0, s>>>(p, tr, 1, static_cast<uint32_t *>(chunk1));
    TRACE("k_tree_border_export", s);
    h->launches += 1;
    CU(cudaGetLastError());
    return LAB_OK;
}

// all chunks 1 in: the short list's ranking, the band's final ranks, orientation, the flood -> chunk 2
int lab_band_tree_flood(lab_grid *h, const void *all1, int root_rank, void *chunk2) {
    Device_guard const guard_(h);
    int rc = band_tree_ready(h, "lab_band_tree_flood");
    if (rc) return rc;
    cudaStream_t s = h->stream;
    Tree const local = band_tree_local(h), border = band_tree_border(h);
    Params p = make_params(h, 1);
    k_tree_border_import<<<h->sm_count, 256, 0, s>>>(p, local, 1, static_cast<uint32_t const *>(all1), h->band_layout, h->band_world, h->band_me, root_rank);
    TRACE("k_tree_border_import", s);
    rc = band_tree_jump<0>(h, border, p, h->band_border_count);
    if (rc) return rc;
    CU(cudaMemcpyAsync(&local.tc->n_arcs, &h->band_crossings, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    k_tree_finish_jump<<<h->sm_count * 2, 256, 0, s>>>(p, local);
    k_tree_orient<<<blocks_for(h, h->g.ntiles, 256), 256, 0, s>>>(p, local);
    TRACE("k_tree_orient", s);
    int const fill_blocks = std::min<long long>((h->g.ntiles + 3) / 4, static_cast<long long>(h->sm_count) * h->tree_flood_per_sm);
    k_tree_flood<<<fill_blocks, 128, 4 * FLOOD_U64 * sizeof(uint64_t), s>>>(p, local);
    TRACE("k_tree_flood", s);
    k_tree_border_export<<<h->sm_count, 256, 0, s>>>(p, local, 2, static_cast<uint32_t *>(chunk2));
    h->launches += 5;
    CU(cudaGetLastError());
    return LAB_OK;
}

// all chunks 2 in: parents of the band's entries, the band's own part of the prefix sums -> chunk 3
int lab_band_tree_prefix_local(lab_grid *h, const void *all2, void *chunk3) {
    Device_guard const guard_(h);
    int rc = band_tree_ready(h, "lab_band_tree_prefix_local");
    if (rc) return rc;
    cudaStream_t s = h->stream;
    Tree local = band_tree_local(h);
    Params p = make_params(h, 1);
    k_tree_border_import<<<h->sm_count, 256, 0, s>>>(p, local, 2, static_cast<uint32_t const *>(all2), h->band_layout, h->band_world, h->band_me, 0);
    TRACE("k_tree_border_import", s);
    local.jump_skip = 2u; // (the parent pass runs at the head of this launch)
    rc = band_tree_jump<1>(h, local, p, h->band_crossings);
    if (rc) return rc;
    k_tree_border_export<<<h->sm_count, 256, 0, s>>>(p, local, 3, static_cast<uint32_t *>(chunk3));
    h->launches += 2;
    CU(cudaGetLastError());
    return LAB_OK;
}

// all chunks 3 in: the short list's prefix sums, the band's final entry levels, the band's final field.
// Every rank has read the same flags, so all reach the same verdict `used`: 1 = the band's field is final;
// 0 = not one tree, the field is wiped (tree_gate_body) and the relax / exchange loop takes over.
int lab_band_tree_levels(lab_grid *h, const void *all3, int32_t *used) {
    Device_guard const guard_(h);
    int rc = band_tree_ready(h, "lab_band_tree_levels");
    if (rc) return rc;
    cudaStream_t s = h->stream;
    Tree const local = band_tree_local(h), border = band_tree_border(h);
    Params p = make_params(h, 1);
    k_tree_border_import<<<h->sm_count, 256, 0, s>>>(p, local, 3, static_cast<uint32_t const *>(all3), h->band_layout, h->band_world, h->band_me, 0);
    TRACE("k_tree_border_import", s);
    rc = band_tree_jump<1>(h, border, p, h->band_border_count);
    if (rc) return rc;
    CU(cudaMemcpyAsync(&local.tc->n_arcs, &h->band_crossings, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    k_tree_finish_jump<<<h->sm_count * 2, 256, 0, s>>>(p, local);
    k_tree_fixup<<<blocks_for(h, static_cast<long long>(h->g.rows) * ((h->g.tiles_x + 3) / 4), 256), 256, 0, s>>>(p, local, h->squares);
    TRACE("k_tree_fixup", s);
    k_tree_gate<<<h->sm_count * 2, 256, 0, s>>>(p, local);
    TRACE("k_tree_gate", s);
    h->launches += 4;
    CU(cudaGetLastError());
    rc = band_tree_read(h, h->tree_last);
    if (rc) return rc;
    if (used) *used = static_cast<int32_t>(h->tree_last.used);
    return LAB_OK;
}

// ---- the solvers across row bands: compositions of the pieces below, driven by bands.py (sharded_hunt / sharded_gather).
// The level field is the sharded spanning-tree pipeline's or the open room's (begin with lab_band_begin_game).
int lab_band_level_at_plane(lab_grid *h, int plane, int r, int c, uint32_t *level) {
    Device_guard const guard_(h);
    if (!h || !h->band_open || !level || !in_grid(h, {r, c}) || plane < 0 || plane >= h->band_desc.nplanes) {
        return fail(LAB_ERR_ARG, "lab_band_level_at: bad arguments");
    }
    CU(cudaMemcpyAsync(level, h->plane[plane].dist + static_cast<size_t>(r) * h->g.cols + c, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return LAB_OK;
}
int lab_band_level_at(lab_grid *h, int r, int c, uint32_t *level) { return lab_band_level_at_plane(h, 0, r, c, level); }

// Paint rules (post.cuh Paint_rule, plane 0): OR bits[k] into the band's squares whose level is below below[k].
int lab_band_paint(lab_grid *h, int n_rules, const uint32_t *below, const uint32_t *bits, uint64_t *painted) {
    uint32_t const plane0[4] = {0u, 0u, 0u, 0u};
    return lab_band_paint_planes(h, n_rules, plane0, below, bits, painted);
}
// ... rule k looks at field plane[k] (corners: colour k's own field)
int lab_band_paint_planes(lab_grid *h, int n_rules, const uint32_t *plane, const uint32_t *below, const uint32_t *bits, uint64_t *painted) {
    Device_guard const guard_(h);
    if (!h || !h->band_open || n_rules < 1 || n_rules > 4 || !plane || !below || !bits) {
        return fail(LAB_ERR_ARG, "lab_band_paint: bad arguments");
    }
    cudaStream_t s = h->stream;
    Result r{};
    r.winner = -1;
    r.n_rules = static_cast<uint32_t>(n_rules);
    for (int k = 0; k < n_rules; k++) {
        if (plane[k] >= static_cast<uint32_t>(h->band_desc.nplanes)) {
            return fail(LAB_ERR_ARG, "lab_band_paint: rule %d names field %u of %d", k, plane[k], h->band_desc.nplanes);
        }
        r.rule[k] = {plane[k], below[k], bits[k], 0u};
    }
    CU(cudaMemcpyAsync(h->res, &r, sizeof r, cudaMemcpyHostToDevice, s));
    Params p = make_params(h, h->band_desc.nplanes);
    k_paint<<<grid_blocks(h), 256, 0, s>>>(p, h->squares, h->res);
    TRACE("k_paint", s);
    h->launches += 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(&r, h->res, sizeof r, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (painted) *painted = r.settled;
    return LAB_OK;
}

// One piece of a canonical path inside this band: from (r, c) -- the finish (include_first = 0) or the square a walk
// taken over from the neighbouring band continues at (1) -- towards the start, until the start or the band's border.
// path_dev: NULL or room for path_cap (row, col) pairs (band rows), written from index 0 (in an open room: from the
// square's index along the whole path, other bands' squares left alone).  state: {steps, row, col, level left}.
int lab_band_walk(lab_grid *h, int r, int c, int include_first, int first_only, uint32_t repaint_bits, void *path_dev, uint64_t path_cap,
                  int64_t state[4]) {
    Device_guard const guard_(h);
    if (!h || !h->band_open || !state) {
        return fail(LAB_ERR_ARG, "lab_band_walk: bad arguments");
    }
    if (h->tree_last.used != 2u && !in_grid(h, {r, c})) { // (only an open room's closed form takes a finish outside the band)
        return fail(LAB_ERR_ARG, "lab_band_walk: (%d,%d) outside the band", r, c);
    }
    cudaStream_t s = h->stream;
    Params p = make_params(h, 1);
    k_band_walk<<<1, 32, 0, s>>>(p, r, c, h->squares, static_cast<int32_t *>(path_dev), path_cap, repaint_bits, first_only, include_first, h->res);
    TRACE("k_band_walk", s);
    h->launches += 1;
    CU(cudaGetLastError());
    Result res{};
    CU(cudaMemcpyAsync(&res, h->res, sizeof res, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    state[0] = static_cast<int64_t>(res.path_len[0]);
    state[1] = res.walk_r[0];
    state[2] = res.walk_c[0];
    state[3] = res.walk_left[0] == INF ? -1 : static_cast<int64_t>(res.walk_left[0]);
    return LAB_OK;
}

// gather's path.front() recolour (bfs_threads.cc:355-364) of one square of this band: paint nibble := bits
int lab_band_recolour(lab_grid *h, int r, int c, uint16_t paint_bits) {
    Device_guard const guard_(h);
    if (!h || !h->band_open || !in_grid(h, {r, c})) {
        return fail(LAB_ERR_ARG, "lab_band_recolour: bad arguments");
    }
    k_recolour<<<1, 1, 0, h->stream>>>(h->squares + static_cast<size_t>(r) * h->g.cols + c, paint_bits);
    h->launches += 1;
    CU(cudaGetLastError());
    return LAB_OK;
}

// closes a sharded game (the distance map closes with lab_band_finish): statistics, errors
int lab_band_finish_game(lab_grid *h) {
    Device_guard const guard_(h);
    if (!h || !h->band_open) {
        return fail(LAB_ERR_ARG, "no band search open");
    }
    CU(cudaEventRecord(h->ev[EV_POST0], h->stream));
    h->band_open = false;
    return finish_run(h);
}

uint64_t lab_last_painted(const lab_grid *h) { return h ? h->last.settled : 0; }

int lab_last_timing(const lab_grid *h, lab_timing *out) {
    Device_guard const guard_(h);
    if (!h || !out) {
        return fail(LAB_ERR_ARG, "null argument");
    }
    *out = h->timing;
    return LAB_OK;
}

int lab_last_stats(const lab_grid *h, lab_stats *out) {
    Device_guard const guard_(h);
    if (!h || !out) {
        return fail(LAB_ERR_ARG, "null argument");
    }
    *out = h->stats;
    return LAB_OK;
}

// ---- host helpers ---------------------------------------------------------------------------
int lab_pick_random_point(int rows, int cols, const uint16_t *squares, uint32_t seed, lab_point *out) {
    if (!squares || !out || rows < 3 || cols < 3) {
        return fail(LAB_ERR_ARG, "bad arguments");
    }
    host::Pt const p = host::pick_random_point(rows, cols, squares, seed);
    *out = {p.row, p.col};
    return p.row < 0 ? fail(LAB_ERR_PLACE, "could not place a point") : LAB_OK;
}

int lab_set_corner_starts(int rows, int cols, const uint16_t *squares, lab_point out[4]) {
    if (!squares || !out || rows < 3 || cols < 3) {
        return fail(LAB_ERR_ARG, "bad arguments");
    }
    host::Pt s[4];
    host::set_corner_starts(rows, cols, squares, s);
    int rc = LAB_OK;
    for (int i = 0; i < 4; i++) {
        out[i] = {s[i].row, s[i].col};
        if (s[i].row < 0) rc = fail(LAB_ERR_PLACE, "could not place corner %d", i);
    }
    return rc;
}

int lab_place_hunt(int rows, int cols, uint16_t *squares, uint32_t seed, lab_point *start, lab_point *finish) {
    if (!squares || !start || !finish || rows < 3 || cols < 3) {
        return fail(LAB_ERR_ARG, "bad arguments");
    }
    host::Pt s, f;
    if (!host::place_hunt(rows, cols, squares, seed, s, f)) {
        return fail(LAB_ERR_PLACE, "could not place start/finish");
    }
    *start = {s.row, s.col};
    *finish = {f.row, f.col};
    return LAB_OK;
}

int lab_place_gather(int rows, int cols, uint16_t *squares, uint32_t seed, lab_point *start, lab_point finishes[4]) {
    if (!squares || !start || !finishes || rows < 3 || cols < 3) {
        return fail(LAB_ERR_ARG, "bad arguments");
    }
    host::Pt s, f[4];
    if (!host::place_gather(rows, cols, squares, seed, s, f)) {
        return fail(LAB_ERR_PLACE, "could not place start/finishes");
    }
    *start = {s.row, s.col};
    for (int i = 0; i < 4; i++) {
        finishes[i] = {f[i].row, f[i].col};
    }
    return LAB_OK;
}

int lab_place_corners(int rows, int cols, uint16_t *squares, uint32_t seed, lab_point starts[4], lab_point *finish) {
    if (!squares || !starts || !finish || rows < 5 || cols < 5) {
        return fail(LAB_ERR_ARG, "bad arguments");
    }
    host::Pt s[4], f;
    if (!host::place_corners(rows, cols, squares, seed, s, f)) {
        return fail(LAB_ERR_PLACE, "could not place corner starts");
    }
    for (int i = 0; i < 4; i++) {
        starts[i] = {s[i].row, s[i].col};
    }
    *finish = {f.row, f.col};
    return LAB_OK;
}

int lab_build_maze(const char *builder, const char *mod, int rows, int cols, uint32_t seed, uint16_t *squares) {
    int const rc = host::build_maze(builder, mod, rows, cols, seed, squares);
    if (rc == -1) return fail(LAB_ERR_ARG, "bad size %dx%d (odd, >= 7) or null pointer", rows, cols);
    if (rc == -2) return fail(LAB_ERR_ARG, "unknown builder '%s'", builder ? builder : "(null)");
    if (rc == -3) return fail(LAB_ERR_ARG, "unknown modification '%s'", mod ? mod : "(null)");
    return LAB_OK;
}

// ---- the raw maze file (host/mazefile.cc) ----------------------------------------------------
static int maze_file_status(int rc, char const *what, char const *path) {
    switch (rc) {
    case 0: return LAB_OK;
    case -1: return fail(LAB_ERR_ARG, "%s: bad argument", what);
    case -2: return fail(LAB_ERR_IO, "%s: cannot open or write '%s'", what, path ? path : "(null)");
    case -3: return fail(LAB_ERR_IO, "%s: '%s' is not a maze file (or is truncated)", what, path ? path : "(null)");
    default: return fail(LAB_ERR_ARG, "%s: the size asked for differs from the one in '%s'", what, path ? path : "(null)");
    }
}
int lab_maze_write(const char *path, int rows, int cols, const uint16_t *squares) {
    return maze_file_status(host::maze_file_write(path, rows, cols, squares), "lab_maze_write", path);
}
int lab_maze_header(const char *path, int *rows, int *cols) {
    return maze_file_status(host::maze_file_header(path, rows, cols), "lab_maze_header", path);
}
int lab_maze_read(const char *path, int rows, int cols, uint16_t *squares) {
    return maze_file_status(host::maze_file_read(path, rows, cols, squares), "lab_maze_read", path);
}

} // extern "C"
