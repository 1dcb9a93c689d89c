// sharded.cc -- ONE process, several GPUs: the distance map of one maze cut into row bands, behind the C ABI.
//
// The reference is a single process and its entry points take one Maze (run_maze/run_maze.cc:182, measure/measure.cc);
// SURVEY.md 8(b) asks for a way to let those entry points use the whole box without a launcher.  This is the host
// side of that: lab_distance_map_sharded() drives the lattice road across bands (csrc/device/lattice.cuh,
// lab_band_lat_*) on N devices of one process.  No NCCL: the devices are peers (cudaDeviceEnablePeerAccess), and
// the two exchanges of the protocol are done BY THE KERNELS -- every band's local ranking stores its slice of the
// border list, and every band's flood its arcs' weights, straight into every device's copy over NVLink
// (lab_band_lat_set_peers); the host only orders the phases (one stream synchronisation per device and phase) and
// moves the bands' facing path rows and 8-word headers.
// Applies to what the lattice road applies to (the reference builders' perfect mazes, odd sizes, start on a cell, at
// least one super-tile row per device); anything else runs on the first device alone (lab_distance_map): still the
// GPU, never the CPU.
#include "../../../include/labyrinth_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_sharded_error;

struct Band {
    int device = 0;
    int row0 = 0, rows = 0;
    lab_grid *grid = nullptr;
    void *gb = nullptr, *wsum = nullptr, *chunk = nullptr, *hdr = nullptr, *all_hdr = nullptr;
    uint8_t *top = nullptr, *bottom = nullptr, *from_above = nullptr, *from_below = nullptr;
    uint32_t *levels = nullptr;
};

bool cuda_ok(cudaError_t e, char const *what) {
    if (e == cudaSuccess) {
        return true;
    }
    g_sharded_error = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}

void release(std::vector<Band> &bands) {
    for (Band &b : bands) {
        cudaSetDevice(b.device);
        if (b.grid) lab_grid_destroy(b.grid);
        void *const arrays[] = {b.gb, b.wsum, b.chunk, b.hdr, b.all_hdr, b.top, b.bottom, b.from_above, b.from_below, b.levels};
        for (void *a : arrays) {
            cudaFree(a);
        }
    }
}

} // namespace

extern "C" {

const char *lab_sharded_last_error(void) { return g_sharded_error.c_str(); }

// *devices_used (may be NULL): on how many devices the map was made (1: the maze did not qualify for the sharded road).
int lab_distance_map_sharded(int ndev, const int *devices, int rows, int cols, uint16_t *squares, int sr, int sc, uint32_t *dist_host,
                             uint64_t *max_out, uint64_t *settled_out, int *devices_used) {
    g_sharded_error.clear();
    if (ndev < 1 || ndev > 8 || !devices || !squares || rows < 3 || cols < 3) {
        g_sharded_error = "lab_distance_map_sharded: bad argument";
        return LAB_ERR_ARG;
    }
    int const lst = lab_band_lat_super_rows();
    int const unit = lst * 64;
    int const sup_rows = (rows + unit - 1) / unit;
    int const per = (sup_rows + ndev - 1) / ndev;
    int world = (sup_rows + per - 1) / per; // (fewer bands than devices if the maze is short)
    bool const qualifies = world >= 2 && (rows & 1) && (cols & 1) && (sr & 1) && (sc & 1);
    if (!qualifies) { // one device does it (any maze, any road)
        int prev = 0;
        cudaGetDevice(&prev);
        if (!cuda_ok(cudaSetDevice(devices[0]), "cudaSetDevice")) return LAB_ERR_CUDA;
        lab_grid *g = nullptr;
        int rc = lab_grid_create(rows, cols, squares, &g);
        if (rc == LAB_OK) rc = lab_distance_map(g, sr, sc, dist_host, max_out, settled_out);
        if (rc == LAB_OK) rc = lab_grid_download(g, squares);
        if (rc != LAB_OK) g_sharded_error = lab_last_error();
        lab_grid_destroy(g);
        cudaSetDevice(prev);
        if (devices_used) *devices_used = 1;
        return rc;
    }
    int prev = 0;
    cudaGetDevice(&prev);
    std::vector<Band> bands(static_cast<size_t>(world));
    int const tiles_x = (cols + 63) / 64;
    uint64_t n_border = 0, n_weight = 0, n_chunk = 0;
    int rc = lab_band_lat_words(tiles_x, per * world, &n_border, &n_weight, &n_chunk);
    auto fail_with = [&](int code) {
        if (g_sharded_error.empty()) g_sharded_error = lab_last_error();
        release(bands);
        cudaSetDevice(prev);
        return code;
    };
    if (rc) return fail_with(rc);
    size_t const padded = static_cast<size_t>(tiles_x) * 64;
    // ---- devices, peers, buffers, the bands' Squares ----
    for (int r = 0; r < world; r++) {
        Band &b = bands[static_cast<size_t>(r)];
        b.device = devices[r];
        b.row0 = r * per * unit;
        b.rows = std::min(rows, (r + 1) * per * unit) - b.row0;
        if (!cuda_ok(cudaSetDevice(b.device), "cudaSetDevice")) return fail_with(LAB_ERR_CUDA);
        for (int q = 0; q < world; q++) {
            if (devices[q] != devices[r]) { // (the same device listed twice: its own memory needs no peering)
                cudaError_t const e = cudaDeviceEnablePeerAccess(devices[q], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    cuda_ok(e, "cudaDeviceEnablePeerAccess (the devices must be NVLink / PCIe peers)");
                    return fail_with(LAB_ERR_CUDA);
                }
                (void)cudaGetLastError();
            }
        }
        if (!cuda_ok(cudaMalloc(&b.gb, n_border * sizeof(uint64_t)), "cudaMalloc") || !cuda_ok(cudaMalloc(&b.wsum, n_weight * sizeof(uint32_t)), "cudaMalloc") ||
            !cuda_ok(cudaMalloc(&b.chunk, n_chunk * sizeof(uint32_t)), "cudaMalloc") || !cuda_ok(cudaMalloc(&b.hdr, 8 * sizeof(uint32_t)), "cudaMalloc") ||
            !cuda_ok(cudaMalloc(&b.all_hdr, 8 * sizeof(uint32_t) * static_cast<size_t>(world)), "cudaMalloc") ||
            !cuda_ok(cudaMalloc(&b.top, padded), "cudaMalloc") || !cuda_ok(cudaMalloc(&b.bottom, padded), "cudaMalloc") ||
            !cuda_ok(cudaMalloc(&b.from_above, padded), "cudaMalloc") || !cuda_ok(cudaMalloc(&b.from_below, padded), "cudaMalloc") ||
            !cuda_ok(cudaMalloc(&b.levels, static_cast<size_t>(b.rows) * cols * sizeof(uint32_t)), "cudaMalloc")) {
            return fail_with(LAB_ERR_CUDA);
        }
        // (the last 8 words of the weights hold the bands' counts: everybody adds into everybody's)
        if (!cuda_ok(cudaMemset(static_cast<uint32_t *>(b.wsum) + (n_weight - 1024 - 8), 0, 8 * sizeof(uint32_t)), "cudaMemset")) return fail_with(LAB_ERR_CUDA);
        rc = lab_grid_create(b.rows, cols, squares + static_cast<size_t>(b.row0) * cols, &b.grid);
        if (rc == LAB_OK) rc = lab_levels_bind(b.grid, 0, b.levels);
        if (rc) return fail_with(rc);
    }
    auto sync_all = [&]() {
        for (Band &b : bands) {
            cudaSetDevice(b.device);
            if (!cuda_ok(cudaDeviceSynchronize(), "cudaDeviceSynchronize")) return false;
        }
        return true;
    };
    // ---- begin, the facing path rows ----
    for (int r = 0; r < world; r++) {
        Band &b = bands[static_cast<size_t>(r)];
        bool const mine = sr >= b.row0 && sr < b.row0 + b.rows;
        rc = lab_band_begin_lattice(b.grid, 0, mine ? sr - b.row0 : -1, mine ? sc : -1, 1);
        if (rc == LAB_OK) rc = lab_band_path_rows(b.grid, b.top, b.bottom);
        if (rc) return fail_with(rc);
    }
    if (!sync_all()) return fail_with(LAB_ERR_CUDA);
    for (int r = 0; r < world; r++) {
        Band &b = bands[static_cast<size_t>(r)];
        cudaSetDevice(b.device);
        if (r > 0 && !cuda_ok(cudaMemcpyPeer(b.from_above, b.device, bands[static_cast<size_t>(r) - 1].bottom, bands[static_cast<size_t>(r) - 1].device, padded), "cudaMemcpyPeer")) {
            return fail_with(LAB_ERR_CUDA);
        }
        if (r + 1 < world && !cuda_ok(cudaMemcpyPeer(b.from_below, b.device, bands[static_cast<size_t>(r) + 1].top, bands[static_cast<size_t>(r) + 1].device, padded), "cudaMemcpyPeer")) {
            return fail_with(LAB_ERR_CUDA);
        }
        rc = lab_band_set_neighbour_paths(b.grid, r > 0 ? b.from_above : nullptr, r + 1 < world ? b.from_below : nullptr);
        if (rc == LAB_OK) rc = lab_band_lat_bind(b.grid, r * per, per, per * world, b.gb, b.wsum, b.chunk);
        if (rc) return fail_with(rc);
        std::vector<void *> gbs, ws; // [0] = this band's own arrays, then the peers'
        gbs.push_back(b.gb);
        ws.push_back(b.wsum);
        for (int q = 0; q < world; q++) {
            if (q != r) {
                gbs.push_back(bands[static_cast<size_t>(q)].gb);
                ws.push_back(bands[static_cast<size_t>(q)].wsum);
            }
        }
        rc = lab_band_lat_set_peers(b.grid, world, gbs.data(), ws.data());
        if (rc) return fail_with(rc);
    }
    // ---- phase 1: link + local ranking; every band's slice of the border list lands in every device's copy ----
    for (Band &b : bands) {
        rc = lab_band_lat_link(b.grid, b.hdr);
        if (rc) return fail_with(rc);
    }
    if (!sync_all()) return fail_with(LAB_ERR_CUDA);
    {
        std::vector<uint32_t> all(static_cast<size_t>(world) * 8);
        for (int r = 0; r < world; r++) {
            cudaSetDevice(bands[static_cast<size_t>(r)].device);
            if (!cuda_ok(cudaMemcpy(all.data() + static_cast<size_t>(r) * 8, bands[static_cast<size_t>(r)].hdr, 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost), "cudaMemcpy")) return fail_with(LAB_ERR_CUDA);
        }
        for (Band &b : bands) {
            cudaSetDevice(b.device);
            if (!cuda_ok(cudaMemcpy(b.all_hdr, all.data(), all.size() * sizeof(uint32_t), cudaMemcpyHostToDevice), "cudaMemcpy")) return fail_with(LAB_ERR_CUDA);
        }
    }
    // ---- phase 2: ranking over the whole list, floods; every band's weights land in every device's copy ----
    for (Band &b : bands) {
        rc = lab_band_lat_flood(b.grid, b.all_hdr, world, 0);
        if (rc) return fail_with(rc);
    }
    if (!sync_all()) return fail_with(LAB_ERR_CUDA);
    // ---- phase 3: scan, levels; results ----
    bool all_used = true;
    for (Band &b : bands) {
        int32_t used = 0;
        rc = lab_band_lat_levels(b.grid, &used);
        if (rc) return fail_with(rc);
        all_used = all_used && used == 1;
    }
    if (!all_used) { // not one lattice tree: nothing was written; one device searches the whole maze
        release(bands);
        bands.clear();
        cudaSetDevice(prev);
        int one = devices[0];
        return lab_distance_map_sharded(1, &one, rows, cols, squares, sr, sc, dist_host, max_out, settled_out, devices_used);
    }
    uint64_t mx = 0, settled = 0;
    for (Band &b : bands) {
        uint64_t m = 0, s = 0;
        rc = lab_band_finish(b.grid, &m, &s);
        if (rc == LAB_OK) rc = lab_grid_download(b.grid, squares + static_cast<size_t>(b.row0) * cols);
        if (rc) return fail_with(rc);
        mx = std::max(mx, m);
        settled += s;
        if (dist_host) {
            cudaSetDevice(b.device);
            if (!cuda_ok(cudaMemcpy(dist_host + static_cast<size_t>(b.row0) * cols, b.levels, static_cast<size_t>(b.rows) * cols * sizeof(uint32_t), cudaMemcpyDeviceToHost), "cudaMemcpy")) {
                return fail_with(LAB_ERR_CUDA);
            }
        }
    }
    if (max_out) *max_out = mx;
    if (settled_out) *settled_out = settled;
    if (devices_used) *devices_used = world;
    release(bands);
    cudaSetDevice(prev);
    return LAB_OK;
}

} // extern "C"
