/* labyrinth_b200.h -- C ABI of the B200 BFS hot path (liblabyrinth_b200.so).
 *
 * The reference (agl-alexglopez/multithreading-with-mazes) has no FFI: its boundary for this path
 * is a set of C++20-module free functions of type void(Maze::Maze&) kept in std::function tables
 *   run_maze/run_maze.cc:97-124   "bfs-hunt" / "bfs-gather" / "bfs-corners"  -> Bfs::hunt/gather/corners
 *   measure/measure.cc:96-101     "distance"                                 -> Distance::paint_distance_from_center
 * include/labyrinth.hh re-creates those namespaces/names/signatures on top of THIS header, so the
 * CLIs dispatch unchanged; INTEGRATION.md shows the binding a maintainer would add upstream.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 on success, a negative lab_status
 *     otherwise, and lab_last_error() gives the message (the reference prints to stderr and
 *     std::abort()s instead: solvers/solve_utilities.cc:200-202,247-251; the C++ mirror keeps that).
 *   - squares: rows*cols uint16_t, dense row-major, exactly Maze::Maze's vector<Square>
 *     (maze/maze.cc:127-144; sizeof(Square) == 2).  Bit layout: maze/maze.cc:39-55,170-173,
 *     solvers/solve_utilities.cc:59-79, painters/rgb.cc:21-22.
 *   - a lab_grid is a maze resident in one GPU's HBM together with the search workspace; it can be
 *     solved/painted repeatedly without touching the host.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     LAB_ERR_CUDA.
 */
#ifndef LABYRINTH_B200_H
#define LABYRINTH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAB_INF 0xFFFFFFFFu /* level of a wall / unreached square */

typedef enum lab_status {
    LAB_OK = 0,
    LAB_ERR_ARG = -1,      /* bad pointer / size / coordinate                                  */
    LAB_ERR_CUDA = -2,     /* CUDA runtime error (no device, out of memory, launch failure)    */
    LAB_ERR_TIMEOUT = -3,  /* the persistent search kernel hit its deadline (LAB_BFS_BUDGET_MS) */
    LAB_ERR_INTERNAL = -4, /* invariant broken inside the search (queue overflow)              */
    LAB_ERR_PLACE = -5,    /* no valid square for a start/finish (reference aborts here)       */
    LAB_ERR_IO = -6        /* maze file cannot be opened / written / is not a maze file        */
} lab_status;

/* Maze::Point, maze/maze.cc:116-119 */
typedef struct lab_point {
    int32_t row;
    int32_t col;
} lab_point;

typedef struct lab_grid lab_grid;

/* Device-side timing of the last search on a grid (CUDA events on the grid's stream), ms. */
typedef struct lab_timing {
    float h2d_ms;    /* lab_grid_upload / create-with-host-data            */
    float setup_ms;  /* workspace reset (memsets)                          */
    float pack_ms;   /* Square grid -> tile bit masks + border masks       */
    float search_ms; /* the persistent tile-relaxation kernel              */
    float post_ms;   /* measure / paint / path kernels                     */
    float total_ms;  /* setup + pack + search + post                       */
    float d2h_ms;    /* lab_grid_download / result copies                  */
} lab_timing;

/* Work counters of the last search (diagnostics; zero cost when unread). */
typedef struct lab_stats {
    uint64_t tiles;         /* tiles in the grid                                         */
    uint64_t activations;   /* tile relaxation passes                                    */
    uint64_t empty;         /* ... that found nothing to do                              */
    uint64_t levels;        /* warp-level BFS steps summed over all passes               */
    uint64_t settled;       /* square settle events (> visited squares if corrections)   */
    uint64_t rechecks;      /* label-correcting look-ups                                 */
    uint64_t resettles;     /* squares whose level was lowered after first being settled */
    uint64_t queue_pushes;  /* tiles handed over through the global queue                */
    uint64_t direct;        /* tiles the finishing warp walked into itself               */
    uint32_t worker_warps;  /* warps of the persistent launch                            */
    uint32_t error;         /* Control.error                                             */
    /* SM cycles summed over worker warps; non-zero only in the -DLAB_PROFILE build */
    uint64_t cyc_load, cyc_loop, cyc_publish, cyc_idle, cyc_step, cyc_emit, batches;
    /* spanning-tree pipeline (csrc/device/tree.cuh): passable squares, adjacent passable pairs (a tree
     * has pairs == squares - 1), tile-border crossings, and which road produced the field: 0 the tile
     * wave, 1 the spanning-tree pipeline, 2 the open room (csrc/device/room.cuh: closed-form levels) */
    uint64_t tree_squares, tree_pairs, tree_crossings, tree_used;
    uint64_t launches; /* kernels launched by the search (reset, pack, search, epilogue) */
} lab_stats;

/* ---- grid lifetime (replaces Maze::Maze construction, maze/maze.cc:349-354, for the GPU side) -- */
/* host_squares may be NULL (grid zero-filled). rows, cols >= 3. */
int lab_grid_create(int rows, int cols, const uint16_t *host_squares, lab_grid **out);
/* Borrow device memory the caller owns (e.g. a torch tensor's data_ptr) as the Square array. */
int lab_grid_create_on_device(int rows, int cols, void *device_squares, lab_grid **out);
int lab_grid_upload(lab_grid *g, const uint16_t *host_squares);
int lab_grid_download(lab_grid *g, uint16_t *host_squares);
int lab_grid_destroy(lab_grid *g);
/* All later work of this grid is enqueued on `cuda_stream` (a cudaStream_t; NULL = default stream). */
/* Options of a grid.  LAB_OPT_SOLVER_LEVELS (default 1): hunt / gather / corners leave their level field(s) behind for
 * lab_levels_download (the animated variants replay from them).  0: a solver may skip writing a field it does not need itself --
 * bfs-gather without paths on the reference builders' perfect mazes then only reads and writes the Squares (4 bytes
 * per square); lab_levels_download afterwards returns LAB_ERR_ARG.  The grid's contents are the same either way. */
enum { LAB_OPT_SOLVER_LEVELS = 1 };
int lab_grid_set_option(lab_grid *g, int option, int value);
int lab_grid_set_stream(lab_grid *g, void *cuda_stream);
void *lab_grid_device_squares(lab_grid *g);
int lab_grid_rows(const lab_grid *g);
int lab_grid_cols(const lab_grid *g);
/* OR `bits` into / AND `mask` onto single squares of the device grid (start/finish placement,
 * corner carving: bfs_threads.cc:248-251,424-433). */
int lab_grid_or_bits(lab_grid *g, const lab_point *cells, const uint16_t *bits, int n);
int lab_grid_and_bits(lab_grid *g, uint16_t mask);

/* ---- Distance::paint_distance_from_center, BFS part (painters/distance.cc:129-153) ------------ */
/* Levels from (sr, sc) over squares with path_bit set and measure bit clear; the start is expanded
 * whatever its bits.  Side effect on the grid: Rgb::measure (bit 9) on every reached square.
 *   dist_host   NULL or rows*cols uint32 (LAB_INF = no map entry); copied D2H when non-NULL
 *   max_out     Distance_map::max (largest level);   settled_out  Distance_map::distances.size()
 * The level field stays resident: lab_distance_device() returns it. */
int lab_distance_map(lab_grid *g, int sr, int sc, uint32_t *dist_host, uint64_t *max_out,
                     uint64_t *settled_out);
/* The reference's start square for a rows x cols maze (distance.cc:131-134). */
lab_point lab_distance_start(int rows, int cols);
/* Device pointer to plane k's level field (rows*cols uint32), or NULL.  It holds the last
 * lab_distance_map's levels; the solvers use the fields as scratch (bounded searches leave squares
 * beyond the end of the game at LAB_INF; in an open room -- csrc/device/room.cuh -- no field is written). */
void *lab_levels_device(lab_grid *g, int plane);
/* Use caller-owned device memory (rows*cols uint32) as plane k's level field from now on. */
/* The level field the last search left in `plane` (hunt, gather, distance: plane 0 from the shared start; corners: plane k
 * from corner k), copied to the host: rows*cols uint32, 0xFFFFFFFF = no level.  The animated variants of the reference
 * (bfs_threads.cc:282-331,371-418,455-514; darkbfs_threads.cc:154-300; distance.cc:158-204) are replayed from it.
 * LAB_ERR_ARG after a search that took the open-room road (no field is written there: levels are |dr| + |dc|). */
int lab_levels_download(lab_grid *g, int plane, uint32_t *levels_host);
int lab_levels_bind(lab_grid *g, int plane, void *device_levels);

/* ---- Distance painter (painters/distance.cc:45-70, `painter`) -------------------------------- */
/* The colour the reference prints for every square of the last lab_distance_map, computed on the
 * device with the reference's double arithmetic: intensity = double(max - level) / double(max),
 * dark = uint8(255.0 * intensity), bright = 128 + uint8(127.0 * intensity); `channel` (0 red, 1 green,
 * 2 blue: the reference draws it at random once per picture, :47-49) gets bright, the other two dark.
 * One pixel per square: R | G << 8 | B << 16 | 0xFF << 24; 0 where the reference prints a wall glyph
 * (no map entry).  pixels_host: NULL or rows*cols uint32 (copied D2H).  The pixels stay resident
 * (lab_pixels_device); lab_pixels_bind makes the painter write into caller-owned device memory. */
int lab_distance_paint(lab_grid *g, int channel, uint32_t *pixels_host);

/* Runs::paint_runs (painters/runs.cc:130-161) -- SURVEY.md 8(f) rank 2.  The same search as the distance map (start, FIFO
 * order N,E,S,W, measure bit on every reached square); the map holds for every reached square the length of the straight
 * run its tree path ends with: 0 at the start, `turn ? 1 : parent's + 1` (:148-151).  runs_host (may be NULL):
 * rows*cols uint32, 0xFFFFFFFF = no entry.  On the GPU for mazes that are ONE TREE (all perfect mazes; the parent of a
 * square is then its one neighbour a level lower); LAB_ERR_ARG for a maze with loops, where the reference's parent
 * depends on its queue order.  lab_runs_paint: the painter's colours (:45-70), as lab_distance_paint. */
int lab_runs_map(lab_grid *g, int start_row, int start_col, uint32_t *runs_host, uint64_t *max_out, uint64_t *settled_out);
int lab_runs_paint(lab_grid *g, int channel, uint32_t *pixels_host);
void *lab_pixels_device(lab_grid *g);
int lab_pixels_bind(lab_grid *g, void *device_pixels);

/* ---- Bfs::hunt / gather / corners, thread part (solvers/bfs_threads.cc:35-85,144-185) --------- */
/* The caller has placed start_bit / finish_bit (and carved the centre for corners) exactly as
 * Bfs::hunt :245-251, Bfs::gather :335-344, Bfs::corners :422-433 do; lab_place_* below do it on a
 * host copy with the reference's rules.  Canonical lock-step semantics.
 * Paths: (row,col) pairs ordered like Bfs_monitor::thread_paths (finish side first, start last,
 * finish excluded); path buffers are host memory and may be NULL; path_len is always written. */
int lab_bfs_hunt(lab_grid *g, lab_point start, lab_point finish, int32_t *winner, lab_point *path,
                 uint64_t path_cap, uint64_t *path_len);
int lab_bfs_gather(lab_grid *g, lab_point start, const lab_point finishes[4], int32_t claimed_by[4],
                   lab_point *const paths[4], uint64_t path_cap, uint64_t path_len[4]);
int lab_bfs_corners(lab_grid *g, const lab_point starts[4], lab_point finish, int32_t *winner,
                    lab_point *path, uint64_t path_cap, uint64_t *path_len);
/* Number of squares that received at least one paint/cache bit in the last solver call. */
uint64_t lab_last_painted(const lab_grid *g);

/* ---- row bands: one lab_grid per GPU holds a band of rows of a taller maze --------------------- */
/* A sharded distance map is a loop the caller drives (one process per GPU; the exchange is the
 * caller's: NCCL send/recv of one row of levels per neighbour and a one-word all-reduce):
 *     lab_band_begin          reset + pack this band (src = -1,-1 if the start is in another band)
 *     lab_band_path_rows      passable flags (1 byte/column, lab_band_cols_padded columns) of the band's
 *                             first and last row -> exchange once with the neighbours
 *     lab_band_set_neighbour_paths   the neighbours' facing rows (NULL = no band there)
 *     repeat { lab_band_relax; lab_band_export_edges; exchange; lab_band_import_edges -> activated;
 *              all-reduce(activated) } until no rank activated a tile
 *     lab_band_finish         measure bits, local max level and local count of settled squares
 * All row buffers are DEVICE pointers (uint8 / uint32, lab_band_cols_padded entries).  A band with a
 * neighbour below must have a multiple of 64 rows.  Levels are global BFS levels. */
int lab_band_cols_padded(const lab_grid *g);
int lab_band_begin(lab_grid *g, int src_r_local, int src_c);
int lab_band_path_rows(lab_grid *g, void *top_dev, void *bottom_dev);
int lab_band_set_neighbour_paths(lab_grid *g, const void *above_bottom_dev, const void *below_top_dev);
int lab_band_relax(lab_grid *g);
int lab_band_export_edges(lab_grid *g, void *top_levels_dev, void *bottom_levels_dev);
int lab_band_import_edges(lab_grid *g, const void *above_bottom_levels_dev, const void *below_top_levels_dev, uint32_t *activated);
int lab_band_finish(lab_grid *g, uint64_t *local_max, uint64_t *local_settled);
/* The solvers across row bands (Bfs::hunt / Bfs::gather, solvers/bfs_threads.cc:243-280,333-369, canonical model as on one
 * GPU), on mazes whose level field the pipeline below or lab_band_room provides (one tree, an open room): compositions of
 *     lab_band_begin_game     like lab_band_begin; game 1 hunt, 2 gather: passable <=> path_bit, no measure bits
 *     lab_band_level_at       level of one square of this band (the finishes' levels: the games' thresholds)
 *     lab_band_paint          the game's paint rules on this band's squares -> squares that received a bit
 *     lab_band_walk           one piece of a canonical path inside this band, from the finish (include_first = 0) or from
 *                             the square a walk taken over from the neighbouring band continues at (1), until the start or
 *                             the band's border; repaint_bits != 0 replaces the paint nibble of its squares (hunt's winner).
 *                             state = {squares walked, row, col where it stopped, level left (0: at the start)}
 *     lab_band_recolour       paint nibble of one square := bits (gather's path.front(), bfs_threads.cc:355-364)
 *     lab_band_finish_game    closes the game
 * multithreading-with-mazes_b200/bands.py (sharded_hunt, sharded_gather) is the driver.  On any other maze (loops:
 * grid, -m cross, -m x) the field comes from the relax / exchange loop above run to its fixpoint after lab_band_begin_game.
 *
 * Bfs::corners across row bands (solvers/bfs_threads.cc:420-453; bands.sharded_corners): FOUR fields at once,
 *     lab_band_begin_corners  src_rc = colour k's start (row, col) in band rows, (-1, -1) if it lies in another band
 *     lab_band_planes         fields of the open band search (1; 4 after lab_band_begin_corners): lab_band_export_edges /
 *                             _import_edges then move lab_band_planes() x lab_band_cols_padded() levels per row buffer,
 *                             field k's at [k * cols_padded, (k + 1) * cols_padded)
 *     lab_band_level_at_plane / lab_band_paint_planes   as above, naming the field (colour k paints from its own) */
int lab_band_begin_game(lab_grid *g, int game, int src_r_local, int src_c);
int lab_band_begin_corners(lab_grid *g, const int32_t src_rc[8]);
int lab_band_planes(const lab_grid *g);
int lab_band_level_at_plane(lab_grid *g, int plane, int r, int c, uint32_t *level);
int lab_band_paint_planes(lab_grid *g, int n_rules, const uint32_t *plane, const uint32_t *below, const uint32_t *bits, uint64_t *painted);
int lab_band_level_at(lab_grid *g, int r, int c, uint32_t *level);
int lab_band_paint(lab_grid *g, int n_rules, const uint32_t *below, const uint32_t *bits, uint64_t *painted);
int lab_band_walk(lab_grid *g, int r, int c, int include_first, int first_only, uint32_t repaint_bits, void *path_dev,
                  uint64_t path_cap, int64_t state[4]);
int lab_band_recolour(lab_grid *g, int r, int c, uint16_t paint_bits);
int lab_band_finish_game(lab_grid *g);

/* The spanning-tree pipeline (csrc/device/tree.cuh) across row bands: what replaces the relax / exchange
 * loop above when the maze is one tree (every reference builder without -m), whose diameter would cross the
 * band borders thousands of times.  Everything heavy is local: the walk, the flood and the fix-up work on the
 * band's own tiles, and each of the two pointer-jumping phases first contracts the band's own part of the
 * list (pointers are followed only inside the band's slots).  What is left of a list then runs over the bands'
 * BORDER TILE ROWS only (the first and last tile row of every band: the only tiles another band's pointers can
 * lead into); the ranks all-gather those rows -- one fixed-size chunk of uint32 per rank and stage, with the
 * rank's error flags and counts in its header -- every rank finishes the short list itself, and a last local
 * pass adds the two.  Three all-gathers per map, whatever its diameter; the ranks agree on "one tree or not"
 * through the flags in the chunks, without a host round trip.
 *     lab_band_tree_bind      the caller's device arrays, indexed by slot = global tile * 256 + side * 64 + position
 *                             (global tile = tile_base + the band's tile): succ0: uint32[ntiles_global*256+2]; sr0, sr1: the
 *                             two ping-pong lists of (pointer, value) pairs, uint32[2*(ntiles_global*256+2)], 8-byte aligned;
 *                             loc: uint32[ntiles_global*256]; inmask: uint64[ntiles_global*4]; alist: uint32[the band's
 *                             tiles * 256]; border_list: the slots of every band's first and last tile row (uint32
 *                             [border_count]); layout: {tile_base, ntiles} of every rank (uint32[world][2])
 *     lab_band_tree_count     -> the band's passable squares, adjacent pairs, border crossings, bounding box.  Over all
 *                             ranks a tree has sum(pairs) == sum(squares) - 1 (anything else: lab_band_room if it
 *                             applies, or stay with the loop above)
 *     lab_band_tree_rank_local   (the sums in) walk + windowed ranking                      -> chunk 1
 *     lab_band_tree_flood        (all ranks' chunks 1 in; root_rank = who holds the start) short list, final ranks,
 *                                orientation, flood                                          -> chunk 2
 *     lab_band_tree_prefix_local (all chunks 2 in) parents + windowed prefix sums            -> chunk 3
 *     lab_band_tree_levels       (all chunks 3 in) short list, the band's final levels -> used: 1 = the field is final
 *                                (lab_band_finish next); 0 = not one tree, the field is wiped, the loop above takes over
 *     lab_band_tree_chunk_words  uint32 words of a stage's chunk (stage 1..3)
 * Results are identical to the loop's, bit for bit. */
int lab_band_tree_bind(lab_grid *g, uint32_t tile_base, uint32_t ntiles_global, void *succ0, void *sr0, void *sr1,
                       void *loc, void *inmask, void *alist, const void *border_list, uint32_t border_count,
                       const void *layout, int world, int me);
int lab_band_tree_chunk_words(const lab_grid *g, int stage);
int lab_band_tree_count(lab_grid *g, uint64_t *squares, uint64_t *pairs, uint32_t *crossings, int32_t box[4]);
/* The open room (csrc/device/room.cuh) across row bands: if the ranks' boxes (lab_band_tree_count: bounding box
 * r0, r1, c0, c1 of the band's passable squares, rows relative to the band) add up to one rectangle that holds
 * exactly sum(squares) squares, there is nothing to search: every band writes the part [r0, r1] x [c0, c1] of the
 * room it holds (r0 > r1: none) from the closed form, the source (src_r, src_c) again relative to the band (it may
 * lie above row 0 or below the band's last row).  lab_band_finish next.  No exchange at all. */
int lab_band_room(lab_grid *g, int r0, int r1, int c0, int c1, int src_r, int src_c);
int lab_band_tree_rank_local(lab_grid *g, uint64_t squares_global, uint64_t pairs_global, void *chunk1);
int lab_band_tree_flood(lab_grid *g, const void *all_chunks1, int root_rank, void *chunk2);
int lab_band_tree_prefix_local(lab_grid *g, const void *all_chunks2, void *chunk3);
int lab_band_tree_levels(lab_grid *g, const void *all_chunks3, int32_t *used);

/* The lattice road (csrc/device/lattice.cuh) across row bands: what the reference builders' perfect mazes take (cells on
 * the odd squares: builders of kruskal.cc, wilson_path_carver.cc, recursive_backtracker.cc ...), tried BEFORE the general
 * pipeline above.  Bands are whole SUPER-TILE rows (lab_band_lat_super_rows() x 64 rows), all of `super_rows_band`
 * super-tile rows but the last, which may be shorter; the whole maze has odd sizes and the start lies on an
 * (odd, odd) square.  Every rank holds three whole-maze arrays of lab_band_lat_words' sizes: the border list
 * (uint64 entries), the weights (uint32 words) and the scan's chunk offsets (uint32).  After lab_band_begin(_game) and
 * lab_band_set_neighbour_paths:
 *     lab_band_lat_bind    the band's place (super-tile rows [base, base + band) of `global`) and the three arrays
 *     lab_band_lat_link    counts, link, local ranking -> the band's slice of the border list (entries
 *                          [rank * slice, (rank + 1) * slice), slice = entries / world) and its 8-word header
 *                          -- the ranks all-gather both --
 *     lab_band_lat_flood   the whole maze's counts from the headers (every rank reads the same ones: they agree without
 *                          a host round trip), the ranking over the whole border list, the band's floods; its arcs'
 *                          weights go to their tour positions in the weights array (zeroed first if zero_weights)
 *                          -- the ranks sum the weights arrays (all-reduce over weight_words - 1024 words) --
 *     lab_band_lat_levels  prefix sums over the whole tour, the band's final levels -> used: 1 = the field is final
 *                          (lab_band_finish next); 0 = not one lattice tree: nothing was written, the pipeline above or
 *                          the loop take over
 * Two exchanges per map whatever its diameter.  lab_band_lat_set_peers (optional, after bind): the other ranks' border
 * lists and weights arrays as device pointers valid on this device ([0] = this rank's own); the kernels then store a
 * band's slice and weights into EVERY rank's arrays themselves (NVLink peer stores) and the two exchanges shrink to
 * barriers (call lab_band_lat_flood with zero_weights = 0 and clear the last 8 weight words before lab_band_lat_link). */
int lab_band_begin_lattice(lab_grid *g, int game, int src_r, int src_c, int lattice_next); /* lab_band_begin_game that leaves the field's reset and the
                                                                                           * measure bits to the lattice road (lattice_next != 0) */
int lab_band_lat_super_rows(void);
int lab_band_lat_words(int tiles_x, int super_rows_global, uint64_t *border_entries, uint64_t *weight_words, uint64_t *chunk_words);
int lab_band_lat_bind(lab_grid *g, int super_row_base, int super_rows_band, int super_rows_global, void *border_list, void *weights,
                      void *chunks);
int lab_band_lat_set_peers(lab_grid *g, int n, void *const *border_lists, void *const *weights);
int lab_band_lat_link(lab_grid *g, void *header_dev);
int lab_band_lat_flood(lab_grid *g, const void *all_headers_dev, int world, int zero_weights);
int lab_band_lat_levels(lab_grid *g, int32_t *used);
/* hunt's winner path on the lattice road across bands WITHOUT hand-overs (the tour's ranks are the whole maze's: every band finds
 * and walks the path's segments in its own rows; csrc/device/lattice.cuh lat_path_*; replaces the walk of
 * solvers/bfs_threads.cc:263-274's repaint).  _rank: the band that holds the finish -> rank of the entry arc of the finish's
 * component (0: a root's component) and the finish's level; the caller tells every band.  _mark: (fin_r, fin_c) / (start_r,
 * start_c) in band rows, -1 if in another band; path_dev: NULL or path_cap (row, col) pairs written at the squares' indices
 * along the whole path (band rows); repaint_bits != 0: the paint nibble of the path's squares := bits. */
int lab_band_lat_path_rank(lab_grid *g, int r, int c, uint32_t *rank, uint32_t *level);
int lab_band_lat_path_mark(lab_grid *g, uint32_t rank, uint32_t level, int fin_r, int fin_c, int start_r, int start_c, uint32_t repaint_bits,
                           void *path_dev, uint64_t path_cap, uint64_t *segments);

/* ---- one process, several GPUs --------------------------------------------------------------- */
/* The reference is one process with one Maze (run_maze/run_maze.cc:182); this lets its entry points use the whole box
 * without a launcher (SURVEY.md 8(b)'s lab_set_devices): the distance map of the host Squares (rows*cols, updated in place
 * as lab_distance_map + lab_grid_download do) cut into row bands over `ndev` devices of THIS process (devices[] = CUDA
 * ordinals; they must be peers).  It is the lattice road across bands above in peer mode: no NCCL, the kernels store their
 * slices of the border list and their weights into every device's copy over NVLink, the host only orders the three phases.
 * A maze the sharded road does not apply to (even sizes, start not on a cell, fewer than two super-tile rows, not one
 * lattice tree) is searched by devices[0] alone -- *devices_used says which it was.  dist_host, max_out, settled_out,
 * devices_used may be NULL.  Errors: lab_sharded_last_error() (or lab_last_error() for the calls underneath). */
int lab_distance_map_sharded(int ndev, const int *devices, int rows, int cols, uint16_t *squares, int start_row, int start_col,
                             uint32_t *dist_host, uint64_t *max_out, uint64_t *settled_out, int *devices_used);
const char *lab_sharded_last_error(void);

int lab_last_timing(const lab_grid *g, lab_timing *out);
int lab_last_stats(const lab_grid *g, lab_stats *out);
const char *lab_last_error(void);

/* ---- host-side helpers shared with the C++ mirror (no GPU needed) ----------------------------- */
/* Sutil::is_valid_start_or_finish / find_nearest_square / pick_random_point / set_corner_starts
 * (solvers/solve_utilities.cc:166-173,229-252,275-285,254-273) on a host copy of the squares.
 * `seed` stands in for the value the reference draws from std::random_device. */
int lab_pick_random_point(int rows, int cols, const uint16_t *squares, uint32_t seed, lab_point *out);
int lab_set_corner_starts(int rows, int cols, const uint16_t *squares, lab_point out[4]);
/* Placement part of each game on a host copy: sets start/finish bits (and carves the centre). */
int lab_place_hunt(int rows, int cols, uint16_t *squares, uint32_t seed, lab_point *start, lab_point *finish);
int lab_place_gather(int rows, int cols, uint16_t *squares, uint32_t seed, lab_point *start, lab_point finishes[4]);
int lab_place_corners(int rows, int cols, uint16_t *squares, uint32_t seed, lab_point starts[4], lab_point *finish);
/* Linear-time seeded builders (builders/{arena,grid,recursive_backtracker,kruskal,
 * wilson_path_carver,mods}.cc): builder in arena|rdfs|grid|kruskal|wilson, mod in none|cross|x. */
int lab_build_maze(const char *builder, const char *mod, int rows, int cols, uint32_t seed, uint16_t *squares);

/* ---- the raw maze file ------------------------------------------------------------------------ */
/* The reference has no file I/O; "identical serialized mazes" (BASELINE.json) need a
 * format: "LABMAZE1", int32 rows, int32 cols (little endian), then rows*cols uint16 Squares, dense
 * row-major exactly as Maze::Maze holds them (maze/maze.cc:127-144).  Data starts at byte 16 (mmap-able).
 * The oracle, the CLIs (LAB_LOAD=path / LAB_DUMP=path) and bench.py's maze cache read and write it. */
int lab_maze_write(const char *path, int rows, int cols, const uint16_t *squares);
int lab_maze_header(const char *path, int *rows, int *cols);
int lab_maze_read(const char *path, int rows, int cols, uint16_t *squares);

/* Library / device probe: fills a short description, returns the number of CUDA devices (0 = none). */
int lab_device_info(char *buf, size_t buflen);

#ifdef __cplusplus
}
#endif
#endif /* LABYRINTH_B200_H */
