// oracle/restate.cc -- TEST INFRASTRUCTURE ONLY.  Nothing under multithreading-with-mazes_b200/
// may include, link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, and only as the checker.
//
// CPU restatement of the reference's BFS hot path with flat arrays instead of the
// std::unordered_map<Point,...> the reference uses (whose hash, maze/maze.cc:397-405, is
// row ^ col and makes every probe O(n): SURVEY.md §3.4).  Same bit semantics, same visit
// order, same RNG call order (libstdc++ 13.3 <random>, as the reference compiled here).
//
// PINNED: tests/test_oracle_vs_reference.py checks every function below against the real
// reference text compiled into oracle/_ref/ (seed ladder, 5 builders, sizes up to 255^2 in the
// quick suite), and tests/golden/ holds vectors produced by that reference so the pin also
// holds on the GPU box where /root/reference does not exist.
//
// Each function cites the reference lines it follows (paths relative to /root/reference).

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

namespace {

// maze/maze.cc:170-173, 205-210
constexpr uint16_t path_bit = 0b0010'0000'0000'0000;
constexpr uint16_t start_bit = 0b0100'0000'0000'0000;
// solvers/solve_utilities.cc:57-79
constexpr int num_threads = 4;
constexpr uint16_t no_winner = UINT16_MAX;
constexpr uint16_t finish_bit = 0b1000'0000'0000'0000;
constexpr uint16_t thread_paint_shift = 4;
constexpr uint16_t thread_paint_mask = 0b1111'0000;
constexpr uint16_t cache_mask = 0b1111'0000'0000;
constexpr uint16_t thread_cache_shift = 8;
constexpr int num_gather_finishes = 4;
// painters/rgb.cc:21-22
constexpr uint16_t measure_bit = 0b10'0000'0000;

constexpr uint32_t INF = 0xFFFFFFFFu;

struct Point {
    int row;
    int col;
};
// north, east, south, west: maze/maze.cc:213-218 == solvers/solve_utilities.cc:124-125
constexpr Point dirs[4] = {{-1, 0}, {0, 1}, {1, 0}, {0, -1}};
// solvers/solve_utilities.cc:127-128 -- SEVEN entries, south-west is missing upstream.
constexpr Point all_dirs[7] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}};

struct Grid {
    int rows;
    int cols;
    uint16_t *sq;
    uint16_t &at(int r, int c) const { return sq[static_cast<size_t>(r) * cols + c]; }
    size_t idx(int r, int c) const { return static_cast<size_t>(r) * cols + c; }
};

// solvers/solve_utilities.cc:166-173
bool is_valid_start_or_finish(Grid const &g, Point p) {
    return p.row > 0 && p.row < g.rows - 1 && p.col > 0 && p.col < g.cols - 1
           && (g.at(p.row, p.col) & path_bit) && !(g.at(p.row, p.col) & finish_bit)
           && !(g.at(p.row, p.col) & start_bit);
}

// solvers/solve_utilities.cc:229-252.  Returns {-1,-1} where the reference would abort.
Point find_nearest_square(Grid const &g, Point choice) {
    for (Point const &p : all_dirs) {
        Point const next = {choice.row + p.row, choice.col + p.col};
        if (is_valid_start_or_finish(g, next)) {
            return next;
        }
    }
    for (int row = 1; row < g.rows - 1; row++) {
        for (int col = 1; col < g.cols - 1; col++) {
            if (is_valid_start_or_finish(g, {row, col})) {
                return {row, col};
            }
        }
    }
    return {-1, -1};
}

// solvers/solve_utilities.cc:275-285
Point pick_random_point(Grid const &g, unsigned seed) {
    std::mt19937 generator(seed);
    std::uniform_int_distribution<int> row_random(1, g.rows - 2);
    std::uniform_int_distribution<int> col_random(1, g.cols - 2);
    // braced-init-list: row drawn first, then col
    int const r = row_random(generator);
    int const c = col_random(generator);
    Point choice = {r, c};
    if (!is_valid_start_or_finish(g, choice)) {
        choice = find_nearest_square(g, choice);
    }
    return choice;
}

// solvers/solve_utilities.cc:254-273
void set_corner_starts(Grid const &g, Point out[4]) {
    Point const want[4]
        = {{1, 1}, {1, g.cols - 2}, {g.rows - 2, 1}, {g.rows - 2, g.cols - 2}};
    for (int i = 0; i < 4; i++) {
        out[i] = want[i];
        if (!(g.at(want[i].row, want[i].col) & path_bit)) {
            out[i] = find_nearest_square(g, want[i]);
        }
    }
}

// One colour's BFS state: Sutil::Bfs_monitor's per-thread map/queue/path
// (solvers/solve_utilities.cc:146-164) on flat arrays.  parent: -2 unseen, -1 = the {-1,-1}
// sentinel stored for the start.
struct Thread_state {
    std::vector<int64_t> parent;
    std::vector<int64_t> queue;
    size_t head = 0;
    std::vector<Point> path;
};

struct Monitor {
    Point starts[4];
    uint16_t winning_index = no_winner;
    Thread_state t[4];
};

void rebuild_path(Grid const &g, Thread_state &ts, Point cur) {
    // bfs_threads.cc:80-84 / :178-183: cur = seen.at(cur); while (cur.row > 0) {push; cur = seen.at(cur)}
    int64_t p = ts.parent[g.idx(cur.row, cur.col)];
    while (p >= 0) {
        Point const q = {static_cast<int>(p / g.cols), static_cast<int>(p % g.cols)};
        if (!(q.row > 0)) {
            break;
        }
        ts.path.push_back(q);
        p = ts.parent[static_cast<size_t>(p)];
    }
}

// solvers/bfs_threads.cc:35-85
void hunter(Grid const &g, Monitor &m, int index) {
    uint16_t const paint = static_cast<uint16_t>((1u << index) << thread_paint_shift);
    Thread_state &ts = m.t[index];
    ts.parent.assign(static_cast<size_t>(g.rows) * g.cols, -2);
    Point const start = m.starts[index];
    ts.parent[g.idx(start.row, start.col)] = -1;
    ts.queue.push_back(static_cast<int64_t>(g.idx(start.row, start.col)));
    Point cur = start;
    while (ts.head < ts.queue.size()) {
        if (m.winning_index != no_winner) {
            break;
        }
        int64_t const ci = ts.queue[ts.head++];
        cur = {static_cast<int>(ci / g.cols), static_cast<int>(ci % g.cols)};
        if (g.at(cur.row, cur.col) & finish_bit) {
            if (m.winning_index == no_winner) {
                m.winning_index = static_cast<uint16_t>(index);
            }
            break;
        }
        g.at(cur.row, cur.col) |= paint;
        for (int count = 0, i = index; count < 4; count++, ++i %= 4) {
            Point const next = {cur.row + dirs[i].row, cur.col + dirs[i].col};
            size_t const ni = g.idx(next.row, next.col);
            if (ts.parent[ni] == -2 && (g.at(next.row, next.col) & path_bit)) {
                ts.parent[ni] = ci;
                ts.queue.push_back(static_cast<int64_t>(ni));
            }
        }
    }
    rebuild_path(g, ts, cur);
}

// solvers/bfs_threads.cc:144-185
void gatherer(Grid const &g, Monitor &m, int index) {
    uint16_t const seen_bit = static_cast<uint16_t>((1u << index) << thread_cache_shift);
    uint16_t const paint = static_cast<uint16_t>((1u << index) << thread_paint_shift);
    Thread_state &ts = m.t[index];
    ts.parent.assign(static_cast<size_t>(g.rows) * g.cols, -2);
    Point const start = m.starts[index];
    ts.parent[g.idx(start.row, start.col)] = -1;
    ts.queue.push_back(static_cast<int64_t>(g.idx(start.row, start.col)));
    Point cur = start;
    while (ts.head < ts.queue.size()) {
        int64_t const ci = ts.queue[ts.head++];
        cur = {static_cast<int>(ci / g.cols), static_cast<int>(ci % g.cols)};
        if ((g.at(cur.row, cur.col) & finish_bit) && !(g.at(cur.row, cur.col) & cache_mask)) {
            g.at(cur.row, cur.col) |= seen_bit;
            break;
        }
        g.at(cur.row, cur.col) |= paint;
        g.at(cur.row, cur.col) |= seen_bit;
        for (int count = 0, i = index; count < 4; count++, ++i %= 4) {
            Point const next = {cur.row + dirs[i].row, cur.col + dirs[i].col};
            size_t const ni = g.idx(next.row, next.col);
            if (ts.parent[ni] == -2 && (g.at(next.row, next.col) & path_bit)) {
                ts.parent[ni] = ci;
                ts.queue.push_back(static_cast<int64_t>(ni));
            }
        }
    }
    rebuild_path(g, ts, cur);
    m.winning_index = static_cast<uint16_t>(index);
}

// Plain level BFS over path_bit from (sr,sc); the start is expanded whether or not it is a
// path square (the reference pushes it unconditionally, bfs_threads.cc:41-43).
void bfs_levels(Grid const &g, Point s, std::vector<uint32_t> &dist) {
    dist.assign(static_cast<size_t>(g.rows) * g.cols, INF);
    std::vector<int64_t> q;
    q.reserve(1024);
    dist[g.idx(s.row, s.col)] = 0;
    q.push_back(static_cast<int64_t>(g.idx(s.row, s.col)));
    for (size_t head = 0; head < q.size(); head++) {
        int64_t const ci = q[head];
        int const r = static_cast<int>(ci / g.cols);
        int const c = static_cast<int>(ci % g.cols);
        for (Point const &d : dirs) {
            int const nr = r + d.row;
            int const nc = c + d.col;
            if (nr < 0 || nr >= g.rows || nc < 0 || nc >= g.cols) {
                continue;
            }
            size_t const ni = g.idx(nr, nc);
            if (dist[ni] == INF && (g.at(nr, nc) & path_bit)) {
                dist[ni] = dist[static_cast<size_t>(ci)] + 1;
                q.push_back(static_cast<int64_t>(ni));
            }
        }
    }
}

// Canonical shortest path (what the GPU path extraction produces): from `finish` step to the
// FIRST neighbour in N,E,S,W order whose level is one less, until level 0.  Output is ordered
// like the reference's thread_paths (bfs_threads.cc:80-84): finish side first, the start last,
// the finish itself excluded; length == dist(finish).
void canonical_path(Grid const &g, std::vector<uint32_t> const &dist, Point finish,
                    std::vector<Point> &path) {
    path.clear();
    uint32_t d = dist[g.idx(finish.row, finish.col)];
    if (d == INF) {
        return;
    }
    Point cur = finish;
    while (d > 0) {
        for (Point const &dd : dirs) {
            Point const n = {cur.row + dd.row, cur.col + dd.col};
            if (n.row < 0 || n.row >= g.rows || n.col < 0 || n.col >= g.cols) {
                continue;
            }
            if (dist[g.idx(n.row, n.col)] == d - 1) {
                cur = n;
                break;
            }
        }
        path.push_back(cur);
        d--;
    }
}

void emit_path(std::vector<Point> const &p, int32_t *out, int64_t max_path, int64_t *len) {
    *len = static_cast<int64_t>(p.size());
    if (!out) {
        return;
    }
    for (int64_t j = 0; j < std::min<int64_t>(max_path, static_cast<int64_t>(p.size())); j++) {
        out[2 * j] = p[static_cast<size_t>(j)].row;
        out[2 * j + 1] = p[static_cast<size_t>(j)].col;
    }
}

} // namespace

extern "C" {

uint32_t orc_inf(void) { return INF; }

// painters/distance.cc:129-153 (the BFS; the painter that follows is orc_painter_rgb).
// dist[r*cols+c] = level (the reference's uint64 map value; < 2^32 for every maze that fits in
// memory), INF where the map has no entry.  grid gets Rgb::measure (bit 9) on reached squares.
int orc_distance_map(int rows, int cols, uint16_t *grid, uint32_t *dist, uint64_t *max_out,
                     uint64_t *settled_out) {
    Grid const g{rows, cols, grid};
    int const row_mid = rows / 2;
    int const col_mid = cols / 2;
    Point const start = {row_mid + 1 - (row_mid % 2), col_mid + 1 - (col_mid % 2)};
    std::fill(dist, dist + static_cast<size_t>(rows) * cols, INF);
    uint64_t max = 0;
    uint64_t settled = 1;
    dist[g.idx(start.row, start.col)] = 0; // Distance_map map(start, 0)  :135
    std::vector<int64_t> q;
    q.push_back(static_cast<int64_t>(g.idx(start.row, start.col)));
    g.at(start.row, start.col) |= measure_bit;
    for (size_t head = 0; head < q.size(); head++) {
        int64_t const ci = q[head];
        int const r = static_cast<int>(ci / cols);
        int const c = static_cast<int>(ci % cols);
        uint32_t const d = dist[static_cast<size_t>(ci)];
        max = std::max<uint64_t>(max, d);
        for (Point const &p : dirs) {
            int const nr = r + p.row;
            int const nc = c + p.col;
            uint16_t const s = g.at(nr, nc);
            if (!(s & path_bit) || (s & measure_bit)) {
                continue;
            }
            g.at(nr, nc) |= measure_bit;
            dist[g.idx(nr, nc)] = d + 1;
            settled++;
            q.push_back(static_cast<int64_t>(g.idx(nr, nc)));
        }
    }
    *max_out = max;
    *settled_out = settled;
    return 0;
}

// painters/runs.cc:130-161 (the BFS of Runs::paint_runs; its painter is orc_painter_rgb with runs for levels: :45-70).
// runs[r*cols+c] = length of the straight run the FIFO search's tree path ends with at that square (0 at the start),
// INF where the map has no entry.  grid gets Rgb::measure on reached squares.  Any maze (FIFO order, N,E,S,W).
int orc_runs_map(int rows, int cols, uint16_t *grid, uint32_t *runs, uint64_t *max_out, uint64_t *settled_out) {
    Grid const g{rows, cols, grid};
    int const row_mid = rows / 2;
    int const col_mid = cols / 2;
    Point const start = {row_mid + 1 - (row_mid % 2), col_mid + 1 - (col_mid % 2)};
    std::fill(runs, runs + static_cast<size_t>(rows) * cols, INF);
    struct Run_point { // runs.cc:39-43
        uint32_t len;
        int64_t prev, cur;
    };
    uint64_t max = 0, settled = 1;
    int64_t const s0 = static_cast<int64_t>(g.idx(start.row, start.col));
    runs[s0] = 0; // Run_map map(start, 0)  :136
    std::vector<Run_point> q;
    q.push_back({0, s0, s0});
    g.at(start.row, start.col) |= measure_bit;
    for (size_t head = 0; head < q.size(); head++) {
        Run_point const cur = q[head];
        int const r = static_cast<int>(cur.cur / cols), c = static_cast<int>(cur.cur % cols);
        int const pr = static_cast<int>(cur.prev / cols), pc = static_cast<int>(cur.prev % cols);
        max = std::max<uint64_t>(max, cur.len);
        for (Point const &p : dirs) {
            int const nr = r + p.row, nc = c + p.col;
            uint16_t const sq = g.at(nr, nc);
            if (!(sq & path_bit) || (sq & measure_bit)) {
                continue;
            }
            uint32_t const len = std::abs(nr - pr) == std::abs(nc - pc) ? 1u : cur.len + 1u; // a turn starts a new run  :148-151
            g.at(nr, nc) |= measure_bit;
            runs[g.idx(nr, nc)] = len;
            settled++;
            q.push_back({len, cur.cur, static_cast<int64_t>(g.idx(nr, nc))});
        }
    }
    *max_out = max;
    *settled_out = settled;
    return 0;
}

// painters/distance.cc:45-70: the three ints Rgb::print_rgb prints per map cell.
// rgb[3*i..] = -1 where the reference prints a wall glyph instead (no map entry).
int orc_painter_rgb(int rows, int cols, uint32_t const *dist, uint64_t max, int color_choice,
                    int32_t *rgb) {
    size_t const n = static_cast<size_t>(rows) * cols;
    for (size_t i = 0; i < n; i++) {
        if (dist[i] == INF) {
            rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = -1;
            continue;
        }
        auto const intensity
            = static_cast<double>(max - dist[i]) / static_cast<double>(max);
        auto const dark = static_cast<uint8_t>(255.0 * intensity);
        auto const bright = static_cast<uint8_t>(128) + static_cast<uint8_t>(127.0 * intensity);
        uint16_t color[3] = {dark, dark, dark};
        color[color_choice] = static_cast<uint16_t>(bright);
        rgb[3 * i] = color[0];
        rgb[3 * i + 1] = color[1];
        rgb[3 * i + 2] = color[2];
    }
    return 0;
}

int orc_bfs_levels(int rows, int cols, uint16_t const *grid, int sr, int sc, uint32_t *dist) {
    Grid const g{rows, cols, const_cast<uint16_t *>(grid)};
    std::vector<uint32_t> d;
    bfs_levels(g, {sr, sc}, d);
    std::memcpy(dist, d.data(), d.size() * sizeof(uint32_t));
    return 0;
}

// Same contract as ref_solver_sequential in oracle/ref_harness.cc.
int orc_solver_sequential(int game, int rows, int cols, uint16_t *grid, int32_t const *starts,
                          int32_t const *order, int n_order, int32_t *paths_out,
                          int64_t max_path, int64_t *path_len_out, int32_t *winner_out) {
    Grid const g{rows, cols, grid};
    Monitor m;
    for (int i = 0; i < 4; i++) {
        m.starts[i] = {starts[2 * i], starts[2 * i + 1]};
    }
    for (int k = 0; k < n_order; k++) {
        if (game == 0) {
            hunter(g, m, order[k]);
        } else {
            gatherer(g, m, order[k]);
        }
    }
    for (int i = 0; i < 4; i++) {
        emit_path(m.t[i].path, paths_out ? paths_out + static_cast<size_t>(i) * max_path * 2 : nullptr,
                  max_path, &path_len_out[i]);
    }
    *winner_out = m.winning_index;
    return 0;
}

int orc_is_valid_start_or_finish(int rows, int cols, uint16_t const *grid, int r, int c) {
    Grid const g{rows, cols, const_cast<uint16_t *>(grid)};
    return is_valid_start_or_finish(g, {r, c}) ? 1 : 0;
}

int orc_find_nearest_square(int rows, int cols, uint16_t const *grid, int r, int c,
                            int32_t *out) {
    Grid const g{rows, cols, const_cast<uint16_t *>(grid)};
    Point const p = find_nearest_square(g, {r, c});
    out[0] = p.row;
    out[1] = p.col;
    return p.row < 0 ? -1 : 0;
}

int orc_pick_random_point(int rows, int cols, uint16_t const *grid, unsigned seed, int32_t *out) {
    Grid const g{rows, cols, const_cast<uint16_t *>(grid)};
    Point const p = pick_random_point(g, seed);
    out[0] = p.row;
    out[1] = p.col;
    return p.row < 0 ? -1 : 0;
}

int orc_set_corner_starts(int rows, int cols, uint16_t const *grid, int32_t *out8) {
    Grid const g{rows, cols, const_cast<uint16_t *>(grid)};
    Point s[4];
    set_corner_starts(g, s);
    for (int i = 0; i < 4; i++) {
        out8[2 * i] = s[i].row;
        out8[2 * i + 1] = s[i].col;
    }
    return 0;
}

// ---- placement: the part of each game that runs before the threads are spawned -------------
// `seed` is the value the FIRST std::random_device{}() of the call returns; each later one
// returns one more (matching oracle/ref_harness.cc's seeded device).

// Bfs::hunt, solvers/bfs_threads.cc:245-251
int orc_place_hunt(int rows, int cols, uint16_t *grid, unsigned seed, int32_t *start,
                   int32_t *finish) {
    Grid const g{rows, cols, grid};
    Point const s = pick_random_point(g, seed);
    if (s.row < 0) {
        return -1;
    }
    g.at(s.row, s.col) |= start_bit;
    Point const f = pick_random_point(g, seed + 1);
    if (f.row < 0) {
        return -1;
    }
    g.at(f.row, f.col) |= finish_bit;
    start[0] = s.row;
    start[1] = s.col;
    finish[0] = f.row;
    finish[1] = f.col;
    return 0;
}

// Bfs::gather, solvers/bfs_threads.cc:335-344
int orc_place_gather(int rows, int cols, uint16_t *grid, unsigned seed, int32_t *start,
                     int32_t *finishes8) {
    Grid const g{rows, cols, grid};
    Point const s = pick_random_point(g, seed);
    if (s.row < 0) {
        return -1;
    }
    g.at(s.row, s.col) |= start_bit;
    start[0] = s.row;
    start[1] = s.col;
    for (int i = 0; i < num_gather_finishes; i++) {
        Point const f = pick_random_point(g, seed + 1 + static_cast<unsigned>(i));
        if (f.row < 0) {
            return -1;
        }
        g.at(f.row, f.col) |= finish_bit;
        finishes8[2 * i] = f.row;
        finishes8[2 * i + 1] = f.col;
    }
    return 0;
}

// Bfs::corners, solvers/bfs_threads.cc:422-439.  starts8 = monitor.starts AFTER the shuffle,
// i.e. starts8[2k..] is where colour k begins.
int orc_place_corners(int rows, int cols, uint16_t *grid, unsigned seed, int32_t *starts8,
                      int32_t *finish) {
    Grid const g{rows, cols, grid};
    Point s[4];
    set_corner_starts(g, s);
    for (Point const &p : s) {
        if (p.row < 0) {
            return -1;
        }
        g.at(p.row, p.col) |= start_bit;
    }
    Point const f = {rows / 2, cols / 2};
    for (Point const &p : dirs) {
        g.at(f.row + p.row, f.col + p.col) |= path_bit;
    }
    g.at(f.row, f.col) |= path_bit;
    g.at(f.row, f.col) |= finish_bit;
    std::vector<Point> v(s, s + 4);
    std::shuffle(v.begin(), v.end(), std::mt19937(seed));
    for (int i = 0; i < 4; i++) {
        starts8[2 * i] = v[static_cast<size_t>(i)].row;
        starts8[2 * i + 1] = v[static_cast<size_t>(i)].col;
    }
    finish[0] = f.row;
    finish[1] = f.col;
    return 0;
}

// ---- canonical ("lock-step") games: the deterministic model the GPU path implements --------
// The reference's threads race; what is schedule-independent is the set of
// cells closer to the start than the finish a colour is looking for.  Canonical rule: colour k
// paints exactly {c : d_k(c) < D_k}, where D_k is the level at which k's game ends.  Any real
// schedule of the reference paints a superset of that and a subset of {d_k(c) <= D_k}; the
// tests check this sandwich against the reference itself.

// hunt: one start, one finish, all four colours share the start (bfs_threads.cc:243-280).
// winner = 0 (every colour reaches the finish at the same level; lowest index wins the CAS in
// lock-step).  winner = -1 and no repaint if the finish is unreachable.
int orc_canon_hunt(int rows, int cols, uint16_t *grid, int32_t const *start,
                   int32_t const *finish, int32_t *winner, int32_t *path, int64_t max_path,
                   int64_t *path_len) {
    Grid const g{rows, cols, grid};
    std::vector<uint32_t> d;
    bfs_levels(g, {start[0], start[1]}, d);
    uint32_t const T = d[g.idx(finish[0], finish[1])];
    size_t const n = static_cast<size_t>(rows) * cols;
    for (size_t i = 0; i < n; i++) {
        if (d[i] < T) { // d == INF is never < T
            grid[i] |= thread_paint_mask;
        }
    }
    std::vector<Point> p;
    if (T != INF) {
        *winner = 0;
        canonical_path(g, d, {finish[0], finish[1]}, p);
        uint16_t const winner_color = static_cast<uint16_t>(1u << thread_paint_shift);
        for (Point const &q : p) { // bfs_threads.cc:263-274
            g.at(q.row, q.col) &= static_cast<uint16_t>(~thread_paint_mask);
            g.at(q.row, q.col) |= winner_color;
        }
    } else {
        *winner = -1;
    }
    emit_path(p, path, max_path, path_len);
    return 0;
}

// gather: one start, four finishes (bfs_threads.cc:333-369).  Finishes sorted by
// (level, row, col); colour k claims the k-th.  claimed_by[j] = colour that claims
// finishes8[j] (or -1 if unreachable).  paths = 4 * max_path (row,col) pairs, path k for colour k.
int orc_canon_gather(int rows, int cols, uint16_t *grid, int32_t const *start,
                     int32_t const *finishes8, int32_t *claimed_by, int32_t *paths,
                     int64_t max_path, int64_t *path_len) {
    Grid const g{rows, cols, grid};
    std::vector<uint32_t> d;
    bfs_levels(g, {start[0], start[1]}, d);
    int order[4] = {0, 1, 2, 3};
    auto key = [&](int j) { return d[g.idx(finishes8[2 * j], finishes8[2 * j + 1])]; };
    std::sort(order, order + 4, [&](int a, int b) {
        if (key(a) != key(b)) {
            return key(a) < key(b);
        }
        if (finishes8[2 * a] != finishes8[2 * b]) {
            return finishes8[2 * a] < finishes8[2 * b];
        }
        return finishes8[2 * a + 1] < finishes8[2 * b + 1];
    });
    uint32_t D[4];
    for (int k = 0; k < 4; k++) {
        D[k] = key(order[k]);
        claimed_by[order[k]] = D[k] == INF ? -1 : k;
    }
    size_t const n = static_cast<size_t>(rows) * cols;
    for (size_t i = 0; i < n; i++) {
        if (d[i] == INF) {
            continue;
        }
        uint16_t bits = 0;
        for (int k = 0; k < 4; k++) {
            if (d[i] < D[k]) { // :163-164 paint and cache on every popped non-claimed square
                bits |= static_cast<uint16_t>((1u << k) << thread_paint_shift);
                bits |= static_cast<uint16_t>((1u << k) << thread_cache_shift);
            }
        }
        grid[i] |= bits;
    }
    for (int k = 0; k < 4; k++) { // :158-161 the claim
        if (D[k] != INF) {
            g.at(finishes8[2 * order[k]], finishes8[2 * order[k] + 1])
                |= static_cast<uint16_t>((1u << k) << thread_cache_shift);
        }
    }
    for (int k = 0; k < 4; k++) { // :355-364 recolour path.front() in thread order
        std::vector<Point> p;
        if (D[k] != INF) {
            canonical_path(g, d, {finishes8[2 * order[k]], finishes8[2 * order[k] + 1]}, p);
        }
        if (!p.empty()) {
            Point const &f = p.front();
            g.at(f.row, f.col) &= static_cast<uint16_t>(~thread_paint_mask);
            g.at(f.row, f.col) |= static_cast<uint16_t>((1u << k) << thread_paint_shift);
        }
        emit_path(p, paths ? paths + static_cast<size_t>(k) * max_path * 2 : nullptr, max_path,
                  &path_len[k]);
    }
    return 0;
}

// corners: colour k starts at starts8[2k..], all hunt the one finish (bfs_threads.cc:420-453).
// T = min_k d_k(finish), winner = lowest k attaining it; colour k paints {d_k < T}.  The static
// game does not repaint the winner's path (:446-452).  path = the winner's.
int orc_canon_corners(int rows, int cols, uint16_t *grid, int32_t const *starts8,
                      int32_t const *finish, int32_t *winner, int32_t *path, int64_t max_path,
                      int64_t *path_len) {
    Grid const g{rows, cols, grid};
    std::vector<uint32_t> d[4];
    uint32_t T = INF;
    int w = -1;
    for (int k = 0; k < num_threads; k++) {
        bfs_levels(g, {starts8[2 * k], starts8[2 * k + 1]}, d[k]);
        uint32_t const dk = d[k][g.idx(finish[0], finish[1])];
        if (dk < T) {
            T = dk;
            w = k;
        }
    }
    size_t const n = static_cast<size_t>(rows) * cols;
    for (size_t i = 0; i < n; i++) {
        uint16_t bits = 0;
        for (int k = 0; k < num_threads; k++) {
            if (d[k][i] < T) {
                bits |= static_cast<uint16_t>((1u << k) << thread_paint_shift);
            }
        }
        grid[i] |= bits;
    }
    *winner = w;
    std::vector<Point> p;
    if (w >= 0) {
        canonical_path(g, d[w], {finish[0], finish[1]}, p);
    }
    emit_path(p, path, max_path, path_len);
    return 0;
}

} // extern "C"
