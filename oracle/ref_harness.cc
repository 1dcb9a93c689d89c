// oracle/ref_harness.cc -- TEST INFRASTRUCTURE ONLY.
//
// White-box C harness around the REAL reference text (see oracle/Makefile for how the text is
// brought in: sed-stripped module directives, one unity TU per family, generated under
// oracle/_ref/ and never committed).  Everything in this file is mine; the reference code it
// drives arrives through the #include of the generated .inc below.
//
// Determinism: the reference seeds every RNG with std::random_device{}() (e.g.
// builders/kruskal.cc:39-40, solvers/solve_utilities.cc:277).  After all std headers are in,
// `random_device` is re-pointed at a counter whose value the harness sets, so the unmodified
// reference code becomes reproducible: the n-th RNG constructed after ref_seed(s) is seeded s+n.

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <iterator>
#include <mutex>
#include <numeric>
#include <optional>
#include <random>
#include <span>
#include <sstream>
#include <stack>
#include <streambuf>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace std {
struct oracle_seeded_device {
    using result_type = unsigned;
    static inline unsigned next_seed = 0;
    unsigned operator()() { return next_seed++; }
};
} // namespace std
#define random_device oracle_seeded_device

#if defined(REF_FAMILY_DIST)
namespace Maze { struct Point; }
// filled by the capture hook the Makefile puts in front of distance.cc:154 `painter(maze, map);`
static uint64_t g_captured_max = 0;
static std::vector<std::pair<std::pair<int, int>, uint64_t>> g_captured;
template <class Map> static void oracle_capture_distance_map(uint64_t max, Map const &m) {
    g_captured_max = max;
    g_captured.clear();
    g_captured.reserve(m.size());
    for (auto const &kv : m) {
        g_captured.push_back({{kv.first.row, kv.first.col}, kv.second});
    }
}
#include "_ref/unity_dist.inc"
#elif defined(REF_FAMILY_RUNS)
namespace Maze { struct Point; }
// filled by the capture hook the Makefile puts in front of runs.cc:158 `painter(maze, map);`
static uint64_t g_captured_max = 0;
static std::vector<std::pair<std::pair<int, int>, uint64_t>> g_captured;
template <class Map> static void oracle_capture_run_map(uint64_t max, Map const &m) {
    g_captured_max = max;
    g_captured.clear();
    g_captured.reserve(m.size());
    for (auto const &kv : m) {
        g_captured.push_back({{kv.first.row, kv.first.col}, kv.second});
    }
}
#include "_ref/unity_runs.inc"
#elif defined(REF_FAMILY_BFS)
#include "_ref/unity_bfs.inc"
#else
#error "define REF_FAMILY_BFS, REF_FAMILY_DIST or REF_FAMILY_RUNS"
#endif
#undef random_device

namespace {

struct Null_buf : std::streambuf {
    int overflow(int c) override { return c; }
    std::streamsize xsputn(char const *, std::streamsize n) override { return n; }
};

// RAII: send std::cout to `to` for the scope (the reference prints the whole maze after every
// build and solve: build_utilities.cc:51-61, solve_utilities.cc:205-214).
struct Cout_redirect {
    std::streambuf *old;
    explicit Cout_redirect(std::streambuf *to) : old(std::cout.rdbuf(to)) {}
    ~Cout_redirect() { std::cout.rdbuf(old); }
};

Maze::Maze make_maze(int rows, int cols, uint16_t const *squares) {
    Maze::Maze_args args;
    args.odd_rows = static_cast<uint64_t>(rows);
    args.odd_cols = static_cast<uint64_t>(cols);
    Maze::Maze maze(args);
    if (squares) {
        for (int r = 0; r < rows; r++) {
            auto row = maze[r];
            for (int c = 0; c < cols; c++) {
                row[c].store(squares[static_cast<size_t>(r) * cols + c]);
            }
        }
    }
    return maze;
}

void read_back(Maze::Maze const &maze, uint16_t *out) {
    int const rows = maze.row_size();
    int const cols = maze.col_size();
    for (int r = 0; r < rows; r++) {
        auto row = maze[r];
        for (int c = 0; c < cols; c++) {
            out[static_cast<size_t>(r) * cols + c] = row[c].load();
        }
    }
}

} // namespace

extern "C" {

void ref_seed(unsigned seed) { std::oracle_seeded_device::next_seed = seed; }

// builder in {"arena","grid","rdfs","kruskal","wilson"}; mod 0 none, 1 cross, 2 x.
// Returns 0, or -1 for an unknown builder.  out = rows*cols squares, row-major.
int ref_build(char const *builder, int rows, int cols, unsigned seed, int mod, uint16_t *out) {
    Null_buf nb;
    Cout_redirect quiet(&nb);
    ref_seed(seed);
    Maze::Maze maze = make_maze(rows, cols, nullptr);
    std::string const b(builder);
    if (b == "arena") {
        Arena::generate_maze(maze);
    } else if (b == "grid") {
        Grid::generate_maze(maze);
    } else if (b == "rdfs") {
        Recursive_backtracker::generate_maze(maze);
    } else if (b == "kruskal") {
        Kruskal::generate_maze(maze);
    } else if (b == "wilson") {
        Wilson_path_carver::generate_maze(maze);
    } else {
        return -1;
    }
    if (mod == 1) {
        Mods::add_cross(maze);
    } else if (mod == 2) {
        Mods::add_x(maze);
    }
    read_back(maze, out);
    return 0;
}

#if defined(REF_FAMILY_DIST)

// Runs the reference's Distance::paint_distance_from_center (painters/distance.cc:129-156) on
// the given squares.  grid is updated in place (measure bit 9 on reached cells).
//   dist_out[r*cols+c] = captured map value, UINT64_MAX where the map has no entry
//   rgb_out[(r*cols+c)*3..] = the three colour ints the un-hooked painter printed for that
//                             cell (distance.cc:55-62), or -1,-1,-1 where it printed a wall glyph
int ref_distance(int rows, int cols, uint16_t *grid, unsigned seed, uint64_t *dist_out,
                 uint64_t *max_out, int32_t *rgb_out) {
    std::stringbuf capture;
    {
        Cout_redirect grab(&capture);
        ref_seed(seed);
        Maze::Maze maze = make_maze(rows, cols, grid);
        Distance::paint_distance_from_center(maze);
        read_back(maze, grid);
    }
    size_t const n = static_cast<size_t>(rows) * cols;
    std::fill(dist_out, dist_out + n, UINT64_MAX);
    for (auto const &e : g_captured) {
        dist_out[static_cast<size_t>(e.first.first) * cols + e.first.second] = e.second;
    }
    *max_out = g_captured_max;
    if (rgb_out) {
        std::fill(rgb_out, rgb_out + 3 * n, -1);
        std::string const s = capture.str();
        // stream = ( ESC[ r+1 ; c+1 f  ( ESC[38;2; R;G;B m <brush> ESC[0m | <wall glyph> ) )*
        size_t i = 0;
        auto read_int = [&](int &v) {
            v = 0;
            bool any = false;
            while (i < s.size() && s[i] >= '0' && s[i] <= '9') {
                v = v * 10 + (s[i] - '0');
                i++;
                any = true;
            }
            return any;
        };
        while (i + 2 < s.size()) {
            if (s[i] != '\033' || s[i + 1] != '[') {
                i++;
                continue;
            }
            size_t const save = i;
            i += 2;
            int r = 0, c = 0;
            if (!read_int(r) || i >= s.size() || s[i] != ';') {
                i = save + 1;
                continue;
            }
            i++;
            if (!read_int(c) || i >= s.size() || s[i] != 'f') {
                i = save + 1;
                continue;
            }
            i++;
            if (s.compare(i, 7, "\033[38;2;") == 0) {
                i += 7;
                int R = 0, G = 0, B = 0;
                read_int(R);
                i++; // ;
                read_int(G);
                i++; // ;
                read_int(B);
                size_t const cell = static_cast<size_t>(r - 1) * cols + (c - 1);
                if (r >= 1 && r <= rows && c >= 1 && c <= cols) {
                    rgb_out[3 * cell + 0] = R;
                    rgb_out[3 * cell + 1] = G;
                    rgb_out[3 * cell + 2] = B;
                }
            }
        }
    }
    return 0;
}

#endif // REF_FAMILY_DIST

#if defined(REF_FAMILY_RUNS)
// Runs the reference's Runs::paint_runs (painters/runs.cc:130-161) on the given squares.  grid is updated in place
// (measure bit 9 on reached squares).  runs_out[r*cols+c] = captured map value, UINT64_MAX where the map has no entry.
int ref_runs(int rows, int cols, uint16_t *grid, unsigned seed, uint64_t *runs_out, uint64_t *max_out) {
    Null_buf nb;
    {
        Cout_redirect quiet(&nb);
        ref_seed(seed);
        Maze::Maze maze = make_maze(rows, cols, grid);
        Runs::paint_runs(maze);
        read_back(maze, grid);
    }
    size_t const n = static_cast<size_t>(rows) * cols;
    std::fill(runs_out, runs_out + n, UINT64_MAX);
    for (auto const &e : g_captured) {
        runs_out[static_cast<size_t>(e.first.first) * cols + e.first.second] = e.second;
    }
    *max_out = g_captured_max;
    return 0;
}
#endif // REF_FAMILY_RUNS

#if defined(REF_FAMILY_BFS)

// White-box: one Bfs_monitor, then the anonymous-namespace hunter (game 0,
// bfs_threads.cc:35-85) or gatherer (game 1, :144-185) is called ONCE PER ENTRY of order[],
// sequentially on this thread (a legal schedule of the reference's 4 threads: thread
// order[0] runs to completion, then order[1], ...).  starts = 4 (row,col) pairs = monitor.starts.
// start/finish bits must already be in grid.  paths_out holds 4 * max_path (row,col) pairs.
int ref_solver_sequential(int game, int rows, int cols, uint16_t *grid, int32_t const *starts,
                          int32_t const *order, int n_order, int32_t *paths_out,
                          int64_t max_path, int64_t *path_len_out, int32_t *winner_out) {
    Null_buf nb;
    Cout_redirect quiet(&nb);
    Maze::Maze maze = make_maze(rows, cols, grid);
    Sutil::Bfs_monitor monitor;
    for (int i = 0; i < 4; i++) {
        monitor.starts.push_back({starts[2 * i], starts[2 * i + 1]});
    }
    for (int k = 0; k < n_order; k++) {
        auto const idx = static_cast<uint16_t>(order[k]);
        Sutil::Thread_id const id{idx, Sutil::thread_bits.at(idx)};
        if (game == 0) {
            hunter(maze, monitor, id);
        } else {
            gatherer(maze, monitor, id);
        }
    }
    read_back(maze, grid);
    for (int i = 0; i < 4; i++) {
        auto const &p = monitor.thread_paths.at(i);
        path_len_out[i] = static_cast<int64_t>(p.size());
        for (int64_t j = 0; j < std::min<int64_t>(max_path, p.size()); j++) {
            paths_out[(i * max_path + j) * 2 + 0] = p[j].row;
            paths_out[(i * max_path + j) * 2 + 1] = p[j].col;
        }
    }
    *winner_out = monitor.winning_index.load();
    return 0;
}

// The full games exactly as run_maze calls them (real std::threads, racy by design):
// game 0 Bfs::hunt (:243-280), 1 Bfs::gather (:333-369), 2 Bfs::corners (:420-453).
int ref_game(int game, int rows, int cols, uint16_t *grid, unsigned seed) {
    Null_buf nb;
    Cout_redirect quiet(&nb);
    ref_seed(seed);
    Maze::Maze maze = make_maze(rows, cols, grid);
    if (game == 0) {
        Bfs::hunt(maze);
    } else if (game == 1) {
        Bfs::gather(maze);
    } else if (game == 2) {
        Bfs::corners(maze);
    } else {
        return -1;
    }
    read_back(maze, grid);
    return 0;
}

// Sutil::pick_random_point (solve_utilities.cc:275-285) with the RNG seeded `seed`.
int ref_pick_random_point(int rows, int cols, uint16_t const *grid, unsigned seed, int32_t *out) {
    Null_buf nb;
    Cout_redirect quiet(&nb);
    ref_seed(seed);
    Maze::Maze maze = make_maze(rows, cols, grid);
    Maze::Point const p = Sutil::pick_random_point(maze);
    out[0] = p.row;
    out[1] = p.col;
    return 0;
}

// Sutil::set_corner_starts (solve_utilities.cc:254-273).
int ref_set_corner_starts(int rows, int cols, uint16_t const *grid, int32_t *out8) {
    Null_buf nb;
    Cout_redirect quiet(&nb);
    Maze::Maze maze = make_maze(rows, cols, grid);
    std::vector<Maze::Point> const s = Sutil::set_corner_starts(maze);
    for (int i = 0; i < 4; i++) {
        out8[2 * i] = s[i].row;
        out8[2 * i + 1] = s[i].col;
    }
    return 0;
}

#endif // REF_FAMILY_BFS

} // extern "C"
