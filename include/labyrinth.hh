// labyrinth.hh -- the reference's C++ interface for the BFS hot path, on top of the B200 library.
//
// The reference exposes this path as C++20-module free functions (module/labyrinth.cc:1-25):
//     Bfs::hunt / gather / corners (Maze::Maze&)                solvers/bfs_threads.cc:22-29
//     Distance::paint_distance_from_center (Maze::Maze&)        painters/distance.cc:23-26
//     <Builder>::generate_maze (Maze::Maze&), Mods::add_cross/add_x
// and keeps them in std::function tables keyed by the CLI strings (run_maze/run_maze.cc:97-124,
// measure/measure.cc:96-101).  This header re-creates those namespaces, names and signatures as a
// plain (non-module) header so that the same call sites compile with g++ 13 / nvcc 12.9, and routes
// the compute to liblabyrinth_b200.so through the C ABI in labyrinth_b200.h:
//
//     Maze::Maze (host vector<Square>)  --lab_grid_create-->  HBM  --lab_bfs_* / lab_distance_map-->
//     --lab_grid_download-->  the same vector<Square>, then the reference's terminal rendering.
//
// Start/finish selection stays on the host with the reference's rules and RNG stream
// (lab_place_*).  There is no CPU fallback: if the library reports an error the wrapper prints it
// and std::abort()s, which is what the reference does on its own fatal paths
// (solvers/solve_utilities.cc:200-202,247-251).
//
// Animated variants (animate_*) keep their signatures; on this path they compute on the GPU and
// then draw the finished frame (the per-square sleep/redraw of the reference is presentation only).
#pragma once

#include <array>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <random>
#include <span>
#include <string>
#include <string_view>
#include <vector>

#include <optional>
#include <thread>
#include <chrono>
#include <algorithm>
#include "labyrinth_b200.h"

namespace Speed {
using Speed_unit = int;
enum class Speed { instant = 0, speed_1, speed_2, speed_3, speed_4, speed_5, speed_6, speed_7 }; // speed/speed.cc:7-16
} // namespace Speed

namespace Maze {

using Square_bits = uint16_t;

// maze/maze.cc:66-94: a Square is one atomic 16-bit word; layout-identical here (sizeof == 2), so a
// row of Squares can be handed to the C ABI as uint16_t*.
class Square {
  public:
    constexpr explicit Square(Square_bits a) : u16(a) {}
    constexpr Square() = default;
    Square(Square const &o) : u16(o.u16.load()) {}
    Square &operator=(Square const &o) noexcept {
        u16.store(o.u16.load());
        return *this;
    }
    Square operator&(Square_bits b) const noexcept { return Square(static_cast<Square_bits>(u16.load() & b)); }
    Square operator|(Square_bits b) const noexcept { return Square(static_cast<Square_bits>(u16.load() | b)); }
    Square &operator|=(Square_bits b) noexcept {
        u16 |= b;
        return *this;
    }
    Square operator&=(Square_bits b) noexcept {
        u16 &= b;
        return *this;
    }
    explicit operator bool() const noexcept { return u16.load() != 0; }
    uint16_t load() const noexcept { return u16.load(std::memory_order_relaxed); }
    void store(Square_bits b) noexcept { u16.store(b); }

  private:
    std::atomic_uint16_t u16{0};
};
static_assert(sizeof(Square) == sizeof(uint16_t), "Square must be layout compatible with uint16_t");

enum class Maze_style : uint8_t { sharp = 0, round, doubles, bold, contrast, spikes };
struct Point {
    int row;
    int col;
};
struct Maze_args {
    uint64_t odd_rows = 31;
    uint64_t odd_cols = 111;
    Maze_style style = Maze_style::sharp;
};

constexpr Square_bits path_bit{0x2000}, start_bit{0x4000}, builder_bit{0x1000}, wall_mask{0x000F};

// maze/maze.cc:127-144
class Maze {
  public:
    explicit Maze(Maze_args const &args)
        : rows_(static_cast<int>(args.odd_rows)), cols_(static_cast<int>(args.odd_cols)),
          maze_(args.odd_rows * args.odd_cols), style_(static_cast<int>(args.style)) {}
    std::span<Square> operator[](uint64_t r) { return {&maze_.at(r * cols_), static_cast<uint64_t>(cols_)}; }
    std::span<Square const> operator[](uint64_t r) const { return {&maze_.at(r * cols_), static_cast<uint64_t>(cols_)}; }
    int row_size() const { return rows_; }
    int col_size() const { return cols_; }
    int style_index() const { return style_; }
    // what the C ABI borrows for the duration of a call
    uint16_t *squares() { return reinterpret_cast<uint16_t *>(maze_.data()); }
    uint16_t const *squares() const { return reinterpret_cast<uint16_t const *>(maze_.data()); }

  private:
    int rows_, cols_;
    std::vector<Square> maze_;
    int style_;
};

} // namespace Maze

// Knobs the reference does not have (it seeds everything from std::random_device and never times).
namespace Lab {
inline bool g_seeded = false;
inline uint32_t g_next_seed = 0;
inline bool g_quiet = false; // skip the terminal rendering (benchmarks, tests)
inline lab_timing g_last_timing{};
inline void seed(uint32_t s) {
    g_seeded = true;
    g_next_seed = s;
}
// stands in for std::random_device{}(): consecutive values when seeded, as in the oracle harness
inline uint32_t device_seed() { return g_seeded ? g_next_seed++ : std::random_device{}(); }
inline void quiet(bool q) { g_quiet = q; }
inline lab_timing const &last_timing() { return g_last_timing; }

[[noreturn]] inline void fatal(char const *what) {
    std::cerr << what << ": " << lab_last_error() << "\n";
    std::abort();
}
inline void check(int rc, char const *what) {
    if (rc != LAB_OK) fatal(what);
}

// RAII device grid for one call
struct Device_maze {
    lab_grid *g = nullptr;
    Maze::Maze &m;
    explicit Device_maze(Maze::Maze &maze) : m(maze) { check(lab_grid_create(m.row_size(), m.col_size(), m.squares(), &g), "lab_grid_create"); }
    void download() {
        check(lab_grid_download(g, m.squares()), "lab_grid_download");
        lab_last_timing(g, &g_last_timing);
    }
    ~Device_maze() { lab_grid_destroy(g); }
};
} // namespace Lab

// ---------------------------------------------------------------------------- terminal rendering
// (solvers/solve_utilities.cc:81-123,175-214,287-345 and maze/maze.cc:148-168; presentation only)
namespace Sutil {
constexpr uint16_t finish_bit{0x8000}, start_bit{0x4000}, thread_paint_mask{0x00F0}, cache_mask{0x0F00};
constexpr int thread_paint_shift = 4;
constexpr uint16_t no_winner = UINT16_MAX;
inline constexpr std::array<int, 16> thread_color_codes = {-1, 1, 2, 3, 4, 201, 87, 121, 183, 204, 106, 105, 57, 89, 60, 231};
inline constexpr std::array<std::string_view, 96> wall_styles = {{
    "■", "╵", "╶", "└", "╷", "│", "┌", "├", "╴", "┘", "─", "┴", "┐", "┤", "┬", "┼", // sharp
    "●", "╵", "╶", "╰", "╷", "│", "╭", "├", "╴", "╯", "─", "┴", "╮", "┤", "┬", "┼", // round
    "◫", "║", "═", "╚", "║", "║", "╔", "╠", "═", "╝", "═", "╩", "╗", "╣", "╦", "╬", // doubles
    "■", "╹", "╺", "┗", "╻", "┃", "┏", "┣", "╸", "┛", "━", "┻", "┓", "┫", "┳", "╋", // bold
    "█", "█", "█", "█", "█", "█", "█", "█", "█", "█", "█", "█", "█", "█", "█", "█", // contrast
    "✸", "╀", "┾", "╊", "╁", "╂", "╆", "╊", "┽", "╃", "┿", "╇", "╅", "╉", "╈", "╋", // spikes
}};
inline void print_color_block(int nibble) {
    if (nibble == 0) {
        std::cout << "\033[38;5;15m\033[48;255;0;0m╳ no thread won..\033[0m";
    } else {
        std::cout << "\033[38;5;" << thread_color_codes.at(nibble) << "m█\033[0m";
    }
}
inline void print_point(Maze::Maze const &maze, Maze::Point p) {
    uint16_t const s = maze[p.row][p.col].load();
    if (s & finish_bit) {
        std::cout << "\033[1m\033[38;5;87mF\033[0m";
    } else if (s & start_bit) {
        std::cout << "\033[1m\033[38;5;87mS\033[0m";
    } else if (s & thread_paint_mask) {
        print_color_block((s & thread_paint_mask) >> thread_paint_shift);
    } else if (!(s & Maze::path_bit)) {
        std::cout << wall_styles.at(maze.style_index() * 16 + (s & Maze::wall_mask));
    } else {
        std::cout << " ";
    }
}
inline void print_maze(Maze::Maze const &maze) {
    for (int r = 0; r < maze.row_size(); r++) {
        for (int c = 0; c < maze.col_size(); c++) {
            print_point(maze, {r, c});
        }
        std::cout << "\n";
    }
    std::cout << std::flush;
}
inline void print_overlap_key() {
    static constexpr std::array<std::string_view, 16> labels = {"", "0", "1", "1|0", "2", "2|0", "2|1", "2|1|0", "3", "3|0", "3|1", "3|1|0", "3|2", "3|2|0", "3|2|1", "3|2|1|0"};
    std::cout << "Overlap Key: 3_THREAD | 2_THREAD | 1_THREAD | 0_THREAD\n";
    for (int i = 1; i < 16; i++) {
        print_color_block(i);
        std::cout << " = " << labels[i] << (i % 5 == 0 ? "\n" : "   ");
    }
    std::cout << "\n";
}
inline void print_hunt_solution_message(uint16_t winner) {
    if (winner == no_winner) {
        print_color_block(0);
        return;
    }
    print_color_block(1 << winner);
    std::cout << " thread won!\n";
}
inline void print_gather_solution_message() {
    for (int k = 0; k < 4; k++) print_color_block(1 << k);
    std::cout << " All threads found their finish squares!\n";
}

// Printer::set_cursor_position (printers/printers.cc:14-17) and Sutil::flush_cursor_path_coordinate (solve_utilities.cc:216-221)
inline void set_cursor_position(Maze::Point p) { std::cout << "\033[" << p.row + 1 << ";" << p.col + 1 << "f"; }
inline void flush_cursor_path_coordinate(Maze::Maze const &maze, Maze::Point p) {
    set_cursor_position(p);
    print_point(maze, p);
    std::cout << std::flush;
}
inline constexpr std::array<int, 8> solver_speeds = {0, 20000, 10000, 5000, 2000, 1000, 500, 250}; // solve_utilities.cc:130-131
inline constexpr int overlap_key_and_message_height = 9;                                           // solve_utilities.cc:129
inline constexpr std::array<Maze::Point, 7> all_dirs = {{{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}}}; // :127-128 (SW is missing upstream)
// Sutil::deluminate_maze (solve_utilities.cc:227 ff.): the dark solvers start from a blank screen
inline void deluminate_maze(Maze::Maze const &maze) {
    for (int r = 0; r < maze.row_size(); r++) {
        for (int c = 0; c < maze.col_size(); c++) {
            set_cursor_position({r, c});
            std::cout << " ";
        }
    }
    std::cout << std::flush;
}
} // namespace Sutil

// ------------------------------------------------------------------------------------- builders
#define LAB_BUILDER_NS(NS, NAME)                                                                     \
    namespace NS {                                                                                   \
    inline void generate_maze(Maze::Maze &maze) {                                                    \
        Lab::check(lab_build_maze(NAME, "none", maze.row_size(), maze.col_size(), Lab::device_seed(), maze.squares()), "lab_build_maze"); \
    }                                                                                                \
    inline void animate_maze(Maze::Maze &maze, Speed::Speed) { generate_maze(maze); }                \
    }
LAB_BUILDER_NS(Recursive_backtracker, "rdfs")
LAB_BUILDER_NS(Wilson_path_carver, "wilson")
LAB_BUILDER_NS(Kruskal, "kruskal")
LAB_BUILDER_NS(Grid, "grid")
LAB_BUILDER_NS(Arena, "arena")
#undef LAB_BUILDER_NS

namespace Mods {
// builders/mods.cc:20-27,41-56.  The C ABI applies a mod as part of a build; applied to an existing
// maze it is re-done here on the host squares (build_path is idempotent and order free).
inline void carve(Maze::Maze &m, int r, int c) {
    if (r - 1 >= 0) m[r - 1][c] &= static_cast<uint16_t>(~0x4);
    if (r + 1 < m.row_size()) m[r + 1][c] &= static_cast<uint16_t>(~0x1);
    if (c - 1 >= 0) m[r][c - 1] &= static_cast<uint16_t>(~0x2);
    if (c + 1 < m.col_size()) m[r][c + 1] &= static_cast<uint16_t>(~0x8);
    m[r][c] |= Maze::path_bit;
}
inline void add_cross(Maze::Maze &m) {
    int const rows = m.row_size(), cols = m.col_size();
    for (int r = 0; r < rows; r++) {
        for (int c = 0; c < cols; c++) {
            if ((r == rows / 2 && c > 1 && c < cols - 2) || (c == cols / 2 && r > 1 && r < rows - 2)) {
                carve(m, r, c);
                if (c + 1 < cols - 2) carve(m, r, c + 1);
            }
        }
    }
}
inline void widen(Maze::Maze &m, int r, int c) {
    int const cols = m.col_size();
    carve(m, r, c);
    if (c + 1 < cols - 2) carve(m, r, c + 1);
    if (c - 1 > 1) carve(m, r, c - 1);
    if (c + 2 < cols - 2) carve(m, r, c + 2);
    if (c - 2 > 1) carve(m, r, c - 2);
}
inline void add_x(Maze::Maze &m) {
    float const row_size = static_cast<float>(m.row_size()) - 2.0F, col_size = static_cast<float>(m.col_size()) - 2.0F;
    float const ps = (2.0F - row_size) / (2.0F - col_size), pb = 2.0F - (2.0F * ps);
    float const ns = (2.0F - row_size) / (col_size - 2.0F), nbv = row_size - (2.0F * ns);
    for (int r = 1; r < m.row_size() - 1; r++) {
        int const on_pos = static_cast<int>((static_cast<float>(r) - pb) / ps);
        int const on_neg = static_cast<int>((static_cast<float>(r) - nbv) / ns);
        for (int c = 1; c < m.col_size() - 1; c++) {
            if (c == on_pos && c < m.col_size() - 2 && c > 1) widen(m, r, c);
            if (c == on_neg && c > 1 && c < m.col_size() - 2 && r < m.row_size() - 2) widen(m, r, c);
        }
    }
}
inline void add_cross_animated(Maze::Maze &m, Speed::Speed) { add_cross(m); }
inline void add_x_animated(Maze::Maze &m, Speed::Speed) { add_x(m); }
} // namespace Mods

// -------------------------------------------------------------------------------------- solvers
namespace Bfs {

// solvers/bfs_threads.cc:243-280
inline void hunt(Maze::Maze &maze) {
    lab_point start{}, finish{};
    Lab::check(lab_place_hunt(maze.row_size(), maze.col_size(), maze.squares(), Lab::device_seed(), &start, &finish), "Bfs::hunt placement");
    Lab::g_next_seed += Lab::g_seeded ? 1 : 0; // the reference draws two device seeds (start, finish)
    int32_t winner = -1;
    uint64_t path_len = 0;
    {
        Lab::Device_maze d(maze);
        Lab::check(lab_bfs_hunt(d.g, start, finish, &winner, nullptr, 0, &path_len), "lab_bfs_hunt");
        d.download();
    }
    if (Lab::g_quiet) return;
    Sutil::print_maze(maze);
    Sutil::print_overlap_key();
    Sutil::print_hunt_solution_message(winner < 0 ? Sutil::no_winner : static_cast<uint16_t>(winner));
    std::cout << "\n";
}

// solvers/bfs_threads.cc:333-369
inline void gather(Maze::Maze &maze) {
    lab_point start{}, finishes[4]{};
    Lab::check(lab_place_gather(maze.row_size(), maze.col_size(), maze.squares(), Lab::device_seed(), &start, finishes), "Bfs::gather placement");
    Lab::g_next_seed += Lab::g_seeded ? 4 : 0; // five device seeds upstream (start + 4 finishes)
    int32_t claimed[4];
    uint64_t lens[4];
    {
        Lab::Device_maze d(maze);
        lab_grid_set_option(d.g, LAB_OPT_SOLVER_LEVELS, 0); // (the static game shows the grid alone: no level field to keep)
        Lab::check(lab_bfs_gather(d.g, start, finishes, claimed, nullptr, 0, lens), "lab_bfs_gather");
        d.download();
    }
    if (Lab::g_quiet) return;
    Sutil::print_maze(maze);
    Sutil::print_overlap_key();
    Sutil::print_gather_solution_message();
    std::cout << "\n";
}

// solvers/bfs_threads.cc:420-453
inline void corners(Maze::Maze &maze) {
    lab_point starts[4]{}, finish{};
    Lab::check(lab_place_corners(maze.row_size(), maze.col_size(), maze.squares(), Lab::device_seed(), starts, &finish), "Bfs::corners placement");
    int32_t winner = -1;
    uint64_t path_len = 0;
    {
        Lab::Device_maze d(maze);
        Lab::check(lab_bfs_corners(d.g, starts, finish, &winner, nullptr, 0, &path_len), "lab_bfs_corners");
        d.download();
    }
    if (Lab::g_quiet) return;
    Sutil::print_maze(maze);
    Sutil::print_overlap_key();
    Sutil::print_hunt_solution_message(winner < 0 ? Sutil::no_winner : static_cast<uint16_t>(winner));
    std::cout << "\n";
}

} // namespace Bfs

// ---------------------------------------------------------------------- animated solvers: replay
// The reference's animated variants (solvers/bfs_threads.cc:282-331,371-418,455-514; dark ones: darkbfs_threads.cc:154-300)
// draw every square the moment a thread paints it and sleep solver_speeds[speed] microseconds per square and thread.  Here
// the search runs on the GPU first (the same entry points as the static variants: same grid afterwards); the frames are
// then REPLAYED from its level fields: the painted squares in level order -- the order a lock-step schedule of the
// reference's four threads reaches them --, each drawn with the colours it carries during the search, then the winner's
// path square by square (hunt, corners) or the claimed finishes' fronts (gather), as the reference does after its join.
namespace Lab {
struct Replay {
    std::vector<uint32_t> order;  // flat indices of the squares to draw, by the level at which they are painted
    std::vector<uint16_t> paint;  // the paint bits a square shows while the search runs
};
// levels[k]: field of colour k (one shared field: pass it four times); limit[k]: colour k paints levels below limit[k]
inline Replay make_replay(int rows, int cols, std::array<uint32_t const *, 4> levels, std::array<uint32_t, 4> limit) {
    size_t const n = static_cast<size_t>(rows) * cols;
    Replay out;
    std::vector<std::pair<uint32_t, uint32_t>> keyed; // (first level at which some colour paints it, index)
    out.paint.assign(n, 0);
    for (size_t i = 0; i < n; i++) {
        uint32_t first = 0xFFFFFFFFu;
        uint16_t bits = 0;
        for (int k = 0; k < 4; k++) {
            uint32_t const d = levels[k][i];
            if (d != 0xFFFFFFFFu && d < limit[k]) {
                bits = static_cast<uint16_t>(bits | (1u << (Sutil::thread_paint_shift + k)));
                first = d < first ? d : first;
            }
        }
        if (bits) {
            out.paint[i] = bits;
            keyed.push_back({first, static_cast<uint32_t>(i)});
        }
    }
    std::stable_sort(keyed.begin(), keyed.end(), [](auto const &a, auto const &b) { return a.first < b.first; });
    out.order.reserve(keyed.size());
    for (auto const &k : keyed) out.order.push_back(k.second);
    return out;
}
inline void nap(int microseconds) {
    if (microseconds > 0) std::this_thread::sleep_for(std::chrono::microseconds(microseconds));
}
// Manhattan levels of an open room (the searches of csrc/device/room.cuh write no field)
inline std::vector<uint32_t> room_levels(Maze::Maze const &maze, lab_point s) {
    int const rows = maze.row_size(), cols = maze.col_size();
    std::vector<uint32_t> d(static_cast<size_t>(rows) * cols, 0xFFFFFFFFu);
    for (int r = 0; r < rows; r++) {
        for (int c = 0; c < cols; c++) {
            if (maze[r][c].load() & Maze::path_bit) d[static_cast<size_t>(r) * cols + c] = static_cast<uint32_t>(std::abs(r - s.row) + std::abs(c - s.col));
        }
    }
    d[static_cast<size_t>(s.row) * cols + s.col] = 0;
    return d;
}
inline std::vector<uint32_t> levels_of(lab_grid *g, int plane, Maze::Maze const &maze, lab_point start) {
    std::vector<uint32_t> d(static_cast<size_t>(maze.row_size()) * maze.col_size());
    lab_stats st{};
    lab_last_stats(g, &st);
    if (st.tree_used == 2) return room_levels(maze, start);
    check(lab_levels_download(g, plane, d.data()), "lab_levels_download");
    return d;
}
inline void draw_replay(Maze::Maze &shown, Replay const &rp, int step_us) {
    int const cols = shown.col_size();
    for (uint32_t i : rp.order) {
        Maze::Point const p{static_cast<int>(i / static_cast<uint32_t>(cols)), static_cast<int>(i % static_cast<uint32_t>(cols))};
        shown[p.row][p.col] |= rp.paint[i];
        Sutil::flush_cursor_path_coordinate(shown, p);
        nap(step_us);
    }
}

inline void animate_hunt(Maze::Maze &maze, Speed::Speed speed, bool dark) {
    int const rows = maze.row_size(), cols = maze.col_size();
    int const us = Sutil::solver_speeds.at(static_cast<size_t>(speed));
    lab_point start{}, finish{};
    check(lab_place_hunt(rows, cols, maze.squares(), device_seed(), &start, &finish), "Bfs::animate_hunt placement");
    g_next_seed += g_seeded ? 1 : 0;
    Maze::Maze shown = maze; // what is on the screen: the placed maze, painted square by square below
    int32_t winner = -1;
    uint64_t path_len = 0;
    std::vector<lab_point> path(static_cast<size_t>(rows) * cols / 2 + 2);
    std::vector<uint32_t> levels;
    {
        Device_maze d(maze);
        check(lab_bfs_hunt(d.g, start, finish, &winner, path.data(), path.size(), &path_len), "lab_bfs_hunt");
        levels = levels_of(d.g, 0, maze, start);
        d.download();
    }
    if (g_quiet) return;
    Sutil::set_cursor_position({rows, 0});
    Sutil::print_overlap_key();
    if (dark) Sutil::deluminate_maze(shown);
    Sutil::flush_cursor_path_coordinate(shown, {finish.row, finish.col});
    nap(us);
    uint32_t const D = levels[static_cast<size_t>(finish.row) * cols + finish.col];
    // four threads share the start: every square closer than the finish carries all four colours while they search
    draw_replay(shown, make_replay(rows, cols, {levels.data(), levels.data(), levels.data(), levels.data()}, {D, D, D, D}), us / 4);
    if (winner >= 0) {
        uint16_t const colour = static_cast<uint16_t>(1u << (Sutil::thread_paint_shift + winner));
        for (uint64_t k = 0; k < path_len && k < path.size(); k++) {
            Maze::Point const p{path[k].row, path[k].col};
            shown[p.row][p.col] &= static_cast<uint16_t>(~Sutil::thread_paint_mask);
            shown[p.row][p.col] |= colour;
            Sutil::flush_cursor_path_coordinate(shown, p);
            nap(us);
        }
    }
    Sutil::set_cursor_position({rows + Sutil::overlap_key_and_message_height, 0});
    Sutil::print_hunt_solution_message(winner < 0 ? Sutil::no_winner : static_cast<uint16_t>(winner));
    std::cout << "\n";
}

inline void animate_gather(Maze::Maze &maze, Speed::Speed speed, bool dark) {
    int const rows = maze.row_size(), cols = maze.col_size();
    int const us = Sutil::solver_speeds.at(static_cast<size_t>(speed));
    lab_point start{}, finishes[4]{};
    check(lab_place_gather(rows, cols, maze.squares(), device_seed(), &start, finishes), "Bfs::animate_gather placement");
    g_next_seed += g_seeded ? 4 : 0;
    Maze::Maze shown = maze;
    int32_t claimed[4];
    uint64_t lens[4];
    std::vector<uint32_t> levels;
    {
        Device_maze d(maze);
        check(lab_bfs_gather(d.g, start, finishes, claimed, nullptr, 0, lens), "lab_bfs_gather");
        levels = levels_of(d.g, 0, maze, start);
        d.download();
    }
    if (g_quiet) return;
    Sutil::set_cursor_position({rows, 0});
    Sutil::print_overlap_key();
    if (dark) Sutil::deluminate_maze(shown);
    for (auto const &f : finishes) Sutil::flush_cursor_path_coordinate(shown, {f.row, f.col});
    // colour k searches until it reaches the finish it claims
    std::array<uint32_t, 4> limit{0, 0, 0, 0};
    for (int j = 0; j < 4; j++) {
        if (claimed[j] >= 0 && claimed[j] < 4) limit[static_cast<size_t>(claimed[j])] = levels[static_cast<size_t>(finishes[j].row) * cols + finishes[j].col];
    }
    draw_replay(shown, make_replay(rows, cols, {levels.data(), levels.data(), levels.data(), levels.data()}, limit), us / 4);
    // what the reference does after the join (bfs_threads.cc:356-364): the final grid's recoloured fronts and claims
    for (int r = 0; r < rows; r++) {
        for (int c = 0; c < cols; c++) {
            if (shown[r][c].load() != maze[r][c].load()) {
                shown[r][c].store(maze[r][c].load());
                Sutil::flush_cursor_path_coordinate(shown, {r, c});
            }
        }
    }
    Sutil::set_cursor_position({rows + Sutil::overlap_key_and_message_height, 0});
    Sutil::print_gather_solution_message();
    std::cout << "\n";
}

inline void animate_corners(Maze::Maze &maze, Speed::Speed speed, bool dark) {
    int const rows = maze.row_size(), cols = maze.col_size();
    int const us = Sutil::solver_speeds.at(static_cast<size_t>(speed));
    lab_point starts[4]{}, finish{};
    check(lab_place_corners(rows, cols, maze.squares(), device_seed(), starts, &finish), "Bfs::animate_corners placement");
    for (Maze::Point const &p : Sutil::all_dirs) { // the animated variant carves SEVEN neighbours of the centre (bfs_threads.cc:469-472; the static one four)
        maze[finish.row + p.row][finish.col + p.col] |= Maze::path_bit;
    }
    Maze::Maze shown = maze;
    int32_t winner = -1;
    uint64_t path_len = 0;
    std::vector<lab_point> path(static_cast<size_t>(rows) * cols / 2 + 2);
    std::array<std::vector<uint32_t>, 4> levels;
    {
        Device_maze d(maze);
        check(lab_bfs_corners(d.g, starts, finish, &winner, path.data(), path.size(), &path_len), "lab_bfs_corners");
        for (int k = 0; k < 4; k++) levels[static_cast<size_t>(k)] = levels_of(d.g, k, maze, starts[k]);
        d.download();
    }
    if (g_quiet) return;
    Sutil::set_cursor_position({rows, 0});
    Sutil::print_overlap_key();
    if (dark) Sutil::deluminate_maze(shown);
    Sutil::flush_cursor_path_coordinate(shown, {finish.row, finish.col});
    uint32_t D = 0xFFFFFFFFu; // the game ends when the first colour is home
    for (int k = 0; k < 4; k++) D = std::min(D, levels[static_cast<size_t>(k)][static_cast<size_t>(finish.row) * cols + finish.col]);
    draw_replay(shown, make_replay(rows, cols, {levels[0].data(), levels[1].data(), levels[2].data(), levels[3].data()}, {D, D, D, D}), us / 4);
    if (winner >= 0) { // (the animated corners repaints the winner's path: bfs_threads.cc:494-507)
        uint16_t const colour = static_cast<uint16_t>(1u << (Sutil::thread_paint_shift + winner));
        for (uint64_t k = 0; k < path_len && k < path.size(); k++) {
            Maze::Point const p{path[k].row, path[k].col};
            shown[p.row][p.col] &= static_cast<uint16_t>(~Sutil::thread_paint_mask);
            shown[p.row][p.col] |= colour;
            Sutil::flush_cursor_path_coordinate(shown, p);
            nap(us);
        }
    }
    Sutil::set_cursor_position({rows + Sutil::overlap_key_and_message_height, 0});
    Sutil::print_hunt_solution_message(winner < 0 ? Sutil::no_winner : static_cast<uint16_t>(winner));
    std::cout << "\n";
}
} // namespace Lab

namespace Bfs {
inline void animate_hunt(Maze::Maze &maze, Speed::Speed speed) { Lab::animate_hunt(maze, speed, false); }
inline void animate_gather(Maze::Maze &maze, Speed::Speed speed) { Lab::animate_gather(maze, speed, false); }
inline void animate_corners(Maze::Maze &maze, Speed::Speed speed) { Lab::animate_corners(maze, speed, false); }
} // namespace Bfs

// solvers/darkbfs_threads.cc:154-300: the same searches drawn on a dark screen (the static slots of run_maze's
// darkbfs-* entries are Bfs::hunt / gather / corners upstream: run_maze.cc:113-115)
namespace Dark_bfs {
inline void animate_hunt(Maze::Maze &maze, Speed::Speed speed) { Lab::animate_hunt(maze, speed, true); }
inline void animate_gather(Maze::Maze &maze, Speed::Speed speed) { Lab::animate_gather(maze, speed, true); }
inline void animate_corners(Maze::Maze &maze, Speed::Speed speed) { Lab::animate_corners(maze, speed, true); }
} // namespace Dark_bfs

// ------------------------------------------------------------------------------------- painters
namespace Lab {
inline constexpr std::array<int, 8> painter_speeds = {0, 10000, 5000, 2000, 1000, 500, 250, 50}; // painters/rgb.cc:24-25
// Rgb::print_rgb / print_wall for one square (painters/rgb.cc:51-66): a pixel of lab_distance_paint / lab_runs_paint or the wall glyph
inline void print_heat_square(Maze::Maze const &maze, std::vector<uint32_t> const &pixels, int r, int c) {
    std::cout << "\033[" << r + 1 << ";" << c + 1 << "f";
    uint32_t const px = pixels[static_cast<size_t>(r) * maze.col_size() + c];
    if (px >> 24) {
        std::cout << "\033[38;2;" << (px & 0xFF) << ";" << ((px >> 8) & 0xFF) << ";" << ((px >> 16) & 0xFF) << "m█\033[0m";
    } else {
        std::cout << Sutil::wall_styles.at(maze.style_index() * 16 + (maze[r][c].load() & Maze::wall_mask));
    }
}
// BFS and the painter's colour arithmetic on the GPU; `runs`: Runs::paint_runs instead of the distance map; speed: the
// animated variant (distance.cc:158-204, runs.cc:163 ff.: four threads repaint the map outwards from the start) replayed
// from the level field -- squares in level order, painter_speeds[speed] / 4 microseconds apart.
inline void paint_map(Maze::Maze &maze, bool runs, std::optional<Speed::Speed> speed) {
    int const rows = maze.row_size(), cols = maze.col_size();
    std::vector<uint32_t> pixels, levels;
    uint64_t max = 0, settled = 0;
    std::mt19937 rng(device_seed());
    std::uniform_int_distribution<int> uid(0, 2);
    int const choice = uid(rng); // (drawn whether or not anything is printed: the same RNG stream either way)
    // LAB_DEVICES=n (n > 1): the map is made on n GPUs of this process, the maze cut into row bands (lab_distance_map_sharded).
    // (The painter's colours need the map on one device: with rendering on, one device makes it.)
    if (char const *nd = std::getenv("LAB_DEVICES"); nd && std::atoi(nd) > 1 && !runs && g_quiet) {
        int const n = std::min(std::atoi(nd), 8);
        int devices[8] = {0, 1, 2, 3, 4, 5, 6, 7};
        lab_point const s = lab_distance_start(rows, cols);
        if (lab_distance_map_sharded(n, devices, rows, cols, maze.squares(), s.row, s.col, nullptr, &max, &settled, nullptr) != LAB_OK) {
            std::cerr << "lab_distance_map_sharded: " << lab_sharded_last_error() << "\n";
            std::abort();
        }
        return;
    }
    {
        Device_maze d(maze);
        lab_point const s = lab_distance_start(rows, cols);
        if (runs) {
            check(lab_runs_map(d.g, s.row, s.col, nullptr, &max, &settled), "lab_runs_map");
        } else {
            check(lab_distance_map(d.g, s.row, s.col, nullptr, &max, &settled), "lab_distance_map");
        }
        if (!g_quiet) {
            pixels.resize(static_cast<size_t>(rows) * cols);
            check(runs ? lab_runs_paint(d.g, choice, pixels.data()) : lab_distance_paint(d.g, choice, pixels.data()), "painter");
            if (speed) levels = levels_of(d.g, 0, maze, s);
        }
        d.download();
    }
    if (g_quiet) return;
    if (!speed) {
        for (int r = 0; r < rows; r++) {
            for (int c = 0; c < cols; c++) print_heat_square(maze, pixels, r, c);
        }
    } else {
        std::vector<std::pair<uint32_t, uint32_t>> keyed;
        for (size_t i = 0; i < levels.size(); i++) {
            if (levels[i] != 0xFFFFFFFFu) keyed.push_back({levels[i], static_cast<uint32_t>(i)});
        }
        std::stable_sort(keyed.begin(), keyed.end(), [](auto const &a, auto const &b) { return a.first < b.first; });
        int const us = painter_speeds.at(static_cast<size_t>(*speed));
        for (auto const &k : keyed) {
            print_heat_square(maze, pixels, static_cast<int>(k.second / static_cast<uint32_t>(cols)), static_cast<int>(k.second % static_cast<uint32_t>(cols)));
            std::cout << std::flush;
            nap(us / 4);
        }
        Sutil::set_cursor_position({rows, 0});
    }
    std::cout << "\n\n";
}
} // namespace Lab

namespace Distance {
// painters/distance.cc:129-156 / :158-204
inline void paint_distance_from_center(Maze::Maze &maze) { Lab::paint_map(maze, false, std::nullopt); }
inline void animate_distance_from_center(Maze::Maze &maze, Speed::Speed speed) { Lab::paint_map(maze, false, speed); }
} // namespace Distance

namespace Runs {
// painters/runs.cc:130-161 / :163 ff. -- on the GPU for mazes that are one tree (every builder without -m; lab_runs_map)
inline void paint_runs(Maze::Maze &maze) { Lab::paint_map(maze, true, std::nullopt); }
inline void animate_runs(Maze::Maze &maze, Speed::Speed speed) { Lab::paint_map(maze, true, speed); }
} // namespace Runs
