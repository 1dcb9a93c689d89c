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

inline void animate_hunt(Maze::Maze &maze, Speed::Speed) { hunt(maze); }
inline void animate_gather(Maze::Maze &maze, Speed::Speed) { gather(maze); }
inline void animate_corners(Maze::Maze &maze, Speed::Speed) { corners(maze); }

} // namespace Bfs

// ------------------------------------------------------------------------------------- painters
namespace Distance {

// painters/distance.cc:129-156: the BFS and the painter's colour arithmetic (distance.cc:45-70) both run on
// the GPU (lab_distance_map, lab_distance_paint); what is left here is the reference's printing.
inline void paint_distance_from_center(Maze::Maze &maze) {
    int const rows = maze.row_size(), cols = maze.col_size();
    std::vector<uint32_t> pixels;
    uint64_t max = 0, settled = 0;
    std::mt19937 rng(Lab::device_seed());
    std::uniform_int_distribution<int> uid(0, 2);
    int const choice = uid(rng); // (drawn whether or not anything is printed: the same RNG stream either way)
    {
        Lab::Device_maze d(maze);
        lab_point const s = lab_distance_start(rows, cols);
        Lab::check(lab_distance_map(d.g, s.row, s.col, nullptr, &max, &settled), "lab_distance_map");
        if (!Lab::g_quiet) {
            pixels.resize(static_cast<size_t>(rows) * cols);
            Lab::check(lab_distance_paint(d.g, choice, pixels.data()), "lab_distance_paint");
        }
        d.download();
    }
    if (Lab::g_quiet) return;
    for (int r = 0; r < rows; r++) {
        for (int c = 0; c < cols; c++) {
            std::cout << "\033[" << r + 1 << ";" << c + 1 << "f";
            uint32_t const px = pixels[static_cast<size_t>(r) * cols + c];
            if (px >> 24) {
                std::cout << "\033[38;2;" << (px & 0xFF) << ";" << ((px >> 8) & 0xFF) << ";" << ((px >> 16) & 0xFF) << "m█\033[0m";
            } else {
                std::cout << Sutil::wall_styles.at(maze.style_index() * 16 + (maze[r][c].load() & Maze::wall_mask));
            }
        }
    }
    std::cout << "\n\n";
}
inline void animate_distance_from_center(Maze::Maze &maze, Speed::Speed) { paint_distance_from_center(maze); }

} // namespace Distance
