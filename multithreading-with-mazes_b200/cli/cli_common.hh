// Shared flag handling of the two command-line programs: the reference's hand-rolled argv loop
// (run_maze/run_maze.cc:161-180, measure/measure.cc:140-158) with its tables, defaults and errors.
#pragma once
#include <cstdlib>
#include <functional>
#include <iostream>
#include <optional>
#include <string>
#include <string_view>
#include <tuple>
#include <unordered_map>
#include <unordered_set>

#include "../../include/labyrinth.hh"

namespace cli {

using Build_function = std::tuple<std::function<void(Maze::Maze &)>, std::function<void(Maze::Maze &, Speed::Speed)>>;
using Call_function = Build_function;

struct Flag_arg {
    std::string_view flag;
    std::string_view arg;
};

inline std::unordered_map<std::string, Build_function> builder_table() {
    return {
        {"rdfs", {Recursive_backtracker::generate_maze, Recursive_backtracker::animate_maze}},
        {"wilson", {Wilson_path_carver::generate_maze, Wilson_path_carver::animate_maze}},
        {"kruskal", {Kruskal::generate_maze, Kruskal::animate_maze}},
        {"grid", {Grid::generate_maze, Grid::animate_maze}},
        {"arena", {Arena::generate_maze, Arena::animate_maze}},
    };
}
// builders the reference has and this drop-in does not (not on the BFS path)
inline std::unordered_set<std::string> const other_builders = {"wilson-walls", "fractal", "eller", "prim"};

inline std::unordered_map<std::string, Build_function> modification_table() {
    return {{"cross", {Mods::add_cross, Mods::add_cross_animated}}, {"x", {Mods::add_x, Mods::add_x_animated}}};
}
inline std::unordered_map<std::string, Maze::Maze_style> style_table() {
    return {{"sharp", Maze::Maze_style::sharp},       {"round", Maze::Maze_style::round}, {"doubles", Maze::Maze_style::doubles},
            {"bold", Maze::Maze_style::bold},         {"contrast", Maze::Maze_style::contrast}, {"spikes", Maze::Maze_style::spikes}};
}
inline std::optional<Speed::Speed> speed_of(std::string const &s) {
    if (s.size() == 1 && s[0] >= '0' && s[0] <= '7') {
        return static_cast<Speed::Speed>(s[0] - '0');
    }
    return std::nullopt;
}

[[noreturn]] inline void print_invalid_arg(Flag_arg const &pairs, void (*usage)()) {
    std::cerr << "Flag was: " << pairs.flag << "\n" << "Invalid argument: " << pairs.arg << "\n";
    usage();
    std::exit(1);
}

// -r / -c: stoi, +1 if even, reject < 7 (run_maze.cc:280-299)
inline void set_odd(uint64_t &field, Flag_arg const &pairs, void (*usage)()) {
    int v = 0;
    try {
        v = std::stoi(std::string{pairs.arg});
    } catch (...) {
        print_invalid_arg(pairs, usage);
    }
    if (v % 2 == 0) {
        v++;
    }
    if (v < 7) {
        print_invalid_arg(pairs, usage);
    }
    field = static_cast<uint64_t>(v);
}

// LAB_SEED=n makes a run reproducible (the reference draws every seed from std::random_device);
// LAB_QUIET=1 skips the terminal rendering.
inline void apply_environment() {
    if (char const *s = std::getenv("LAB_SEED")) {
        Lab::seed(static_cast<uint32_t>(std::strtoul(s, nullptr, 0)));
    }
    if (char const *q = std::getenv("LAB_QUIET")) {
        Lab::quiet(q[0] != '0');
    }
}

// LAB_DUMP=path writes the final Squares as a raw maze file (include/labyrinth_b200.h: "LABMAZE1", int32 rows,
// int32 cols, rows*cols uint16 little endian) so a run can be compared with the oracle -- and fed back in:
inline void dump_if_requested(Maze::Maze const &maze) {
    char const *path = std::getenv("LAB_DUMP");
    if (!path) {
        return;
    }
    if (lab_maze_write(path, maze.row_size(), maze.col_size(), maze.squares()) != LAB_OK) {
        std::cerr << lab_last_error() << "\n";
        std::exit(1);
    }
}

// LAB_LOAD=path: the maze comes from a raw maze file instead of the builder (-b / -m are then ignored, -r / -c
// must match the file or be left at their defaults: the file's size wins).  This is how "identical serialized
// mazes" (BASELINE.json) reach the CLIs: a maze built once -- here, by the oracle, by bench.py -- is solved as it is.
inline bool load_requested(int &rows, int &cols) {
    char const *path = std::getenv("LAB_LOAD");
    if (!path) {
        return false;
    }
    if (lab_maze_header(path, &rows, &cols) != LAB_OK) {
        std::cerr << lab_last_error() << "\n";
        std::exit(1);
    }
    return true;
}
inline void load_into(Maze::Maze &maze) {
    if (lab_maze_read(std::getenv("LAB_LOAD"), maze.row_size(), maze.col_size(), maze.squares()) != LAB_OK) {
        std::cerr << lab_last_error() << "\n";
        std::exit(1);
    }
}

inline void print_build(Maze::Maze const &maze) {
    if (!Lab::g_quiet) {
        std::cout << "\033[2J\033[1;1H"; // Printer::clear_screen, printers/printers.cc:9-12
        Sutil::print_maze(maze);        // == Butil::clear_and_flush_grid for a finished build
    }
}

} // namespace cli
