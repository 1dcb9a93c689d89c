// run_maze -- the reference's solver CLI (run_maze/run_maze.cc) with the bfs-* (and static
// darkbfs-*) entries dispatched to the B200 library through include/labyrinth.hh.
// Flags are the reference's: -r -c -b -m -s -d -sa -ba -h.  The DFS / flood-fill families are
// outside this drop-in (inherently sequential stack walks): asking for them is
// an argument error here.
#include "cli_common.hh"

namespace {

void print_usage() {
    std::cout << "run_maze (B200 drop-in for the BFS solvers)\n"
                 "  -r rows   -c cols          odd sizes >= 7 (even values are bumped by one)\n"
                 "  -b rdfs|wilson|kruskal|grid|arena\n"
                 "  -m cross|x                 builder modification\n"
                 "  -s bfs-hunt|bfs-gather|bfs-corners|darkbfs-hunt|darkbfs-gather|darkbfs-corners\n"
                 "  -d sharp|round|doubles|bold|contrast|spikes\n"
                 "  -sa 0-7  -ba 0-7           animation speeds (frames replayed from the GPU search's level fields)\n"
                 "  -h                         this message\n"
                 "  environment: LAB_SEED=n reproducible run, LAB_QUIET=1 no rendering\n";
}

struct Maze_runner {
    Maze::Maze_args args;
    bool builder_animated = false, solver_animated = false;
    Speed::Speed builder_speed{}, solver_speed{};
    cli::Build_function builder{Recursive_backtracker::generate_maze, Recursive_backtracker::animate_maze};
    std::optional<cli::Build_function> modder;
    cli::Call_function solver{Bfs::hunt, Bfs::animate_hunt};
};

} // namespace

int main(int argc, char **argv) {
    auto const builders = cli::builder_table();
    auto const mods = cli::modification_table();
    auto const styles = cli::style_table();
    std::unordered_map<std::string, cli::Call_function> const solvers = {
        {"bfs-hunt", {Bfs::hunt, Bfs::animate_hunt}},
        {"bfs-gather", {Bfs::gather, Bfs::animate_gather}},
        {"bfs-corners", {Bfs::corners, Bfs::animate_corners}},
        // run_maze.cc:113-115: the static slots of the dark variants are the plain Bfs functions
        {"darkbfs-hunt", {Bfs::hunt, Dark_bfs::animate_hunt}},
        {"darkbfs-gather", {Bfs::gather, Dark_bfs::animate_gather}},
        {"darkbfs-corners", {Bfs::corners, Dark_bfs::animate_corners}},
    };
    std::unordered_set<std::string> const flags = {"-r", "-c", "-b", "-s", "-h", "-g", "-d", "-m", "-sa", "-ba"};
    Maze_runner runner;
    cli::apply_environment();
    bool process_current = false;
    std::string_view prev_flag;
    for (int i = 1; i < argc; i++) {
        std::string_view const arg = argv[i];
        if (!process_current) {
            if (!flags.contains(std::string(arg))) {
                std::cerr << "Invalid argument flag: " << arg << "\n";
                print_usage();
                return 1;
            }
            if (arg == "-h") {
                print_usage();
                return 0;
            }
            process_current = true;
            prev_flag = arg;
            continue;
        }
        process_current = false;
        cli::Flag_arg const pairs{prev_flag, arg};
        std::string const data(arg);
        if (prev_flag == "-r") {
            cli::set_odd(runner.args.odd_rows, pairs, print_usage);
        } else if (prev_flag == "-c") {
            cli::set_odd(runner.args.odd_cols, pairs, print_usage);
        } else if (prev_flag == "-b") {
            auto const f = builders.find(data);
            if (f == builders.end()) cli::print_invalid_arg(pairs, print_usage);
            runner.builder = f->second;
        } else if (prev_flag == "-m") {
            auto const f = mods.find(data);
            if (f == mods.end()) cli::print_invalid_arg(pairs, print_usage);
            runner.modder = f->second;
        } else if (prev_flag == "-s") {
            auto const f = solvers.find(data);
            if (f == solvers.end()) cli::print_invalid_arg(pairs, print_usage);
            runner.solver = f->second;
        } else if (prev_flag == "-d") {
            auto const f = styles.find(data);
            if (f == styles.end()) cli::print_invalid_arg(pairs, print_usage);
            runner.args.style = f->second;
        } else if (prev_flag == "-sa") {
            auto const sp = cli::speed_of(data);
            if (!sp) cli::print_invalid_arg(pairs, print_usage);
            runner.solver_speed = *sp;
            runner.solver_animated = true;
        } else if (prev_flag == "-ba") {
            auto const sp = cli::speed_of(data);
            if (!sp) cli::print_invalid_arg(pairs, print_usage);
            runner.builder_speed = *sp;
            runner.builder_animated = true;
        } else {
            cli::print_invalid_arg(pairs, print_usage); // "-g" is listed upstream but unhandled (run_maze.cc:73,277)
        }
    }

    int file_rows = 0, file_cols = 0;
    bool const from_file = cli::load_requested(file_rows, file_cols); // LAB_LOAD: a raw maze file instead of -b / -m
    if (from_file) {
        runner.args.odd_rows = static_cast<uint64_t>(file_rows);
        runner.args.odd_cols = static_cast<uint64_t>(file_cols);
    }
    Maze::Maze maze(runner.args);
    if (from_file) {
        cli::load_into(maze);
    } else if (runner.builder_animated) {
        std::get<1>(runner.builder)(maze, runner.builder_speed);
        if (runner.modder) std::get<1>(*runner.modder)(maze, runner.builder_speed);
    } else {
        std::get<0>(runner.builder)(maze);
        if (runner.modder) std::get<0>(*runner.modder)(maze);
    }
    cli::print_build(maze);
    if (!Lab::g_quiet) std::cout << "\033[1;1f"; // Printer::set_cursor_position({0,0}), run_maze.cc:202
    if (runner.solver_animated) {
        std::get<1>(runner.solver)(maze, runner.solver_speed);
    } else {
        std::get<0>(runner.solver)(maze);
    }
    cli::dump_if_requested(maze);
    if (std::getenv("LAB_TIMING")) {
        lab_timing const &t = Lab::last_timing();
        std::cerr << "[lab] h2d " << t.h2d_ms << " ms, setup " << t.setup_ms << ", pack " << t.pack_ms << ", search " << t.search_ms
                  << ", post " << t.post_ms << ", d2h " << t.d2h_ms << " ms\n";
    }
    return 0;
}
