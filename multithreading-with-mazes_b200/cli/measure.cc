// measure -- the reference's painter CLI (measure/measure.cc) with "-p distance" and "-p runs" dispatched to the
// B200 library.  Flags are the reference's: -r -c -b -m -p -d -pa -ba -h.
#include "cli_common.hh"

namespace {

void print_usage() {
    std::cout << "measure (B200 drop-in for the distance painter)\n"
                 "  -r rows   -c cols          odd sizes >= 7 (even values are bumped by one)\n"
                 "  -b rdfs|wilson|kruskal|grid|arena\n"
                 "  -m cross|x                 builder modification\n"
                 "  -p distance|runs           painter (runs: perfect mazes, i.e. builders without -m)\n"
                 "  -d sharp|round|doubles|bold|contrast|spikes\n"
                 "  -pa 0-7  -ba 0-7           animation speeds (frames replayed from the GPU search's level field)\n"
                 "  -h                         this message\n"
                 "  environment: LAB_SEED=n reproducible run, LAB_QUIET=1 no rendering, LAB_TIMING=1 device timings\n";
}

} // namespace

int main(int argc, char **argv) {
    auto const builders = cli::builder_table();
    auto const mods = cli::modification_table();
    auto const styles = cli::style_table();
    std::unordered_map<std::string, cli::Call_function> const painters = {
        {"distance", {Distance::paint_distance_from_center, Distance::animate_distance_from_center}},
        {"runs", {Runs::paint_runs, Runs::animate_runs}},
    };
    std::unordered_set<std::string> const flags = {"-r", "-c", "-b", "-p", "-h", "-d", "-m", "-pa", "-ba"};
    Maze::Maze_args margs;
    bool builder_animated = false, painter_animated = false;
    Speed::Speed builder_speed{}, painter_speed{};
    cli::Build_function builder{Recursive_backtracker::generate_maze, Recursive_backtracker::animate_maze};
    std::optional<cli::Build_function> modder;
    cli::Call_function painter = painters.at("distance");
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
            cli::set_odd(margs.odd_rows, pairs, print_usage);
        } else if (prev_flag == "-c") {
            cli::set_odd(margs.odd_cols, pairs, print_usage);
        } else if (prev_flag == "-b") {
            auto const f = builders.find(data);
            if (f == builders.end()) cli::print_invalid_arg(pairs, print_usage);
            builder = f->second;
        } else if (prev_flag == "-m") {
            auto const f = mods.find(data);
            if (f == mods.end()) cli::print_invalid_arg(pairs, print_usage);
            modder = f->second;
        } else if (prev_flag == "-p") {
            auto const f = painters.find(data);
            if (f == painters.end()) cli::print_invalid_arg(pairs, print_usage);
            painter = f->second;
        } else if (prev_flag == "-d") {
            auto const f = styles.find(data);
            if (f == styles.end()) cli::print_invalid_arg(pairs, print_usage);
            margs.style = f->second;
        } else if (prev_flag == "-pa") {
            auto const sp = cli::speed_of(data);
            if (!sp) cli::print_invalid_arg(pairs, print_usage);
            painter_speed = *sp;
            painter_animated = true;
        } else if (prev_flag == "-ba") {
            auto const sp = cli::speed_of(data);
            if (!sp) cli::print_invalid_arg(pairs, print_usage);
            builder_speed = *sp;
            builder_animated = true;
        } else {
            cli::print_invalid_arg(pairs, print_usage);
        }
    }
    int file_rows = 0, file_cols = 0;
    bool const from_file = cli::load_requested(file_rows, file_cols); // LAB_LOAD: a raw maze file instead of -b / -m
    if (from_file) {
        margs.odd_rows = static_cast<uint64_t>(file_rows);
        margs.odd_cols = static_cast<uint64_t>(file_cols);
    }
    Maze::Maze maze(margs);
    if (from_file) {
        cli::load_into(maze);
    } else if (builder_animated) {
        std::get<1>(builder)(maze, builder_speed);
        if (modder) std::get<1>(*modder)(maze, builder_speed);
    } else {
        std::get<0>(builder)(maze);
        if (modder) std::get<0>(*modder)(maze);
    }
    cli::print_build(maze);
    if (!Lab::g_quiet) std::cout << "\033[1;1f";
    if (painter_animated) {
        std::get<1>(painter)(maze, painter_speed);
    } else {
        std::get<0>(painter)(maze);
    }
    cli::dump_if_requested(maze);
    if (std::getenv("LAB_TIMING")) {
        lab_timing const &t = Lab::last_timing();
        std::cerr << "[lab] h2d " << t.h2d_ms << " ms, setup " << t.setup_ms << ", pack " << t.pack_ms << ", search " << t.search_ms
                  << ", post " << t.post_ms << ", d2h " << t.d2h_ms << " ms\n";
    }
    return 0;
}
