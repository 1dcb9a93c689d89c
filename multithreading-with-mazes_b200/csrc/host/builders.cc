// Host-side maze builders: the step BEFORE the BFS hot path.
//
// These produce the packed Square grids the solvers/painters consume.  They are linear-time
// restatements on a flat uint16_t array of the reference generators that BASELINE.json's
// configs name, and -- given the same seed for the RNG the reference seeds from
// std::random_device -- they emit byte-identical grids (tests/test_oracle_vs_reference.py checks this
// against the real reference text on a size/seed ladder; tests/golden/ pins it on the GPU box).
//
//   arena    builders/arena.cc:17-26
//   grid     builders/grid.cc:31-47,73-104
//   rdfs     builders/recursive_backtracker.cc:27-75
//   kruskal  builders/kruskal.cc:23-86 (+ builders/disjoint_set.cc) -- the reference tags cells
//            through an unordered_map<Point,int> whose hash collides (maze/maze.cc:397-405,
//            17 s at 1023^2); here a cell's id is computed from its coordinates.
//   wilson   builders/wilson_path_carver.cc:185-296 -- the reference rescans the grid from row 1
//            for the next unbuilt cell on every walk (build_utilities.cc:604-615, quadratic);
//            builder bits are only ever added, so a monotone cursor returns the same cell.
//   mods     builders/mods.cc:20-27,41-56 (+ build_utilities.cc:479-561)
//
// RNG streams are libstdc++'s (std::mt19937, std::uniform_int_distribution, std::shuffle), as
// in the reference when compiled with this image's g++ 13.3.
#include "builders.hh"

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <random>
#include <string>
#include <vector>

namespace lab::host {

namespace {

// maze/maze.cc:170-210
constexpr uint16_t path_bit = 0x2000;
constexpr uint16_t start_bit = 0x4000;
constexpr uint16_t builder_bit = 0x1000;
constexpr uint16_t markers_mask = 0x00F0;
constexpr int marker_shift = 4;
constexpr uint16_t from_north = 0x0010, from_east = 0x0020, from_south = 0x0030, from_west = 0x0040;
constexpr uint16_t north_wall = 0x1, east_wall = 0x2, south_wall = 0x4, west_wall = 0x8;

struct P {
    int r, c;
    bool operator==(P const &o) const { return r == o.r && c == o.c; }
    bool operator!=(P const &o) const { return !(*this == o); }
};
constexpr P build_dirs[4] = {{-2, 0}, {0, 2}, {2, 0}, {0, -2}};           // maze.cc:219-224
constexpr P backtracking_marks[5] = {{0, 0}, {-2, 0}, {0, 2}, {2, 0}, {0, -2}}; // maze.cc:188-194

struct G {
    int rows, cols;
    uint16_t *sq;
    uint16_t &at(P p) const { return sq[static_cast<size_t>(p.r) * cols + p.c]; }
    uint16_t &at(int r, int c) const { return sq[static_cast<size_t>(r) * cols + c]; }
    bool interior(P p) const { return p.r > 0 && p.r < rows - 1 && p.c > 0 && p.c < cols - 1; }
};

// build_utilities.cc:261-278,313-320: every square becomes a wall that connects to each
// neighbour that exists.
void fill_with_walls(G const &g) {
    for (int r = 0; r < g.rows; r++) {
        uint16_t const ns = static_cast<uint16_t>((r > 0 ? north_wall : 0) | (r + 1 < g.rows ? south_wall : 0));
        uint16_t *row = g.sq + static_cast<size_t>(r) * g.cols;
        for (int c = 0; c < g.cols; c++) {
            uint16_t const ew = static_cast<uint16_t>((c > 0 ? west_wall : 0) | (c + 1 < g.cols ? east_wall : 0));
            row[c] = static_cast<uint16_t>((row[c] | ns | ew) & ~path_bit);
        }
    }
}

// build_utilities.cc:165-180: neighbours stop drawing a wall segment toward p; p becomes path.
inline void open_square(G const &g, P p, uint16_t extra) {
    if (p.r - 1 >= 0) g.at(p.r - 1, p.c) &= static_cast<uint16_t>(~south_wall);
    if (p.r + 1 < g.rows) g.at(p.r + 1, p.c) &= static_cast<uint16_t>(~north_wall);
    if (p.c - 1 >= 0) g.at(p.r, p.c - 1) &= static_cast<uint16_t>(~east_wall);
    if (p.c + 1 < g.cols) g.at(p.r, p.c + 1) &= static_cast<uint16_t>(~west_wall);
    g.at(p) |= static_cast<uint16_t>(path_bit | extra);
}
inline void build_path(G const &g, P p) { open_square(g, p, 0); }
inline void carve(G const &g, P p) { open_square(g, p, builder_bit); } // carve_path_walls :333-348

inline P between(P a, P b) { return {(a.r + b.r) / 2, (a.c + b.c) / 2}; }

// build_utilities.cc:435-456
inline void join_squares(G const &g, P cur, P next) {
    carve(g, cur);
    carve(g, between(cur, next));
    carve(g, next);
}

inline bool can_build_new_square(G const &g, P p) { // build_utilities.cc:84-89
    return g.interior(p) && !(g.at(p) & builder_bit);
}

// ------------------------------------------------------------------------------------- arena
void arena(G const &g) {
    fill_with_walls(g);
    for (int r = 1; r < g.rows - 1; r++) {
        for (int c = 1; c < g.cols - 1; c++) {
            build_path(g, {r, c});
        }
    }
}

// -------------------------------------------------------------------------------------- rdfs
void rdfs(G const &g, uint32_t seed) {
    fill_with_walls(g);
    std::mt19937 gen(seed);
    std::uniform_int_distribution<int> row_random(1, g.rows - 2);
    std::uniform_int_distribution<int> col_random(1, g.cols - 2);
    int const r0 = row_random(gen);
    int const c0 = col_random(gen);
    P const start = {2 * (r0 / 2) + 1, 2 * (c0 / 2) + 1};
    std::vector<int> order(4);
    std::iota(order.begin(), order.end(), 0);
    P cur = start;
    for (;;) {
        std::shuffle(order.begin(), order.end(), gen);
        bool advanced = false;
        for (int i : order) {
            P const next = {cur.r + build_dirs[i].r, cur.c + build_dirs[i].c};
            if (!can_build_new_square(g, next)) {
                continue;
            }
            // carve_path_markings, build_utilities.cc:383-405: next remembers where it came from
            uint16_t mark = 0;
            if (next.r < cur.r) mark = from_south;
            else if (next.r > cur.r) mark = from_north;
            else if (next.c < cur.c) mark = from_east;
            else mark = from_west;
            g.at(next) |= mark;
            carve(g, cur);
            carve(g, next);
            carve(g, between(cur, next));
            cur = next;
            advanced = true;
            break;
        }
        if (advanced) {
            continue;
        }
        if (cur == start) {
            break;
        }
        int const dir = (g.at(cur) & markers_mask) >> marker_shift;
        g.at(cur) &= static_cast<uint16_t>(~markers_mask);
        cur = {cur.r + backtracking_marks[dir].r, cur.c + backtracking_marks[dir].c};
    }
}

// -------------------------------------------------------------------------------------- grid
void grid(G const &g, uint32_t seed) {
    constexpr int run_limit = 4;
    fill_with_walls(g);
    std::mt19937 gen(seed);
    std::uniform_int_distribution<int> row_random(1, g.rows - 2);
    std::uniform_int_distribution<int> col_random(1, g.cols - 2);
    int const r0 = row_random(gen);
    int const c0 = col_random(gen);
    std::vector<P> dfs;
    dfs.push_back({2 * (r0 / 2) + 1, 2 * (c0 / 2) + 1});
    std::vector<int> order(4);
    std::iota(order.begin(), order.end(), 0);
    while (!dfs.empty()) {
        P const cur = dfs.back();
        std::shuffle(order.begin(), order.end(), gen);
        bool branched = false;
        for (int i : order) {
            P const d = build_dirs[i];
            if (!can_build_new_square(g, {cur.r + d.r, cur.c + d.c})) {
                continue;
            }
            // complete_run, grid.cc:31-47: keep going the same way, over old paths too
            P at = cur;
            P next = {cur.r + d.r, cur.c + d.c};
            for (int run = 0; g.interior(next) && run < run_limit; run++) {
                join_squares(g, at, next);
                at = next;
                dfs.push_back(next);
                next = {next.r + d.r, next.c + d.c};
            }
            branched = true;
            break;
        }
        if (!branched) {
            dfs.pop_back();
        }
    }
}

// ----------------------------------------------------------------------------------- kruskal
struct Dsu {
    std::vector<uint32_t> parent;
    std::vector<uint8_t> rank;
    explicit Dsu(size_t n) : parent(n), rank(n, 0) { std::iota(parent.begin(), parent.end(), 0u); }
    uint32_t find(uint32_t x) {
        uint32_t root = x;
        while (parent[root] != root) root = parent[root];
        while (parent[x] != root) {
            uint32_t const nx = parent[x];
            parent[x] = root;
            x = nx;
        }
        return root;
    }
    bool unite(uint32_t a, uint32_t b) {
        a = find(a);
        b = find(b);
        if (a == b) return false;
        if (rank[a] > rank[b]) parent[b] = a;
        else if (rank[a] < rank[b]) parent[a] = b;
        else {
            parent[a] = b;
            rank[b]++;
        }
        return true;
    }
};

void kruskal(G const &g, uint32_t seed) {
    fill_with_walls(g);
    // Walls are listed exactly as kruskal.cc:23-42 does: first the walls left/right of cells
    // (odd row, even col), row-major, then the walls above/below (even row, odd col).
    // A wall is stored as its index in that listing; shuffle only depends on the count.
    int const cell_rows = (g.rows - 1) / 2; // odd rows 1..rows-2
    int const cell_cols = (g.cols - 1) / 2;
    int const h_per_row = cell_cols - 1;    // even cols 2..cols-3
    int const v_rows = cell_rows - 1;       // even rows 2..rows-3
    uint64_t const n_h = static_cast<uint64_t>(cell_rows) * h_per_row;
    uint64_t const n_v = static_cast<uint64_t>(v_rows) * cell_cols;
    std::vector<uint32_t> walls(n_h + n_v);
    std::iota(walls.begin(), walls.end(), 0u);
    std::shuffle(walls.begin(), walls.end(), std::mt19937(seed));
    Dsu sets(static_cast<size_t>(cell_rows) * cell_cols);
    auto cell_id = [&](int r, int c) { return static_cast<uint32_t>((r / 2) * cell_cols + (c / 2)); };
    for (uint32_t w : walls) {
        P a, b;
        if (w < n_h) {
            int const r = 1 + 2 * static_cast<int>(w / h_per_row);
            int const c = 2 + 2 * static_cast<int>(w % h_per_row);
            a = {r, c - 1};
            b = {r, c + 1};
        } else {
            uint64_t const v = w - n_h;
            int const r = 2 + 2 * static_cast<int>(v / cell_cols);
            int const c = 1 + 2 * static_cast<int>(v % cell_cols);
            a = {r - 1, c};
            b = {r + 1, c};
        }
        if (sets.unite(cell_id(a.r, a.c), cell_id(b.r, b.c))) {
            join_squares(g, a, b);
        }
    }
}

// ------------------------------------------------------------------------------------ wilson
void wilson(G const &g, uint32_t seed) {
    fill_with_walls(g);
    std::mt19937 gen(seed);
    std::uniform_int_distribution<int> row_rand(2, g.rows - 2);
    std::uniform_int_distribution<int> col_rand(2, g.cols - 2);
    int const r0 = row_rand(gen);
    int const c0 = col_rand(gen);
    P const root = {2 * (r0 / 2) + 1, 2 * (c0 / 2) + 1};
    build_path(g, root);
    g.at(root) |= builder_bit;

    // first (odd,odd) square without builder_bit in row-major order; monotone because builder
    // bits are never removed (== choose_arbitrary_point(odd), build_utilities.cc:604-615)
    P cursor = {1, 1};
    auto next_unbuilt = [&]() -> P {
        while (cursor.r < g.rows - 1) {
            if (!(g.at(cursor) & builder_bit)) {
                return cursor;
            }
            cursor.c += 2;
            if (cursor.c >= g.cols - 1) {
                cursor.c = 1;
                cursor.r += 2;
            }
        }
        return {0, 0};
    };
    auto step_back = [&](P p) -> P { // follow p's backtrack marker one cell (two squares)
        int const m = (g.at(p) & markers_mask) >> marker_shift;
        return {p.r + backtracking_marks[m].r, p.c + backtracking_marks[m].c};
    };
    auto carve_link = [&](P a, P b) { // build_marks, wilson_path_carver.cc:38-59
        g.at(a) &= static_cast<uint16_t>(~start_bit);
        g.at(b) &= static_cast<uint16_t>(~start_bit);
        carve(g, a);
        carve(g, b);
        carve(g, between(a, b));
    };

    P prev = {0, 0};
    P walk = {1, 1};
    g.at(walk) &= static_cast<uint16_t>(~markers_mask);
    std::vector<int> order(4);
    std::iota(order.begin(), order.end(), 0);
    for (;;) {
        g.at(walk) |= start_bit; // start_bit = "on the current walk"
        std::shuffle(order.begin(), order.end(), gen);
        for (int i : order) {
            P const next = {walk.r + build_dirs[i].r, walk.c + build_dirs[i].c};
            if (!g.interior(next) || next == prev) {
                continue;
            }
            if (g.at(next) & builder_bit) {
                // the walk reached the maze: carve it in, following the markers home
                carve_link(walk, next);
                P cur = walk;
                while (g.at(cur) & markers_mask) {
                    P const back = step_back(cur);
                    carve_link(cur, back);
                    g.at(cur) &= static_cast<uint16_t>(~markers_mask);
                    cur = back;
                }
                g.at(cur) &= static_cast<uint16_t>(~(start_bit | markers_mask));
                carve(g, cur);
                walk = next_unbuilt();
                if (walk.r == 0) {
                    return;
                }
                g.at(walk) &= static_cast<uint16_t>(~markers_mask);
                prev = {0, 0};
            } else if (g.at(next) & start_bit) {
                // the walk bit its own tail: erase the loop back to `next`
                P cur = walk;
                while (cur != next) {
                    g.at(cur) &= static_cast<uint16_t>(~start_bit);
                    P const back = step_back(cur);
                    g.at(cur) &= static_cast<uint16_t>(~markers_mask);
                    cur = back;
                }
                walk = next;
                prev = step_back(walk);
            } else {
                // mark_origin, build_utilities.cc:216-228
                if (next.r > walk.r) g.at(next) |= from_north;
                else if (next.r < walk.r) g.at(next) |= from_south;
                else if (next.c < walk.c) g.at(next) |= from_east;
                else g.at(next) |= from_west;
                prev = walk;
                walk = next;
            }
            break;
        }
    }
}

// -------------------------------------------------------------------------------------- mods
void add_cross(G const &g) { // mods.cc:41-56
    int const mid_r = g.rows / 2;
    int const mid_c = g.cols / 2;
    auto widen = [&](int r, int c) {
        build_path(g, {r, c});
        if (c + 1 < g.cols - 2) {
            build_path(g, {r, c + 1});
        }
    };
    // the reference visits squares row-major; build_path is idempotent and order-free
    for (int r = 0; r < g.rows; r++) {
        if (r == mid_r) {
            for (int c = 2; c < g.cols - 2; c++) widen(r, c);
        } else if (r > 1 && r < g.rows - 2) {
            widen(r, mid_c);
        }
    }
}

void widen_x(G const &g, P p) { // build_utilities.cc:492-507
    build_path(g, p);
    if (p.c + 1 < g.cols - 2) build_path(g, {p.r, p.c + 1});
    if (p.c - 1 > 1) build_path(g, {p.r, p.c - 1});
    if (p.c + 2 < g.cols - 2) build_path(g, {p.r, p.c + 2});
    if (p.c - 2 > 1) build_path(g, {p.r, p.c - 2});
}

void add_x(G const &g) { // mods.cc:20-27 + build_utilities.cc:479-561 (float maths kept as is)
    float const row_size = static_cast<float>(g.rows) - 2.0F;
    float const col_size = static_cast<float>(g.cols) - 2.0F;
    float const pos_slope = (2.0F - row_size) / (2.0F - col_size);
    float const pos_b = 2.0F - (2.0F * pos_slope);
    float const neg_slope = (2.0F - row_size) / (col_size - 2.0F);
    float const neg_b = row_size - (2.0F * neg_slope);
    for (int r = 1; r < g.rows - 1; r++) {
        float const cur_row = static_cast<float>(r);
        int const on_pos = static_cast<int>((cur_row - pos_b) / pos_slope);
        int const on_neg = static_cast<int>((cur_row - neg_b) / neg_slope);
        for (int c = 1; c < g.cols - 1; c++) {
            if (c == on_pos && c < g.cols - 2 && c > 1) {
                widen_x(g, {r, c});
            }
            if (c == on_neg && c > 1 && c < g.cols - 2 && r < g.rows - 2) {
                widen_x(g, {r, c});
            }
        }
    }
}

} // namespace

int build_maze(char const *builder, char const *mod, int rows, int cols, uint32_t seed,
               uint16_t *squares) {
    if (!builder || !squares || rows < 7 || cols < 7 || !(rows & 1) || !(cols & 1)) {
        return -1; // run_maze.cc:280-299 only ever hands odd sizes >= 7 to a builder
    }
    G const g{rows, cols, squares};
    std::memset(squares, 0, static_cast<size_t>(rows) * cols * sizeof(uint16_t)); // maze.cc:349-354
    std::string const b(builder);
    if (b == "arena") arena(g);
    else if (b == "rdfs") rdfs(g, seed);
    else if (b == "grid") grid(g, seed);
    else if (b == "kruskal") kruskal(g, seed);
    else if (b == "wilson") wilson(g, seed);
    else return -2;
    std::string const m(mod ? mod : "none");
    if (m == "cross") add_cross(g);
    else if (m == "x") add_x(g);
    else if (m != "none" && !m.empty()) return -3;
    return 0;
}

} // namespace lab::host
