// Host-side start/finish placement (see points.hh).  These run on the caller's copy of the maze
// before a search is launched; they are the reference's rules, RNG stream included, so that a
// seed picks the same squares here, in the reference and in the oracle.
#include "points.hh"

#include <algorithm>
#include <random>
#include <vector>

namespace lab::host {

namespace {
constexpr uint16_t path_bit = 0x2000, start_bit = 0x4000, finish_bit = 0x8000;
// solve_utilities.cc:127-128 lists seven directions: S, SE, E, NE, N, NW, W (no SW).
constexpr Pt fan[7] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}};
constexpr Pt nesw[4] = {{-1, 0}, {0, 1}, {1, 0}, {0, -1}};

inline uint16_t at(int cols, uint16_t const *sq, Pt p) { return sq[static_cast<size_t>(p.row) * cols + p.col]; }
inline uint16_t &at(int cols, uint16_t *sq, Pt p) { return sq[static_cast<size_t>(p.row) * cols + p.col]; }
} // namespace

bool is_valid_start_or_finish(int rows, int cols, uint16_t const *sq, Pt p) {
    if (p.row <= 0 || p.row >= rows - 1 || p.col <= 0 || p.col >= cols - 1) {
        return false;
    }
    uint16_t const s = at(cols, sq, p);
    return (s & path_bit) && !(s & (finish_bit | start_bit));
}

Pt find_nearest_square(int rows, int cols, uint16_t const *sq, Pt choice) {
    for (Pt const &d : fan) {
        Pt const n = {choice.row + d.row, choice.col + d.col};
        if (is_valid_start_or_finish(rows, cols, sq, n)) {
            return n;
        }
    }
    for (int r = 1; r < rows - 1; r++) {
        for (int c = 1; c < cols - 1; c++) {
            if (is_valid_start_or_finish(rows, cols, sq, {r, c})) {
                return {r, c};
            }
        }
    }
    return {-1, -1};
}

Pt pick_random_point(int rows, int cols, uint16_t const *sq, uint32_t seed) {
    std::mt19937 gen(seed);
    std::uniform_int_distribution<int> row_random(1, rows - 2);
    std::uniform_int_distribution<int> col_random(1, cols - 2);
    int const r = row_random(gen); // row first, then column, as the braced initialiser upstream
    int const c = col_random(gen);
    Pt p = {r, c};
    if (!is_valid_start_or_finish(rows, cols, sq, p)) {
        p = find_nearest_square(rows, cols, sq, p);
    }
    return p;
}

void set_corner_starts(int rows, int cols, uint16_t const *sq, Pt out[4]) {
    Pt const want[4] = {{1, 1}, {1, cols - 2}, {rows - 2, 1}, {rows - 2, cols - 2}};
    for (int i = 0; i < 4; i++) {
        out[i] = (at(cols, sq, want[i]) & path_bit) ? want[i] : find_nearest_square(rows, cols, sq, want[i]);
    }
}

bool place_hunt(int rows, int cols, uint16_t *sq, uint32_t seed, Pt &start, Pt &finish) {
    start = pick_random_point(rows, cols, sq, seed);
    if (start.row < 0) {
        return false;
    }
    at(cols, sq, start) |= start_bit;
    finish = pick_random_point(rows, cols, sq, seed + 1);
    if (finish.row < 0) {
        return false;
    }
    at(cols, sq, finish) |= finish_bit;
    return true;
}

bool place_gather(int rows, int cols, uint16_t *sq, uint32_t seed, Pt &start, Pt finishes[4]) {
    start = pick_random_point(rows, cols, sq, seed);
    if (start.row < 0) {
        return false;
    }
    at(cols, sq, start) |= start_bit;
    for (int i = 0; i < 4; i++) {
        finishes[i] = pick_random_point(rows, cols, sq, seed + 1 + static_cast<uint32_t>(i));
        if (finishes[i].row < 0) {
            return false;
        }
        at(cols, sq, finishes[i]) |= finish_bit;
    }
    return true;
}

bool place_corners(int rows, int cols, uint16_t *sq, uint32_t seed, Pt starts[4], Pt &finish) {
    Pt s[4];
    set_corner_starts(rows, cols, sq, s);
    for (Pt const &p : s) {
        if (p.row < 0) {
            return false;
        }
        at(cols, sq, p) |= start_bit;
    }
    finish = {rows / 2, cols / 2};
    for (Pt const &d : nesw) {
        at(cols, sq, Pt{finish.row + d.row, finish.col + d.col}) |= path_bit;
    }
    at(cols, sq, finish) |= static_cast<uint16_t>(path_bit | finish_bit);
    std::vector<Pt> v(s, s + 4);
    std::shuffle(v.begin(), v.end(), std::mt19937(seed)); // colours start from shuffled corners
    std::copy(v.begin(), v.end(), starts);
    return true;
}

} // namespace lab::host
