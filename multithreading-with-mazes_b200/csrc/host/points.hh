// Start/finish placement rules of the reference's solvers, on a host copy of the squares.
#pragma once
#include <cstdint>

namespace lab::host {

struct Pt {
    int row, col;
};

// solvers/solve_utilities.cc:166-173
bool is_valid_start_or_finish(int rows, int cols, uint16_t const *sq, Pt p);
// solvers/solve_utilities.cc:229-252; {-1,-1} where the reference would abort
Pt find_nearest_square(int rows, int cols, uint16_t const *sq, Pt choice);
// solvers/solve_utilities.cc:275-285 with mt19937(seed)
Pt pick_random_point(int rows, int cols, uint16_t const *sq, uint32_t seed);
// solvers/solve_utilities.cc:254-273
void set_corner_starts(int rows, int cols, uint16_t const *sq, Pt out[4]);
// placement halves of Bfs::hunt/gather/corners (solvers/bfs_threads.cc:245-251,335-344,422-439);
// return false where the reference would abort
bool place_hunt(int rows, int cols, uint16_t *sq, uint32_t seed, Pt &start, Pt &finish);
bool place_gather(int rows, int cols, uint16_t *sq, uint32_t seed, Pt &start, Pt finishes[4]);
bool place_corners(int rows, int cols, uint16_t *sq, uint32_t seed, Pt starts[4], Pt &finish);

} // namespace lab::host
