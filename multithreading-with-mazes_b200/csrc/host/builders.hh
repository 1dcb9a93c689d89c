// Host-side maze builders (see builders.cc).  Not on the GPU hot path: they make its inputs.
#pragma once
#include <cstdint>

namespace lab::host {

// builder in {"arena","rdfs","grid","kruskal","wilson"}; mod in {nullptr,"none","cross","x"}.
// squares = rows*cols dense row-major Square bits (maze/maze.cc:127-144 layout), overwritten.
// `seed` stands in for the value the reference draws from std::random_device for the builder's RNG.
// Returns 0; -1 bad size/pointer (rows, cols must be odd and >= 7, run_maze.cc:280-299);
// -2 unknown builder; -3 unknown mod.
int build_maze(char const *builder, char const *mod, int rows, int cols, uint32_t seed,
               uint16_t *squares);

} // namespace lab::host
