// The raw maze file (see mazefile.cc): "LABMAZE1", int32 rows, int32 cols, rows*cols uint16 LE row-major.
#pragma once
#include <cstdint>

namespace lab::host {

constexpr int MAZE_FILE_HEADER_BYTES = 16;

// 0 ok; -1 bad argument; -2 cannot open / write; -3 not a maze file (or truncated); -4 size differs from the file's
int maze_file_write(char const *path, int rows, int cols, uint16_t const *squares);
int maze_file_header(char const *path, int *rows, int *cols);
int maze_file_read(char const *path, int rows, int cols, uint16_t *squares);

} // namespace lab::host
