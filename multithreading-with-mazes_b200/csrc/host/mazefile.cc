// mazefile.cc -- the raw maze file: what makes "identical serialized mazes" (BASELINE.json north_star) a file.
//
// The reference has no file I/O at all (`grep fstream|fopen` finds nothing); every maze lives
// and dies in one process.  Here a built maze can be written once and read by the oracle, the CLIs
// (LAB_LOAD / LAB_DUMP) and bench.py, so that all of them consume the same bytes and a 90 s host build
// (Kruskal 32767 x 32767) is paid once per box.
//
//   offset  0  "LABMAZE1"                       8 bytes
//   offset  8  rows, cols                       2 x int32, little endian
//   offset 16  rows * cols Squares              uint16 little endian, dense row-major (maze/maze.cc:127-144)
//
// The data starts 16-byte aligned, so the file can be memory-mapped and handed to the C ABI as it is.
#include "mazefile.hh"

#include <cerrno>
#include <cstdio>
#include <cstring>
#include <string>

#include <unistd.h>

namespace lab::host {

namespace {
constexpr char MAGIC[8] = {'L', 'A', 'B', 'M', 'A', 'Z', 'E', '1'};
struct Header {
    char magic[8];
    int32_t rows, cols;
};
static_assert(sizeof(Header) == MAZE_FILE_HEADER_BYTES, "maze file header is 16 bytes");
} // namespace

int maze_file_write(char const *path, int rows, int cols, uint16_t const *squares) {
    if (!path || !squares || rows < 1 || cols < 1) {
        return -1;
    }
    // written under a temporary name and renamed: a reader never sees half a maze (several ranks share one cache)
    std::string const tmp = std::string(path) + ".tmp." + std::to_string(static_cast<long>(getpid()));
    std::FILE *f = std::fopen(tmp.c_str(), "wb");
    if (!f) {
        return -2;
    }
    Header h{};
    std::memcpy(h.magic, MAGIC, sizeof MAGIC);
    h.rows = rows;
    h.cols = cols;
    size_t const n = static_cast<size_t>(rows) * static_cast<size_t>(cols);
    bool ok = std::fwrite(&h, sizeof h, 1, f) == 1 && std::fwrite(squares, sizeof(uint16_t), n, f) == n;
    ok = std::fclose(f) == 0 && ok;
    if (!ok || std::rename(tmp.c_str(), path) != 0) {
        std::remove(tmp.c_str());
        return -2;
    }
    return 0;
}

int maze_file_header(char const *path, int *rows, int *cols) {
    if (!path || !rows || !cols) {
        return -1;
    }
    std::FILE *f = std::fopen(path, "rb");
    if (!f) {
        return -2;
    }
    Header h{};
    bool const ok = std::fread(&h, sizeof h, 1, f) == 1;
    std::fclose(f);
    if (!ok || std::memcmp(h.magic, MAGIC, sizeof MAGIC) != 0 || h.rows < 1 || h.cols < 1) {
        return -3;
    }
    *rows = h.rows;
    *cols = h.cols;
    return 0;
}

int maze_file_read(char const *path, int rows, int cols, uint16_t *squares) {
    if (!squares) {
        return -1;
    }
    int fr = 0, fc = 0;
    int const rc = maze_file_header(path, &fr, &fc);
    if (rc) {
        return rc;
    }
    if (fr != rows || fc != cols) {
        return -4;
    }
    std::FILE *f = std::fopen(path, "rb");
    if (!f) {
        return -2;
    }
    size_t const n = static_cast<size_t>(rows) * static_cast<size_t>(cols);
    bool const ok = std::fseek(f, MAZE_FILE_HEADER_BYTES, SEEK_SET) == 0 && std::fread(squares, sizeof(uint16_t), n, f) == n;
    std::fclose(f);
    return ok ? 0 : -3;
}

} // namespace lab::host
