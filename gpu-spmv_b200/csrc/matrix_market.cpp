// matrix_market.cpp -- Matrix Market coordinate files <-> CSRMatrix (host arrays).
//
// The reference's requirements list real-matrix (SuiteSparse / Matrix Market) input
// (.kiro/specs/.../requirements.md:90) but the code has no loader: its only ways to fill
// a CSRMatrix are csr_from_dense (src/csr_matrix.cpp:50-95) and its own binary format
// (csr_deserialize, :231-279).  SURVEY 8f rank 4.
//
// Reads `%%MatrixMarket matrix coordinate {real|integer|pattern} {general|symmetric|
// skew-symmetric}`: 1-based indices, `%` comment lines, pattern entries get the value 1,
// symmetric files are expanded (the mirrored entry of an off-diagonal (i, j, v) is
// (j, i, v), (j, i, -v) for skew-symmetric).  The result follows csr_from_dense's
// conventions: host arrays re-allocated and owned, device arrays left alone, entries
// sorted by (row, column); duplicates are kept, in file order (a stable counting sort by
// row, then a stable sort of every row by column).  Errors: unreadable file -> FILE_IO;
// `array`, `complex`, `hermitian`, a malformed header / size line / entry, an index outside
// the matrix or fewer entries than announced -> INVALID_FORMAT (the matrix is unchanged).
#include "internal.hpp"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace spmv {
namespace b200 {
namespace {

std::string lower(std::string s) {
    for (char& ch : s) ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
    return s;
}

struct Entry {
    int row, col;
    float val;
};

}  // namespace

int csr_load_matrix_market(CSRMatrix* out, const char* filename) {
    if (!out || !filename) return static_cast<int>(SpMVError::INVALID_ARGUMENT);
    std::ifstream f(filename);
    if (!f) return static_cast<int>(SpMVError::FILE_IO);
    const int bad = static_cast<int>(SpMVError::INVALID_FORMAT);

    std::string line;
    if (!std::getline(f, line)) return bad;
    std::istringstream header(line);
    std::string banner, object, format, field, symmetry;
    header >> banner >> object >> format >> field >> symmetry;
    if (banner != "%%MatrixMarket" || lower(object) != "matrix" || lower(format) != "coordinate") return bad;
    field = lower(field);
    symmetry = lower(symmetry);
    const bool pattern = field == "pattern";
    if (!pattern && field != "real" && field != "integer") return bad;
    const bool general = symmetry == "general";
    const bool skew = symmetry == "skew-symmetric";
    if (!general && !skew && symmetry != "symmetric") return bad;

    do {  // comments and blank lines before the size line
        if (!std::getline(f, line)) return bad;
    } while (line.empty() || line[0] == '%' || line.find_first_not_of(" \t\r") == std::string::npos);
    long long rows = 0, cols = 0, announced = 0;
    {
        std::istringstream size(line);
        if (!(size >> rows >> cols >> announced)) return bad;
    }
    if (rows < 0 || cols < 0 || announced < 0 || rows > 0x7fffffffll || cols > 0x7fffffffll) return bad;

    std::vector<Entry> entries;
    // the announced count comes from the file: clamp the reservation, push_back grows the vector if it was honest
    entries.reserve(static_cast<size_t>(std::min<long long>(general ? announced : 2 * announced, 1ll << 26)));
    for (long long k = 0; k < announced; ++k) {
        long long i = 0, j = 0;
        double v = 1.0;
        if (!(f >> i >> j)) return bad;
        if (!pattern && !(f >> v)) return bad;
        if (i < 1 || i > rows || j < 1 || j > cols) return bad;
        entries.push_back({static_cast<int>(i - 1), static_cast<int>(j - 1), static_cast<float>(v)});
        if (!general && i != j) {
            if (j > rows || i > cols) return bad;  // a symmetric file must be square enough to mirror
            entries.push_back({static_cast<int>(j - 1), static_cast<int>(i - 1), static_cast<float>(skew ? -v : v)});
        }
    }
    if (entries.size() > 0x7fffffffull) return bad;

    // stable counting sort by row, then a stable sort by column inside every row
    const int n_rows = static_cast<int>(rows);
    const int nnz = static_cast<int>(entries.size());
    std::vector<int> row_ptrs(static_cast<size_t>(n_rows) + 1, 0);
    for (const Entry& e : entries) ++row_ptrs[static_cast<size_t>(e.row) + 1];
    for (int r = 0; r < n_rows; ++r) row_ptrs[r + 1] += row_ptrs[r];
    std::vector<Entry> sorted(entries.size());
    {
        std::vector<int> at(row_ptrs.begin(), row_ptrs.end() - 1);
        for (const Entry& e : entries) sorted[static_cast<size_t>(at[e.row]++)] = e;
    }
    for (int r = 0; r < n_rows; ++r)
        std::stable_sort(sorted.begin() + row_ptrs[r], sorted.begin() + row_ptrs[r + 1],
                         [](const Entry& a, const Entry& b) { return a.col < b.col; });

    if (out->owns_host_memory) {
        delete[] out->values;
        delete[] out->col_indices;
        delete[] out->row_ptrs;
    }
    out->num_rows = n_rows;
    out->num_cols = static_cast<int>(cols);
    out->nnz = nnz;
    out->values = nnz ? new float[nnz] : nullptr;
    out->col_indices = nnz ? new int[nnz] : nullptr;
    out->row_ptrs = new int[static_cast<size_t>(n_rows) + 1];
    out->owns_host_memory = true;
    std::copy(row_ptrs.begin(), row_ptrs.end(), out->row_ptrs);
    for (int k = 0; k < nnz; ++k) {
        out->values[k] = sorted[static_cast<size_t>(k)].val;
        out->col_indices[k] = sorted[static_cast<size_t>(k)].col;
    }
    return 0;
}

// `%%MatrixMarket matrix coordinate real general`, one entry per stored non-zero in CSR order,
// values with 9 significant digits (enough to round-trip fp32 exactly).
int csr_save_matrix_market(const CSRMatrix* m, const char* filename) {
    if (!m || !filename) return static_cast<int>(SpMVError::INVALID_ARGUMENT);
    if (!m->row_ptrs || (m->nnz > 0 && (!m->values || !m->col_indices))) return static_cast<int>(SpMVError::INVALID_FORMAT);
    std::FILE* f = std::fopen(filename, "w");
    if (!f) return static_cast<int>(SpMVError::FILE_IO);
    bool ok = std::fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n%d %d %d\n", m->num_rows, m->num_cols,
                           m->nnz) > 0;
    for (int r = 0; ok && r < m->num_rows; ++r)
        for (int k = m->row_ptrs[r]; ok && k < m->row_ptrs[r + 1]; ++k)
            ok = std::fprintf(f, "%d %d %.9g\n", r + 1, m->col_indices[k] + 1, static_cast<double>(m->values[k])) > 0;
    ok = (std::fclose(f) == 0) && ok;
    return ok ? 0 : static_cast<int>(SpMVError::FILE_IO);
}

}  // namespace b200
}  // namespace spmv
