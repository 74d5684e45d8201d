// formats.cpp -- host side of the CSR / ELL containers: construction,
// dense <-> sparse conversion, element lookup, host <-> device transfer, raw
// (de)serialisation and row statistics.
//
// Pure integer / copy logic; results are bit-identical to the reference
// (src/csr_matrix.cpp, src/ell_matrix.cpp) -- tests/test_host_formats.py
// checks that against the oracle and against oracle/_ref.
#include "internal.hpp"

#include <algorithm>
#include <climits>
#include <cstring>
#include <fstream>

namespace spmv {
namespace {

constexpr int kOk = static_cast<int>(SpMVError::SUCCESS);
constexpr int kBadArg = static_cast<int>(SpMVError::INVALID_ARGUMENT);
constexpr int kFileIo = static_cast<int>(SpMVError::FILE_IO);

template <typename T>
T* alloc_zeroed(size_t n) { return n ? new T[n]() : nullptr; }
template <typename T>
T* alloc_raw(size_t n) { return n ? new T[n] : nullptr; }

void drop_host(CSRMatrix* m) {
    if (!m->owns_host_memory) return;
    delete[] m->values;
    delete[] m->col_indices;
    delete[] m->row_ptrs;
}
void drop_host(ELLMatrix* m) {
    if (!m->owns_host_memory) return;
    delete[] m->values;
    delete[] m->col_indices;
}

// (re)allocates host arrays for the given shape; contents uninitialised
void reshape_host(CSRMatrix* m, int rows, int cols, int nnz) {
    drop_host(m);
    m->num_rows = rows;
    m->num_cols = cols;
    m->nnz = nnz;
    m->values = alloc_raw<float>(nnz > 0 ? nnz : 0);
    m->col_indices = alloc_raw<int>(nnz > 0 ? nnz : 0);
    m->row_ptrs = new int[static_cast<size_t>(rows) + 1];
    m->owns_host_memory = true;
}

size_t ell_slots(const ELLMatrix* m) {
    return static_cast<size_t>(m->num_rows) * static_cast<size_t>(m->max_nnz_per_row);
}

// (re)allocates host arrays and writes the padding pattern (col -1, value 0)
void reshape_host_padded(ELLMatrix* m, int rows, int cols, int width) {
    drop_host(m);
    m->num_rows = rows;
    m->num_cols = cols;
    m->max_nnz_per_row = width;
    const size_t n = ell_slots(m);
    m->values = alloc_zeroed<float>(n);
    m->col_indices = alloc_raw<int>(n);
    if (n) std::fill_n(m->col_indices, n, -1);
    m->owns_host_memory = true;
}

template <typename T>
bool write_pod(std::ofstream& f, const T* p, size_t n) {
    f.write(reinterpret_cast<const char*>(p), static_cast<std::streamsize>(n * sizeof(T)));
    return static_cast<bool>(f);
}
template <typename T>
bool read_pod(std::ifstream& f, T* p, size_t n) {
    f.read(reinterpret_cast<char*>(p), static_cast<std::streamsize>(n * sizeof(T)));
    return static_cast<bool>(f);
}

}  // namespace

// ------------------------------------------------------------------ CSR ----

// reference: src/csr_matrix.cpp:10-32
CSRMatrix* csr_create(int rows, int cols, int nnz) {
    if (rows < 0 || cols < 0 || nnz < 0) return nullptr;
    CSRMatrix* m = new CSRMatrix();
    m->num_rows = rows;
    m->num_cols = cols;
    m->nnz = nnz;
    m->values = alloc_zeroed<float>(nnz);
    m->col_indices = alloc_zeroed<int>(nnz);
    m->row_ptrs = new int[static_cast<size_t>(rows) + 1]();
    m->owns_host_memory = true;
    return m;  // device pointers and owns_device_memory are zero from value-init
}

// reference: src/csr_matrix.cpp:34-48
void csr_destroy(CSRMatrix* m) {
    if (!m) return;
    drop_host(m);
    if (m->owns_device_memory) csr_free_gpu(m);
    delete m;
}

// reference: src/csr_matrix.cpp:50-95.  Keeps entries with v != 0.0f, columns
// ascending inside a row; the element count is an int product like the
// reference's (matrices above 2^31 dense elements are outside its domain).
int csr_from_dense(CSRMatrix* csr, const float* dense, int rows, int cols) {
    if (!csr || !dense || rows <= 0 || cols <= 0) return kBadArg;
    const int total = rows * cols;
    int nnz = 0;
    for (int i = 0; i < total; ++i) nnz += (dense[i] != 0.0f);

    reshape_host(csr, rows, cols, nnz);
    int out = 0;
    const float* row = dense;
    for (int r = 0; r < rows; ++r, row += cols) {
        csr->row_ptrs[r] = out;
        for (int c = 0; c < cols; ++c) {
            if (row[c] != 0.0f) {
                csr->values[out] = row[c];
                csr->col_indices[out] = c;
                ++out;
            }
        }
    }
    csr->row_ptrs[rows] = nnz;
    return kOk;
}

// reference: src/csr_matrix.cpp:97-114 (zero fill, then scatter; a repeated
// (row, col) keeps the last value)
int csr_to_dense(const CSRMatrix* csr, float* dense) {
    if (!csr || !dense) return kBadArg;
    std::memset(dense, 0, sizeof(float) * csr->num_rows * csr->num_cols);
    for (int r = 0; r < csr->num_rows; ++r) {
        float* out_row = dense + r * csr->num_cols;
        for (int p = csr->row_ptrs[r]; p < csr->row_ptrs[r + 1]; ++p)
            out_row[csr->col_indices[p]] = csr->values[p];
    }
    return kOk;
}

// reference: src/csr_matrix.cpp:116-136 (forward scan that relies on sorted
// columns to stop early)
float csr_get_element(const CSRMatrix* m, int row, int col) {
    if (!m || row < 0 || col < 0 || row >= m->num_rows || col >= m->num_cols) return 0.0f;
    for (int p = m->row_ptrs[row]; p < m->row_ptrs[row + 1]; ++p) {
        const int c = m->col_indices[p];
        if (c == col) return m->values[p];
        if (c > col) break;
    }
    return 0.0f;
}

// reference: src/csr_matrix.cpp:138-165
int csr_to_gpu(CSRMatrix* m) {
    if (!m) return kBadArg;
    csr_free_gpu(m);
    const size_t nnz = m->nnz > 0 ? static_cast<size_t>(m->nnz) : 0;
    const size_t nptr = static_cast<size_t>(m->num_rows) + 1;
    if (nnz) {
        CUDA_CHECK(cudaMalloc(&m->d_values, nnz * sizeof(float)));
        CUDA_CHECK(cudaMalloc(&m->d_col_indices, nnz * sizeof(int)));
    }
    CUDA_CHECK(cudaMalloc(&m->d_row_ptrs, nptr * sizeof(int)));
    if (nnz) {
        CUDA_CHECK(cudaMemcpy(m->d_values, m->values, nnz * sizeof(float), cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(m->d_col_indices, m->col_indices, nnz * sizeof(int), cudaMemcpyHostToDevice));
    }
    CUDA_CHECK(cudaMemcpy(m->d_row_ptrs, m->row_ptrs, nptr * sizeof(int), cudaMemcpyHostToDevice));
    m->owns_device_memory = true;
    b200::note_device_csr(m);  // spmv_csr(MERGE_PATH) may attach a column plan to this upload
    {   // the row-owner launchers steer by the longest row: known here for free, so the first
        // stream-ordered launch on this upload does not have to measure it (allocate + synchronise)
        int longest = 0;
        for (int r = 0; r < m->num_rows; ++r) longest = std::max(longest, m->row_ptrs[r + 1] - m->row_ptrs[r]);
        b200::seed_longest_row(m->d_row_ptrs, m->num_rows, m->nnz, longest);
    }
    return kOk;
}

// reference: src/csr_matrix.cpp:167-183
int csr_from_gpu(CSRMatrix* m) {
    if (!m || !m->d_row_ptrs) return kBadArg;
    if (m->nnz > 0 && m->d_values && m->d_col_indices) {
        const size_t nnz = static_cast<size_t>(m->nnz);
        CUDA_CHECK(cudaMemcpy(m->values, m->d_values, nnz * sizeof(float), cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(m->col_indices, m->d_col_indices, nnz * sizeof(int), cudaMemcpyDeviceToHost));
    }
    CUDA_CHECK(cudaMemcpy(m->row_ptrs, m->d_row_ptrs,
                          (static_cast<size_t>(m->num_rows) + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    return kOk;
}

// reference: src/csr_matrix.cpp:185-200
void csr_free_gpu(CSRMatrix* m) {
    if (!m) return;
    b200::forget_device_csr(m->d_col_indices);
    b200::forget_longest_row(m->d_row_ptrs, m->num_rows, m->nnz);
    if (m->d_values) cudaFree(m->d_values);
    if (m->d_col_indices) cudaFree(m->d_col_indices);
    if (m->d_row_ptrs) cudaFree(m->d_row_ptrs);
    m->d_values = nullptr;
    m->d_col_indices = nullptr;
    m->d_row_ptrs = nullptr;
    m->owns_device_memory = false;
}

// File layout (native endian, no magic): i32 rows, cols, nnz; f32 values[nnz];
// i32 col_indices[nnz]; i32 row_ptrs[rows+1].  reference: src/csr_matrix.cpp:202-229
int csr_serialize(const CSRMatrix* m, const char* filename) {
    if (!m || !filename) return kBadArg;
    std::ofstream f(filename, std::ios::binary);
    if (!f) return kFileIo;
    const int header[3] = {m->num_rows, m->num_cols, m->nnz};
    write_pod(f, header, 3);
    if (m->nnz > 0) {
        write_pod(f, m->values, m->nnz);
        write_pod(f, m->col_indices, m->nnz);
    }
    write_pod(f, m->row_ptrs, static_cast<size_t>(m->num_rows) + 1);
    return f ? kOk : kFileIo;
}

// reference: src/csr_matrix.cpp:231-279.  A short file leaves the struct
// re-shaped and reports FILE_IO, as the reference does.
// A corrupt header must not turn into a multi-gigabyte new[] (the plain C++ API has no guard against
// bad_alloc): a payload the file cannot hold AND larger than 1 GiB is refused before anything is
// allocated.  Smaller short files keep the reference's behaviour -- the struct is reshaped to the
// header, then the read fails with FILE_IO (reference src/csr_matrix.cpp:247-276).
static bool implausible_payload(std::ifstream& f, unsigned long long payload_bytes) {
    if (payload_bytes <= (1ull << 30)) return false;
    const std::streampos here = f.tellg();
    f.seekg(0, std::ios::end);
    const std::streampos end = f.tellg();
    f.seekg(here);
    return !f || end < here || static_cast<unsigned long long>(end - here) < payload_bytes;
}

int csr_deserialize(CSRMatrix* m, const char* filename) {
    if (!m || !filename) return kBadArg;
    std::ifstream f(filename, std::ios::binary);
    if (!f) return kFileIo;
    int header[3] = {0, 0, 0};
    if (!read_pod(f, header, 3) || header[0] < 0 || header[1] < 0 || header[2] < 0) return kFileIo;
    // payload announced by the header: values + col_indices + row_ptrs
    if (implausible_payload(f, 8ull * static_cast<unsigned long long>(header[2]) + 4ull * (static_cast<unsigned long long>(header[0]) + 1)))
        return kFileIo;
    reshape_host(m, header[0], header[1], header[2]);
    if (m->nnz > 0) {
        read_pod(f, m->values, m->nnz);
        read_pod(f, m->col_indices, m->nnz);
    }
    read_pod(f, m->row_ptrs, static_cast<size_t>(m->num_rows) + 1);
    return f ? kOk : kFileIo;
}

// reference: src/csr_matrix.cpp:281-300.  The two fp32 expressions are part of
// the selector contract and must not be re-associated:
//   avg  = float(nnz) / rows          (rows converted to float by the division)
//   skew = float(max) / (min + 1)
CSRStats csr_compute_stats(const CSRMatrix* m) {
    CSRStats s = {0.0f, 0, 0, 0.0f};
    if (!m || m->num_rows == 0) return s;
    int longest = 0, shortest = INT_MAX;
    for (int r = 0; r < m->num_rows; ++r) {
        const int len = m->row_ptrs[r + 1] - m->row_ptrs[r];
        longest = std::max(longest, len);
        shortest = std::min(shortest, len);
    }
    s.avg_nnz_per_row = static_cast<float>(m->nnz) / m->num_rows;
    s.max_nnz_per_row = longest;
    s.min_nnz_per_row = shortest;
    s.skewness = static_cast<float>(longest) / (shortest + 1);
    return s;
}

// ------------------------------------------------------------------ ELL ----

// reference: src/ell_matrix.cpp:8-36
ELLMatrix* ell_create(int rows, int cols, int max_nnz_per_row) {
    if (rows < 0 || cols < 0 || max_nnz_per_row < 0) return nullptr;
    ELLMatrix* m = new ELLMatrix();
    m->owns_host_memory = false;  // nothing to drop yet
    reshape_host_padded(m, rows, cols, max_nnz_per_row);
    return m;
}

// reference: src/ell_matrix.cpp:38-51
void ell_destroy(ELLMatrix* m) {
    if (!m) return;
    drop_host(m);
    if (m->owns_device_memory) ell_free_gpu(m);
    delete m;
}

// reference: src/ell_matrix.cpp:53-109
int ell_from_dense(ELLMatrix* ell, const float* dense, int rows, int cols) {
    if (!ell || !dense || rows <= 0 || cols <= 0) return kBadArg;
    int width = 0;
    for (int r = 0; r < rows; ++r) {
        const float* row = dense + r * cols;
        int len = 0;
        for (int c = 0; c < cols; ++c) len += (row[c] != 0.0f);
        width = std::max(width, len);
    }
    reshape_host_padded(ell, rows, cols, width);
    for (int r = 0; r < rows; ++r) {
        const float* row = dense + r * cols;
        int k = 0;
        for (int c = 0; c < cols; ++c) {
            if (row[c] == 0.0f) continue;
            const int slot = ell_index(r, k++, rows);
            ell->values[slot] = row[c];
            ell->col_indices[slot] = c;
        }
    }
    return kOk;
}

// reference: src/ell_matrix.cpp:111-159
int ell_from_csr(ELLMatrix* ell, const CSRMatrix* csr) {
    if (!ell || !csr) return kBadArg;
    const int rows = csr->num_rows;
    int width = 0;
    for (int r = 0; r < rows; ++r) width = std::max(width, csr->row_ptrs[r + 1] - csr->row_ptrs[r]);
    reshape_host_padded(ell, rows, csr->num_cols, width);
    for (int r = 0; r < rows; ++r) {
        int k = 0;
        for (int p = csr->row_ptrs[r]; p < csr->row_ptrs[r + 1]; ++p, ++k) {
            const int slot = ell_index(r, k, rows);
            ell->values[slot] = csr->values[p];
            ell->col_indices[slot] = csr->col_indices[p];
        }
    }
    return kOk;
}

// reference: src/ell_matrix.cpp:161-182
int ell_to_dense(const ELLMatrix* ell, float* dense) {
    if (!ell || !dense) return kBadArg;
    std::memset(dense, 0, sizeof(float) * ell->num_rows * ell->num_cols);
    for (int r = 0; r < ell->num_rows; ++r)
        for (int k = 0; k < ell->max_nnz_per_row; ++k) {
            const int slot = ell_index(r, k, ell->num_rows);
            const int c = ell->col_indices[slot];
            if (c >= 0) dense[r * ell->num_cols + c] = ell->values[slot];
        }
    return kOk;
}

// reference: src/ell_matrix.cpp:184-200 (gives up at the first padding slot)
float ell_get_element(const ELLMatrix* m, int row, int col) {
    if (!m || row < 0 || col < 0 || row >= m->num_rows || col >= m->num_cols) return 0.0f;
    for (int k = 0; k < m->max_nnz_per_row; ++k) {
        const int slot = ell_index(row, k, m->num_rows);
        const int c = m->col_indices[slot];
        if (c == col) return m->values[slot];
        if (c < 0) break;
    }
    return 0.0f;
}

// reference: src/ell_matrix.cpp:202-222
int ell_to_gpu(ELLMatrix* m) {
    if (!m) return kBadArg;
    ell_free_gpu(m);
    const size_t n = ell_slots(m);
    if (n) {
        CUDA_CHECK(cudaMalloc(&m->d_values, n * sizeof(float)));
        CUDA_CHECK(cudaMalloc(&m->d_col_indices, n * sizeof(int)));
        CUDA_CHECK(cudaMemcpy(m->d_values, m->values, n * sizeof(float), cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(m->d_col_indices, m->col_indices, n * sizeof(int), cudaMemcpyHostToDevice));
    }
    m->owns_device_memory = true;
    return kOk;
}

// reference: src/ell_matrix.cpp:224-238
int ell_from_gpu(ELLMatrix* m) {
    if (!m) return kBadArg;
    const size_t n = ell_slots(m);
    if (n && m->d_values && m->d_col_indices) {
        CUDA_CHECK(cudaMemcpy(m->values, m->d_values, n * sizeof(float), cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemcpy(m->col_indices, m->d_col_indices, n * sizeof(int), cudaMemcpyDeviceToHost));
    }
    return kOk;
}

// reference: src/ell_matrix.cpp:240-252
void ell_free_gpu(ELLMatrix* m) {
    if (!m) return;
    if (m->d_values) cudaFree(m->d_values);
    if (m->d_col_indices) cudaFree(m->d_col_indices);
    m->d_values = nullptr;
    m->d_col_indices = nullptr;
    m->owns_device_memory = false;
}

// File layout: i32 rows, cols, width; f32 values[rows*width]; i32 cols[rows*width]
// (column-major).  reference: src/ell_matrix.cpp:254-281
int ell_serialize(const ELLMatrix* m, const char* filename) {
    if (!m || !filename) return kBadArg;
    std::ofstream f(filename, std::ios::binary);
    if (!f) return kFileIo;
    const int header[3] = {m->num_rows, m->num_cols, m->max_nnz_per_row};
    write_pod(f, header, 3);
    const size_t n = ell_slots(m);
    if (n) {
        write_pod(f, m->values, n);
        write_pod(f, m->col_indices, n);
    }
    return f ? kOk : kFileIo;
}

// reference: src/ell_matrix.cpp:283-324 (arrays are NOT pre-padded on load)
int ell_deserialize(ELLMatrix* m, const char* filename) {
    if (!m || !filename) return kBadArg;
    std::ifstream f(filename, std::ios::binary);
    if (!f) return kFileIo;
    int header[3] = {0, 0, 0};
    if (!read_pod(f, header, 3) || header[0] < 0 || header[1] < 0 || header[2] < 0) return kFileIo;
    if (implausible_payload(f, 8ull * static_cast<unsigned long long>(header[0]) * static_cast<unsigned long long>(header[2])))
        return kFileIo;
    drop_host(m);
    m->num_rows = header[0];
    m->num_cols = header[1];
    m->max_nnz_per_row = header[2];
    const size_t n = ell_slots(m);
    m->values = alloc_raw<float>(n);
    m->col_indices = alloc_raw<int>(n);
    m->owns_host_memory = true;
    if (n) {
        read_pod(f, m->values, n);
        read_pod(f, m->col_indices, n);
    }
    return f ? kOk : kFileIo;
}

}  // namespace spmv
