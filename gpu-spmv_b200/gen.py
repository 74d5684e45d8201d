"""Synthetic inputs for the BASELINE configurations.

The reference ships no generator beyond a dense mt19937 one (test_utils.h), so
the large inputs are defined here.  Everything is a pure function of
(seed, index) through a counter-based 32-bit integer hash evaluated with torch
int64 arithmetic, so the SAME code gives bit-identical matrices on the CPU
(small sizes, for the oracle) and on the GPU (BASELINE sizes, for bench.py).
Index arrays are int32, values float32, as in the reference's CSRMatrix.
"""
import torch

M32 = 0xFFFFFFFF


def _mix32(x):
    """lowbias32 finaliser on int64 tensors holding uint32 values."""
    x = ((x ^ (x >> 16)) * 0x7FEB352D) & M32
    x = ((x ^ (x >> 15)) * 0x846CA68B) & M32
    return x ^ (x >> 16)


def hash32(seed, idx, stream=0):
    """uint32 hash of (seed, stream, idx); idx is an int64 tensor (any magnitude < 2^62)."""
    key = (seed * 0x9E3779B1 + stream * 0x85EBCA77 + 0x165667B1) & M32
    hi = _mix32(((idx >> 32) & M32) ^ key)
    return _mix32((idx & M32) ^ hi ^ ((key * 0xC2B2AE3D) & M32))


def uniform_pm1(seed, idx, stream=0):
    """float32 U[-1, 1) with 24 random bits (exact in fp32)."""
    u = (hash32(seed, idx, stream) >> 8).to(torch.float32) * (1.0 / 16777216.0)
    return u * 2.0 - 1.0


def uniform_01_open_low(seed, idx, stream=0):
    """float32 U(0, 1] with 24 random bits."""
    return ((hash32(seed, idx, stream) >> 8) + 1).to(torch.float32) * (1.0 / 16777216.0)


def vector_pm1(n, seed, device):
    return uniform_pm1(seed, torch.arange(n, dtype=torch.int64, device=device), stream=7)


# ---------------------------------------------------------------- config 2 ----

def laplacian_2d_csr(grid, device):
    """5-point stencil on a grid x grid mesh, row-major node order: columns
    ascending {i-grid, i-1, i, i+1, i+grid} where inside the mesh; diagonal 4,
    off-diagonals -1.  Returns (row_ptrs i32 [n+1], col_indices i32, values f32)."""
    n = grid * grid
    i = torch.arange(n, dtype=torch.int64, device=device)
    gy, gx = i // grid, i % grid
    cand = torch.stack([i - grid, i - 1, i, i + 1, i + grid], dim=1)
    keep = torch.stack([gy > 0, gx > 0, torch.ones_like(gx, dtype=torch.bool), gx < grid - 1, gy < grid - 1], dim=1)
    vals = torch.tensor([-1.0, -1.0, 4.0, -1.0, -1.0], dtype=torch.float32, device=device).expand(n, 5)
    row_ptrs = torch.zeros(n + 1, dtype=torch.int64, device=device)
    row_ptrs[1:] = torch.cumsum(keep.sum(dim=1), dim=0)
    return row_ptrs.to(torch.int32), cand[keep].to(torch.int32), vals[keep].contiguous()


def laplacian_band_csr(grid_x, rows_lo, rows_hi, grid_y, device):
    """Rows [rows_lo, rows_hi) of the 5-point stencil on a grid_y x grid_x mesh
    (global column ids), for row-sharded runs."""
    i = torch.arange(rows_lo, rows_hi, dtype=torch.int64, device=device)
    gy, gx = i // grid_x, i % grid_x
    cand = torch.stack([i - grid_x, i - 1, i, i + 1, i + grid_x], dim=1)
    keep = torch.stack([gy > 0, gx > 0, torch.ones_like(gx, dtype=torch.bool), gx < grid_x - 1, gy < grid_y - 1], dim=1)
    n = rows_hi - rows_lo
    vals = torch.tensor([-1.0, -1.0, 4.0, -1.0, -1.0], dtype=torch.float32, device=device).expand(n, 5)
    row_ptrs = torch.zeros(n + 1, dtype=torch.int64, device=device)
    row_ptrs[1:] = torch.cumsum(keep.sum(dim=1), dim=0)
    return row_ptrs.to(torch.int32), cand[keep].to(torch.int32), vals[keep].contiguous()


# ---------------------------------------------------------------- config 3 ----

def short_rows_with_outliers_csr(rows, seed, device, outlier_rows=None, outlier_nnz=1_000_000):
    """rows x rows matrix: row length = hash(seed, i) mod 7 (0..6, mean 3) except
    `outlier_rows` (default: 0, rows/4, rows/2, 3*rows/4) which hold
    `outlier_nnz` entries.  Entry k of a row of length L sits in column
    k*(rows//L) + hash mod (rows//L): ascending and distinct by construction.
    Values are U(0, 1]."""
    if outlier_rows is None:
        outlier_rows = [0, rows // 4, rows // 2, (3 * rows) // 4]
    outlier_nnz = min(outlier_nnz, rows)
    i = torch.arange(rows, dtype=torch.int64, device=device)
    lens = hash32(seed, i, stream=1) % 7
    if len(outlier_rows):
        lens[torch.tensor(outlier_rows, dtype=torch.int64, device=device)] = outlier_nnz
    row_ptrs = torch.zeros(rows + 1, dtype=torch.int64, device=device)
    row_ptrs[1:] = torch.cumsum(lens, dim=0)
    nnz = int(row_ptrs[-1].item())
    j = torch.arange(nnz, dtype=torch.int64, device=device)
    row_of = torch.repeat_interleave(i, lens)
    k = j - row_ptrs[row_of]
    seg = rows // lens[row_of]
    cols = k * seg + hash32(seed, j, stream=2) % seg
    vals = uniform_01_open_low(seed, j, stream=3)
    return row_ptrs.to(torch.int32), cols.to(torch.int32), vals


# ------------------------------------------------------------ configs 4, 5 ----

RMAT_A, RMAT_B, RMAT_C = 0.57, 0.19, 0.19  # Graph500; D = 0.05


def rmat_edges(scale, edge_factor, seed, device, chunk=1 << 24):
    """Graph500 R-MAT edge list (src, dst), int64, duplicates and self loops kept,
    no vertex permutation.  Level l of edge e draws u = hash(seed, e, l):
    u < A -> (0,0); < A+B -> dst bit; < A+B+C -> src bit; else both."""
    n_edges = edge_factor << scale
    ta = int(RMAT_A * 4294967296.0)
    tb = int((RMAT_A + RMAT_B) * 4294967296.0)
    tc = int((RMAT_A + RMAT_B + RMAT_C) * 4294967296.0)
    src = torch.empty(n_edges, dtype=torch.int64, device=device)
    dst = torch.empty(n_edges, dtype=torch.int64, device=device)
    for lo in range(0, n_edges, chunk):
        hi = min(lo + chunk, n_edges)
        e = torch.arange(lo, hi, dtype=torch.int64, device=device)
        s = torch.zeros_like(e)
        d = torch.zeros_like(e)
        for level in range(scale):
            u = hash32(seed, e, stream=16 + level)
            s = (s << 1) | (u >= tb).to(torch.int64)
            d = (d << 1) | (((u >= ta) & (u < tb)) | (u >= tc)).to(torch.int64)
        src[lo:hi] = s
        dst[lo:hi] = d
    return src, dst


def rmat_pagerank_csr(scale, edge_factor, seed, device):
    """Column-normalised adjacency matrix of the R-MAT graph in the layout
    pagerank() expects (reference include/spmv/pagerank.h:28): edge src->dst is
    stored at row = dst, col = src with value 1/outdeg(src); entries sorted by
    (row, col).  Returns (n, row_ptrs i32, col_indices i32, values f32)."""
    n = 1 << scale
    src, dst = rmat_edges(scale, edge_factor, seed, device)
    key = (dst << 32) | src
    del dst
    key, _ = torch.sort(key)
    outdeg = torch.bincount(src, minlength=n)
    del src
    rows = key >> 32
    cols = key & M32
    del key
    row_ptrs = torch.zeros(n + 1, dtype=torch.int64, device=device)
    row_ptrs[1:] = torch.cumsum(torch.bincount(rows, minlength=n), dim=0)
    del rows
    vals = torch.ones((), dtype=torch.float32, device=device) / outdeg.to(torch.float32)[cols]
    return n, row_ptrs.to(torch.int32), cols.to(torch.int32), vals


def relabel(v, scale):
    """Bijective pseudo-random relabelling of vertex ids in [0, 2^scale) (xorshift, odd multiply,
    xorshift), the role of Graph500's vertex permutation: it removes R-MAT's id/degree correlation."""
    mask = (1 << scale) - 1
    v = v ^ (v >> (scale // 2))
    v = (v * 0x9E3779B1 + 0x7F4A7C15) & mask
    return v ^ (v >> (scale // 2 + 1))


def partition_bounds(row_ptrs, parts, row_weight=0):
    """Contiguous row split balancing work(row) = nnz(row) + row_weight: bounds[p] = first row whose
    prefix work row_ptrs[i] + i * row_weight reaches p * total / parts (same rule as
    spmv_b200_partition_rows_weighted).  row_ptrs: int64 tensor [rows + 1] on any device."""
    rows = row_ptrs.numel() - 1
    prefix = row_ptrs[:rows].to(torch.int64) + torch.arange(rows, dtype=torch.int64, device=row_ptrs.device) * row_weight
    total = int(row_ptrs[-1].item()) + rows * row_weight
    targets = torch.tensor([(total * p) // parts for p in range(1, parts)], dtype=torch.int64, device=row_ptrs.device)
    inner = torch.searchsorted(prefix, targets, right=False).tolist() if parts > 1 else []
    bounds = [0] + [int(b) for b in inner] + [rows]
    for p in range(1, parts + 1):
        bounds[p] = max(bounds[p], bounds[p - 1])
    return bounds


def rmat_pagerank_shard(scale, edge_factor, seed, rank, world, device, chunk=1 << 24, row_weight=1, relabelled=False):
    """Rank `rank`'s row shard of rmat_pagerank_csr(scale, edge_factor, seed) WITHOUT materialising the
    whole graph: pass 1 counts in/out degrees of every edge (bincount), pass 2 keeps the edges whose
    destination falls in this rank's row range (work(row) = nnz + row_weight balanced) and sorts only
    those.  Returns (n, bounds, row_ptrs i32 rebased to 0, col_indices i32 global, values f32, n_edges);
    concatenating the shards of all ranks gives exactly rmat_pagerank_csr's arrays."""
    n = 1 << scale
    n_edges = edge_factor << scale
    indeg = torch.zeros(n, dtype=torch.int64, device=device)
    outdeg = torch.zeros(n, dtype=torch.int64, device=device)
    ta = int(RMAT_A * 4294967296.0)
    tb = int((RMAT_A + RMAT_B) * 4294967296.0)
    tc = int((RMAT_A + RMAT_B + RMAT_C) * 4294967296.0)

    def edges(lo, hi):
        e = torch.arange(lo, hi, dtype=torch.int64, device=device)
        s = torch.zeros_like(e)
        d = torch.zeros_like(e)
        for level in range(scale):
            u = hash32(seed, e, stream=16 + level)
            s = (s << 1) | (u >= tb).to(torch.int64)
            d = (d << 1) | (((u >= ta) & (u < tb)) | (u >= tc)).to(torch.int64)
        if relabelled:
            s, d = relabel(s, scale), relabel(d, scale)
        return s, d

    for lo in range(0, n_edges, chunk):
        s, d = edges(lo, min(lo + chunk, n_edges))
        indeg += torch.bincount(d, minlength=n)
        outdeg += torch.bincount(s, minlength=n)
    row_ptrs = torch.zeros(n + 1, dtype=torch.int64, device=device)
    row_ptrs[1:] = torch.cumsum(indeg, dim=0)
    bounds = partition_bounds(row_ptrs, world, row_weight)
    r_lo, r_hi = bounds[rank], bounds[rank + 1]
    keys = []
    for lo in range(0, n_edges, chunk):
        s, d = edges(lo, min(lo + chunk, n_edges))
        m = (d >= r_lo) & (d < r_hi)
        keys.append(((d[m] << 32) | s[m]))
    key = torch.cat(keys) if keys else torch.zeros(0, dtype=torch.int64, device=device)
    del keys
    key, _ = torch.sort(key)
    cols = (key & 0xFFFFFFFF)
    vals = torch.ones((), dtype=torch.float32, device=device) / outdeg.to(torch.float32)[cols]
    rp_local = (row_ptrs[r_lo:r_hi + 1] - row_ptrs[r_lo]).to(torch.int32)
    return n, bounds, rp_local, cols.to(torch.int32), vals, n_edges


def random_csr(rows, cols, avg_nnz, seed, device, skew=0.0):
    """General random test matrix: row lengths around avg_nnz (optionally a
    heavy tail when skew > 0), columns from the hash, values U[-1, 1).
    Columns are sorted within a row; duplicates may occur (legal CSR)."""
    i = torch.arange(rows, dtype=torch.int64, device=device)
    base = hash32(seed, i, stream=1) % (2 * avg_nnz + 1)
    if skew > 0:
        heavy = (hash32(seed, i, stream=4) % 1000) < int(skew * 1000)
        base = torch.where(heavy, base * 64, base)
    lens = torch.clamp(base, max=cols)
    row_ptrs = torch.zeros(rows + 1, dtype=torch.int64, device=device)
    row_ptrs[1:] = torch.cumsum(lens, dim=0)
    nnz = int(row_ptrs[-1].item())
    j = torch.arange(nnz, dtype=torch.int64, device=device)
    row_of = torch.repeat_interleave(i, lens)
    c = hash32(seed, j, stream=2) % cols
    key, _ = torch.sort((row_of << 32) | c)
    vals = uniform_pm1(seed, j, stream=3)
    return row_ptrs.to(torch.int32), (key & M32).to(torch.int32), vals
