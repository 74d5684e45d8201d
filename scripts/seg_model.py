#!/usr/bin/env python
"""Executable model of the segmented-stream kernel (csr_seg_kernels.cu), lane by lane.

The kernel's index logic (runs of 4 non-zeros per lane, two steps per warp, head counts, lead /
tail / closed segments, cross-warp and cross-tile carries) is modelled here with plain Python
loops in the same order as the CUDA code, and checked against a direct row-by-row sum on random
CSR structures with empty rows, long rows and ragged ends.  Development aid (no GPU needed):

    python scripts/seg_model.py
"""
import numpy as np

RUN = 4          # non-zeros per lane and step (one 128-bit load)
STEPS = 2        # steps per warp and tile
WARPS = 8        # warps per worker
TILE = 32 * RUN * STEPS * WARPS  # 2048 non-zeros


def build_plan(rows, rp):
    nnz = int(rp[-1])
    head = np.zeros(nnz, bool)
    nonempty = rp[1:] > rp[:-1]
    rows_nz = np.nonzero(nonempty)[0]
    head[rp[:-1][nonempty]] = True
    num_tiles = (nnz + TILE - 1) // TILE
    starts = rp[:-1][nonempty]
    tile_head_base = np.searchsorted(starts, np.arange(num_tiles + 1) * TILE, side="left")
    return head, rows_nz, tile_head_base, num_tiles


def run_tile(t, nnz, prod, head, rows_nz, thb, y, tile_lead, tile_tail):
    base = t * TILE
    # pass 0: head counts per warp (known before the gathers return)
    warp_heads = []
    for w in range(WARPS):
        lo, hi = base + w * 256, min(base + (w + 1) * 256, nnz)
        warp_heads.append(int(head[lo:hi].sum()) if hi > lo else 0)
    warp_off = np.concatenate([[0], np.cumsum(warp_heads)])
    # pass 1: every warp on its own
    warp_info = []  # (has_head, lead, tail, lead_segment_local_index)
    for w in range(WARPS):
        open_sum, seen = 0.0, False
        heads_before = thb[t] + warp_off[w]   # global index of the next head
        warp_lead = 0.0
        for c in range(STEPS):
            lane_cnt, lane_lead, lane_tail, lane_closed = [], [], [], []
            for lane in range(32):
                j0 = base + w * 256 + c * 128 + lane * RUN
                acc, cnt, lead, closed = 0.0, 0, 0.0, []
                for k in range(RUN):
                    j = j0 + k
                    if j >= nnz:
                        break
                    if head[j]:
                        if cnt == 0:
                            lead = acc
                        else:
                            closed.append(acc)
                        cnt += 1
                        acc = 0.0
                    acc += prod[j]
                lane_cnt.append(cnt); lane_lead.append(lead); lane_tail.append(acc); lane_closed.append(closed)
            # exclusive add-scan of cnt, segmented exclusive scan of (cnt > 0, tail)
            hb = heads_before
            carry, carry_has_head = open_sum, seen
            for lane in range(32):
                cnt = lane_cnt[lane]
                if cnt > 0:
                    total = carry + lane_lead[lane]
                    if carry_has_head or False:
                        y[rows_nz[hb - 1]] = total          # segment closed by this lane's first head
                    else:
                        warp_lead = total                  # the segment open at the warp's start
                    for j, s in enumerate(lane_closed[lane]):
                        y[rows_nz[hb + j]] = s
                    hb += cnt
                    carry, carry_has_head = lane_tail[lane], True
                else:
                    carry += lane_tail[lane]
            open_sum, seen = carry, carry_has_head
            heads_before = hb
        warp_info.append((seen, warp_lead, open_sum))
    # pass 2: fold the warps (as the CUDA code does through shared memory)
    carry, has = 0.0, False
    lead_tile = 0.0
    g = thb[t]
    for w in range(WARPS):
        seen, wlead, wtail = warp_info[w]
        if seen:
            total = carry + wlead
            if has:
                y[rows_nz[thb[t] + warp_off[w] - 1]] = total
            else:
                lead_tile = total
            carry, has = wtail, True
        else:
            carry += wtail
    tile_lead[t] = lead_tile
    tile_tail[t] = carry
    return has


def spmv_model(rows, rp, prod):
    nnz = int(rp[-1])
    head, rows_nz, thb, num_tiles = build_plan(rows, rp)
    y = np.zeros(rows)
    tile_lead, tile_tail = np.zeros(num_tiles), np.zeros(num_tiles)
    for t in range(num_tiles):
        run_tile(t, nnz, prod, head, rows_nz, thb, y, tile_lead, tile_tail)
    # fix-up: the segment open at the end of every tile that has a head
    for t in range(num_tiles):
        if thb[t + 1] > thb[t]:
            total = tile_tail[t]
            u = t + 1
            while u < num_tiles and thb[u + 1] == thb[u]:
                total += tile_tail[u]
                u += 1
            if u < num_tiles:
                total += tile_lead[u]
            y[rows_nz[thb[t + 1] - 1]] = total
    return y


def main():
    rng = np.random.default_rng(0)
    cases = 0
    for trial in range(60):
        rows = int(rng.integers(1, 4000))
        kind = trial % 6
        if kind == 0:
            lens = rng.integers(0, 8, rows)
        elif kind == 1:
            lens = np.where(rng.random(rows) < 0.7, 0, rng.integers(1, 40, rows))
        elif kind == 2:
            lens = rng.integers(0, 3, rows); lens[rng.integers(0, rows, 3)] = rng.integers(3000, 9000, 3)
        elif kind == 3:
            lens = np.zeros(rows, int); lens[rows // 2] = int(rng.integers(1, 20000))
        elif kind == 4:
            lens = np.full(rows, 4)
        else:
            lens = rng.integers(0, 600, rows)
        rp = np.zeros(rows + 1, np.int64)
        rp[1:] = np.cumsum(lens)
        nnz = int(rp[-1])
        prod = rng.integers(-8, 9, nnz).astype(np.float64)  # exact in any order
        expect = np.array([prod[rp[i]:rp[i + 1]].sum() for i in range(rows)])
        got = spmv_model(rows, rp, prod) if nnz else np.zeros(rows)
        assert np.array_equal(got, expect), (trial, kind, rows, nnz, np.nonzero(got != expect)[0][:5])
        cases += 1
    print(f"segmented-stream model: {cases} structures OK")


if __name__ == "__main__":
    main()
