"""The combinatorics of the Newton-3 pair kernel (csrc/mmm_pair_n3.cu), replayed in numpy: the
lane <-> bead mapping covers every pair of a 64 x 32 tile exactly once, the select-free
reduce-scatter leaves lane l holding the total for j-bead l, the work items cover every unordered
stage pair once, and round-robin sharding over ranks partitions them."""
import numpy as np


def lanes():
    for lane in range(32):
        yield lane, lane >> 2, lane & 3  # a = i-group, b = j-subset


def test_every_pair_of_a_tile_is_computed_by_exactly_one_lane():
    seen = np.zeros((64, 32), dtype=int)
    for lane, a, b in lanes():
        for jj in range(8):
            jl = ((jj ^ a) << 2) | b  # j-bead of register jj
            for ii in range(8):
                seen[a * 8 + ii, jl] += 1
    assert (seen == 1).all()


def test_reduce_scatter_delivers_bead_l_to_lane_l():
    rng = np.random.default_rng(0)
    # part[lane][jj] = this lane's partial force on the j-bead its register jj holds
    part = rng.normal(size=(32, 8))
    bead_of = np.array([[((jj ^ (lane >> 2)) << 2) | (lane & 3) for jj in range(8)] for lane in range(32)])
    want = np.zeros(32)
    for lane in range(32):
        for jj in range(8):
            want[bead_of[lane, jj]] += part[lane, jj]

    v = part.copy()

    def shfl_xor(col, m):  # value of register `col` on lane ^ m
        return np.array([v[lane ^ m, col] for lane in range(32)])

    # the kernel's order: groups (0,1),(2,3) combine over lane bit 3, (4,5),(6,7) likewise, then bit 4, then bit 2
    p = np.stack([v[:, r] + shfl_xor(r + 2, 8) for r in (0, 1)], axis=1)
    q = np.stack([v[:, 4 + r] + shfl_xor(6 + r, 8) for r in (0, 1)], axis=1)
    v = np.concatenate([p, q], axis=1)  # columns 0,1 = p; 2,3 = q
    s = np.stack([v[:, r] + shfl_xor(2 + r, 16) for r in (0, 1)], axis=1)
    v = s
    out = v[:, 0] + shfl_xor(1, 4)
    assert np.allclose(out, want)  # lane l holds j-bead l


def build_items(npad, cj):
    nib, njs = npad // 512, npad // 256
    items = []
    for i in range(nib):
        js = 2 * i
        while js < njs:
            cnt = min(cj, njs - js)
            items.append((i, js, cnt))
            js += cnt
    return items


def test_items_cover_every_unordered_stage_pair_once_and_shard_cleanly():
    npad, cj = 512 * 7, 3
    items = build_items(npad, cj)
    njs = npad // 256
    cover = np.zeros((npad // 512, njs), dtype=int)
    for i, js, cnt in items:
        cover[i, js:js + cnt] += 1
    for i in range(npad // 512):
        assert (cover[i, :2 * i] == 0).all() and (cover[i, 2 * i:] == 1).all()
    # the two diagonal stages of block i are evaluated as ordered pairs with half weight; every
    # other (block, stage) pair is an unordered pair seen once: total weight = N (N - 1) / 2 pairs
    pairs = 0.0
    for i, js, cnt in items:
        for s in range(js, js + cnt):
            pairs += 512 * 256 * (0.5 if s // 2 == i else 1.0)
    assert pairs - 0.5 * npad == npad * (npad - 1) / 2  # ordered diagonal pairs include the masked self pairs
    for world in (2, 3, 8):
        shards = [items[r::world] for r in range(world)]
        assert sorted(sum(shards, [])) == sorted(items)
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


def test_self_pair_detection_in_diagonal_stages():
    """pairs16: self iff jl - ii == self_base - 32 step, self_base = ibase + iw - 256 js."""
    iblk = 3
    ibase = iblk * 512
    hits = 0
    for js in (2 * iblk, 2 * iblk + 1):
        for warp in range(8):
            for lane, a, b in lanes():
                iw = warp * 64 + a * 8
                self_base = ibase + iw - js * 256
                for step in range(8):
                    for jj in range(8):
                        jl = ((jj ^ a) << 2) | b
                        for ii in range(8):
                            is_self = (jl - ii) == self_base - step * 32
                            gi, gj = ibase + iw + ii, js * 256 + step * 32 + jl
                            assert is_self == (gi == gj)
                            hits += is_self
    assert hits == 512  # every bead of the block meets itself exactly once


def test_fixed_point_accumulation_is_order_independent():
    rng = np.random.default_rng(1)
    f = rng.normal(0, 300.0, size=20000).astype(np.float32)
    q = np.rint(f.astype(np.float64) * 2.0 ** 24).astype(np.int64)
    a = int(q.sum())
    b = int(q[rng.permutation(len(q))].sum())
    assert a == b  # integer sums commute: the force does not depend on which CTA ran which item
    assert abs(a / 2.0 ** 24 - float(f.astype(np.float64).sum())) <= len(f) * 2.0 ** -25


def test_counting_sort_by_cell_reproduces_the_stable_order():
    """csrc/mmm_cells.cu: histogram -> exclusive scan -> unstable atomic placement -> id-rank inside
    each cell.  Whatever order the atomics land in, the result is the (key, id) order of a stable
    sort — the oracle's definition of the cell list."""
    rng = np.random.default_rng(4)
    n, ncodes = 5000, 512
    keys = rng.integers(0, ncodes, size=n)
    keys[rng.integers(0, n, size=200)] = 17  # a crowded cell
    count = np.bincount(keys, minlength=ncodes)
    cstart = np.concatenate([[0], np.cumsum(count)[:-1]])
    cend = cstart + count
    want = np.lexsort((np.arange(n), keys))  # stable sort by key, ties by id
    for trial in range(3):
        arrival = rng.permutation(n)  # the order in which the atomics happen to execute
        cursor = np.zeros(ncodes, dtype=int)
        slot_id = np.full(n, -1)
        for i in arrival:
            slot_id[cstart[keys[i]] + cursor[keys[i]]] = i
            cursor[keys[i]] += 1
        order = np.full(n, -1)
        for s in range(n):
            i = slot_id[s]
            k = keys[i]
            rank = int((slot_id[cstart[k]:cend[k]] < i).sum())
            order[cstart[k] + rank] = i
        assert np.array_equal(order, want)
