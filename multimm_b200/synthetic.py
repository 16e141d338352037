"""Synthetic loop (.bedpe) and compartment (.bed) tracks in the reference's input formats
(README.md:265-288), so the same files can be read by the reference's own loaders
(utils.py:220, 425) and by multimm_b200.loaders.  Recipe: SURVEY.md section 8(d).

* bedpe: 7 tab-separated columns, no header; 10 kb anchors; per chromosome the number of loops
  is proportional to its length; span = 10 kb * (3 + floor(Exp(mean 30))); count ~ U[13, 200];
  one loop always ends in the last 10 kb of the last chromosome so the genome-wide resolution
  (max(col5) // N, utils.py:474) is ~ genome / N.
* bed: CALDER-like, >= 4 columns; run lengths 50 kb * Geom(1/8); labels uniform over the eight
  sub-compartment names -> Cs is ~25 % each of {2, 1, -1, -2}.
"""
from __future__ import annotations

import numpy as np

from .loaders import CHROM_NAMES, CHROM_SIZES

ANCHOR = 10_000
LABELS = ("A.1.1", "A.1.2", "A.2.1", "A.2.2", "B.1.1", "B.1.2", "B.2.1", "B.2.2")


def _chrom_list(chrom=None, n_chroms=22):
    if chrom is not None:
        return [chrom]
    return [CHROM_NAMES[i] for i in range(n_chroms)]


def write_bedpe(path, n_loops, seed=0, chrom=None, region=None, n_chroms=22):
    """Write a synthetic loop file; returns the number of rows written."""
    rng = np.random.default_rng(seed)
    chroms = _chrom_list(chrom, n_chroms)
    sizes = np.array([CHROM_SIZES[c] for c in chroms], dtype=np.int64)
    lo = np.zeros(len(chroms), dtype=np.int64)
    hi = sizes.copy()
    if region is not None:
        lo[:] = region[0]
        hi[:] = region[1]
    share = (hi - lo) / float((hi - lo).sum())
    per = np.maximum(1, np.round(share * n_loops).astype(int))
    rows = []
    for c, name in enumerate(chroms):
        span = ANCHOR * (3 + np.floor(rng.exponential(30.0, size=per[c])).astype(np.int64))
        span = np.minimum(span, (hi[c] - lo[c]) // 2)
        start = lo[c] + 1 + (rng.random(per[c]) * (hi[c] - lo[c] - span - 2 * ANCHOR - 2)).astype(np.int64)
        count = rng.integers(13, 201, size=per[c]).astype(float)
        for s0, sp, ct in zip(start, span, count):
            rows.append((name, s0, s0 + ANCHOR, name, s0 + sp, s0 + sp + ANCHOR, ct))
    # the closing loop that pins the genome-wide resolution
    last = chroms[-1]
    end = int(hi[-1]) - 1
    rows.append((last, end - 60 * ANCHOR, end - 59 * ANCHOR, last, end - ANCHOR, end, 50.0))
    with open(path, "w") as f:
        for r in rows:
            f.write(f"{r[0]}\t{r[1]}\t{r[2]}\t{r[3]}\t{r[4]}\t{r[5]}\t{r[6]}\n")
    return len(rows)


def write_bed(path, seed=0, chrom=None, n_chroms=22, bin_size=50_000):
    """Write a synthetic sub-compartment file; returns the number of rows written."""
    rng = np.random.default_rng(seed + 7919)
    n = 0
    with open(path, "w") as f:
        for name in _chrom_list(chrom, n_chroms):
            pos, size = 0, CHROM_SIZES[name]
            while pos < size:
                run = bin_size * int(rng.geometric(1.0 / 8.0))
                end = min(pos + run, size)
                f.write(f"{name}\t{pos}\t{end}\t{LABELS[int(rng.integers(0, 8))]}\n")
                pos = end
                n += 1
    return n


# The sizes of BASELINE.json's configs: loops requested per config.
CONFIGS = {
    "S1_region": dict(n_beads=10_000, chrom="chr1", region=(10_000_000, 110_000_000), n_loops=400),
    "S2_chrom": dict(n_beads=50_000, chrom="chr1", region=None, n_loops=2_000),
    "S3_gw": dict(n_beads=200_000, chrom=None, region=None, n_loops=10_000),
    "S5_highres": dict(n_beads=2_000_000, chrom=None, region=None, n_loops=100_000),
}
