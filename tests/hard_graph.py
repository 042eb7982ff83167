"""A structurally hard test graph at a size where the library's default schedule (sliding window + warp-coherent
steps) really engages: reference fixture DRB1-3123.gfa tiled end to end and perturbed.

DRB1-3123 is the one structurally real graph the reference ships (4955 nodes, 12 paths, 35 059 steps, one path
fully reverse, stress floor 0.176) but it is far too small for the sweep schedule (records > 64 MB).  Here T copies
of it are chained; output path k walks tile after tile and, per tile, (a) follows a RANDOM one of DRB1's 12 paths
(haplotypes recombine between tiles), (b) with probability p_inv walks the tile backwards with flipped orientations
(an inversion of the whole tile), (c) inverts random sub-segments of its walk (nested inversions: an inverted
segment inside an inverted tile), (d) with probability p_rep walks the tile twice in a row (a tandem repeat: the
path revisits nodes, the graph has cycles).  Deterministic in (tiles, seed)."""
import os

import numpy as np

from conftest import DATA


def tiled_drb1(gfs, tiles: int = 150, seed: int = 7, p_inv: float = 0.10, p_rep: float = 0.10, p_sub: float = 0.30):
    """Returns (step_handles u64[S], path_first u64[P+1], node_len u32[N]) in the C-ABI's flat form."""
    g = gfs.load_gfa(os.path.join(DATA, "DRB1-3123.gfa"))
    h0, f0, l0 = g.dense()
    n0, p0 = len(l0), len(f0) - 1
    walks = [h0[int(f0[k]):int(f0[k + 1])] for k in range(p0)]
    rng = np.random.default_rng(seed)
    one = np.uint64(1)

    def flip(w):
        return (w[::-1] ^ one).copy()

    paths = []
    for k in range(p0):
        parts = []
        for t in range(tiles):
            w = walks[int(rng.integers(p0))].copy()
            while rng.random() < p_sub:                              # nested inversions of sub-segments
                a = int(rng.integers(0, len(w) - 20))
                b = a + int(rng.integers(20, min(400, len(w) - a)))
                w[a:b] = flip(w[a:b])
            if rng.random() < p_inv:                                 # the whole tile walked backwards
                w = flip(w)
            w = w + (np.uint64(t * n0) << one)
            parts.append(w)
            if rng.random() < p_rep:                                 # tandem repeat: walk the tile again
                parts.append(w)
        paths.append(np.concatenate(parts))
    first = np.zeros(p0 + 1, dtype=np.uint64)
    np.cumsum([len(p) for p in paths], out=first[1:])
    return np.concatenate(paths).astype(np.uint64), first, np.tile(l0, tiles).astype(np.uint32)
