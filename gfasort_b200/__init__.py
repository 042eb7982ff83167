"""gfasort_b200 — B200-native (sm_100a) path-guided SGD for pangenome graphs.

One hot path of pangenome/gfasort — the 1D `Y` sort, the nD `L` layout and the path index they
sample from — as hand-written CUDA behind a C ABI (include/gfasort_cuda.h), plus this thin
host-side mirror of the reference's own interface (names as in reference src/sgd.rs, src/ygs.rs,
src/layout.rs).  No CPU fallback: importing works without a GPU, computing does not.
"""
from ._cabi import GfsError, LaunchCfg, Stats, lib  # noqa: F401
from .graph import BidirectedGraph, load_gfa  # noqa: F401
from .layout import Layout  # noqa: F401
from .sgd import (LayoutSGDParams, PathIndex, PathSGDParams, YgsParams, calculate_layout_stress,  # noqa: F401
                  initial_layout, initial_positions, layout_stress, path_linear_sgd, path_linear_sgd_array,
                  path_linear_sgd_layout, path_sgd_sort, sgd_sort_only, sort_positions, sort_stress)
from .io import load_gfa_flat, write_gfa, write_layout_tsv  # noqa: F401
from .synth import SynthGraph  # noqa: F401
from .ygs import (apply_grooming_with_reorder, count_edge_directions, exact_odgi_topological_order,  # noqa: F401
                  find_head_nodes, groom, groom_only, topological_sort_only, ygs_sort)

__version__ = "0.1.0"
