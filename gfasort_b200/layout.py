"""`Layout` — the nD layout value type, same storage and TSV format as reference src/layout.rs.

coords[node * 2 * dimensions + end * dimensions + dim], end 0 = '+', 1 = '-' (src/layout.rs:14-24).
"""
from __future__ import annotations

import io
from dataclasses import dataclass

import numpy as np


def _dim_name(dim: int) -> str:
    # src/layout.rs:248-256 dim_name: x, y, z, w, then a bare "d"
    return "xyzw"[dim] if dim < 4 else "d"


def _rust_f64(v: float) -> str:
    """Rust's `{}` (Display) for f64, as `write_tsv` uses it (src/layout.rs:150-160): shortest
    round-trip digits, never exponent notation, 1.0 prints as "1", NaN / inf / -inf."""
    v = float(v)
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "inf" if v > 0 else "-inf"
    r = repr(v)
    if "e" in r or "E" in r:
        r = np.format_float_positional(v, trim="-")
    if r.endswith(".0"):
        r = r[:-2]
    return r


@dataclass
class Layout:
    dimensions: int
    num_nodes: int
    coords: np.ndarray

    @staticmethod
    def new(dimensions: int, num_nodes: int) -> "Layout":
        return Layout(dimensions, num_nodes, np.zeros(num_nodes * 2 * dimensions))

    @staticmethod
    def from_vectors(coord_vecs) -> "Layout":
        """[dim][2*node+end] -> interleaved (src/layout.rs:39-69)."""
        dims = len(coord_vecs)
        assert dims > 0, "Must have at least 1 dimension"
        entries = len(coord_vecs[0])
        assert entries % 2 == 0, "Must have even number of entries (2 per node)"
        for v in coord_vecs:
            assert len(v) == entries, "All dimension vectors must have same length"
        coords = np.stack([np.asarray(v, dtype=np.float64) for v in coord_vecs], axis=1).reshape(-1)
        return Layout(dims, entries // 2, coords)

    def index(self, node: int, end: int, dim: int) -> int:
        return node * 2 * self.dimensions + end * self.dimensions + dim

    def get(self, node: int, end: int, dim: int) -> float:
        return float(self.coords[self.index(node, end, dim)])

    def set(self, node: int, end: int, dim: int, value: float) -> None:
        self.coords[self.index(node, end, dim)] = value

    def get_coords(self, node: int, end: int) -> np.ndarray:
        s = self.index(node, end, 0)
        return self.coords[s:s + self.dimensions]

    def x_plus(self, node): return self.get(node, 0, 0)
    def y_plus(self, node): return self.get(node, 0, 1)
    def x_minus(self, node): return self.get(node, 1, 0)
    def y_minus(self, node): return self.get(node, 1, 1)

    def distance(self, node_a: int, end_a: int, node_b: int, end_b: int) -> float:
        d = self.get_coords(node_a, end_a) - self.get_coords(node_b, end_b)
        return float(np.sqrt(np.sum(d * d)))

    def calculate_stress(self, target_distances) -> float:
        """sqrt(sum(w*(d_layout-d_target)^2)/sum(w)), w = 1/d_target^2 (src/layout.rs:224-245)."""
        ws = wt = 0.0
        for node_a, end_a, node_b, end_b, target in target_distances:
            if target == 0.0:
                continue
            w = 1.0 / (target * target)
            err = self.distance(node_a, end_a, node_b, end_b) - target
            ws += err * err * w
            wt += w
        return float(np.sqrt(ws / wt)) if wt > 0 else 0.0

    def write_tsv(self, writer) -> None:
        """idx, then the + end's dims, then the - end's (src/layout.rs:138-163)."""
        hdr = ["idx"] + [f"{_dim_name(d)}+" for d in range(self.dimensions)] + \
              [f"{_dim_name(d)}-" for d in range(self.dimensions)]
        buf = io.StringIO()
        buf.write("\t".join(hdr) + "\n")
        c = self.coords.reshape(self.num_nodes, 2 * self.dimensions) if self.num_nodes else self.coords.reshape(0, 2 * self.dimensions)
        for node in range(self.num_nodes):
            buf.write(str(node) + "\t" + "\t".join(_rust_f64(v) for v in c[node]) + "\n")
        writer.write(buf.getvalue())

    @staticmethod
    def read_tsv(reader) -> "Layout":
        lines = [l for l in reader.read().split("\n") if l.strip()]
        header = lines[0].split("\t")
        dims = (len(header) - 1) // 2
        rows = [[float(x) for x in l.split("\t")[1:1 + 2 * dims]] for l in lines[1:]]
        coords = np.array(rows, dtype=np.float64).reshape(-1)
        return Layout(dims, len(rows), coords)
