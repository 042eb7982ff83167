"""A SECOND, independent restatement of the reference's 1D path-guided SGD, in pure Python.

Test infrastructure.  Written from the reference source (/root/reference/src/sgd.rs, cited by line) and SURVEY.md
Appendix A only — it deliberately shares no code with oracle/gfs_oracle.cpp, whose job it is to cross-check: if the
C++ oracle and this file, driven by the same xoshiro256+ stream on one thread with exact-count epochs, end with
bit-identical positions, then the oracle's reading of `PathIndex`, `fast_precise_pow`, `DirtyZipfian`, the zeta
table, the eta schedule, the cooling rule, the lazy draw order and the update arithmetic is confirmed by a second
reading.  What neither can pin is the arithmetic of the un-vendored crates (rand 0.9 `Uniform`, rand_xoshiro 0.7):
both follow the published algorithms spelled out in tests/test_oracle.py — "parity unpinned" narrows to that boundary.

Pure Python floats are IEEE-754 doubles with round-to-nearest and no fused multiply-add, like Rust's f64.
"""
import math
import struct

M64 = (1 << 64) - 1


# ---- rand_xoshiro 0.7: Xoshiro256Plus::seed_from_u64 (SplitMix64 fills the state), next_u64 -----------------
class Xoshiro256Plus:
    def __init__(self, seed: int):
        s, z = [], seed & M64
        for _ in range(4):                       # SplitMix64 (Vigna)
            z = (z + 0x9E3779B97F4A7C15) & M64
            x = z
            x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
            x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
            s.append(x ^ (x >> 31))
        self.s = s

    def next_u64(self) -> int:
        s = self.s
        out = (s[0] + s[3]) & M64
        t = (s[1] << 17) & M64
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = ((s[3] << 45) | (s[3] >> 19)) & M64
        return out

    def f64(self) -> float:                      # rng.random::<f64>(): 53 high bits * 2^-53
        return (self.next_u64() >> 11) * (1.0 / 9007199254740992.0)

    def below(self, n: int) -> int:              # Uniform::new(0, n).sample: widening multiply with rejection
        if n <= 1 << 32:
            thresh = ((1 << 32) - n) % n
            while True:
                m = (self.next_u64() >> 32) * n
                if (m & 0xFFFFFFFF) >= thresh:
                    return m >> 32
        thresh = ((1 << 64) - n) % n
        while True:
            m = self.next_u64() * n
            if (m & M64) >= thresh:
                return m >> 64


# ---- casts with Rust's semantics ---------------------------------------------------------------------------
def as_i32(v: float) -> int:                     # `as i32`: truncate, saturate, NaN -> 0
    if v != v:
        return 0
    if v >= 2147483647.0:
        return 2147483647
    if v <= -2147483648.0:
        return -2147483648
    return int(v)


def as_u64(v: float) -> int:                     # `as u64`
    if v != v or v <= 0.0:
        return 0
    if v >= 18446744073709551616.0:
        return M64
    return int(v)


def fast_precise_pow(a: float, b: float) -> float:                      # sgd.rs:155-182
    e = as_i32(b)
    bits = struct.unpack("<Q", struct.pack("<d", a))[0]
    high = bits >> 32
    if high >= 1 << 31:
        high -= 1 << 32                                                 # (bits >> 32) as i32
    diff = high - 1072632447
    diff = (diff + (1 << 31)) % (1 << 32) - (1 << 31)                    # i32 wrapping (release build)
    new_high = as_i32((b - float(e)) * float(diff) + 1072632447.0)
    frac = struct.unpack("<d", struct.pack("<Q", ((new_high & 0xFFFFFFFF) << 32) & M64))[0]   # (new_high as u64) << 32
    base, r, ex = a, 1.0, e
    while ex != 0:
        if ex & 1:
            r *= base
        base *= base
        ex >>= 1                                                        # e >= 0 at every call site
    return r * frac


def dirty_zipf(rng, zmin: int, zmax: int, theta: float, zeta: float, zeta2theta: float) -> int:   # sgd.rs:128-150
    n = zmax - zmin + 1
    alpha = 1.0 / (1.0 - theta)
    eta = (1.0 - fast_precise_pow(2.0 / float(n), 1.0 - theta)) / (1.0 - zeta2theta / zeta)
    u = rng.f64()
    uz = u * zeta
    if uz < 1.0:
        return zmin
    if uz < 1.0 + fast_precise_pow(0.5, theta):
        return zmin + 1
    result = float(zmin) + float(n) * fast_precise_pow(eta * u - eta + 1.0, alpha)
    return min(as_u64(result), zmax)


def schedule(w_min, w_max, iter_max, iter_with_max_lr, eps):            # sgd.rs:617-638
    eta_max = 1.0 / w_min
    eta_min = eps / w_max
    lam = math.log(eta_max / eta_min) / (float(iter_max) - 1.0)
    return [eta_max * math.exp(-lam * float(abs(t - iter_with_max_lr))) for t in range(iter_max + 1)]


def zeta_table(space, space_max, q, theta):                             # sgd.rs:311-331
    size = (space if space <= space_max else space_max + (space - space_max) // q + 1) + 1
    zetas = [0.0] * size
    z = 0.0
    for i in range(1, space + 1):
        z += fast_precise_pow(1.0 / float(i), theta)
        if i <= space_max:
            zetas[i] = z
        if i >= space_max and (i - space_max) % q == 0:
            idx = space_max + 1 + (i - space_max) // q
            if idx < size:
                zetas[idx] = z
    return zetas


def path_index(steps, path_first, present, seq_len):                    # sgd.rs:34-71
    pos, path_of, rank = [], [], []
    infos = []
    for p in range(len(path_first) - 1):
        position = 0
        lo, hi = int(path_first[p]), int(path_first[p + 1])
        for k, s in enumerate(range(lo, hi)):
            node = int(steps[s]) >> 1
            pos.append(position)
            path_of.append(p)
            rank.append(k)
            if node < len(present) and present[node]:
                position += int(seq_len[node])                          # missing node => +0 (:52-54)
        infos.append((hi - lo, position, lo))
    return pos, path_of, rank, infos


def path_linear_sgd_single_thread(g, params) -> list:
    """sgd.rs:237-614 with ONE worker (tid 0, xoshiro256+(seed)) and exact-count epochs: epoch e = 0..iter_max uses
    etas[e], cools iff e > floor(cooling_start * iter_max), and applies exactly min_term_updates updates
    (the reference applies "at least" that many, by a 1 ms polling race — SURVEY.md §3.4).
    g: oracle.Graph (arrays only); params: any object with the PathSGDParams fields.  Returns X by dense idx."""
    steps = [int(v) for v in g.steps]
    pos, path_of, rank, infos = path_index(steps, g.path_first, g.present, g.seq_len)
    total_steps = len(steps)
    node_ids = [int(v) for v in g.node_ids()]                            # :276-284
    h2i, X, cum, idx = {}, [], 0, 0
    for nid in node_ids:                                                 # :286-293: idx advances for live nodes only
        if nid < len(g.present) and g.present[nid]:
            X.append(float(cum))
            h2i[nid] = idx
            cum += int(g.seq_len[nid])
            idx += 1
    if not any(c > 1 for c, _, _ in infos):
        return X                                                         # :250-261
    first_cooling = int(math.floor(params.cooling_start * float(params.iter_max)))       # :297
    etas = schedule(1.0 / params.eta_max, 1.0, params.iter_max, params.iter_with_max_learning_rate, params.eps)
    space, space_max, q = params.space, params.space_max, params.space_quantization_step
    zetas = zeta_table(space, space_max, q, params.theta)
    rng = Xoshiro256Plus(params.seed + 0)                                # :431-432
    for e in range(params.iter_max + 1):
        eta = etas[e]
        cooling = e > first_cooling                                      # :393-396 (strict)
        theta = 0.001 if cooling else params.theta
        applied = 0
        while applied < params.min_term_updates:
            s = rng.below(total_steps)                                   # :444
            p = path_of[s]
            n = infos[p][0]
            if n == 1:
                continue
            ra = rank[s]
            rb = ra
            if cooling or rng.below(2) == 1:                             # :456 (lazy: no draw while cooling)
                if ra > 0 and (rng.below(2) == 1 or ra == n - 1):        # :460 (lazy: no draw at rank 0)
                    J = min(space, ra)
                    k = space_max + (J - space_max) // q + 1 if J > space_max else J
                    k = min(k, len(zetas) - 1)
                    z = dirty_zipf(rng, 1, J, theta, zetas[k], 1.0 + fast_precise_pow(0.5, theta))
                    rb = max(ra - z, 0)                                  # saturating_sub
                elif ra < n - 1:
                    J = min(space, n - ra - 1)
                    k = space_max + (J - space_max) // q + 1 if J > space_max else J
                    k = min(k, len(zetas) - 1)
                    z = dirty_zipf(rng, 1, J, theta, zetas[k], 1.0 + fast_precise_pow(0.5, theta))
                    rb = min(ra + z, n - 1)
            else:
                rb = rng.below(n)                                        # :493-494
            if ra == rb:
                continue
            a, b = infos[p][2] + ra, infos[p][2] + rb
            d = abs(float(pos[a]) - float(pos[b]))                       # :509-513
            if d == 0.0:
                continue
            mu = min(eta * (1.0 / d), 1.0)                               # :517-520
            i, j = h2i.get(steps[a] >> 1), h2i.get(steps[b] >> 1)        # :525-538
            if i is None or j is None:
                continue
            dx = X[i] - X[j]
            if dx == 0.0:
                dx = 1e-9                                                # :546-548
            mag = abs(dx)
            delta = mu * (mag - d) / 2.0                                 # :552
            r_x = (delta / mag) * dx
            X[i] = X[i] - r_x                                            # :575 (re-load, then store)
            X[j] = X[j] + r_x                                            # :576
            applied += 1
    return X
