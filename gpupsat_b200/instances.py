"""Synthetic instance generators and literal-encoding helpers (host side, data only).

Literal encoding everywhere in this repo is the reference's (SATSolver/SolverTypes.cu:6-32):
``x = 2*var + (1 if positive else 0)``, vars 0-based; DIMACS ``k`` -> ``mk_lit(|k|-1, k>0)``
(FileManager/CnfReader.cpp:109-112).

Generators follow SURVEY.md section 8(d) so that any language reproduces the same instance from a seed:
a splitmix64 stream; clause i draws vars ``next() % n`` rejecting repeats inside the clause until k distinct,
then one sign bit ``next() & 1`` per literal (1 = positive).
"""
from __future__ import annotations

import numpy as np

MASK64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed: int):
        self.s = seed & MASK64

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & MASK64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
        return z ^ (z >> 31)


def mk_lit(var: int, positive: bool) -> int:
    return 2 * var + (1 if positive else 0)


def lit_to_dimacs(x: int) -> int:
    return (x >> 1) + 1 if (x & 1) else -((x >> 1) + 1)


def dimacs_to_lit(k: int) -> int:
    return mk_lit(abs(k) - 1, k > 0)


def random_ksat(n: int, m: int, seed: int, k: int = 3):
    """Uniform random k-SAT as CSR (offsets int64[m+1], lits int32[m*k]) in reference encoding."""
    rng = SplitMix64(seed)
    lits = np.empty(m * k, dtype=np.int32)
    pos = 0
    for _ in range(m):
        vs = []
        while len(vs) < k:
            v = rng.next() % n
            if v not in vs:
                vs.append(v)
        for v in vs:
            lits[pos] = 2 * v + (rng.next() & 1)
            pos += 1
    offsets = np.arange(0, m * k + 1, k, dtype=np.int64)
    return offsets, lits


def _splitmix_vec(seed: int, count: int) -> np.ndarray:
    """count consecutive splitmix64 outputs, vectorised (same stream as SplitMix64(seed))."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, count + 1, dtype=np.uint64)
        z = np.uint64(seed & MASK64) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def random_3sat_large(n: int, m: int, seed: int):
    """Vectorised generator for the large-database sweep (config 4).

    Not stream-identical to ``random_ksat`` (rejection sampling is replaced by a fixed 3-draw scheme:
    v0 = r0 % n, v1 = (v0 + 1 + r1 % (n-1)) % n, v2 drawn likewise and nudged off v0/v1), but deterministic
    in ``seed`` and uniform over clauses with 3 distinct vars.
    """
    r = _splitmix_vec(seed, 6 * m).reshape(m, 6)
    n64 = np.uint64(n)
    v0 = r[:, 0] % n64
    v1 = (v0 + np.uint64(1) + r[:, 1] % np.uint64(n - 1)) % n64
    v2 = r[:, 2] % np.uint64(n - 2)
    lo = np.minimum(v0, v1)
    hi = np.maximum(v0, v1)
    v2 = v2 + (v2 >= lo).astype(np.uint64)
    v2 = v2 + (v2 >= hi).astype(np.uint64)
    vs = np.stack([v0, v1, v2], axis=1).astype(np.int64)
    sg = (r[:, 3:6] & np.uint64(1)).astype(np.int64)
    lits = (2 * vs + sg).astype(np.int32).reshape(-1)
    offsets = np.arange(0, 3 * m + 1, 3, dtype=np.int64)
    return offsets, lits


def pigeonhole(p: int, h: int):
    """PHP(p,h): var(i,j) = i*h + j (0-based); p clauses OR_j x(i,j); then per hole j all pairs
    (~x(a,j) | ~x(b,j)), a<b.  Layout reproduces tests/cnf/unsat_php_3_2.cnf:5-13 of the reference."""
    offs = [0]
    lits = []
    for i in range(p):
        for j in range(h):
            lits.append(mk_lit(i * h + j, True))
        offs.append(len(lits))
    for j in range(h):
        for a in range(p):
            for b in range(a + 1, p):
                lits.append(mk_lit(a * h + j, False))
                lits.append(mk_lit(b * h + j, False))
                offs.append(len(lits))
    return np.asarray(offs, dtype=np.int64), np.asarray(lits, dtype=np.int32)


def to_dimacs(offsets, lits, n_vars: int | None = None, comment: str | None = None) -> str:
    m = len(offsets) - 1
    if n_vars is None:
        n_vars = int(lits.max() >> 1) + 1 if len(lits) else 0
    out = []
    if comment:
        out.append("c " + comment)
    out.append(f"p cnf {n_vars} {m}")
    for c in range(m):
        row = [str(lit_to_dimacs(int(x))) for x in lits[offsets[c]:offsets[c + 1]]]
        out.append(" ".join(row + ["0"]))
    return "\n".join(out) + "\n"


def parse_dimacs_text(text: str):
    """Minimal DIMACS reader for test fixtures (the product reader is the C++ one behind gpsat_read_dimacs)."""
    offs = [0]
    lits = []
    n_vars = 0
    for line in text.splitlines():
        s = line.strip()
        if not s or s[0] in "c%":
            continue
        if s[0] == "p":
            n_vars = int(s.split()[2])
            continue
        for tok in s.split():
            k = int(tok)
            if k == 0:
                offs.append(len(lits))
            else:
                lits.append(dimacs_to_lit(k))
    return np.asarray(offs, dtype=np.int64), np.asarray(lits, dtype=np.int32), n_vars


def check_model(offsets, lits, model) -> bool:
    """Independent, sound model check against the ORIGINAL CNF (the reference's own printed check is
    unsound: SATSolver/Results.cu:143-155 ignores signs).  model[v] in {0,1}."""
    model = np.asarray(model)
    val = (model[lits >> 1] == (lits & 1))
    sat = np.add.reduceat(val.astype(np.int64), offsets[:-1]) if len(lits) else np.zeros(0)
    lens = np.diff(offsets)
    if (lens == 0).any():
        return False
    return bool((sat > 0).all())


def planted_3sat_large(n: int, m: int, seed: int):
    """random_3sat_large with a planted solution: the hidden assignment is drawn from the same stream and every
    clause it would falsify (1 in 8) gets one sign flipped, so the formula is satisfiable and any subset of the
    planted assignment propagates without conflict (config 4: an under-constrained instance whose BCP throughput
    can be swept with trails of any length).  Returns (offsets, lits, planted[n] in {0,1})."""
    offsets, lits = random_3sat_large(n, m, seed)
    planted = (_splitmix_vec(seed ^ 0x5EED5EED, n) & np.uint64(1)).astype(np.int32)
    l3 = lits.reshape(m, 3)
    true3 = planted[l3 >> 1] == (l3 & 1)
    dead = ~true3.any(axis=1)
    which = (_splitmix_vec(seed ^ 0xF11BF11B, m) % np.uint64(3)).astype(np.int64)
    rows = np.nonzero(dead)[0]
    l3[rows, which[rows]] ^= 1
    return offsets, l3.reshape(-1).astype(np.int32), planted


def sweep_trails(n: int, n_jobs: int, length: int, seed: int, planted=None):
    """Trails for the large-database BCP sweep (config 4): job j assigns `length` literals over DISTINCT variables,
    v_i = (a_j * i + b_j) mod n with a_j coprime to n (an affine permutation seeded by splitmix64(seed ^ j)); signs
    follow `planted` when given (conflict-free trails), else come from the same stream.
    Returns CSR (offsets int64[n_jobs+1], lits int32[n_jobs*length])."""
    import math
    lits = np.empty((n_jobs, length), dtype=np.int32)
    i = np.arange(length, dtype=np.int64)
    for j in range(n_jobs):
        rng = SplitMix64(seed ^ (j * 0x9E3779B1 + 1))
        a = int(rng.next() % n) | 1
        while math.gcd(a, n) != 1:
            a = (a + 2) % n
        b = int(rng.next() % n)
        v = (a * i + b) % n
        if planted is not None:
            sg = planted[v].astype(np.int64)
        else:
            sg = (_splitmix_vec(rng.next(), length) & np.uint64(1)).astype(np.int64)
        lits[j] = (2 * v + sg).astype(np.int32)
    offsets = np.arange(0, n_jobs * length + 1, length, dtype=np.int64)
    return offsets, lits.reshape(-1)
