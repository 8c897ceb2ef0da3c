"""Seeded synthetic inputs shared by the tests, bench.py and the device-side generator.

The synthetic reference of SURVEY.md section 8(d): ``nrec`` records of ``reclen`` bases, iid
uniform over ACGT.  Base number ``g`` (0-based over all records, newlines not counted) is

    "ACGT"[ splitmix64_finalizer(seed + g * 0x9E3779B97F4A7C15) >> 62 ]

so any window of the text can be produced without materialising the rest; the CUDA generator
in ``csrc/dg_build.cu`` (kernel ``k_synth_text``) evaluates the same expression.  The text the
index is built over follows dicey's dump format (reference ``src/index.h:96-115``): records
joined by ``\\n`` with one trailing ``\\n``; SDSL appends the ``\\0`` sentinel.
"""
from __future__ import annotations

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = ord("N")
for _a, _b in zip(b"ACGTN", b"TGCAN"):
    _COMP[_a] = _b


def _mix(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.uint64, copy=True)
    z ^= z >> np.uint64(30)
    z *= _M1
    z ^= z >> np.uint64(27)
    z *= _M2
    z ^= z >> np.uint64(31)
    return z


def base_codes(seed: int, start: int, count: int) -> np.ndarray:
    """2-bit codes (0..3 = A,C,G,T) of bases [start, start+count)."""
    with np.errstate(over="ignore"):
        g = np.arange(start, start + count, dtype=np.uint64)
        z = np.uint64(seed) + g * GOLDEN
        return (_mix(z) >> np.uint64(62)).astype(np.uint8)


def bases(seed: int, start: int, count: int) -> np.ndarray:
    """ASCII bases [start, start+count) as a uint8 array."""
    return ACGT[base_codes(seed, start, count)]


def text(seed: int, nrec: int, reclen: int) -> np.ndarray:
    """The dump text (records joined by '\\n', trailing '\\n'), without the sentinel."""
    out = np.empty(nrec * (reclen + 1), dtype=np.uint8)
    view = out.reshape(nrec, reclen + 1)
    for r in range(nrec):
        view[r, :reclen] = bases(seed, r * reclen, reclen)
        view[r, reclen] = 10
    return out


def records(nrec: int, reclen: int):
    """(names, seqlen+1) exactly as reference util.h:183-206 would return them."""
    return [f"chr{i + 1}" for i in range(nrec)], [reclen + 1] * nrec


def revcomp(seq: bytes) -> bytes:
    a = np.frombuffer(seq, dtype=np.uint8)
    return _COMP[a][::-1].tobytes()


def primers(seed: int, nrec: int, reclen: int, count: int, length: int, distance: int,
            indel: bool, rng_seed: int = 7, planted_frac: float = 0.5) -> list[bytes]:
    """Primer set of SURVEY.md section 8(d): ``planted_frac`` planted (a locus copied from the
    reference with k in 0..distance random edits, half of them reverse-complemented), the rest
    uniform random."""
    rng = np.random.default_rng(rng_seed)
    out: list[bytes] = []
    for i in range(count):
        if rng.random() < planted_frac:
            rec = int(rng.integers(0, nrec))
            off = int(rng.integers(0, reclen - length - 4))
            s = bytearray(bases(seed, rec * reclen + off, length + 4).tobytes())
            k = int(rng.integers(0, distance + 1))
            s = s[:length]
            for _ in range(k):
                op = int(rng.integers(0, 3)) if indel else 0
                p = int(rng.integers(0, len(s)))
                c = b"ACGT"[int(rng.integers(0, 4))]
                if op == 0:
                    s[p] = c
                elif op == 1 and len(s) > 10:
                    del s[p]
                else:
                    s.insert(p, c)
            seq = bytes(s)
            if rng.random() < 0.5:
                seq = revcomp(seq)
        else:
            seq = ACGT[rng.integers(0, 4, size=length)].tobytes()
        out.append(seq)
    return out


def primers_fast(seed: int, nrec: int, reclen: int, count: int, length: int, distance: int,
                 indel: bool, rng_seed: int = 7, return_truth: bool = False):
    """Vectorised fixed-length variant for the large bench batches: returns a (count, length)
    uint8 array.  Even rows are planted loci with k in 0..distance substitutions (and, in
    edit mode, optionally one 1-base shift emulating an indel at the 5' end), odd rows are
    uniform random; half of the planted rows are reverse-complemented."""
    rng = np.random.default_rng(rng_seed)
    out = ACGT[rng.integers(0, 4, size=(count, length))]
    npl = (count + 1) // 2
    rec = rng.integers(0, nrec, size=npl).astype(np.uint64)
    off = rng.integers(0, reclen - length - 1, size=npl).astype(np.uint64)
    start = rec * np.uint64(reclen) + off
    with np.errstate(over="ignore"):
        g = start[:, None] + np.arange(length, dtype=np.uint64)[None, :]
        z = np.uint64(seed) + g * GOLDEN
        planted = ACGT[(_mix(z) >> np.uint64(62)).astype(np.uint8)]
    k = rng.integers(0, distance + 1, size=npl)
    for j in range(distance):
        sel = np.nonzero(k > j)[0]
        pos = rng.integers(0, length, size=sel.size)
        planted[sel, pos] = ACGT[rng.integers(0, 4, size=sel.size)]
    if indel:
        # a deletion of the first base + a random base appended at the 3' end for a quarter
        # of the planted rows (costs one edit; keeps the array rectangular)
        sel = np.nonzero((rng.random(npl) < 0.25) & (k < max(distance, 1)))[0]
        planted[sel, :-1] = planted[sel, 1:]
        planted[sel, -1] = ACGT[rng.integers(0, 4, size=sel.size)]
    rc = rng.random(npl) < 0.5
    planted[rc] = _COMP[planted[rc]][:, ::-1]
    out[0::2] = planted
    out = np.ascontiguousarray(out)
    if return_truth:
        # row 2 i was copied from record rec[i], 0-based offset off[i]; rc[i]: reverse-complemented;
        # shifted[i]: the 5' base was dropped (an indel), so the locus starts one base later
        shifted = np.zeros(npl, dtype=bool)
        if indel:
            shifted[sel] = True
        return out, {"rec": rec.astype(np.int64), "off": off.astype(np.int64), "rc": rc, "shifted": shifted, "edits": k}
    return out
