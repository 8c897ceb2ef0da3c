"""Helpers shared by the tests: readers of the golden files written by tests/golden/make_golden.py."""
import os

from dicey_b200.api import HuntParams

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def read_queries(path):
    out = []
    for line in open(path):
        line = line.rstrip("\n")
        if not line:
            continue
        if "\t" in line:
            n, s = line.split("\t", 1)
        else:
            n, s = "", line
        out.append((n, s))
    return out


def read_records(path):
    """-> list of dict(seq, distance, msgs, push, sorted, work) per query."""
    qs = []
    for line in open(path):
        f = line.rstrip("\n").split("\t")
        if f[0] == "Q":
            qs.append({"seq": f[2], "distance": int(f[3]), "msgs": [], "push": [], "sorted": [], "work": None})
        elif f[0] == "M":
            qs[-1]["msgs"].append(f[1])
        elif f[0] in ("P", "S"):
            rec = (int(f[1]), int(f[2]), int(f[3]), f[4], f[5], f[6])
            qs[-1]["push" if f[0] == "P" else "sorted"].append(rec)
        elif f[0] == "W":
            qs[-1]["work"] = tuple(int(x) for x in f[1:])
    return qs


def read_rec_tsv(path):
    names, lens = [], []
    for line in open(path):
        n, l = line.split()
        names.append(n)
        lens.append(int(l) + 1)  # util.h:201
    return names, lens


def params_from_flags(flags: str, search=False) -> HuntParams:
    p = HuntParams()
    if search:
        p.maxmatches = 10000
    t = flags.split()
    i = 0
    while i < len(t):
        if t[i] == "-d":
            p.distance = int(t[i + 1]); i += 2
        elif t[i] == "-m":
            p.maxmatches = int(t[i + 1]); i += 2
        elif t[i] == "-x":
            p.max_neighborhood = int(t[i + 1]); i += 2
        elif t[i] == "-k":
            p.seed_len = int(t[i + 1]); i += 2
        elif t[i] == "-n":
            p.hamming = True; i += 1
        elif t[i] == "-f":
            p.forward_only = True; i += 1
        else:
            i += 1
    return p


HUNT_CASES = [
    ("cfg1_d0", "t1m"), ("t1m_e1", "t1m"), ("t1m_h1", "t1m"), ("t1m_h2", "t1m"), ("t1m_e2", "t1m"),
    ("t1m_e1_fwd", "t1m"), ("t1m_e0", "t1m"), ("stress_e1", "stress"), ("stress_h1", "stress"),
    ("stress_e1_m7", "stress"), ("stress_h2_m50", "stress"), ("stress_e2", "stress"),
    # the reference truncates the neighbourhood at -x (neighbors.h:50) in these
    ("t1m_e1_x50", "t1m"), ("t1m_e1_x150", "t1m"), ("t1m_h2_x300", "t1m"), ("t1m_h1_x40", "t1m"), ("t1m_e2_x500", "t1m"),
    ("t1m_e2_x5000", "t1m"), ("t1m_e2_long", "t1m"), ("stress_e2_x2000", "stress"), ("t1m_e1_m0", "t1m"),
    # distance 3: searched from host-made neighbour lists
    ("t1m_h3", "t1m"), ("t1m_e3_x3000", "t1m"), ("t1m_d12", "t1m"),
]


def fm9_sections(path):
    """Byte ranges of a .fm9: dict name -> bytes."""
    import struct
    data = open(path, "rb").read()
    pos = 0
    out = {}

    def int_vector():
        nonlocal pos
        (h,) = struct.unpack_from("<Q", data, pos)
        bits = h & ((1 << 56) - 1)
        n = 8 + ((bits + 63) >> 6) * 8
        pos += n
        return n

    def select():
        nonlocal pos
        (cnt,) = struct.unpack_from("<Q", data, pos)
        pos += 8
        if cnt:
            int_vector()
            int_vector()
            for _ in range((cnt + 4095) >> 12):
                int_vector()

    def take(name, fn):
        nonlocal pos
        a = pos
        fn()
        out[name] = data[a:pos]

    def hdr():
        nonlocal pos
        pos += 16

    def tree():
        nonlocal pos
        (nn,) = struct.unpack_from("<Q", data, pos)
        pos += 8 + nn * 22 + 256 * 2 + 256 * 8

    def alphabet():
        nonlocal pos
        int_vector(); int_vector(); int_vector()
        pos += 2

    take("header", hdr)
    take("bv", int_vector)
    take("rank", int_vector)
    take("select1", select)
    take("select0", select)
    take("tree", tree)
    take("sa", int_vector)
    take("isa", int_vector)
    take("alphabet", alphabet)
    assert pos == len(data), (pos, len(data))
    return out


def write_primer3_config(dirpath, dump_path=None):
    """A primer3_config directory (what `dicey search -i` reads) rebuilt from the committed dump of
    the tables the reference held (tests/golden/thal.params.tsv): the GPU box has no /root/reference.
    Values are written with repr(), which strtod reads back to the same double."""
    import struct
    dump_path = dump_path or os.path.join(GOLDEN, "thal.params.tsv")
    tabs = {}
    for line in open(dump_path):
        f = line.split()
        tabs[f[0]] = [struct.unpack("<d", struct.pack("<Q", int(x, 16)))[0] for x in f[2:]]
    os.makedirs(dirpath, exist_ok=True)

    def num(x):
        return "inf" if x >= 999999.0 / 2 else repr(x)

    def four(name, key):
        t = tabs[key]
        with open(os.path.join(dirpath, name), "w") as fh:
            for i in range(4):
                for ii in range(4):
                    for j in range(4):
                        for jj in range(4):
                            fh.write(num(t[((i * 5 + ii) * 5 + j) * 5 + jj]) + "\n")

    four("stack.ds", "stackEntropies"); four("stack.dh", "stackEnthalpies")
    four("stackmm.ds", "stackint2Entropies"); four("stackmm.dh", "stackint2Enthalpies")
    four("tstack_tm_inf.ds", "tstackEntropies"); four("tstack.dh", "tstackEnthalpies")
    four("tstack2.ds", "tstack2Entropies"); four("tstack2.dh", "tstack2Enthalpies")
    for ext, k3, k5 in ((".ds", "dangleEntropies3", "dangleEntropies5"), (".dh", "dangleEnthalpies3", "dangleEnthalpies5")):
        with open(os.path.join(dirpath, "dangle" + ext), "w") as fh:
            for i in range(4):
                for j in range(4):
                    for k in range(4):
                        fh.write(num(tabs[k3][(i * 5 + k) * 5 + j]) + "\n")      # dangle3[i][k][j]
            for i in range(4):
                for j in range(4):
                    for k in range(4):
                        fh.write(num(tabs[k5][(i * 5 + j) * 5 + k]) + "\n")      # dangle5[i][j][k]
    for ext, ki, kb in ((".ds", "interiorLoopEntropies", "bulgeLoopEntropies"), (".dh", "interiorLoopEnthalpies", "bulgeLoopEnthalpies")):
        with open(os.path.join(dirpath, "loops" + ext), "w") as fh:
            for k in range(30):
                fh.write(f"{k + 1}\t{num(tabs[ki][k])}\t{num(tabs[kb][k])}\t0\n")
    for name in ("tetraloop.ds", "tetraloop.dh", "triloop.ds", "triloop.dh"):
        open(os.path.join(dirpath, name), "w").close()
    return dirpath
