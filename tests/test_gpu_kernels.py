"""GPU parity tests (-m gpu): the CUDA kernels, called through the C ABI (include/ema_b200.h),
against the oracle on the same seeded inputs and against the committed golden vectors.
Bar: bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers
from helpers import _p

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def unpack(flat, off):
    return [flat[off[i]:off[i + 1]] for i in range(len(off) - 1)]


@pytest.fixture(scope="module")
def emab():
    import ema_b200
    return ema_b200


@pytest.fixture(scope="module", params=[0, 1], ids=["lanes", "warp"])
def ctx(emab, request):
    """both SW kernel families: one thread per task (ksw_lanes.cuh) and one warp per task (ksw_warp.cuh)"""
    c = emab.Context()
    emab.set_sw_mode(c, request.param)
    return c


@pytest.fixture(scope="module")
def tiny(emab):
    ix = emab.Index(os.path.join(G, "tiny_rep", "ref.fa"))
    return ix, emab.Context(ix)


def test_extend_golden(emab, ctx):
    g = np.load(os.path.join(G, "sw_golden.npz"))
    out, cells = emab.extend_batch(ctx, unpack(g["q"], g["qo"]), unpack(g["t"], g["to"]), g["h0"])
    bad = np.nonzero((out != g["ext"]).any(axis=1))[0]
    assert len(bad) == 0, f"{len(bad)} mismatches, first {bad[:5]}: {out[bad[:3]]} vs {g['ext'][bad[:3]]}"
    assert ctx.last_launches == 1 and ctx.last_kernel_ms > 0


@pytest.mark.parametrize("params", [(100, 5, 100), (200, 5, 100), (7, 0, 0), (100, 5, 15)])
def test_extend_vs_oracle(emab, ctx, port_lib, params):
    w, eb, zd = params
    qs, ts, h0 = helpers.random_extend_tasks(4000, 100 + w, max_q=256)
    a, cells_gpu = emab.extend_batch(ctx, qs, ts, h0, w, eb, zd)
    b, cells_cpu = helpers.sw_extend(port_lib, "orc", qs, ts, h0, w, eb, zd)
    bad = np.nonzero((a != b).any(axis=1))[0]
    assert len(bad) == 0, f"{len(bad)} mismatches, first {bad[:5]}: {a[bad[:3]]} vs {b[bad[:3]]}"
    assert cells_gpu == cells_cpu, "the kernel must visit exactly the reference's DP cells"


def test_extend_edge_cases(emab, ctx, port_lib):
    qs = [np.array([0], np.uint8), np.array([0, 1, 2, 3] * 8, np.uint8), np.full(256, 2, np.uint8), np.full(40, 4, np.uint8),
          np.array([1] * 33, np.uint8), np.array([1] * 32, np.uint8), np.array([1] * 31, np.uint8)]
    ts = [np.array([0], np.uint8), np.array([0, 1, 2, 3] * 100, np.uint8), np.full(700, 2, np.uint8), np.full(50, 4, np.uint8),
          np.array([1] * 64, np.uint8), np.array([1] * 400, np.uint8), np.array([3], np.uint8)]
    h0 = np.array([1, 19, 150, 30, 5, 7, 9], np.int32)
    a, ca = emab.extend_batch(ctx, qs, ts, h0)
    b, cb = helpers.sw_extend(port_lib, "orc", qs, ts, h0)
    assert np.array_equal(a, b) and ca == cb
    # ragged batch sizes around the 32-task warp granularity, mixed lengths in one warp
    for n in (1, 31, 33, 65):
        qs2, ts2, h02 = helpers.random_extend_tasks(n, 7 + n, max_q=256)
        a, ca = emab.extend_batch(ctx, qs2, ts2, h02)
        b, cb = helpers.sw_extend(port_lib, "orc", qs2, ts2, h02)
        assert np.array_equal(a, b) and ca == cb
    out, cells = emab.extend_batch(ctx, [], [], np.zeros(0, np.int32))
    assert out.shape == (0, 6) and cells == 0
    with pytest.raises(emab.EmabError):
        emab.extend_batch(ctx, [np.zeros(300, np.uint8)], [np.zeros(10, np.uint8)], [5])


def test_global_golden_and_oracle(emab, ctx, port_lib):
    g = np.load(os.path.join(G, "sw_golden.npz"))
    out, cig, _ = emab.global_batch(ctx, unpack(g["q"], g["qo"]), unpack(g["t"], g["to"]), g["ws"], max_cigar=512)
    assert np.array_equal(out, g["glo"])
    assert np.array_equal(cig, g["gcig"])
    qs, ts, _ = helpers.random_extend_tasks(3000, 77, max_q=256)
    ws = np.array([max(int(w), abs(len(q) - len(t)) + 3) for w, q, t in
                   zip(np.random.default_rng(3).integers(0, 120, size=len(qs)), qs, ts)], dtype=np.int32)
    a, ca, cg = emab.global_batch(ctx, qs, ts, ws, max_cigar=600)
    b, cb, cc = helpers.sw_global(port_lib, "orc", qs, ts, ws, max_cigar=600)
    assert np.array_equal(a, b) and np.array_equal(ca, cb) and cg == cc


def test_local_golden_and_oracle(emab, ctx, port_lib):
    g = np.load(os.path.join(G, "sw_golden.npz"))
    out, _ = emab.local_batch(ctx, unpack(g["lq"], g["lqo"]), unpack(g["lt"], g["lto"]))
    bad = np.nonzero((out != g["loc"]).any(axis=1))[0]
    assert len(bad) == 0, f"{len(bad)} mismatches, first {bad[:5]}: {out[bad[:3]]} vs {g['loc'][bad[:3]]}"
    qs, ts, _ = helpers.random_extend_tasks(3000, 78, max_q=256)
    lq = [q for q in qs if len(q) >= 2]
    lt = [np.concatenate([t, q[::-1], t[: len(t) // 2]])[:1000] for q, t in zip(qs, ts) if len(q) >= 2]
    a, _ = emab.local_batch(ctx, lq, lt)
    b, _ = helpers.sw_local(port_lib, "orc", lq, lt)
    bad = np.nonzero((a != b).any(axis=1))[0]
    assert len(bad) == 0, f"{len(bad)} mismatches, first {bad[:5]}: {a[bad[:3]]} vs {b[bad[:3]]}"


def test_index_and_sa(emab, tiny):
    ix, c = tiny
    g = np.load(os.path.join(G, "fm_golden.npz"))
    assert np.array_equal(ix.info, g["info"])
    assert np.array_equal(emab.sa_batch(c, g["ks"], mode=0), g["sa"]), "dense SA must reproduce bwt_sa"
    assert np.array_equal(emab.sa_batch(c, g["ks"], mode=1), g["sa"]), "LF-walk bwt_sa"
    allk = np.arange(0, ix.seq_len + 1, dtype=np.int64)
    assert np.array_equal(emab.sa_batch(c, allk, mode=0), emab.sa_batch(c, allk, mode=1)), "every SA slot"
    sa = emab.sa_batch(c, allk[1:], mode=0)
    assert np.array_equal(np.sort(sa), np.arange(ix.seq_len)), "SA[1..] is a permutation of the text positions"


def test_smem_golden_and_oracle(emab, tiny, port_lib):
    """the exact seeding forms (EMAB_SEED_MODE 1, 3, 4): golden intervals with all four fields and the oracle's count of
    64-byte Occ-block loads"""
    ix, c = tiny
    g = np.load(os.path.join(G, "fm_golden.npz"))
    reads = unpack(g["reads"], g["roff"])
    pix = port_lib.orc_index_load(os.path.join(G, "tiny_rep", "ref.fa").encode())
    port_lib.orc_touches(C.c_void_p(pix), 1)
    for s in reads:
        s = np.ascontiguousarray(s)
        ob = np.zeros((256, 4), np.int64)
        port_lib.orc_collect_intv_flat(C.c_void_p(pix), len(s), _p(s, C.c_uint8), _p(ob, C.c_int64), 256)
    want_touches = port_lib.orc_touches(C.c_void_p(pix), 1)
    try:
        for mode in (1, 3, 4):
            emab.set_seed_mode(c, mode)
            ivs, touches = emab.smem_batch(c, reads)
            pos = 0
            for i, (iv, n) in enumerate(zip(ivs, g["n_intv"])):
                assert len(iv) == n and np.array_equal(iv, g["intv"][pos:pos + n]), f"mode {mode} read {i}"
                pos += n
            assert touches == want_touches, "64-byte Occ block touches must equal the reference's"
    finally:
        emab.set_seed_mode(c, 0)


def test_smem_default_form_golden(emab, tiny):
    """the default seeding form (seed_hot.cuh: one-hot Occ blocks, k-mer start table, text comparison at a unique locus):
    the reference's intervals on the coordinates mem_chain reads (x0, x2, info), x1 = 0; then the same against the exact
    form on mutated, N-bearing, truncated and text-end reads"""
    ix, c = tiny
    g = np.load(os.path.join(G, "fm_golden.npz"))
    reads = unpack(g["reads"], g["roff"])
    ivs, sectors = emab.smem_batch(c, reads)
    pos = 0
    for i, (iv, n) in enumerate(zip(ivs, g["n_intv"])):
        assert len(iv) == n and np.array_equal(iv[:, [0, 2, 3]], g["intv"][pos:pos + n][:, [0, 2, 3]]) and not iv[:, 1].any(), f"read {i}"
        pos += n
    assert sectors > 0
    ivs, _ = emab.smem_batch(c, [np.zeros(0, np.uint8), np.full(50, 4, np.uint8), reads[0][:10]])
    assert [len(x) for x in ivs] == [0, 0, 0]
    rng = np.random.default_rng(21)
    more = []
    for s in reads[:300]:
        t = s.copy(); t[rng.integers(0, len(t), size=int(rng.integers(1, 6)))] = 4; more.append(t)
        u = s.copy(); u[rng.integers(0, len(u), size=int(rng.integers(1, 4)))] ^= 1; more.append(np.minimum(u, 3).astype(np.uint8))
        more.append(np.ascontiguousarray(s[:int(rng.integers(1, len(s)))]))
    ref = helpers.read_fasta_nt4(os.path.join(G, "tiny_rep", "ref.fa"))
    cat = np.concatenate(ref)
    both = np.concatenate([cat, (3 - cat)[::-1]])
    L = len(cat)
    for a in (0, 1, L - 150, L - 75, L - 10, 2 * L - 151, 2 * L - 100):
        more.append(np.ascontiguousarray(both[a:a + 151]))
    more.append(rng.integers(0, 4, 151).astype(np.uint8))
    got, _ = emab.smem_batch(c, more)
    try:
        emab.set_seed_mode(c, 3)
        want, _ = emab.smem_batch(c, more)
    finally:
        emab.set_seed_mode(c, 0)
    for i, (a, b) in enumerate(zip(got, want)):
        assert len(a) == len(b) and np.array_equal(a[:, [0, 2, 3]], b[:, [0, 2, 3]]), f"read {i}"


def test_bucket_parser_on_the_device(emab):
    """emab_parse_bucket (read_special_fastq, src/align.c:759-806) against a plain restatement: lines sorted stably by their first
    BC_LEN bytes, fields split at single whitespace characters, '@' dropped from the id, barcodes 2-bit encoded (src/util.c:41-60);
    with and without a final newline, with barcode ties (file order must be kept), tabs as separators, and the error cases."""
    import random
    rnd = random.Random(5)
    ctx = emab.Context(None)
    bcs = ["".join(rnd.choice("ACGT") for _ in range(16)) for _ in range(40)]
    lines = []
    for i in range(3000):
        bc = rnd.choice(bcs)
        r1 = "".join(rnd.choice("ACGTN") for _ in range(rnd.randint(30, 151)))
        r2 = "".join(rnd.choice("ACGT") for _ in range(rnd.randint(30, 151)))
        sep = "\t" if i % 97 == 0 else " "
        lines.append(sep.join([bc, "@read%d" % i, r1, "I" * len(r1), r2, "J" * len(r2)]))
    for tail in ("\n", ""):
        data = ("\n".join(lines) + tail).encode()
        tab, codes = emab.parse_bucket(ctx, data)
        order = sorted(range(len(lines)), key=lambda k: lines[k][:16])     # Python's sort is stable
        assert len(tab) == len(lines)
        for k, li in enumerate(order):
            f = lines[li].replace("\t", " ").split(" ")
            t = tab[k]
            get = lambda o, l: data[int(o):int(o) + int(l)].decode()
            assert get(t["id_off"][0], t["id_len"][0]) == f[1][1:] and get(t["id_off"][1], t["id_len"][1]) == f[1][1:]
            assert get(t["read_off"][0], t["read_len"][0]) == f[2] and get(t["qual_off"][0], t["qual_len"][0]) == f[3]
            assert get(t["read_off"][1], t["read_len"][1]) == f[4] and get(t["qual_off"][1], t["qual_len"][1]) == f[5]
            code = 0
            for ch in reversed(f[0][:16]):
                code = code << 2 | "ACGT".index(ch)
            assert int(codes[k]) == code
    with pytest.raises(emab.EmabError, match="malformed barcode"):
        emab.parse_bucket(ctx, b"ACGTNACGTACGTACG @x ACGT IIII ACGT IIII\n")
    with pytest.raises(emab.EmabError, match="malformed barcode"):
        emab.parse_bucket(ctx, ("\n".join(lines[:5]) + "\n\n" + lines[6] + "\n").encode())   # an empty line
    with pytest.raises(emab.EmabError, match="MAX_READ_LEN"):
        emab.parse_bucket(ctx, (bcs[0] + " @x " + "A" * 201 + " " + "I" * 201 + " ACGT IIII\n").encode())
    tab, _ = emab.parse_bucket(ctx, b"")
    assert len(tab) == 0
