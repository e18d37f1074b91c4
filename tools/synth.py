"""Synthetic workloads for the `ema align` hot path.

Generates, deterministically from a seed (numpy ``default_rng``):

* a reference FASTA (+ ``.fai``) of iid uniform ACGT contigs, optionally with planted duplications
  so that multi-mapping, mate rescue, XA and the EM all fire;
* simulated barcoded read clouds in the three input formats ``ema align`` accepts:
  the preprocessed bucket ("special FASTQ", one line per pair, the format written by the
  reference's cpp/correct.cc:511-612 and read by src/align.c:759-806), and interleaved FASTQ for
  the ``tru`` / ``tellseq`` platforms (src/techs.c:33-61);
* ``ksw_extend2`` microbenchmark tasks (BASELINE.json config 5).

This is workload tooling for tests and bench.py; it is not part of the product path.
"""
from __future__ import annotations

import os
import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = np.array([3, 2, 1, 0], dtype=np.uint8)


# ----------------------------------------------------------------------------------------------
# reference
# ----------------------------------------------------------------------------------------------
def make_reference(n_contigs: int, contig_len: int, seed: int, dup_every: int = 0):
    """Return a list of uint8 code arrays (0..3), one per contig.

    ``dup_every`` > 0 plants L/dup_every duplications (2-10 kb, 0.2-2 % divergence) per genome.
    """
    rng = np.random.default_rng(seed)
    contigs = [rng.integers(0, 4, size=contig_len, dtype=np.uint8) for _ in range(n_contigs)]
    if dup_every:
        total = n_contigs * contig_len
        n_dup = total // dup_every
        for _ in range(n_dup):
            ln = int(rng.integers(2000, 10001))
            ln = min(ln, contig_len // 4)
            src_c = int(rng.integers(0, n_contigs))
            dst_c = int(rng.integers(0, n_contigs))
            src = int(rng.integers(0, contig_len - ln))
            dst = int(rng.integers(0, contig_len - ln))
            seg = contigs[src_c][src:src + ln].copy()
            div = rng.uniform(0.002, 0.02)
            mut = rng.random(ln) < div
            seg[mut] = (seg[mut] + rng.integers(1, 4, size=int(mut.sum()), dtype=np.uint8)) & 3
            if rng.random() < 0.5:
                seg = COMP[seg[::-1]]
            contigs[dst_c][dst:dst + ln] = seg
    return contigs


def write_fasta(path: str, contigs, names=None, width: int = 60):
    names = names or [f"chr{i + 1}" for i in range(len(contigs))]
    offset = 0
    with open(path, "wb") as f, open(path + ".fai", "w") as fai:
        for name, c in zip(names, contigs):
            hdr = f">{name}\n".encode()
            f.write(hdr)
            offset += len(hdr)
            seq = ACGT[c]
            n = len(seq)
            full = n // width
            body = np.empty(full * (width + 1), dtype=np.uint8)
            body.reshape(full, width + 1)[:, :width] = seq[: full * width].reshape(full, width)
            body.reshape(full, width + 1)[:, width] = 10
            f.write(body.tobytes())
            rem = n - full * width
            if rem:
                f.write(seq[full * width:].tobytes() + b"\n")
            fai.write(f"{name}\t{n}\t{offset}\t{width}\t{width + 1}\n")
            offset += n + full + (1 if rem else 0)
    return names


# ----------------------------------------------------------------------------------------------
# reads
# ----------------------------------------------------------------------------------------------
def _mutate(reads: np.ndarray, rng, sub: float):
    """In-place substitutions on a (n, L) code matrix."""
    m = rng.random(reads.shape) < sub
    k = int(m.sum())
    if k:
        reads[m] = (reads[m] + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3


def _indel_row(row: np.ndarray, src_tail: np.ndarray, rng, n_ins: int, n_del: int):
    """Apply n_ins insertions and n_del deletions to one read, keeping its length by
    consuming/dropping template bases at the 3' end (src_tail supplies extra template)."""
    L = len(row)
    seq = list(row) + list(src_tail)
    for _ in range(n_del):
        p = int(rng.integers(5, L - 5))
        del seq[p]
    for _ in range(n_ins):
        p = int(rng.integers(5, L - 5))
        seq.insert(p, int(rng.integers(0, 4)))
    return np.array(seq[:L], dtype=np.uint8)


def simulate_pairs(contigs, n_barcodes: int, pairs_per_bc: int, seed: int, *, r1_len=127, r2_len=150,
                   sub=0.005, indel=0.0, mol_len=(20000, 80000), pairs_per_mol=(3, 11),
                   insert=(250, 500)):
    """Simulate linked-read pairs.  Returns dict with code matrices r1 (n,r1_len), r2 (n,r2_len),
    bc_idx (n,), and truth columns (contig, pos, strand)."""
    rng = np.random.default_rng(seed)
    n_contigs = len(contigs)
    clen = np.array([len(c) for c in contigs])
    n = n_barcodes * pairs_per_bc
    bc_idx = np.repeat(np.arange(n_barcodes), pairs_per_bc)
    # molecules: each barcode draws molecules until pairs_per_bc pairs are placed
    mol_c = np.empty(n, dtype=np.int64)
    mol_s = np.empty(n, dtype=np.int64)
    mol_l = np.empty(n, dtype=np.int64)
    i = 0
    while i < n:
        b_end = (i // pairs_per_bc + 1) * pairs_per_bc
        k = min(int(rng.integers(pairs_per_mol[0], pairs_per_mol[1] + 1)), b_end - i)
        c = int(rng.integers(0, n_contigs))
        ml = int(rng.integers(mol_len[0], mol_len[1] + 1))
        ml = min(ml, clen[c] - 1)
        s = int(rng.integers(0, clen[c] - ml))
        mol_c[i:i + k] = c
        mol_s[i:i + k] = s
        mol_l[i:i + k] = ml
        i += k
    ins = rng.integers(insert[0], insert[1] + 1, size=n)
    tail = 8  # spare template bases for indel reads
    frag_off = (rng.random(n) * (mol_l - ins - tail)).astype(np.int64)
    frag_s = mol_s + np.maximum(frag_off, 0)
    strand = rng.random(n) < 0.5  # True: R1 on the reverse strand
    r1 = np.empty((n, r1_len + tail), dtype=np.uint8)
    r2 = np.empty((n, r2_len + tail), dtype=np.uint8)
    for c in range(n_contigs):
        sel = np.nonzero(mol_c == c)[0]
        if not len(sel):
            continue
        ref = contigs[c]
        fs = frag_s[sel][:, None]
        fe = (frag_s[sel] + ins[sel])[:, None]
        st = strand[sel]
        # forward-strand fragment: left read = ref[fs : fs+L], right read = revcomp(ref[fe-L : fe])
        ar1 = np.arange(r1_len + tail)[None, :]
        ar2 = np.arange(r2_len + tail)[None, :]
        hi = len(ref) - 1
        left1 = ref[np.clip(fs + ar1, 0, hi)]
        right1 = COMP[ref[np.clip(fe - 1 - ar1, 0, hi)]]
        left2 = ref[np.clip(fs + ar2, 0, hi)]
        right2 = COMP[ref[np.clip(fe - 1 - ar2, 0, hi)]]
        r1[sel] = np.where(st[:, None], right1, left1)
        r2[sel] = np.where(st[:, None], left2, right2)
    _mutate(r1, rng, sub)
    _mutate(r2, rng, sub)
    if indel > 0:
        for mat, L in ((r1, r1_len), (r2, r2_len)):
            n_ins = rng.binomial(L, indel, size=n)
            n_del = rng.binomial(L, indel, size=n)
            for idx in np.nonzero((n_ins + n_del) > 0)[0]:
                mat[idx, :L] = _indel_row(mat[idx, :L], mat[idx, L:], rng, int(n_ins[idx]), int(n_del[idx]))
    return dict(r1=r1[:, :r1_len].copy(), r2=r2[:, :r2_len].copy(), bc_idx=bc_idx, contig=mol_c,
                pos=frag_s, strand=strand, n_barcodes=n_barcodes, seed=seed)


def random_barcodes(n: int, length: int, rng) -> list[bytes]:
    seen = set()
    out = []
    while len(out) < n:
        b = ACGT[rng.integers(0, 4, size=length)].tobytes()
        if b not in seen and b != b"A" * length:
            seen.add(b)
            out.append(b)
    return out


def haplotag_barcodes(n: int, rng) -> list[bytes]:
    seen = set()
    out = []
    while len(out) < n:
        a, c, b, d = (int(x) for x in rng.integers(1, 97, size=4))
        s = f"A{a:02d}C{c:02d}B{b:02d}D{d:02d}".encode()
        if s not in seen:
            seen.add(s)
            out.append(s)
    return out


def _lines(sim, names, barcodes, shuffle_rng=None):
    r1 = ACGT[sim["r1"]]
    r2 = ACGT[sim["r2"]]
    n = r1.shape[0]
    s1 = r1.view(f"S{r1.shape[1]}").ravel()
    s2 = r2.view(f"S{r2.shape[1]}").ravel()
    q1 = b"I" * r1.shape[1]
    q2 = b"I" * r2.shape[1]
    order = np.arange(n)
    if shuffle_rng is not None:
        shuffle_rng.shuffle(order)
    return order, s1, s2, q1, q2


def write_bucket(path: str, sim, *, platform="10x", shuffle=True):
    """Write one preprocessed bucket file (one line per pair):
    ``BC @name read1 qual1 read2 qual2`` — fields as parsed by src/align.c:778-794."""
    rng = np.random.default_rng(sim["seed"] + 7919)
    nb = sim["n_barcodes"]
    if platform == "haplotag":
        bcs = haplotag_barcodes(nb, rng)
    else:
        bcs = random_barcodes(nb, {"10x": 16, "dbs": 20, "tellseq": 18}.get(platform, 16), rng)
    order, s1, s2, q1, q2 = _lines(sim, None, bcs, rng if shuffle else None)
    bc_idx = sim["bc_idx"]
    per_bc = np.zeros(nb, dtype=np.int64)
    serial = np.empty(len(bc_idx), dtype=np.int64)
    for i, b in enumerate(bc_idx):
        serial[i] = per_bc[b]
        per_bc[b] += 1
    with open(path, "wb") as f:
        out = []
        for i in order:
            b = bcs[bc_idx[i]]
            out.append(b + b" @r" + b + b"_" + str(serial[i]).encode() + b" " + s1[i] + b" " + q1 + b" " + s2[i] + b" " + q2 + b"\n")
        f.write(b"".join(out))
    return bcs


def write_raw_10x_fastq(path: str, whitelist_path: str, sim, *, n_whitelist=2000, err=0.03):
    """RAW 10x reads as `ema count` / `ema preproc` take them (interleaved FASTQ, pairs in random order): read 1 = the
    16-base barcode + 7 bases that preproc trims + the insert's first mate.  A fraction `err` of the pairs carries one
    substitution in the barcode (low quality there), a few an N or a barcode that is not on the list.  Also writes the
    whitelist (the used barcodes among `n_whitelist` random ones, shuffled)."""
    rng = np.random.default_rng(sim["seed"] + 15485863)
    nb = sim["n_barcodes"]
    bcs = random_barcodes(nb, 16, rng)
    wl = set(bcs)
    while len(wl) < max(n_whitelist, nb):
        wl.add(ACGT[rng.integers(0, 4, 16)].tobytes())
    wl = sorted(wl)
    with open(whitelist_path, "wb") as f:
        for i in rng.permutation(len(wl)):
            f.write(wl[i] + b"\n")
    order, s1, s2, q1, q2 = _lines(sim, None, bcs, rng)
    bc_idx = sim["bc_idx"]
    with open(path, "wb") as f:
        out = []
        for i in order:
            bc = bytearray(bcs[bc_idx[i]])
            qb = bytearray(b"I" * 16)
            u = rng.random()
            if u < err:
                k = int(rng.integers(0, 16)); bc[k] = b"ACGT"[(b"ACGT".index(bc[k]) + int(rng.integers(1, 4))) % 4]; qb[k] = ord("%")
            elif u < err + 0.005:
                k = int(rng.integers(0, 16)); bc[k] = ord("N"); qb[k] = ord("#")
            elif u < err + 0.01:
                bc = bytearray(ACGT[rng.integers(0, 4, 16)].tobytes())
            trim = ACGT[rng.integers(0, 4, 7)].tobytes()
            rid = b"@r" + str(int(i)).encode()
            out.append(rid + b" 1:N:0\n" + bytes(bc) + trim + s1[i] + b"\n+\n" + bytes(qb) + b"IIIIIII" + q1 + b"\n" +
                       rid + b" 2:N:0\n" + s2[i] + b"\n+\n" + q2 + b"\n")
        f.write(b"".join(out))
    return bcs


def write_interleaved_fastq(path: str, sim, *, platform="tru"):
    """Barcode-sorted interleaved FASTQ for `-1` mode.
    tru:     id ``@<int>_<name>`` (barcode = atoi, src/techs.c:57-61)
    cpt:     id ``@name:bc<int>`` (barcode = atoi after the last ':' + 2 characters, src/techs.c:63-69)
    tellseq: id ``@name:<18-mer>`` (src/techs.c:33-55)
    10x:     id ``@name:<16-mer>`` (src/techs.c:19-30)"""
    rng = np.random.default_rng(sim["seed"] + 104729)
    nb = sim["n_barcodes"]
    if platform in ("tru", "cpt"):
        bcs = [str(i + 1).encode() for i in range(nb)]
    else:
        bcs = random_barcodes(nb, 18 if platform == "tellseq" else 16, rng)
    order, s1, s2, q1, q2 = _lines(sim, None, bcs, None)
    bc_idx = sim["bc_idx"]
    with open(path, "wb") as f:
        out = []
        for i in order:  # already grouped by barcode
            b = bcs[bc_idx[i]]
            if platform == "tru":
                rid = b"@" + b + b"_r" + str(i).encode()
            elif platform == "cpt":   # extract_bc_cptseq (src/techs.c:63-69): atoi of what follows ":xx" at the end of the id
                rid = b"@r" + str(i).encode() + b":bc" + b
            else:
                rid = b"@r" + str(i).encode() + b":" + b
            out.append(rid + b"\n" + s1[i] + b"\n+\n" + q1 + b"\n" + rid + b"\n" + s2[i] + b"\n+\n" + q2 + b"\n")
        f.write(b"".join(out))
    return bcs


# ----------------------------------------------------------------------------------------------
# named configurations (BASELINE.json "configs")
# ----------------------------------------------------------------------------------------------
CONFIGS = {
    # name: (n_contigs, contig_len, ref_seed, dup_every, n_barcodes, pairs_per_bc, indel)
    "tiny":     (2, 30_000, 7, 0, 8, 40, 0.0),
    "tiny_rep": (2, 60_000, 8, 15_000, 12, 40, 0.001),
    "c1":       (5, 1_000_000, 42, 0, 200, 50, 0.0),
    "c1_rep":   (5, 1_000_000, 43, 25_000, 200, 50, 0.001),
    "c2":       (10, 10_000_000, 44, 25_000, 5000, 200, 0.001),
    "c2_small": (10, 10_000_000, 44, 25_000, 500, 200, 0.001),
    # references whose BWT does not fit the 126 MB L2: 1 Gbp, and the hg38-sized 3.1 Gbp of BASELINE configs[2]
    # (24 contigs; 2 x l_pac > 2^32, so the u64 suffix-array paths are the ones that run)
    "g1":       (8, 125_000_000, 46, 25_000, 5000, 200, 0.001),
    "c3":       (24, 129_166_664, 45, 25_000, 100_000, 200, 0.001),
}


def build_config(name: str, root: str, bwa_bin: str | None = None, platform: str = "10x", indexer=None, n_barcodes: int | None = None,
                 tag: str = ""):
    """Materialise a named config under ``root/name``; returns dict of paths.
    The FM index is built with the reference's own ``bwa index`` when ``bwa_bin`` is given and the index is absent, or by
    ``indexer(fasta_path)`` (ema_b200.index_build: the same files, on the GPU) when that is given instead.
    ``n_barcodes`` overrides the config's bucket size (the platform sweep uses small buckets on a large reference)."""
    import subprocess
    n_contigs, clen, rseed, dup, nbc, ppb, indel = CONFIGS[name]
    if n_barcodes:
        nbc = n_barcodes
    d = os.path.join(root, name + tag)
    os.makedirs(d, exist_ok=True)
    fa = os.path.join(d, "ref.fa")
    bucket = os.path.join(d, f"ema-bin-000.{platform}")
    contigs = None
    if not os.path.exists(fa + ".fai"):
        contigs = make_reference(n_contigs, clen, rseed, dup)
        write_fasta(fa, contigs)
    if not os.path.exists(bucket):
        if contigs is None:
            contigs = make_reference(n_contigs, clen, rseed, dup)
        sim = simulate_pairs(contigs, nbc, ppb, rseed + 1000, indel=indel)
        if platform in ("10x", "haplotag", "dbs"):
            write_bucket(bucket, sim, platform=platform)
        else:
            write_interleaved_fastq(bucket, sim, platform=platform)
    if not os.path.exists(fa + ".sa"):
        if indexer is not None:
            indexer(fa)
        elif bwa_bin:
            subprocess.run([bwa_bin, "index", fa], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return dict(dir=d, fasta=fa, bucket=bucket, n_pairs=nbc * ppb)


# ----------------------------------------------------------------------------------------------
# ksw_extend2 microbenchmark tasks (config 5)
# ----------------------------------------------------------------------------------------------
def extend_tasks(n: int, qlen: int, seed: int = 88172645463325252 & 0xFFFFFFFF, pad: int = 100,
                 sub=0.01, ins=0.001, dele=0.001):
    """qlen-bp queries; target = query with sub/ins/del noise, padded with random bases to
    tlen = qlen + pad.  Returns (q (n,qlen) u8, t (n,qlen+pad) u8) of codes 0..3."""
    rng = np.random.default_rng(seed)
    q = rng.integers(0, 4, size=(n, qlen), dtype=np.uint8)
    tlen = qlen + pad
    t = rng.integers(0, 4, size=(n, tlen), dtype=np.uint8)
    t[:, :qlen] = q
    _mutate(t[:, :qlen], rng, sub)
    n_ev = rng.binomial(qlen, ins + dele, size=n)
    for idx in np.nonzero(n_ev)[0]:
        row = list(t[idx])
        for _ in range(int(n_ev[idx])):
            p = int(rng.integers(1, qlen - 1))
            if rng.random() < 0.5:
                del row[p]
                row.append(int(rng.integers(0, 4)))
            else:
                row.insert(p, int(rng.integers(0, 4)))
                row.pop()
        t[idx] = row
    return q, t
