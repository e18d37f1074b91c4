"""Regenerates tests/golden/* from the UNMODIFIED reference (oracle/_ref, built from /root/reference
by oracle/Makefile).  Run from the repo root:  python tests/golden/make_golden.py

Outputs
  tiny_rep/ref.fa{,.fai,.amb,.ann,.bwt,.pac,.sa}  120 kbp reference with planted repeats + its `bwa index`
  tiny_rep/ema-bin-000.10x                          480 read pairs in 12 barcodes (bucket format)
  tiny_rep/ref.sam                                  `ema align -s ... -p 10x -t 1` output of the reference
  sw_golden.npz     ksw_extend2 / ksw_global2 / ksw_align2 inputs + reference outputs
  fm_golden.npz     mem_collect_intv intervals and bwt_sa values of the reference on tiny_rep reads
  cand_golden.npz   per pair of the tiny_rep bucket (file order): the candidate SAMRecords append_alignments builds
                    (chrom,pos,rev,mate,mapq,score_mapq,clip,clip_edit_dist,NM,n_cigar,unique ; EM score ; CIGAR)
"""
import ctypes as C
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import helpers  # noqa: E402
from helpers import _p  # noqa: E402
from tools import synth  # noqa: E402


def main():
    R = helpers.ref()
    tmp = "/tmp/emab_golden"
    shutil.rmtree(tmp, ignore_errors=True)
    p = synth.build_config("tiny_rep", tmp, helpers.ref_bin("bwa"))
    dst = os.path.join(HERE, "tiny_rep")
    os.makedirs(dst, exist_ok=True)
    for ext in ("", ".fai", ".amb", ".ann", ".bwt", ".pac", ".sa"):
        shutil.copy(p["fasta"] + ext, os.path.join(dst, "ref.fa" + ext))
    shutil.copy(p["bucket"], os.path.join(dst, "ema-bin-000.10x"))
    subprocess.run([helpers.ref_bin("ema"), "align", "-s", os.path.join(dst, "ema-bin-000.10x"), "-r", os.path.join(dst, "ref.fa"),
                    "-p", "10x", "-t", "1", "-o", os.path.join(dst, "ref.sam")], check=True, cwd=dst, stderr=subprocess.DEVNULL)
    # SW vectors
    qs, ts, h0 = helpers.random_extend_tasks(600, 20261017)
    ext, _ = helpers.sw_extend(R, "ref", qs, ts, h0)
    ws = np.array([max(int(w), abs(len(q) - len(t)) + 3) for w, q, t in
                   zip(np.random.default_rng(1).integers(0, 60, size=len(qs)), qs, ts)], dtype=np.int32)
    glo, gcig, _ = helpers.sw_global(R, "ref", qs, ts, ws, max_cigar=512)
    lq = [q for q in qs if len(q) >= 20]
    lt = [np.concatenate([t, q[::-1], t[: len(t) // 2]])[:1000] for q, t in zip(qs, ts) if len(q) >= 20]
    loc, _ = helpers.sw_local(R, "ref", lq, lt)
    q, qo = helpers.pack(qs)
    t, to = helpers.pack(ts)
    lqf, lqo = helpers.pack(lq)
    ltf, lto = helpers.pack(lt)
    np.savez_compressed(os.path.join(HERE, "sw_golden.npz"), q=q, qo=qo, t=t, to=to, h0=h0, ext=ext, ws=ws, glo=glo, gcig=gcig,
                        lq=lqf, lqo=lqo, lt=ltf, lto=lto, loc=loc)
    # FM-index vectors
    ri = R.ref_idx_load(os.path.join(dst, "ref.fa").encode())
    reads = []
    for ln in open(os.path.join(dst, "ema-bin-000.10x")).read().split("\n")[:150]:
        if ln:
            f = ln.split(" ")
            reads += [helpers.nt4(f[2]), helpers.nt4(f[4])]
    rng = np.random.default_rng(9)
    for r in reads[::7]:
        r[rng.integers(0, len(r))] = 4
    ivs, cnt = [], []
    for s in reads:
        ob = np.zeros((256, 4), np.int64)
        nb = R.ref_collect_intv(C.c_void_p(ri), len(s), _p(s, C.c_uint8), _p(ob, C.c_int64), 256)
        ivs.append(ob[:nb].copy())
        cnt.append(nb)
    info = np.zeros(12, np.int64)
    R.ref_idx_info(C.c_void_p(ri), _p(info, C.c_int64))
    ks = rng.integers(0, info[3] + 1, size=4000).astype(np.int64)
    ks[:3] = [0, info[3], info[2]]
    sa = np.zeros_like(ks)
    R.ref_sa_batch(C.c_void_p(ri), len(ks), _p(ks, C.c_int64), _p(sa, C.c_int64))
    rf, ro = helpers.pack(reads)
    np.savez_compressed(os.path.join(HERE, "fm_golden.npz"), reads=rf, roff=ro, intv=np.concatenate(ivs), n_intv=np.array(cnt),
                        ks=ks, sa=sa, info=info)
    # candidate alignments per pair exactly as append_alignments (src/align.c:986) produces them
    R.ref_ema_init.argtypes = [C.c_char_p, C.c_char_p]
    assert R.ref_ema_init(os.path.join(dst, "ref.fa").encode(), b"10x") == 0
    ints_all, sc_all, cg_all, cnt = [], [], [], []
    for ln in open(os.path.join(dst, "ema-bin-000.10x")).read().split("\n"):
        if not ln:
            continue
        f = ln.split(" ")
        ints = np.zeros((4096, 11), np.int64)
        sc = np.zeros(4096, np.float64)
        cg = np.zeros((4096, 64), np.uint32)
        n = R.ref_ema_candidates(f[1][1:].encode(), f[2].encode(), f[3].encode(), f[4].encode(), f[5].encode(),
                                 _p(ints, C.c_int64), _p(sc, C.c_double), _p(cg, C.c_uint32), 64, 4096)
        ints_all.append(ints[:n].copy()); sc_all.append(sc[:n].copy()); cg_all.append(cg[:n].copy()); cnt.append(n)
    np.savez_compressed(os.path.join(HERE, "cand_golden.npz"), ints=np.concatenate(ints_all), score=np.concatenate(sc_all),
                        cigar=np.concatenate(cg_all), n=np.array(cnt))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
