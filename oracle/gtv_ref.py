"""The GTV segmentation spec (Greta.pdf p.48, Image290) composed from oracle operators.

TEST INFRASTRUCTURE ONLY (see vx_oracle.c).  Used by tests/, smoke() and bench.py's CPU
baseline to produce the reference result and CPU timing of the whole spec for ONE volume.
Derived operators per SURVEY.md section 8a row A0; constants per Q5.
"""
import numpy as np

import oracle as orc


def gtv_oracle(o, flair, conn=26, spacing=None, stages=None):
    """flair: f32 [nz][ny][nx].  Returns every named intermediate of the spec."""
    r = {}
    inten = np.ascontiguousarray(flair, dtype=np.float32)
    r["background"] = o.touch(o.cmp_scalar(inten, "<", 0.1), o.border(inten), conn)
    r["brain"] = o.not_(r["background"])
    r["normFlair"] = o.percentiles(inten, r["brain"], 0.0)
    r["hyperIntense"] = o.flt(5.0, o.cmp_scalar(r["normFlair"], ">", 0.95), spacing)
    r["veryIntense"] = o.flt(2.0, o.cmp_scalar(r["normFlair"], ">", 0.86), spacing)
    r["growTum"] = o.grow(r["hyperIntense"], r["veryIntense"], conn)
    if stages == "pre_similarity":
        return r
    m, M = o.min_(inten), o.max_(inten)
    r["tumSim"] = o.cross_correlation(5.0, inten, inten, r["growTum"], m, M, 100, spacing)
    r["tumStatCC"] = o.flt(2.0, o.cmp_scalar(r["tumSim"], ">", 0.6), spacing)
    r["gtv"] = o.grow(r["growTum"], r["tumStatCC"], conn)
    return r


def texture_oracle(o, mods, conn=26, spacing=None):
    """BASELINE config 4 (voxlogica_b200.imgql.TEXTURE_SPEC) composed from oracle operators.
    mods: {"flair","t1","t1c","t2"} -> f32 [nz][ny][nx]."""
    r = gtv_oracle(o, mods["flair"], conn, spacing, stages="pre_similarity")
    sims = []
    for name in ("flair", "t1", "t1c", "t2"):
        a = np.ascontiguousarray(mods[name], dtype=np.float32)
        sims.append(o.cross_correlation(5.0, a, a, r["growTum"], o.min_(a), o.max_(a), 100, spacing))
    r["simFlair"], r["simT1"], r["simT1c"], r["simT2"] = sims
    acc = o.arith(o.arith(o.arith(sims[0], "+", sims[1]), "+", sims[2]), "+", sims[3])
    r["texture"] = o.arith_scalar(acc, "/", 4.0)
    r["similar"] = o.flt(2.0, o.and_(o.cmp_scalar(r["texture"], ">", 0.6), r["brain"]), spacing)
    r["tumour"] = o.grow(r["growTum"], r["similar"], conn)
    return r
