"""Stage-by-stage comparison of two engines that speak the dvp_mvs.h ABI (product vs reference oracle).

Used by tests/ (-m gpu) and tools/parity_report.py.  Protocol: the reference engine runs
the kernel sequence; before every stage the product engine is loaded with the reference's pre-stage device
state, runs the same stage, and the post-stage buffers are compared.  Chaotic divergence therefore never
accumulates: every stage is judged from identical input state.
"""
from __future__ import annotations

import numpy as np

from ._lib import Engine, STAGES

STATE_BUFS = ("planes", "costs", "selected", "weak", "radius", "view_weight", "rand", "fit_planes", "edge_neigh",
              "nearest_strong", "weak_reliable")
WEAK_BUFS = ("neighbours", "label_boundary", "complex", "candidate")

# which buffers each stage writes (what gets compared after it)
STAGE_OUTPUTS = {
    "K1_INIT_RANDOM_STATES": ("rand",),
    "K2_GEN_EDGE_INFORM": ("edge_neigh", "weak", "complex", "label_boundary"),
    "K3_FIND_NEAREST_STRONG": ("nearest_strong",),
    "K4_GEN_NEIGHBOURS": ("neighbours", "weak_reliable", "rand"),
    "K5_NEIGHBOUR_UPDATE": ("weak",),
    "K6_RANDOM_INITIALIZATION": ("planes", "costs", "selected", "rand"),
    "K7_BLACK_STRONG": ("planes", "costs", "selected", "view_weight", "rand"),
    "K8_RED_STRONG": ("planes", "costs", "selected", "view_weight", "rand"),
    "K9_RANSAC_FIT_PLANE": ("fit_planes", "radius", "rand"),
    "K10_BLACK_WEAK": ("planes", "costs", "selected", "view_weight", "rand", "radius"),
    "K11_RED_WEAK": ("planes", "costs", "selected", "view_weight", "rand", "radius"),
    "K12_DEPTH_NORMAL": ("planes",),
    "K13_BLACK_FILTER": ("planes",),
    "K14_RED_FILTER": ("planes",),
    "K15_DEPTH_TO_WEAK": ("weak", "radius"),
    "K16_LOCAL_REFINE": ("planes",),
}


def sequence(max_iterations: int):
    seq = [(s, 0) for s in STAGES[:6]]
    for it in range(max_iterations):
        seq += [(s, it) for s in STAGES[6:11]]
    seq += [(s, 0) for s in STAGES[11:]]
    return seq


def compare(name: str, a: np.ndarray, b: np.ndarray, rtol: float = 1e-4, mask: np.ndarray | None = None) -> dict:
    """Per-pixel comparison. Float buffers: |a-b| <= rtol*max(|a|,|b|) (NaN == NaN); others: exact."""
    assert a.shape == b.shape, (name, a.shape, b.shape)
    if a.dtype.kind == "f":
        both_nan = np.isnan(a) & np.isnan(b)
        with np.errstate(invalid="ignore"):
            close = np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b))
        ok = close | both_nan | (a == b)
        exact = (a == b) | both_nan
    else:
        ok = a == b
        exact = ok
    # reduce trailing component axes to one verdict per item (pixel, or WEAK-pixel slot)
    lead = 1 if name in ("neighbours", "label_boundary", "complex") else 2
    item_shape = a.shape[:lead]
    okp = ok.reshape(item_shape + (-1,)).all(-1)
    exp = exact.reshape(item_shape + (-1,)).all(-1)
    if mask is not None and okp.shape == mask.shape:
        okp = okp | ~mask
        exp = exp | ~mask
        n = int(mask.sum())
    else:
        n = okp.size
    bad = int((~okp).sum())
    return dict(buffer=name, pixels=n, mismatched=bad, frac=bad / max(n, 1), not_bit_exact=int((~exp).sum()),
                first_bad=[int(v) for v in np.argwhere(~okp)[0]] if bad else None)


def step_compare(ref: Engine, prod: Engine, max_iterations: int, stages=None, mask_fn=None, log=None):
    """Run `ref` through the sequence; before each stage copy its state into `prod`, run, compare."""
    results = []
    have_weak = ref.weak_count() > 0
    bufs = STATE_BUFS + (WEAK_BUFS if have_weak else ())
    for stage, it in sequence(max_iterations):
        pre = {n: ref.get(n) for n in bufs}
        ref.run_stage(stage, it)
        if stages is not None and stage not in stages:
            continue
        for n, arr in pre.items():
            prod.set(n, arr)
        try:
            prod.run_stage(stage, it)
        except Exception as e:  # unsupported stage etc.
            results.append(dict(stage=stage, iter=it, error=str(e)))
            if log:
                log(f"{stage}[{it}] ERROR {e}")
            continue
        for n in STAGE_OUTPUTS[stage]:
            if n in WEAK_BUFS and not have_weak:
                continue
            a, b = ref.get(n), prod.get(n)
            mask = mask_fn(stage, n, pre) if mask_fn else None
            r = compare(n, a, b, mask=mask)
            r.update(stage=stage, iter=it)
            results.append(r)
            if log:
                log(f"{stage}[{it}] {n:12s} mismatched {r['mismatched']:8d}/{r['pixels']} ({100*r['frac']:.4f}%) "
                    f"not-bit-exact {r['not_bit_exact']} first {r['first_bad']}")
    return results


def lockstep_compare(ref: Engine, prod: Engine, max_iterations: int, racy=(), resync_extra=(), stages=None, log=None, before_stage=None, keep=None):
    """Same protocol as step_compare for large images: both engines start from the same upload and advance together;
    after each stage only the buffers that stage WRITES are read back (the rest of the state is identical by induction),
    compared, and — where they differ at all — overwritten in `prod` with the reference's, so a stage is always judged
    from the reference's pre-stage state.  `racy`: stages whose differences are expected (they are re-synchronised like
    any other).  `resync_extra`: {stage: (buffers,)} copied from ref to prod after a stage although they are not
    compared (K2's `candidate`, which holds uninitialised data in the reference where no pixel sees a view, B17).
    `before_stage(stage, it)` is called ahead of every launch; `keep[(stage, it)]` (a dict) receives both engines' outputs
    of that launch under "_ref_out" / "_prod_out"."""
    results = []
    have_weak = ref.weak_count() > 0
    for stage, it in sequence(max_iterations):
        if stages is not None and stage not in stages:
            continue
        if before_stage:
            before_stage(stage, it)
        ref.run_stage(stage, it)
        prod.run_stage(stage, it)
        kept = keep.get((stage, it)) if keep else None       # caller wants both engines' outputs of this launch
        if kept is not None:
            kept["_ref_out"], kept["_prod_out"], kept["_rest"] = {}, {}, None
        for n in STAGE_OUTPUTS[stage]:
            if n in WEAK_BUFS and not have_weak:
                continue
            a, b = ref.get(n), prod.get(n)
            if kept is not None:
                kept["_ref_out"][n], kept["_prod_out"][n] = a, b
            r = compare(n, a, b)
            r.update(stage=stage, iter=it, racy=stage in racy)
            results.append(r)
            if log:
                log(f"{stage}[{it}] {n:12s} mismatched {r['mismatched']:8d}/{r['pixels']} ({100*r['frac']:.4f}%) not-bit-exact {r['not_bit_exact']}")
            if r["not_bit_exact"]:
                prod.set(n, a)
        for n in dict(resync_extra).get(stage, ()):
            prod.set(n, ref.get(n))
    return results
