"""Helpers shared by CPU and GPU tests (not collected by pytest)."""
import hashlib
import json
import os

import numpy as np
import torch

RTOL, ATOL = 1e-5, 1e-6           # the north_star fp32 band: |a-b| <= 1e-6 + 1e-5*|b|
U32 = 2.0 ** -24                  # unit roundoff of fp32
# worst-case relative error of one product in the tcgen05 split-tf32 GEMM of pair_conv.cu (the derivation is the comment above
# pc_rna; dw_tc.cu truncates both operands' parts: 3 * 2^-20, passed as mma_bound by its tests): streamed operand a = trunc_tf32(a) + rn_tf32(rest) + r_a, |r_a| <= 2^-21 |a|; resident / rewritten
# operand w = rn_tf32(w) + rn_tf32(rest) + r_w, |r_w| <= 2^-22 |w|; the dropped lo*lo product <= 2^-21 |a*w|.
MMA4 = 1.25 * 2.0 ** -20


def sha(t) -> str:
    a = np.ascontiguousarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t)
    return hashlib.sha256(a.astype("<i8").tobytes()).hexdigest()


def fb_split(fb, s):
    """(pos_edge, pred_edge, pos1, y) of split s rebuilt from the golden edge lists the
    same way BaseGraph.preprocess does (TwoWL/operators/datasets.py:44-92)."""
    ep, en = fb["edge_pos"].astype(np.int64), fb["edge_neg"].astype(np.int64)
    npos, nneg = fb["num_pos"], fb["num_neg"]
    ei = ep[:, :npos[0]] if s < 2 else ep[:, :npos[0] + npos[1]]
    pos_e = [ep[:, :npos[0]], ep[:, npos[0]:npos[0] + npos[1]], ep[:, ep.shape[1] - npos[2]:]]
    neg_e = [en[:, :nneg[0]], en[:, nneg[0]:nneg[0] + nneg[1]], en[:, en.shape[1] - nneg[2]:]]
    pred = neg_e[0] if s == 0 else np.concatenate([pos_e[s], neg_e[s]], axis=1)
    pos1 = np.concatenate([ei.T, pred.T], axis=0)
    return ei, pred, pos1


def state_dict_from(fb, prefix="sd/"):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in fb.items() if k.startswith(prefix)}


def r32(t):
    """An fp64 tensor whose values are exactly representable in fp32: kernel inputs made this way carry no input-rounding
    error, so the fp64 reference and the fp32 kernel start from the same numbers."""
    return t.float().double()


def assert_close(a, b, rtol=RTOL, atol=ATOL, what="", absum=None, nterms=0, mma=False, mma_bound=None):
    """The north_star fp32 band |a-b| <= atol + rtol*|b| (SURVEY.md 7 hard part 3).

    For an element that is a SUM (a segmented reduction, a GEMM), relative error against the result is not a property of any
    fp32 evaluation - the terms cancel - so such tests also pass `absum` (the same sum over the ABSOLUTE values of the terms,
    computed next to the reference) and `nterms` (the number of terms of the longest sum, plus the roundings inside one term):
    the band is widened by the a-priori forward error bound of fp32 summation IN ANY ORDER (Higham, Accuracy and Stability of
    Numerical Algorithms, 2nd ed., eq. 4.4: |fl(sum) - sum| <= (n-1) u sum|x_i| + O(u^2)), `nterms * 2^-24 * absum`, and, for
    the tensor-core GEMMs (`mma=True`), by the split-tf32 product bound `1.25 * 2^-20 * absum` derived at MMA4 above. Nothing else is
    allowed: no floor relative to the tensor's largest element."""
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs()
    bound = atol + rtol * b.abs()
    if absum is not None:
        bound = bound + (nterms * U32 + ((mma_bound or MMA4) if mma else 0.0)) * torch.as_tensor(absum).detach().cpu().double()
    bad = err > bound
    assert not bool(bad.any()), (f"{what}: {int(bad.sum())}/{bad.numel()} outside tolerance; "
                                 f"max err {float(err.max()):.3e}, max err/bound {float((err / bound).max()):.3f}, "
                                 f"max |ref| {float(b.abs().max()):.3e}")


# ------------------------------------------------------------------------------ parity ledger
#
# End-to-end parity checks (logits, loss, gradients of a whole train step) record HOW every element passed, so that the number
# of elements that needed anything beyond the north_star band is a reported, asserted quantity instead of a silent relaxation.

class _Ledger:
    def __init__(self):
        self.rows = []

    def add(self, **kw):
        self.rows.append(kw)

    def totals(self):
        keys = ("numel", "band_fp32ref", "band_fp64_only", "ref_error_clause", "scale_floor")
        return {k: int(sum(r.get(k, 0) for r in self.rows)) for k in keys}

    def dump(self, path):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump({"band": "|a-b| <= 1e-6 + 1e-5*|b|", "totals": self.totals(), "tensors": self.rows}, f, indent=1)


LEDGER = _Ledger()
REF_FACTOR = 4.0


def parity(got, ref32, ref64, what, scale_floor=0.0, allow_relaxed=None):
    """One tensor of an end-to-end result against the fp32 reference (the reference implementation's / the oracle's own output;
    None when only an fp64 evaluation exists) and the fp64 evaluation of the same formulas. Every element is classified by the
    FIRST clause it satisfies:
      band_fp32ref      within the north_star band of the fp32 reference;
      band_fp64_only    not that, but within the band of the fp64 evaluation (the SURVEY 4 tie-breaker: the fp32 reference is
                        itself only an approximation of the formulas);
      ref_error_clause  neither, but no farther from fp64 than REF_FACTOR = 4 x the fp32 reference's OWN worst distance from fp64
                        on this tensor: where the reference's fp32 arithmetic itself leaves the band, the band is a property of
                        the conditioning, not of an implementation, and two fp32 evaluations of the same formulas in different
                        summation orders draw their worst element from the same error distribution (tools/diag_parity.py:
                        every code path here, the exact-fp32 SIMT one included, lands within 1-4 x of the oracle's own worst
                        error; the ledger counts how many of these elements lie beyond 1 x and beyond 2 x);
      scale_floor       neither, but within `scale_floor` x the tensor's largest fp64 magnitude (only where the caller passes
                        one and says why).
    An element in none of the classes fails the test, and so do more than `allow_relaxed` elements in the last two classes
    (default: 5 % of the tensor, at least 4 elements; the full-size tests state their own); the counts of every class - and how
    many of the ref_error_clause elements lie beyond 1 x and 2 x the reference's error - go to the parity ledger
    (gpurun_out/parity_report.json, printed at the end of the session)."""
    g = got.detach().cpu().double()
    r64 = torch.as_tensor(ref64).detach().cpu().double()
    assert g.shape == r64.shape, f"{what}: shape {tuple(g.shape)} vs {tuple(r64.shape)}"
    err64 = (g - r64).abs()
    in64 = err64 <= ATOL + RTOL * r64.abs()
    if ref32 is not None:
        r32_ = torch.as_tensor(ref32).detach().cpu().double()
        in32 = (g - r32_).abs() <= ATOL + RTOL * r32_.abs()
        ref_err = float((r32_ - r64).abs().max())
    else:
        in32 = torch.zeros_like(in64)
        ref_err = 0.0
    rest = ~(in32 | in64)
    by_ref = rest & (err64 <= REF_FACTOR * ref_err)
    beyond_1x = int((by_ref & (err64 > ref_err)).sum())
    beyond_2x = int((by_ref & (err64 > 2 * ref_err)).sum())
    rest = rest & ~by_ref
    floor = scale_floor * float(r64.abs().max())
    by_floor = rest & (err64 <= floor + RTOL * r64.abs())
    rest = rest & ~by_floor
    test = os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0].split("::")[-1]
    row = dict(test=test, what=what, numel=g.numel(), band_fp32ref=int(in32.sum()), band_fp64_only=int((in64 & ~in32).sum()),
               ref_error_clause=int(by_ref.sum()), ref_error_clause_beyond_1x=beyond_1x, ref_error_clause_beyond_2x=beyond_2x,
               scale_floor=int(by_floor.sum()), failed=int(rest.sum()),
               max_err_vs_fp64=float(err64.max()) if g.numel() else 0.0, fp32ref_max_err_vs_fp64=ref_err,
               max_abs_ref=float(r64.abs().max()) if g.numel() else 0.0)
    LEDGER.add(**row)
    if os.environ.get("TWOWL_PARITY_REPORT_ONLY"):       # survey mode: collect the ledger without enforcing anything
        return row
    assert not bool(rest.any()), f"{what}: {row}"
    relaxed = row["ref_error_clause"] + row["scale_floor"]
    if allow_relaxed is None:
        allow_relaxed = max(4, -(-g.numel() // 20))
    assert relaxed <= allow_relaxed, f"{what}: {relaxed} elements needed a relaxation (allowed {allow_relaxed}): {row}"
    return row
