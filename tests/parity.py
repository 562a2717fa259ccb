"""Shared parity helpers: the stated tolerance contract of SURVEY.md §8(d), in one place.

End to end, the bf16 product is compared with the fp32 oracle *relative to what PyTorch's own bf16 autocast does on
the same oracle graph*:
    logits   max_abs <= 3e-2 * max|logit|
    loss     |delta| <= 2e-3
    per-parameter gradient cosine >= autocast cosine - 0.02
    per-parameter gradient rel_L2 <= 1.25 x autocast rel_L2  (+ REL_FLOOR, see below)
REL_FLOOR = 2e-3 covers the parameters whose autocast error is itself ~0 because torch keeps them in fp32 (biases of
fp32 linears): 1.25 x 0 would demand a bit-exact fp32 sum order. Parameters whose true gradient is zero (a conv bias
in front of a train-mode BatchNorm) are skipped, as is any other quantity with |ref| < 1e-6.
"""
import torch
import torch.nn.functional as F

LOGIT_TOL = 3e-2
LOSS_TOL = 2e-3
COS_MARGIN = 0.02
REL_FACTOR = 1.25
REL_FLOOR = 2e-3


def cos(a, b):
    return float(F.cosine_similarity(a.double().flatten(), b.double().flatten(), dim=0))


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def to_cuda(p):
    return {k: v.cuda() for k, v in p.items()}


def oracle_fp32_and_autocast(O, kind, p, inputs, labels, **kw):
    """(logits, loss, grads, new_buffers) of the fp32 oracle on the GPU, and the same graph under bf16 autocast."""
    pg = to_cuda(p)
    ins = tuple(t.cuda() if t is not None else None for t in inputs)
    lab = labels.cuda()
    ref = O.loss_and_grads(kind, pg, ins, lab, training=True, **kw)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ac = O.loss_and_grads(kind, pg, ins, lab, training=True, **kw)
    return ref, ac


def assert_logits_loss(logits, loss, ref_logits, ref_loss):
    lmax = float(ref_logits.abs().max())
    lerr = float((logits.detach().float() - ref_logits).abs().max())
    assert lerr <= LOGIT_TOL * lmax + 1e-6, ("logits", lerr, lmax)
    assert abs(float(loss) - float(ref_loss)) <= LOSS_TOL, ("loss", float(loss), float(ref_loss))


def assert_grad(name, g, g_ref, g_ac, report=None):
    """One parameter's gradient against the contract; returns False when the quantity is (analytically) zero."""
    if float(g_ref.abs().max()) < 1e-6:
        return False
    c_o, c_a = cos(g, g_ref), cos(g_ac, g_ref)
    r_o, r_a = rel(g, g_ref), rel(g_ac, g_ref)
    if report is not None:
        report.append((c_o - c_a, name, c_o, c_a, r_o, r_a))
    assert c_o >= c_a - COS_MARGIN, (name, "cosine", c_o, "autocast", c_a)
    assert r_o <= REL_FACTOR * r_a + REL_FLOOR, (name, "rel_L2", r_o, "autocast", r_a)
    return True


def print_worst(report, k=10):
    for dlt, name, c_o, c_a, r_o, r_a in sorted(report)[:k]:
        print(f"  {name:44s} cos ours {c_o:.4f} autocast {c_a:.4f} | rel_l2 ours {r_o:.4f} autocast {r_a:.4f}")
