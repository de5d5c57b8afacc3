"""TEST INFRASTRUCTURE -- CPU restatement of the selector filtering and alpha-combination that plonky2
wraps around Gate::eval_unfiltered_base_batch.  Only tests/, smoke() and bench.py's
cpu_baseline leg may import this.

Source status: the code restated here lives in the `plonky2` crate, a git dependency of the reference
without a pinned revision (Cargo.toml:17) and NOT under /root/reference -- PARITY UNPINNED against the
dependency itself.  Restated from the published crate (plonky2/src/gates/gate.rs `eval_filtered_base_batch`,
`compute_filter`; plonky2/src/plonk/vanishing_poly.rs `evaluate_gate_constraints_base_batch`,
`eval_vanishing_poly_base_batch`; plonky2/src/plonk/plonk_common.rs `reduce_with_powers_multi`;
plonky2/src/gates/selectors.rs `UNUSED_SELECTOR = u32::MAX`).  The reference reaches it through every
`builder.build::<C>()` / `data.prove(pw)` of its tests (e.g. gates/mul_nonnative.rs:558-578 exercise the
in-tree gates' eval functions that this step consumes).
"""
GL_P = 0xFFFFFFFF00000001
UNUSED_SELECTOR = (1 << 32) - 1


def compute_filter(gate_index, group_range, s, many_selectors):
    """prod_{i in group_range, i != gate_index} (i - s) * (UNUSED_SELECTOR - s if many_selectors)."""
    f = 1
    for i in range(group_range[0], group_range[1]):
        if i != gate_index:
            f = f * ((i - s) % GL_P) % GL_P
    if many_selectors:
        f = f * ((UNUSED_SELECTOR - s) % GL_P) % GL_P
    return f


def eval_filtered(constraints, gate_index, group_range, s, many_selectors):
    """Gate::eval_filtered: the unfiltered constraint vector of one row times the row's filter."""
    f = compute_filter(gate_index, group_range, s, many_selectors)
    return [c * f % GL_P for c in constraints]


def evaluate_gate_constraints(per_gate_filtered):
    """evaluate_gate_constraints: position-wise sum of the gates' filtered constraint vectors."""
    n = max((len(v) for v in per_gate_filtered), default=0)
    out = [0] * n
    for v in per_gate_filtered:
        for j, c in enumerate(v):
            out[j] = (out[j] + c) % GL_P
    return out


def reduce_with_powers_multi(terms, alphas):
    """For every alpha: sum_t terms[t] * alpha^t (Horner from the last term, as upstream does)."""
    res = []
    for a in alphas:
        acc = 0
        for t in reversed(terms):
            acc = (acc * a + t) % GL_P
        res.append(acc)
    return res


def gate_terms_combined(constraints, gate_index, group_range, s, many_selectors, alphas, first_power=0):
    """What one p2e_eval_filtered_combined call adds to out[a][row]."""
    f = eval_filtered(constraints, gate_index, group_range, s, many_selectors)
    return reduce_with_powers_multi([0] * first_power + f, alphas)
