"""Tolerance policy shared by the parity tests.

Same contract as the reference's ``mojo_opset/utils/acc.py:12-61`` (``check_tol_diff``):
tuple outputs are compared element-wise with per-index tolerances, the default criterion is
``assert_close`` on fp32 casts, ``ptol < 1`` switches to a fraction-of-elements criterion and
``mixed_tol`` to 2**-6 absolute below 1 / relative above.
"""

import torch


def _nth(value, index):
    if isinstance(value, (tuple, list)):
        if index >= len(value):
            raise IndexError(f"Tolerance tuple/list index {index} out of range for value {value}.")
        return value[index]
    return value


def check_tol_diff(norm, ref, atol=1e-2, rtol=1e-2, ptol=1.0, mixed_tol=False):
    if isinstance(norm, (tuple, list)):
        for i, (n_i, r_i) in enumerate(zip(norm, ref)):
            check_tol_diff(n_i, r_i, _nth(atol, i), _nth(rtol, i), _nth(ptol, i), _nth(mixed_tol, i))
        return

    if mixed_tol:
        small = ref.abs() < 1.0
        tol = 2.0**-6
        torch.testing.assert_close(norm[small], ref[small], atol=tol, rtol=0)
        torch.testing.assert_close(norm[~small], ref[~small], atol=0, rtol=tol)
        return

    if ptol != 1.0:
        assert ptol < 1.0, f"{ptol=} should <= 1.0"
        ok = torch.isclose(norm, ref, rtol=rtol, atol=atol)
        ratio = int(ok.sum()) / max(ok.numel(), 1)
        assert ratio >= ptol, f"match ratio {ratio:.5%} of {ok.numel()} elements is under ptol={ptol:%}"
        return

    torch.testing.assert_close(norm.to(torch.float32), ref.to(torch.float32), atol=atol, rtol=rtol)
