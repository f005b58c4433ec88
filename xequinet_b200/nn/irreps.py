"""Minimal irreps bookkeeping for the XPaiNN path: "AxOe + Bx1o + Cx2e" (nn/model.py:58)."""
from __future__ import annotations

import re
from typing import Iterable, Tuple


def parse_irreps(irreps) -> Tuple[int, int, int]:
    """Multiplicities (mul0, mul1, mul2) of an lmax <= 2 irreps spec in the order 0e, 1o, 2e.
    Accepts the e3nn string form or an iterable of (mul, "le") / (mul, (l, p)) pairs."""
    items = []
    if isinstance(irreps, str):
        for tok in irreps.split("+"):
            m = re.fullmatch(r"\s*(?:(\d+)\s*x\s*)?(\d+)([eo])\s*", tok)
            if not m:
                raise ValueError(f"cannot parse irreps term {tok!r}")
            items.append((int(m.group(1) or 1), int(m.group(2)), 1 if m.group(3) == "e" else -1))
    else:
        for mul, ir in irreps:
            if isinstance(ir, str):
                m = re.fullmatch(r"\s*(\d+)([eo])\s*", ir)
                items.append((int(mul), int(m.group(1)), 1 if m.group(2) == "e" else -1))
            else:
                l, p = ir
                items.append((int(mul), int(l), int(p)))
    muls = [0, 0, 0]
    expect = [(0, 1), (1, -1), (2, 1)]
    pos = 0
    for mul, l, p in items:
        while pos < 3 and expect[pos] != (l, p):
            pos += 1
        if pos == 3:
            raise NotImplementedError(
                f"irreps {irreps!r}: the B200 path supports 'Ax0e + Bx1o + Cx2e' (lmax <= 2) in that order")
        muls[pos] += mul
    return tuple(muls)


def irreps_dim(muls) -> int:
    return muls[0] + 3 * muls[1] + 5 * muls[2]


def num_irreps(muls) -> int:
    return sum(muls)
