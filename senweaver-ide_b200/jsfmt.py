"""JavaScript string / number formatting rules the reference's wire formats depend on.

The reference builds its payloads and prompt text with `String.prototype.substring`, `.length`
(both count UTF-16 code units) and `Number.prototype.toFixed` (decimal expansion of the exact
binary64 value, ties away from zero — not Python's round-half-even on the same expansion)."""
from __future__ import annotations

import math
from decimal import ROUND_HALF_UP, Decimal


def js_length(s: str) -> int:
    return len(s.encode("utf-16-le", "surrogatepass")) // 2


def js_substring(s: str, start: int, end: int | None = None) -> str:
    if s.isascii():
        return s[start:end]
    raw = s.encode("utf-16-le", "surrogatepass")
    return raw[2 * start: None if end is None else 2 * end].decode("utf-16-le", "surrogatepass")


def js_to_fixed(x: float, digits: int) -> str:
    """Number.prototype.toFixed for |x| < 1e21: "Let n be an integer for which n / 10^f - x is as close to zero as
    possible. If there are two such n, pick the larger n" (larger in magnitude for the sign-stripped value)."""
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if abs(x) >= 1e21:
        return repr(x)
    q = Decimal(abs(x)).quantize(Decimal(1).scaleb(-digits), rounding=ROUND_HALF_UP)   # exact value, ties up
    txt = f"{q:.{digits}f}"
    return "-" + txt if x < 0 else txt          # (-0).toFixed() is "0"; (-0.0004).toFixed(3) is "-0.000"
