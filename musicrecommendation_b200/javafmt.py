"""java.lang.Double.toString for the model text format `user\\tsong\\tscore\\n` (MusicRecommender.scala:493).

Shortest digit string that round-trips (what JDK >= 19 prints; older JDKs print the same digits except for a few
well-known non-shortest cases), laid out by Java's rules: plain decimal for 1e-3 <= |x| < 1e7, otherwise
"computerised scientific" d.dddE[-]n; always at least one digit after the point.
"""
from __future__ import annotations

import math
from decimal import Decimal


def double_to_string(x: float) -> str:
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    sign = "-" if x < 0 else ""
    t = Decimal(repr(abs(x))).as_tuple()            # repr() is the shortest round-tripping digit string
    digits = "".join(map(str, t.digits)).rstrip("0") or "0"
    point = len(t.digits) + t.exponent              # value = 0.d1d2... * 10**point
    a = abs(x)
    if 1e-3 <= a < 1e7:
        if point <= 0:
            body = "0." + "0" * (-point) + digits
        elif point >= len(digits):
            body = digits + "0" * (point - len(digits)) + ".0"
        else:
            body = digits[:point] + "." + digits[point:]
        return sign + body
    frac = digits[1:] or "0"
    return f"{sign}{digits[0]}.{frac}E{point - 1}"
