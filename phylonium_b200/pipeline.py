"""Host-side mirror of the reference interface around the hot path, same names and
argument meaning as /root/reference/src/process.h:12, src/evo_model.h and
src/io.cxx:141-163, so that callers and tests read like the reference's own.

The numbers come from the CUDA library through capi.Context; the only arithmetic done
here is the estimator on two integers (evo_model.cxx:100-131) and text formatting.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Sequence

from .capi import DIST_ANI, DIST_JC, DIST_RAW, Context


@dataclass
class EvoModel:
    """evo_model, src/evo_model.h:13-50: two counters and the estimators."""

    substitutions: int = 0
    homologs: int = 0

    def total(self) -> int:
        return self.homologs

    def estimate_raw(self, zero_on_error: bool = False) -> float:
        if self.homologs == 0:
            return 0.0 if zero_on_error else math.nan
        return self.substitutions / float(self.homologs)

    def estimate_ani(self, zero_on_error: bool = False) -> float:
        if self.homologs == 0:
            return 0.0 if zero_on_error else math.nan
        return (1.0 - self.substitutions / float(self.homologs)) * 100

    def estimate_JC(self, zero_on_error: bool = False) -> float:
        dist = self.estimate_raw(zero_on_error)
        if math.isnan(dist):
            return dist
        x = 1.0 - (4.0 / 3.0) * dist
        if x <= 0.0:
            # log of a non-positive number, as libm answers it (src/evo_model.cxx:124-131 calls
            # log() unguarded): log(0) = -inf -> dist = +inf; log(x < 0) = -nan on glibc/x86, and
            # the multiplication by -0.75 hands the NaN through with its sign, so the reference
            # (and the C++ host here) print "-nan" for a saturated pair (raw distance > 0.75)
            dist = math.inf if x == 0.0 else math.copysign(math.nan, -1.0)
        else:
            dist = -0.75 * math.log(x)
        return 0.0 if dist <= 0.0 else dist

    def coverage(self, length: int) -> float:
        return self.homologs / length


def process(subject_index: int, queries: Sequence[bytes], flags: int = 0, ctx: Context | None = None) -> List[EvoModel]:
    """process(queries[subject_index], queries) -> row-major N*N evo_model matrix."""
    own = ctx is None
    ctx = ctx or Context()
    try:
        subst, homol = ctx.process(queries, subject_index, flags)
    finally:
        if own:
            ctx.close()
    N = len(queries)
    return [EvoModel(int(subst[i, j]), int(homol[i, j])) for i in range(N) for j in range(N)]


def _fmt(x: float, kind: int) -> str:
    if math.isnan(x):
        return "-nan" if math.copysign(1.0, x) < 0 else "nan"
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    return ("%.4g" if kind == DIST_ANI else "%.4e") % x


def format_matrix(names: Sequence[str], matrix: Sequence[EvoModel], kind: int = DIST_JC) -> str:
    """just_print(), src/io.cxx:141-163: PHYLIP text, diagonal forced to zero."""
    N = len(names)
    getter = {DIST_RAW: EvoModel.estimate_raw, DIST_JC: EvoModel.estimate_JC, DIST_ANI: EvoModel.estimate_ani}[kind]
    lines = [str(N)]
    for i in range(N):
        cells = [names[i]]
        for j in range(N):
            d = 0.0 if i == j else getter(matrix[i * N + j])
            cells.append(_fmt(d, kind))
        lines.append("  ".join(cells))
    return "\n".join(lines) + "\n"
