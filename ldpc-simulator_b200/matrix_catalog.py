"""Registry of the ALIST files of a code database, keyed by (family, rate, n).

Semantics of python_ldpc_app/matrix_catalog.py:9-203: file names are parsed by
family-specific patterns (first match wins, in the reference's order), with the
ALIST header as fallback; ``k = round(n * rate)`` comes from the NAME for the
wimax / wifi families; ``*.txt`` files are accepted, not only ``*.alist.txt``;
the list is sorted by (family, rate, n).
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple


@dataclass
class MatrixInfo:
    path: str
    name: str
    n: int
    k: int
    m: int
    rate: float
    family: str      # wimax | ccsds | bch | wifi | wran | wigig | custom | unknown


def _ratio(k, n, fallback=0):
    return k / n if n > 0 else fallback


# (pattern, family, extractor(match) -> (n, k, rate)) in matching order
_RULES: List[Tuple[re.Pattern, str, Callable]] = [
    (re.compile(r"wimax_(\d+)_([\d.]+[A-B]?)\.alist\.txt"), "wimax",
     lambda g: (lambda n, r: (n, int(round(n * r)), r))(int(g[1]), float(re.sub(r"[A-Za-z]", "", g[2])))),
    (re.compile(r"CCSDS_ldpc_n(\d+)_k(\d+)\.alist\.txt"), "ccsds",
     lambda g: (int(g[1]), int(g[2]), _ratio(int(g[2]), int(g[1])))),
    (re.compile(r"wifi_(\d+)_r(\d+)\.alist\.txt"), "wifi",
     lambda g: (lambda n, r: (n, int(round(n * r)), r))(int(g[1]), int(g[2]) / 100.0)),
    (re.compile(r"wigig_R(\d+)_N(\d+)_K(\d+)\.alist\.txt"), "wigig",
     lambda g: (int(g[2]), int(g[3]), _ratio(int(g[3]), int(g[2]), int(g[1]) / 100.0))),
    (re.compile(r"WRAN_N(\d+)_K(\d+)_P\d+_R(\d+)\.txt"), "wran",
     lambda g: (int(g[1]), int(g[2]), _ratio(int(g[2]), int(g[1])))),
    (re.compile(r"BCH_(\d+)_(\d+)_\d+"), "bch",
     lambda g: (int(g[1]), int(g[2]), _ratio(int(g[2]), int(g[1])))),
    (re.compile(r"Tanner_(\d+)_(\d+)\.alist\.txt"), "custom",
     lambda g: (int(g[1]), int(g[2]), _ratio(int(g[2]), int(g[1])))),
    (re.compile(r"LDPC_N(\d+)_K(\d+)"), "custom",
     lambda g: (int(g[1]), int(g[2]), _ratio(int(g[2]), int(g[1])))),
]


class MatrixCatalog:
    def __init__(self, base_dir: str):
        self.matrices: List[MatrixInfo] = []
        self._scan_directory(base_dir)
        self.matrices.sort(key=lambda mi: (mi.family, mi.rate, mi.n))

    def _scan_directory(self, base_dir: str) -> None:
        for root, _dirs, files in os.walk(base_dir):
            for fname in files:
                if not fname.endswith(".txt"):          # covers *.alist.txt as well
                    continue
                info = self._parse_filename(os.path.join(root, fname), fname)
                if info:
                    self.matrices.append(info)

    def _parse_filename(self, filepath: str, fname: str) -> Optional[MatrixInfo]:
        for pattern, family, extract in _RULES:
            hit = pattern.match(fname)
            if hit:
                n, k, rate = extract(hit)
                return MatrixInfo(path=filepath, name=fname, n=n, k=k, m=n - k, rate=rate, family=family)
        return self._parse_alist_header(filepath, fname)

    def _parse_alist_header(self, filepath: str, fname: str) -> Optional[MatrixInfo]:
        try:
            with open(filepath, "r") as fh:
                fields = fh.readline().split()
            if len(fields) >= 2:
                n, m = int(fields[0]), int(fields[1])
                return MatrixInfo(path=filepath, name=fname, n=n, k=n - m, m=m,
                                  rate=_ratio(n - m, n), family="unknown")
        except (ValueError, IOError):
            pass
        return None

    # ---- queries ---------------------------------------------------------------
    def get_by_rate_range(self, min_rate: float, max_rate: float) -> List[MatrixInfo]:
        return [mi for mi in self.matrices if min_rate <= mi.rate <= max_rate]

    def get_by_family(self, family: str) -> List[MatrixInfo]:
        return [mi for mi in self.matrices if mi.family == family]

    def get_nearest_rate(self, target_rate: float, family: str = None, block_size: int = None):
        pool = [mi for mi in self.matrices
                if (not family or mi.family == family) and (not block_size or mi.n == block_size)]
        return min(pool, key=lambda mi: abs(mi.rate - target_rate)) if pool else None

    def _neighbour(self, current: MatrixInfo, better: Callable[[float], bool], pick):
        same_family = [mi for mi in self.matrices if mi.family == current.family and better(mi.rate)]
        pool = [mi for mi in same_family if mi.n == current.n] or same_family
        return pick(pool, key=lambda mi: mi.rate) if pool else None

    def get_lower_rate(self, current: MatrixInfo) -> Optional[MatrixInfo]:
        return self._neighbour(current, lambda r: r < current.rate, max)

    def get_higher_rate(self, current: MatrixInfo) -> Optional[MatrixInfo]:
        return self._neighbour(current, lambda r: r > current.rate, min)

    def __len__(self):
        return len(self.matrices)

    def __repr__(self):
        tally = {}
        for mi in self.matrices:
            tally[mi.family] = tally.get(mi.family, 0) + 1
        body = ", ".join(f"{fam}={cnt}" for fam, cnt in sorted(tally.items()))
        return f"MatrixCatalog({len(self.matrices)} matrices: {body})"
