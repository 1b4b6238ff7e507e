"""Loader for the committed golden vectors (tests/golden, made by oracle/gen_golden.py)."""
import json
import os
import re
from collections import OrderedDict

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(GOLDEN, "INDEX.json")))["cases"]
FIXTURES = [c for c in CASES if c.startswith("chrT_")]

_TRACE = re.compile(r"^\s+(-?[\d.]+|nan|inf)\s+(-?[\d.]+|nan|inf)\s+(-?[\d.]+|nan|inf)\s+(\d+)\s+(-?[\d.]+|nan|inf)$")


class Case(object):
    def __init__(self, name):
        self.name = name
        self.out = json.load(open(os.path.join(GOLDEN, name + ".json")))
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.size = int(z["size"])
        self.rows, self.cols, self.vals = z["rows"], z["cols"], z["vals"]
        ch = self.out["chromosomes"]
        self.chromosomes = OrderedDict((k, int(v)) for k, v in ch) if ch else None
        self.symmetricized = self.out["symmetricized"]

    @staticmethod
    def bads(d):
        return dict((int(k), v) for k, v in d.items())

    @staticmethod
    def trace_lines(stdout):
        """The per-iteration lines printed at normalize_hic.py:219-220."""
        return [l for l in stdout.splitlines() if _TRACE.match(l)]


_cache = {}


def load(name):
    if name not in _cache:
        _cache[name] = Case(name)
    return _cache[name]
