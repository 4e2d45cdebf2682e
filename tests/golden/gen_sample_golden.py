#!/usr/bin/env python
"""Writes tests/golden/sample_stream.json: the first candidates of the free-state sampling stream specified in
oracle/sample.c (Philox4x32-10 counter layout, u53, lo + u (hi - lo)) and a few Morton keys, as hex floats.
The stream has no counterpart in the reference (Julia's global RNG); this file pins OUR specification so that a
later change to the counter layout or the key definition cannot go unnoticed.   python tests/golden/gen_sample_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as orc  # noqa: E402

spaces = {"unit2": ([0.0, 0.0], [1.0, 1.0]), "box3": ([-1.0, 2.0, 0.0], [1.0, 5.0, 0.5]),
          "di4": ([0.0, 0.0, -1.5, -1.5], [1.0, 1.0, 1.5, 1.5])}
out = {"candidates": {}, "morton": {}}
for name, (lo, hi) in spaces.items():
    S = orc.StateSpace(lo, hi)
    for seed in (1, (1 << 40) + 7):
        for c in (0, 1, 2, 12345, (1 << 33) + 5):
            x = orc.sample_candidate(S, seed, c)
            out["candidates"]["%s/%d/%d" % (name, seed, c)] = [float(v).hex() for v in x]
    for c in (0, 1, 2, 3):
        x = orc.sample_candidate(S, 1, c)
        out["morton"]["%s/%d" % (name, c)] = orc.morton_key(S, x)
with open(os.path.join(HERE, "sample_stream.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print("wrote", len(out["candidates"]), "candidates,", len(out["morton"]), "keys")
