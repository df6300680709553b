"""Generates tests/golden/golden.json from the UNMODIFIED reference (oracle/_ref, built by
`make -C oracle ref` from /root/reference/src). Run in the build container only.

The reference ships no golden vectors for this path, so these pins are
outputs of the reference itself run here: ISAAC known answers, xform order / cumulative weights /
index multipliers, and seeded buffer digests + statistics for every example flame and for a
flame per variation. tests/test_golden.py checks the oracle restatement against them anywhere
(no reference needed); tests/test_oracle_vs_reference.py re-derives them live when _ref exists.
"""
import hashlib
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import pyoracle as po  # noqa: E402
import flames  # noqa: E402

ex = importlib.import_module("flame-fractal-renderer_b200.examples")

CHAINS, LEN, LAST, SEED = 24, 1500, 700, 7


def digest(buf):
    return hashlib.sha256(np.ascontiguousarray(buf).tobytes()).hexdigest()


def clean(st):
    def f(x):
        if isinstance(x, float):
            return repr(x)
        if isinstance(x, list):
            return [f(y) for y in x]
        return x
    return {k: f(v) for k, v in st.items() if k not in ("bad_xf", "bad_pt")}


def render_case(text):
    buf, st, ok = po.ref_render(text, CHAINS, LEN, base_seed=SEED, last_len=LAST, bv_limit=1 << 20)
    info = po.ref_flame_info(text)
    return {"sha256": digest(buf), "stats": clean(st), "ok": ok,
            "ids": info["ids"], "cw": [repr(x) for x in info["cw"]],
            "mult_d": [repr(x) for x in info["mult_d"]], "mult_i": info["mult_i"],
            "cells": info["cells"], "cell_size": info["cell_size"],
            "hist_sum": int(buf.view(np.uint64).reshape(-1, info["cell_size"])[:, 0].sum())}


def main():
    g = {"params": {"chains": CHAINS, "chain_len": LEN, "last_len": LAST, "base_seed": SEED},
         "isaac": {}, "examples": {}, "variations": {}, "raw_pins": {}}
    for seed in (1, 2, 12345, 2**64 - 1):
        g["isaac"][str(seed)] = ["%016x" % int(x) for x in po.ref_isaac_words(seed, 40)]
    for name in sorted(ex.EXAMPLES):
        size = [48, 48, 48] if name.endswith("3d") else None
        g["examples"][name] = render_case(ex.example_json(name, size=size))
    for name in flames.ALL_VARIATIONS:
        dims = [2, 3] + ([1] if name in flames.PARAMS_ND else [])
        for d in dims:
            g["variations"]["%s/%d" % (name, d)] = render_case(
                flames.variation_flame(name, dims=d, final=(d == 3)))
    for nm, text in (("divergent", flames.divergent_flame()), ("one_d", flames.one_d_flame()),
                     ("many_xforms", flames.many_xforms_flame())):
        g["variations"][nm] = render_case(text)
    # the survey's raw-seed pins: one stream across all batches
    text = ex.example_json("sierpinski_triangle")
    for seed in (12345, 999):
        buf, st, _ = po.ref_render_raw(text, seed, 10_000_000, 1_048_576)
        g["raw_pins"][str(seed)] = {"md5": hashlib.md5(buf.tobytes()).hexdigest(),
                                    "max": int(buf.max())}
    with open(os.path.join(os.path.dirname(__file__), "golden.json"), "w") as f:
        json.dump(g, f, indent=0, sort_keys=True)
    print("raw pins:", g["raw_pins"])


if __name__ == "__main__":
    main()
