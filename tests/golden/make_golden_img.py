"""Generates tests/golden/golden_img.json from reference-compiled code (oracle/_ref/
libffr_refimg.so, built by `make -C oracle refimg` from /root/reference/src). Run in the build
container only.

Pins for SURVEY 8 row f1 and the stderr echo:
  tonemap  sha256 of the pixels the reference's own render_image() (ffr_img.cpp:199-309 over
           renderers/image_renderer.hpp:112-192) produces for a seeded buffer, per flame x mode x
           bit depth x gamma, plus the histogram bounds it prints. The buffers are rendered by
           the reference itself (ref_render), so the fixture does not depend on our oracle.
  echo     the text `std::cerr << "flame: " << json_flame` prints (ffr_buf.cpp:129) for every
           example flame and a few flames with awkward numbers.
tests/test_golden.py checks oracle_tonemap and ffr_flame_json_echo against them anywhere.
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

TONE_CASES = [("tkoz_test3", [160, 90], 3), ("csci6360_project", [192, 108], 0),
              ("barnsley_fern", [96, 64], 0)]
MODES = [(1, 8, 1.0), (2, 8, 1.0), (2, 8, 2.2), (2, 16, 0.5), (2, 16, 4.0),
         (3, 8, 1.0), (3, 8, 2.0), (3, 16, 2.2), (3, 16, 0.5)]
CHAINS, LEN, SEED = 400, 1000, 12


def echo_texts():
    out = {n: ex.example_json(n) for n in sorted(ex.EXAMPLES)}
    out["var:julian"] = flames.variation_flame("julian", dims=2, final=True)
    out["var:mobius3"] = flames.variation_flame("mobius", dims=3, final=True)
    out["numbers"] = json.dumps({"dimensions": 2, "a": [1e-7, 1.5e300, -0.0, 0.1, 100.0, 1e15, 1e16,
                                                     123456789012345678, -5, 2.5e-5, 1e-4, 3.0e0,
                                                     0.30000000000000004, 5e-324, 1.7976931348623157e308],
                                 "s": "a\"b\\c\n\té", "t": True, "n": None, "o": {"z": 1, "a": 2}})
    return out


def main():
    g = {"params": {"chains": CHAINS, "chain_len": LEN, "base_seed": SEED}, "tonemap": {}, "echo": {}}
    for name, size, cd in TONE_CASES:
        text = ex.example_json(name, size=size)
        raw, st, ok = po.ref_render(text, CHAINS, LEN, base_seed=SEED)
        g["tonemap"][name] = {"size": size, "buffer_sha256": hashlib.sha256(raw.tobytes()).hexdigest(),
                              "cases": {}}
        for mode, bits, gamma in MODES:
            if mode == 3 and cd != 3:
                continue
            img, info = po.ref_tonemap(text, raw, mode, bits, gamma)
            g["tonemap"][name]["cases"]["%d/%d/%r" % (mode, bits, gamma)] = {
                "sha256": hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest(),
                "hist_min": int(info["hist_min"]), "hist_max": int(info["hist_max"])}
    for key, text in echo_texts().items():
        g["echo"][key] = po.ref_flame_echo(text)
    with open(os.path.join(os.path.dirname(__file__), "golden_img.json"), "w") as f:
        json.dump(g, f, indent=0, sort_keys=True)
    print("tonemap cases:", sum(len(v["cases"]) for v in g["tonemap"].values()), "echo:", len(g["echo"]))


if __name__ == "__main__":
    main()
