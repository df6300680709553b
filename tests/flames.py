"""Synthetic test flames: one per variation of the reference factory
(reference: src/variations/variations.hpp:2387-2612), plus edge-case flames.

None of the 15 shipped examples uses an RNG-consuming variation, so
the draw-order contract of those is only exercised by the flames written here.
"""

import json

# JSON parameters (the keys each reference constructor reads) with values that keep the
# iteration reasonably bounded. D-vectors are given per dimension count below.
def _vec(d, vals):
    return list(vals[:d])

PARAMS_ND = {
    "linear": lambda d: {},
    "sinusoidal": lambda d: {},
    "spherical": lambda d: {},
    "bent": lambda d: {"scales_neg": _vec(d, [2.0, 0.5, 1.5]), "scales_pos": _vec(d, [1.0, 0.75, 1.25])},
    "rectangles": lambda d: {"params": _vec(d, [0.4, 0.0, 0.7])},
    "fisheye": lambda d: {"addval": 1.0},
    "bubble": lambda d: {"addval": 4.0},
    "noise": lambda d: {},
    "blur": lambda d: {},
    "gaussian_blur": lambda d: {},
    "square_noise": lambda d: {},
    "separation": lambda d: {"params": _vec(d, [0.3, 0.6, 0.2]), "inside": _vec(d, [0.1, -0.2, 0.3])},
    "splits": lambda d: {"params": _vec(d, [0.2, -0.3, 0.1])},
    "pre_blur": lambda d: {},
    "modulus": lambda d: {"params": _vec(d, [0.8, 1.3, 0.6])},
    "celln": lambda d: {"sizes": _vec(d, [0.5, 0.8, 0.3])},
    "spherical_p": lambda d: {"norm": 3.0},
    "unit_sphere": lambda d: {},
    "unit_sphere_p": lambda d: {"norm": 1.5},
    "unit_cube": lambda d: {},
}

PARAMS_2D = {
    "swirl": {}, "horseshoe": {}, "polar": {}, "polar2": {}, "handkerchief": {}, "heart": {},
    "disc": {},
    "disc2": {"rotation": 0.7, "twist": 0.4},
    "waves": {"xfreq": 2.0, "xscale": 0.3, "yfreq": 1.5, "yscale": -0.2},
    "fan": {"x": 0.6, "y": 0.3},
    "rings": {"value": 0.5},
    "spiral": {}, "hyperbolic": {}, "diamond": {}, "ex": {}, "julia": {}, "exponential": {},
    "power": {}, "cosine": {},
    "blob": {"low": 0.3, "high": 1.2, "waves": 5.0},
    "pdj": {"a": 1.1, "b": -0.7, "c": 0.9, "d": 1.3},
    "cylinder": {},
    "perspective": {"distance": 2.5, "angle": 0.6},
    "julian": {"power": 3.0, "dist": 1.2},
    "juliascope": {"power": -2.0, "dist": 0.9},
    "radial_blur": {"angle": 0.4, "flam3_weight": 0.5},
    "pie": {"slices": 6.0, "rotation": 0.3, "thickness": 0.5},
    "ngon": {"sides": 5.0, "power": 2.0, "corners": 1.0, "circle": 0.8},
    "curl": {"c1": 0.5, "c2": 0.2},
    "arch": {"flam3_weight": 0.8},
    "tangent": {},
    "rays": {"flam3_weight": 0.7},
    "blade": {"flam3_weight": 0.9},
    "secant": {"flam3_weight": 0.6},
    "twintrian": {"flam3_weight": 0.5},
    "cross": {}, "exp": {}, "log": {}, "sin": {}, "cos": {}, "tan": {}, "sec": {}, "csc": {},
    "cot": {}, "sinh": {}, "cosh": {}, "tanh": {}, "sech": {}, "csch": {}, "coth": {},
    "auger": {"freq": 3.0, "flam3_weight": 0.5, "scale": 0.4, "sym": 0.3},
    "flux": {"spread": 0.5, "flam3_weight": 0.6},
    "mobius": {"a": [0.8, 0.1], "b": [0.2, -0.3], "c": [0.1, 0.2], "d": [1.0, 0.4]},
    "scry": {"flam3_weight": 0.7},
    "split": {"xsize": 0.8, "ysize": 1.3},
    "stripes": {"space": 0.3, "warp": 0.5},
    "wedge": {"swirl": 0.2, "count": 3.0, "angle": 0.5, "hole": 0.1},
    "wedge_julia": {"angle": 0.4, "count": 2.0, "power": 3.0, "dist": 1.1},
    "wedge_sph": {"angle": 0.3, "count": 4.0, "swirl": 0.1, "hole": 0.2},
    "whorl": {"inside": 0.6, "outside": 0.3, "flam3_weight": 0.9},
    "supershape": {"n1": 2.0, "m": 5.0, "n2": 1.5, "n3": 2.5, "rnd": 0.3, "holes": 0.1},
    "flower": {"petals": 5.0, "holes": 0.2},
    "conic": {"eccen": 0.7, "holes": 0.1},
    "parabola": {"height": 0.8, "width": 0.6},
    "bipolar": {"shift": 0.3},
    "boarders": {"prob": 0.75},
    "butterfly": {},
    "cell": {"size": 0.6},
    "cpow": {"r": 1.2, "i": 0.3, "power": 3.0},
    "curve": {"xamp": 0.4, "yamp": 0.3, "xlen": 0.8, "ylen": 1.1},
    "edisc": {}, "elliptic": {},
    "escher": {"beta": 0.5},
    "foci": {},
    "lazysusan": {"x": 0.2, "y": -0.1, "spin": 0.7, "twist": 0.4, "space": 0.3, "flam3_weight": 0.9},
    "loonie": {"flam3_weight": 0.8},
    "oscope": {"frequency": 1.5, "amplitude": 0.8, "damping": 0.3, "separation": 0.4},
    "popcorn": {"x": 0.3, "y": -0.2, "c": 2.0},
}

ALL_VARIATIONS = list(PARAMS_ND) + list(PARAMS_2D)
assert len(ALL_VARIATIONS) == 98

# variations whose calc() uses only + - * / sqrt floor rint trunc copysign fabs max and
# comparisons (SURVEY Q6 class ii): correctly rounded on both sides => bit-exact on the GPU.
# The RNG-consuming ones of the class draw through randNum only (no sincos in
# randDirection for dims == 1, but 2-d noise/blur use sincos, so they are excluded).
IEEE_EXACT = [
    "linear", "spherical", "bent", "rectangles", "fisheye", "bubble", "square_noise",
    "separation", "splits", "modulus", "celln", "horseshoe", "hyperbolic", "perspective",
    "curl", "cross", "mobius", "scry", "stripes", "conic", "boarders", "butterfly", "cell",
    "loonie", "unit_sphere", "unit_cube",
]

RNG_VARIATIONS = [
    "noise", "blur", "gaussian_blur", "square_noise", "pre_blur", "julia", "julian",
    "juliascope", "radial_blur", "pie", "arch", "rays", "blade", "twintrian", "wedge_julia",
    "supershape", "flower", "conic", "parabola", "boarders", "cpow",
]


def _affine(d, scale, rot, b):
    A = [[0.0] * d for _ in range(d)]
    for i in range(d):
        A[i][i] = scale
    if d >= 2:
        A[0][1] = rot
        A[1][0] = -rot
    if d >= 3:
        A[1][2] = 0.5 * rot
        A[2][0] = -0.5 * rot
    return {"A": A, "b": list(b[:d])}


def variation_entry(name, weight, dims, axes=(0, 1)):
    v = {"name": name, "weight": weight}
    if name in PARAMS_ND:
        v.update(PARAMS_ND[name](dims))
    else:
        v.update(PARAMS_2D[name])
        if dims > 2:
            v["axis_x"], v["axis_y"] = axes
    return v


def variation_flame(name, dims=2, size=None, color=True, final=False):
    """A 3-xform flame whose xforms 1 and 2 run `name`; returns JSON text."""
    if name in PARAMS_2D and dims < 2:
        raise ValueError("2-d variation needs dims >= 2")
    if size is None:
        size = {1: [4096], 2: [256, 256], 3: [48, 48, 48]}[dims]
    fl = {
        "dimensions": dims,
        "size": size,
        "bounds": [[-3, 3]] * dims,
        "xforms": [
            {"weight": 1.0,
             "variations": [variation_entry("linear", 1.0, dims)],
             "pre_affine": _affine(dims, 0.5, 0.1, [0.3, -0.2, 0.1])},
            {"weight": 0.7,
             "variations": [variation_entry(name, 0.6, dims, (2, 0)),
                            variation_entry("linear", 0.25, dims)],
             "pre_affine": _affine(dims, 0.8, -0.3, [-0.2, 0.4, 0.3]),
             "post_affine": _affine(dims, 0.7, 0.1, [0.1, 0.1, -0.1])},
            {"weight": 0.4,
             "variations": [variation_entry(name, 0.5, dims, (1, 2))],
             "pre_affine": _affine(dims, 1.1, 0.2, [0.5, 0.5, -0.4])},
        ],
    }
    if color:
        fl["color_dimensions"] = 2
        fl["color_speed"] = 0.4
        fl["xforms"][0]["color"] = [1.0, 0.0]
        fl["xforms"][1]["color"] = [0.0, 1.0]
        fl["xforms"][1]["color_speed"] = 0.8
    if final:
        fl["final_xform"] = {
            "variations": [variation_entry("linear", 0.9, dims),
                           variation_entry(name, 0.1, dims, (0, 2))],
            "post_affine": _affine(dims, 0.9, 0.05, [0.0, 0.1, 0.0]),
        }
        if color:
            fl["final_xform"]["color"] = [0.5, 0.5]
            fl["final_xform"]["color_speed"] = 0.25
    return json.dumps(fl, indent=1)


def divergent_flame():
    """Expanding affine map with a small weight: produces bad values (|x| > 1e20) regularly,
    exercising the bad value record + chain re-initialisation path
    (reference: buffer_renderer.hpp:175-186)."""
    fl = {
        "dimensions": 2,
        "size": [128, 128],
        "bounds": [[-2, 2], [-2, 2]],
        "color_dimensions": 1,
        "xforms": [
            {"weight": 1.0, "color": [0.2],
             "variations": [{"name": "linear", "weight": 1.0}],
             "pre_affine": {"A": [[0.5, 0.0], [0.0, 0.5]], "b": [0.25, -0.25]}},
            {"weight": 0.9, "color": [0.9],
             "variations": [{"name": "linear", "weight": 1.0}],
             "pre_affine": {"A": [[1.0e4, 0.0], [0.0, -1.0e4]], "b": [0.1, 0.0]}},
        ],
        "final_xform": {"variations": [{"name": "linear", "weight": 1.0}],
                        "pre_affine": {"A": [[0.5, 0.5], [-0.5, 0.5]], "b": [0.0, 0.0]}},
    }
    return json.dumps(fl, indent=1)


def one_d_flame():
    fl = {
        "dimensions": 1,
        "size": [65535],
        "bounds": [[-1.5, 1.5]],
        "color_dimensions": 1,
        "xforms": [
            {"weight": 2.0, "color": [0.0],
             "variations": [{"name": "linear", "weight": 1.0}],
             "pre_affine": {"A": [[0.5]], "b": [0.5]}},
            {"weight": 1.0, "color": [1.0],
             "variations": [{"name": "linear", "weight": 0.6},
                            {"name": "spherical", "weight": 0.05}],
             "pre_affine": {"A": [[-0.6]], "b": [-0.3]}},
            {"weight": 0.0,
             "variations": [{"name": "linear", "weight": 1.0}]},
        ],
    }
    return json.dumps(fl, indent=1)


def many_xforms_flame(k=20):
    """More than 16 xforms: std::sort leaves insertion-sort territory (flame.hpp:80-82) and
    the xform selection scan gets long; includes equal weights and a zero weight."""
    xfs = []
    for i in range(k):
        w = [1.0, 0.5, 0.25, 2.0, 0.0, 1.5, 0.5][i % 7]
        xfs.append({"weight": w,
                    "variations": [{"name": "linear", "weight": 1.0}],
                    "pre_affine": {"A": [[0.3, 0.02 * i], [-0.02 * i, 0.3]],
                                   "b": [0.6 * ((i * 7) % 11) / 11.0 - 0.3,
                                         0.6 * ((i * 5) % 13) / 13.0 - 0.3]}})
    fl = {"dimensions": 2, "size": [200, 150], "bounds": [[-1, 1], [-1, 1]], "xforms": xfs}
    return json.dumps(fl, indent=1)


def multi_variation_flame(names, dims=2, size=None, color=True, final_name=None):
    """One xform per entry of `names` (<= 7), each running that variation next to a damped
    linear term, plus a contracting linear xform: many variations per compiled kernel for
    the run-time compiled path (tests/test_gpu_jit.py)."""
    if size is None:
        size = {1: [4096], 2: [128, 128], 3: [32, 32, 32]}[dims]
    xfs = [{"weight": 1.0,
            "variations": [variation_entry("linear", 1.0, dims)],
            "pre_affine": _affine(dims, 0.5, 0.1, [0.3, -0.2, 0.1])}]
    for i, name in enumerate(names):
        xfs.append({"weight": 0.5 + 0.1 * i,
                    "variations": [variation_entry(name, 0.6, dims, ((i + 2) % 3, i % 3) if dims > 2 else (0, 1)),
                                   variation_entry("linear", 0.25, dims)],
                    "pre_affine": _affine(dims, 0.8, -0.3 + 0.05 * i, [-0.2, 0.4, 0.3]),
                    "post_affine": _affine(dims, 0.7, 0.1, [0.1, 0.1 - 0.02 * i, -0.1])})
    fl = {"dimensions": dims, "size": size, "bounds": [[-3, 3]] * dims, "xforms": xfs}
    if color:
        fl["color_dimensions"] = 2
        fl["color_speed"] = 0.4
        for i, xf in enumerate(xfs):
            if i % 2 == 0:
                xf["color"] = [i / 8.0, 1.0 - i / 8.0]
        xfs[1]["color_speed"] = 0.8
    if final_name is not None:
        fl["final_xform"] = {
            "variations": [variation_entry("linear", 0.9, dims),
                           variation_entry(final_name, 0.1, dims, (0, 2))],
            "post_affine": _affine(dims, 0.9, 0.05, [0.0, 0.1, 0.0]),
        }
        if color:
            fl["final_xform"]["color"] = [0.5, 0.5]
            fl["final_xform"]["color_speed"] = 0.25
    return json.dumps(fl, indent=1)


def affine_flame(dims=2, nx=3, pre="general", post="identity", weights=(1.0,), size=None,
                 bounds=None, neg_zero=None, vary_weights=False):
    """Pure-affine flames for the flame-specialised affine kernel (K1e) and its exact
    simplifier (csrc/ffr_jit_host.cuh): every combination of coefficient shapes the generator
    treats differently -- coefficients identical across xforms or not, 0 / +1 / -1 entries,
    offsets that are +0, -0.0 or vary, one or several `linear` variations, present / absent
    pre and post affines (3-d only: absent means skipped, reference xform.hpp:216,222).
      pre / post: "general" (all entries differ per xform), "scale" (0.5*I, offsets differ),
                  "identity", "perm" (a signed permutation: entries 0, +1, -1), "none" (omit)
      neg_zero:   ("pre"|"post", row, xform) puts -0.0 into that offset
    """
    if size is None:
        size = {1: [4096], 2: [160, 120], 3: [40, 36, 32]}[dims]
    if bounds is None:
        bounds = [[-1.5, 1.5]] * dims

    def mat(kind, k):
        if kind == "general":
            A = [[(0.45 if i == j else 0.0) + 0.07 * ((i * 3 + j * 5 + k * 2) % 5 - 2) for j in range(dims)]
                 for i in range(dims)]
            b = [0.5 * (((k * 7 + i * 3) % 9) / 4.0 - 1.0) for i in range(dims)]
        elif kind == "scale":
            A = [[0.5 if i == j else 0.0 for j in range(dims)] for i in range(dims)]
            b = [0.5 * ((k >> i) & 1) - (0.25 if i == 0 else 0.0) for i in range(dims)]
        elif kind == "identity":
            A = [[1.0 if i == j else 0.0 for j in range(dims)] for i in range(dims)]
            b = [0.0] * dims
        elif kind == "perm":
            A = [[0.0] * dims for _ in range(dims)]
            for i in range(dims):
                A[i][(i + 1) % dims] = -1.0 if i == 0 else 1.0
            b = [0.0] * dims
        else:
            raise ValueError(kind)
        return A, b

    xfs = []
    for k in range(nx):
        ws = [w * (1.0 + 0.1 * k) if vary_weights else w for w in weights]
        xf = {"weight": 1.0 + 0.25 * (k % 3),
              "variations": [{"name": "linear", "weight": w} for w in ws]}
        for which, kind in (("pre", pre), ("post", post)):
            if kind == "none":
                continue
            A, b = mat(kind, k)
            if neg_zero is not None and neg_zero[0] == which and neg_zero[2] == k:
                b[neg_zero[1]] = -0.0
            xf[which + "_affine"] = {"A": A, "b": b}
        xfs.append(xf)
    fl = {"dimensions": dims, "size": size, "bounds": bounds, "xforms": xfs}
    return json.dumps(fl, indent=1)
