"""Parity metrics of BASELINE.json's north star, shared by the emulation and GPU tests.

Tolerances (written here once, asserted by the tests):
  * first-hit primitive id equal on >= 99.9 % of pixels
  * every 8-bit channel within +-1 LSB on >= 99.9 % of pixels
  * FP32 hit distance within 1e-4 relative wherever both sides hit the same primitive
The two fractions are also bounded over the HIT pixels alone (pixels where either side hit something): on a frame
that is mostly background a bar over all pixels would tolerate a much larger share of wrong hits.
"""
import numpy as np

ID_MATCH_MIN = 0.999
COLOUR_MATCH_MIN = 0.999
T_REL_TOL = 1e-4


def compare(rgb, prim, t, ref_rgb, ref_prim, ref_t):
    npix = ref_prim.size
    id_equal = (prim == ref_prim)
    both = id_equal & (ref_prim >= 0)
    rel = np.zeros(1)
    if both.any():
        rel = np.abs(t[both].astype(np.float64) - ref_t[both].astype(np.float64)) / np.abs(ref_t[both].astype(np.float64))
    diff = np.abs(rgb.astype(np.int16) - ref_rgb.astype(np.int16)).max(axis=-1)
    any_hit = (ref_prim >= 0) | (prim >= 0)
    n_hit = int(any_hit.sum())
    return {
        "any_hit_pixels": n_hit,
        "id_mismatches_on_hits": int((~id_equal & any_hit).sum()),
        "colour_bad_on_hits": int(((diff > 1) & any_hit).sum()),
        "colour_bad_off_hits": int(((diff > 1) & ~any_hit).sum()),
        "pixels": int(npix),
        "hit_pixels": int((ref_prim >= 0).sum()),
        "id_match": float(id_equal.mean()),
        "id_mismatches": int((~id_equal).sum()),
        "colour_within_1": float((diff <= 1).mean()),
        "colour_exact": float((diff == 0).mean()),
        "colour_bad_pixels": int((diff > 1).sum()),
        "colour_max_diff": int(diff.max()),
        "t_max_rel": float(rel.max()),
    }


def hit_budget(m, frac):
    """Mismatches allowed among the hit pixels: the north star's share of them (two at least: one epsilon tie on a
    shared edge usually shows in both neighbouring pixels)."""
    return max(2, int(np.ceil(frac * m["any_hit_pixels"])))


def assert_parity(m, tag=""):
    assert m["id_match"] >= ID_MATCH_MIN, (tag, m)
    assert m["colour_within_1"] >= COLOUR_MATCH_MIN, (tag, m)
    assert m["t_max_rel"] <= T_REL_TOL, (tag, m)
    assert m["id_mismatches_on_hits"] <= hit_budget(m, 1.0 - ID_MATCH_MIN), (tag, m)
    assert m["colour_bad_on_hits"] <= hit_budget(m, 1.0 - COLOUR_MATCH_MIN), (tag, m)
    assert m["colour_bad_off_hits"] == 0, (tag, m)          # background pixels carry exactly the background colour
