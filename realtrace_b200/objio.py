"""OBJ + texture loading for the harness.

Restates the reference's loader, load_image_from_obj
(/root/reference/Serial/lumina.cpp:195-290): whitespace tokenising, `f a/b/c`
triplets split on '/', only the first three corners of a face are read,
vertices scaled by SCALING_FACTOR 15 (:43,:275), `vt` keeps two components
(:279-281).  The 2000-triangle cap (:266) is a parameter (`max_faces`).

Texture fetch: the reference reads texels through DevIL
(get_value_by_coordinate, lumina.cpp:175-187), a dependency that is neither
vendored nor version-pinned, and its shipped main never passes a texture
(:366).  That fetch is therefore "parity unpinned" (SURVEY §8c); the harness
defines it here in two modes and always feeds the SAME vertex colours to the
CPU oracle and to the GPU, so the hot path stays fully pinned:
  normalised  texel/255, vt index-1, u -> column, v -> row measured from the
              bottom of the image (the usual OBJ convention)
  literal     the arithmetic of :175-187 on a PIL-decoded RGBA8 image with a
              top-left origin, including the un-decremented vt index of :249,
              the swapped row/column and the missing /255
"""
from __future__ import annotations

import os

import numpy as np

SCALING_FACTOR = 15.0  # lumina.cpp:43

ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")


def parse_obj(path, max_faces=None):
    """Returns (vertices (V,3) f64 scaled, texcoords (T,2) f64, faces list of ((v,vt|None),)*3)."""
    verts, tex, faces = [], [], []
    with open(path) as f:
        tokens_by_line = [ln.split() for ln in f]
    # The reference tokenises the whole stream (`is >> c`), not lines; for the
    # well-formed files it ships the two are equivalent.
    for tok in tokens_by_line:
        if not tok:
            continue
        c = tok[0]
        if c == "f":
            corners = []
            for data in tok[1:4]:
                parts = data.split("/")
                vi = int(parts[0])
                ti = int(parts[1]) if len(parts) >= 2 and parts[1] != "" else None
                corners.append((vi, ti))
            if max_faces is None or len(faces) < max_faces:
                faces.append(tuple(corners))
        elif c == "v":
            verts.append((float(tok[1]) * SCALING_FACTOR, float(tok[2]) * SCALING_FACTOR,
                          float(tok[3]) * SCALING_FACTOR))
        elif c == "vt":
            tex.append((float(tok[1]), float(tok[2])))
    return np.asarray(verts, np.float64).reshape(-1, 3), np.asarray(tex, np.float64).reshape(-1, 2), faces


def triangles_from_obj(path, max_faces=None):
    """(N,9) float32 triangle soup in face order, plus the parsed pieces."""
    verts, tex, faces = parse_obj(path, max_faces)
    idx = np.asarray([[c[0] - 1 for c in f] for f in faces], np.int64).reshape(-1, 3)
    tri = verts[idx].reshape(-1, 9).astype(np.float32)
    return tri, verts, tex, faces


def load_texture(path):
    from PIL import Image
    im = Image.open(path).convert("RGBA")
    return np.asarray(im, np.uint8)  # (H, W, 4), row 0 = top


def vertex_colours(faces, tex, image, mode="normalised"):
    """Three RGB colours per face -> (N,9) float32, or None if a face lacks vt."""
    h, w = image.shape[0], image.shape[1]
    out = np.zeros((len(faces), 9), np.float32)
    flat = image.reshape(-1)
    for n, f in enumerate(faces):
        for k, (_, ti) in enumerate(f):
            if ti is None:
                return None
            if mode == "normalised":
                u, v = tex[ti - 1]
                col = min(max(int(np.floor(u * w)), 0), w - 1)
                row = min(max(int(np.floor((1.0 - v) * h)), 0), h - 1)
                out[n, 3 * k:3 * k + 3] = image[row, col, :3].astype(np.float32) / np.float32(255.0)
            elif mode == "literal":
                if ti >= len(tex):           # texture_vertices[idx] one past the end: UB in the reference
                    out[n, 3 * k:3 * k + 3] = (0.8, 0.1, 0.0)
                    continue
                u, v = tex[ti]               # lumina.cpp:249 (no -1)
                i = int(np.floor(u * w))     # :180-183
                j = int(np.floor(v * h))
                if 0 <= i < h and 0 <= j < w:  # :184 (compares i with height, j with width)
                    o = (i * w + j) * 4
                    out[n, 3 * k:3 * k + 3] = flat[o:o + 3].astype(np.float32)
                else:
                    out[n, 3 * k:3 * k + 3] = (0.8, 0.1, 0.0)  # :186
            else:
                raise ValueError(mode)
    return out
