#!/usr/bin/env python3
"""Build oracle/_ref/libserial_ref.so from the reference's own Serial/ sources.

TEST INFRASTRUCTURE ONLY.  The reference sources are read where they lie under
/root/reference/Serial, staged into a throw-away temp directory, line-patched
there (hunks below, each verified against the text it expects), compiled with
plain g++ together with oracle/ref_harness.cpp, and only the resulting shared
object is written into oracle/_ref/ (git-ignored, NOT gpurun-ignored, so it
travels to the GPU box).  No reference source is ever copied into this repo and
the reference's own Makefile is not used (it needs GL/GLUT/GLEW/DevIL, absent).

Hunks (file:line refer to /root/reference/Serial):
  H1 utilities.h:5, utilities.cpp:11-15   drop the global `double abs(double)`
     that collides with libstdc++ under g++ 13 (same semantics as std::abs).
  H2 world.h:11   RECURSION_DEPTH becomes the runtime int g_oracle_depth
     (configs need depth 1/3/5; the reference hard-codes 10).
  H3 world.cpp:7-15   firstIntersection: count the call; walk the grid (when
     there are triangles), then test the non-triangle objects linearly with the
     reference's own Object::intersect — the objectList loop the reference left
     commented out, restricted to analytic objects because the shipped grid
     holds triangles only.  Mode TRUE_NEAREST runs that loop over every object.
     Order matters: UniformGrid::intersect overwrites Ray::t (uniform-grid.cpp
     :190,:224), so the grid goes first.
  H4 uniform-grid.cpp:143-146   silence the per-voxel list-length dump.
  H5 utilities.h:16   BBox::axis_max initialiser selectable at run time:
     numeric_limits<double>::min() as shipped, lowest() in mode BBOX_FIXED.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("REALTRACE_REFERENCE", "/root/reference/Serial")
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "libserial_ref.so")

CORE = ["color", "utilities", "uniform-grid", "camera", "vector3D", "world", "sphere",
        "triangle", "plane", "cylinder", "renderengine", "ray", "material"]  # Serial/Makefile:17 minus gl_utils
HEADERS = ["color.h", "utilities.h", "uniform-grid.h", "camera.h", "vector3D.h", "world.h", "sphere.h",
           "triangle.h", "plane.h", "cylinder.h", "renderengine.h", "ray.h", "material.h", "object.h",
           "lightsource.h", "pointlightsource.h"]


def _lines(path):
    with open(path) as f:
        return f.read().split("\n")


def _expect(lines, lineno, pattern, what):
    text = lines[lineno - 1]
    if not re.search(pattern, text):
        raise SystemExit(f"build_ref: {what}: line {lineno} is {text!r}, expected /{pattern}/ — "
                         "the reference changed; refusing to patch blindly")


def patch(stage):
    # H1 + H5
    p = os.path.join(stage, "utilities.h")
    L = _lines(p)
    _expect(L, 5, r"double\s+abs\s*\(\s*double", "H1 utilities.h")
    _expect(L, 16, r"axis_max\[0\].*numeric_limits.*min\(\)", "H5 utilities.h")
    L[4] = "extern int g_oracle_bbox_fixed;"
    L[15] = ("\t\taxis_max[0] = axis_max[1] = axis_max[2] = g_oracle_bbox_fixed ? "
             "std::numeric_limits < double >::lowest() : std::numeric_limits < double >::min();")
    open(p, "w").write("\n".join(L))

    p = os.path.join(stage, "utilities.cpp")
    L = _lines(p)
    _expect(L, 11, r"double\s+abs\s*\(\s*double", "H1 utilities.cpp")
    _expect(L, 15, r"^\}", "H1 utilities.cpp")
    del L[10:15]
    open(p, "w").write("\n".join(L))

    # H2
    p = os.path.join(stage, "world.h")
    L = _lines(p)
    _expect(L, 11, r"#define\s+RECURSION_DEPTH\s+10", "H2 world.h")
    L[10] = "extern int g_oracle_depth;\n#define RECURSION_DEPTH g_oracle_depth"
    open(p, "w").write("\n".join(L))

    # H3
    p = os.path.join(stage, "world.cpp")
    L = _lines(p)
    _expect(L, 5, r"float\s+World::firstIntersection", "H3 world.cpp")
    _expect(L, 15, r"uniform_grid\.intersect\(ray\)", "H3 world.cpp")
    _expect(L, 16, r"return\s+ray\.getParameter", "H3 world.cpp")
    body = [
        "\t++g_oracle_rays;",
        "\tif(g_oracle_mode == 2) {",
        "\t\tfor(int i=0; i<(int)objectList.size(); i++) if(objectList[i]->intersect(ray)) ray.setIdx(i);",
        "\t} else {",
        "\t\tif(g_oracle_has_grid) uniform_grid.intersect(ray);",
        "\t\tfor(int i : g_oracle_analytic) if(objectList[i]->intersect(ray)) ray.setIdx(i);",
        "\t}",
    ]
    L[6:15] = body
    L.insert(4, "extern thread_local unsigned long long g_oracle_rays; extern int g_oracle_mode; "
                "extern bool g_oracle_has_grid; extern std::vector<int> g_oracle_analytic;")
    open(p, "w").write("\n".join(L))

    # H4
    p = os.path.join(stage, "uniform-grid.cpp")
    L = _lines(p)
    _expect(L, 143, r"for\s*\(int i = 0; i < nv", "H4 uniform-grid.cpp")
    _expect(L, 144, r"cout\s*<<\s*voxels\[i\]", "H4 uniform-grid.cpp")
    _expect(L, 146, r"cout\s*<<\s*endl", "H4 uniform-grid.cpp")
    del L[142:146]
    open(p, "w").write("\n".join(L))


def build(verbose=True, opt="-O3"):
    if not os.path.isdir(REF):
        if verbose:
            print(f"build_ref: {REF} not present — keeping any prebuilt {OUT}")
        return os.path.exists(OUT)
    os.makedirs(OUT_DIR, exist_ok=True)
    stage = tempfile.mkdtemp(prefix="serial_ref_stage_")
    try:
        for name in HEADERS + [c + ".cpp" for c in CORE]:
            shutil.copy(os.path.join(REF, name), os.path.join(stage, name))
            os.chmod(os.path.join(stage, name), 0o644)
        patch(stage)
        srcs = [os.path.join(stage, c + ".cpp") for c in CORE] + [os.path.join(HERE, "ref_harness.cpp")]
        # plain build: no -march=native, no -ffast-math; -w: the reference is warning-noisy
        cmd = ["g++", "-std=c++14", opt, "-fPIC", "-shared", "-pthread", "-w", "-I", stage, "-I", HERE,
               "-o", OUT] + srcs
        if verbose:
            print("build_ref:", " ".join(cmd[:9]), "...")
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(stage, ignore_errors=True)
    if verbose:
        print("build_ref: wrote", OUT)
    return True


if __name__ == "__main__":
    ok = build(opt=os.environ.get("ORACLE_OPT", "-O3"))
    sys.exit(0 if ok else 1)
