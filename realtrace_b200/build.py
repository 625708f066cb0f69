"""Compiles the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librealtrace_b200.so")
SOURCES = ["api.cu", "render.cu", "bvh_build.cu", "radix_sort.cu", "multi_device.cu", "microbench.cu"]
HEADERS = ["rt_hd.h", "rt_scene.h", "rt_intersect.h", "rt_traverse.h", "rt_shade.h", "rt_bvh.h", "rt_context.h",
           os.path.join("..", "..", "include", "realtrace_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_all(force=False, verbose=False):
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    if force or _stale(LIB, deps):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        if not os.path.exists(nvcc):
            nvcc = "nvcc"
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
              [os.path.join(CSRC, s) for s in SOURCES]
        subprocess.check_call(cmd, cwd=CSRC)
    build_host(force)
    return LIB


def build_variant(name, defines):
    """An A/B build of the CUDA library with extra -D flags (tuning runs: RT_LIB_PATH selects it in the Python
    harness).  Not part of build_all(); the shipped library is the one without extra flags."""
    out = os.path.join(HERE, f"librt_variant_{name}.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd, cwd=CSRC)
    return out


def build_host(force=False):
    """The C++ mirror of the reference's scene API (realtrace_b200/host) + its demo driver."""
    host = os.path.join(HERE, "host")
    src = os.path.join(host, "realtrace_host.cpp")
    if not os.path.exists(src):
        return None
    out = os.path.join(HERE, "librealtrace_host.so")
    deps = [os.path.join(host, f) for f in os.listdir(host)] + [LIB]
    if force or _stale(out, deps):
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-I", host,
                               "-I", os.path.join(HERE, "..", "include"), "-o", out, src,
                               "-L", HERE, "-lrealtrace_b200", "-lz", "-Wl,-rpath,$ORIGIN"], cwd=host)
        # the headless lumina-compatible driver (same includes and calls as Serial/lumina.cpp)
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-I", host, "-o", os.path.join(HERE, "lumina_headless"),
                               os.path.join(host, "lumina_headless.cpp"), "-L", HERE, "-lrealtrace_host",
                               "-lrealtrace_b200", "-Wl,-rpath,$ORIGIN"], cwd=host)
    return out


if __name__ == "__main__":
    print(build_all(force=True, verbose=False))
