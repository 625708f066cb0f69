"""ctypes binding of the C ABI (include/realtrace_b200.h) for the Python harness.

There is no fallback of any kind here: if the CUDA library has not been built, or no CUDA
device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RT_LIB_PATH") or os.path.join(HERE, "librealtrace_b200.so")   # RT_LIB_PATH: tuning builds only

FLAG_BRUTE_FORCE, FLAG_COUNT_WORK, FLAG_PACKED_TILES, FLAG_WARP_TIMES = 1, 2, 4, 8
COMMIT_BUILD, COMMIT_REFIT = 0, 1


class RtMaterial(C.Structure):
    _fields_ = [("color", C.c_float * 3), ("ka", C.c_float), ("kd", C.c_float), ("ks", C.c_float),
                ("kr", C.c_float), ("kt", C.c_float), ("eta", C.c_float), ("flags", C.c_uint32)]


class RtCamera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("u", C.c_float * 3), ("v", C.c_float * 3), ("w", C.c_float * 3),
                ("focal_distance", C.c_float), ("aspect", C.c_float), ("width", C.c_int32), ("height", C.c_int32)]


class RtRenderParams(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("tile_w", C.c_int32), ("tile_h", C.c_int32), ("rank", C.c_int32),
                ("world_size", C.c_int32), ("flags", C.c_uint32), ("steal_pool_div", C.c_int32),
                ("frame_index", C.c_uint32), ("steal_cursor", C.c_void_p)]


class RtAuxOut(C.Structure):
    _fields_ = [("prim_id", C.c_void_p), ("t", C.c_void_p)]


class RtFrameStats(C.Structure):
    _fields_ = [("rays_primary", C.c_uint64), ("rays_shadow", C.c_uint64), ("rays_secondary", C.c_uint64),
                ("node_visits", C.c_uint64), ("tri_tests", C.c_uint64), ("shadow_node_visits", C.c_uint64),
                ("shadow_tri_tests", C.c_uint64), ("waves", C.c_uint32), ("tiles", C.c_uint32),
                ("kernel_launches", C.c_uint32), ("max_queue", C.c_uint32), ("stolen_blocks", C.c_uint32),
                ("reserved0", C.c_uint32), ("ms_device", C.c_float),
                ("ms_trace", C.c_float), ("ms_shadow", C.c_float), ("ms_shade", C.c_float),
                ("ms_secondary", C.c_float), ("ms_resolve", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class RtBuildStats(C.Structure):
    _fields_ = [("n_triangles", C.c_uint32), ("n_large_triangles", C.c_uint32), ("n_nodes", C.c_uint32),
                ("sort_passes", C.c_uint32), ("leaf_size", C.c_uint32), ("ms_build", C.c_float), ("ms_refit", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class RtMicrobench(C.Structure):
    _fields_ = [("l2_read_gbs", C.c_double), ("l2_random_node_gbs", C.c_double), ("l2_dependent_fetch_ns", C.c_double),
                ("hbm_read_gbs", C.c_double), ("fma_lane_instr_per_s", C.c_double), ("issue_warp_instr_per_s", C.c_double),
                ("implied_sm_mhz", C.c_double), ("sm_count", C.c_int32), ("l2_buffer_mib", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


ABI_SYMBOLS = [
    "rt_create_multi", "rt_device_count", "rt_host_alloc", "rt_host_free", "rt_render_enqueue", "rt_render_wait", "rt_microbench",
    "rt_create", "rt_destroy", "rt_last_error", "rt_set_stream", "rt_scene_set_triangles", "rt_scene_set_spheres",
    "rt_scene_set_planes", "rt_scene_set_cylinders", "rt_scene_set_materials", "rt_scene_set_lights",
    "rt_scene_set_environment", "rt_scene_commit", "rt_scene_update_vertices", "rt_scene_update_vertices_device", "rt_scene_build_stats", "rt_render",
    "rt_render_device", "rt_tile_layout", "rt_assemble_tiles", "rt_trace_rays", "rt_shade_rays", "rt_bvh_download",
    "rt_debug_sort_pairs", "rt_debug_warp_times", "rt_debug_frame_launches", "rt_synchronize", "rt_peer_sync", "rt_peer_barrier", "rt_render_push", "rt_shared_buffer_create", "rt_shared_buffer_open", "rt_download",
]

_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: the CUDA extension has not been built "
                           "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, u32p, fp = C.c_void_p, C.c_void_p, C.c_void_p
    lib.rt_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.rt_create_multi.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int]
    lib.rt_device_count.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.rt_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_uint64]
    lib.rt_host_free.argtypes = [vp]
    lib.rt_render_enqueue.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtRenderParams), vp, C.c_int32]
    lib.rt_render_wait.argtypes = [vp, C.c_int32]
    lib.rt_microbench.argtypes = [vp, C.POINTER(RtMicrobench)]
    lib.rt_destroy.argtypes = [vp]
    lib.rt_last_error.argtypes = [vp]
    lib.rt_last_error.restype = C.c_char_p
    lib.rt_set_stream.argtypes = [vp, vp]
    lib.rt_scene_set_triangles.argtypes = [vp, fp, u32p, fp, u32p, C.c_uint32]
    for n in ("rt_scene_set_spheres", "rt_scene_set_planes", "rt_scene_set_cylinders"):
        getattr(lib, n).argtypes = [vp, fp, u32p, u32p, C.c_uint32]
    lib.rt_scene_set_materials.argtypes = [vp, vp, C.c_uint32]
    lib.rt_scene_set_lights.argtypes = [vp, fp, C.c_uint32]
    lib.rt_scene_set_environment.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.rt_scene_commit.argtypes = [vp, C.c_int]
    lib.rt_scene_update_vertices.argtypes = [vp, fp, C.c_uint32]
    lib.rt_scene_update_vertices_device.argtypes = [vp, vp, C.c_uint32]
    lib.rt_scene_build_stats.argtypes = [vp, C.POINTER(RtBuildStats)]
    lib.rt_render.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtRenderParams), vp, C.POINTER(RtAuxOut),
                              C.POINTER(RtFrameStats)]
    lib.rt_render_device.argtypes = lib.rt_render.argtypes
    lib.rt_tile_layout.argtypes = [C.c_int32] * 6 + [C.POINTER(C.c_uint32)] * 3
    lib.rt_assemble_tiles.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]
    lib.rt_trace_rays.argtypes = [vp, fp, C.c_uint32, C.c_uint32, vp, vp]
    lib.rt_shade_rays.argtypes = [vp, fp, C.c_uint32, C.c_int32, C.c_uint32, vp]
    lib.rt_bvh_download.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.rt_debug_sort_pairs.argtypes = [vp, vp, vp, C.c_uint32]
    lib.rt_debug_warp_times.argtypes = [vp, vp, C.POINTER(C.c_uint32)]
    lib.rt_debug_frame_launches.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.rt_synchronize.argtypes = [vp]
    lib.rt_peer_sync.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_uint32, C.c_int32]
    lib.rt_peer_barrier.argtypes = [vp, vp, C.c_int32, C.c_uint32]
    lib.rt_render_push.argtypes = [vp, C.POINTER(RtCamera), C.POINTER(RtRenderParams), vp, vp, vp, C.c_uint32]
    lib.rt_shared_buffer_create.argtypes = [vp, C.c_uint64, C.POINTER(C.c_void_p), C.c_char_p]
    lib.rt_shared_buffer_open.argtypes = [vp, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.rt_download.argtypes = [vp, vp, vp, C.c_uint64]
    for n in ABI_SYMBOLS:
        if n != "rt_last_error":
            getattr(lib, n).restype = C.c_int
    _lib = lib
    return lib


class RtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"realtrace_b200 error {code}: {msg}")
        self.code = code


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


def camera_struct(cam) -> RtCamera:
    """Flatten a scene.Camera exactly like Camera's constructor does (camera.cpp:4-25).  An RtCamera is
    passed through, so per-frame callers can build it once."""
    if isinstance(cam, RtCamera):
        return cam
    u, v, w, focal, aspect = cam.basis()
    c = RtCamera()
    c.pos = (C.c_float * 3)(*[float(x) for x in cam.pos])
    c.u = (C.c_float * 3)(*[float(x) for x in u])
    c.v = (C.c_float * 3)(*[float(x) for x in v])
    c.w = (C.c_float * 3)(*[float(x) for x in w])
    c.focal_distance = float(focal)
    c.aspect = float(aspect)
    c.width, c.height = int(cam.width), int(cam.height)
    return c


def tile_layout(width, height, tile_w=0, tile_h=0, rank=0, world=1):
    lib = load_library()
    a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
    rc = lib.rt_tile_layout(width, height, tile_w, tile_h, rank, world, C.byref(a), C.byref(b), C.byref(c))
    if rc != 0:
        raise RtError(rc, "rt_tile_layout: bad arguments")
    return a.value, b.value, c.value


def host_alloc(nbytes):
    """Page-locked host buffer (rt_host_alloc) as a numpy uint8 array; free with host_free(array)."""
    lib = load_library()
    p = C.c_void_p()
    rc = lib.rt_host_alloc(C.byref(p), nbytes)
    if rc != 0:
        raise RtError(rc, "rt_host_alloc failed")
    arr = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p.value))
    return arr


def host_free(arr):
    load_library().rt_host_free(arr.ctypes.data)


class Context:
    """One context (rt_ctx): one GPU (device = ordinal) or several (devices = list of ordinals, or "all")."""

    def __init__(self, device=0, devices=None):
        self.lib = load_library()
        h = C.c_void_p()
        if devices is None:
            rc = self.lib.rt_create(C.byref(h), device)
        elif devices == "all":
            rc = self.lib.rt_create_multi(C.byref(h), None, 0)
        else:
            ids = (C.c_int * len(devices))(*devices)
            rc = self.lib.rt_create_multi(C.byref(h), ids, len(devices))
        if rc != 0:
            raise RtError(rc, self.lib.rt_last_error(None).decode())
        self.h = h
        self._keep = []

    def device_count(self):
        n = C.c_int32()
        self._check(self.lib.rt_device_count(self.h, C.byref(n)))
        return n.value

    def microbench(self):
        r = RtMicrobench()
        self._check(self.lib.rt_microbench(self.h, C.byref(r)))
        return r.as_dict()

    def render_enqueue(self, cam, max_depth, out, slot):
        """Pipelined host frame: enqueue render + copy into `out` (page-locked numpy array), do not wait."""
        c = camera_struct(cam)
        p = self._params(max_depth)
        self._check(self.lib.rt_render_enqueue(self.h, C.byref(c), C.byref(p), out.ctypes.data, slot))

    def render_wait(self, slot):
        self._check(self.lib.rt_render_wait(self.h, slot))

    def close(self):
        if getattr(self, "h", None):
            self.lib.rt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RtError(rc, self.lib.rt_last_error(self.h).decode())

    def set_stream(self, cuda_stream_ptr):
        self._check(self.lib.rt_set_stream(self.h, cuda_stream_ptr))

    # ---- scene
    def set_scene(self, s):
        lib, h = self.lib, self.h
        self._check(lib.rt_scene_set_triangles(h, _ptr(s.tri_v), _ptr(s.tri_material), _ptr(s.tri_rgb),
                                               _ptr(s.tri_object_id), len(s.tri_v)))
        self._check(lib.rt_scene_set_spheres(h, _ptr(s.sph), _ptr(s.sph_material), _ptr(s.sph_object_id), len(s.sph)))
        self._check(lib.rt_scene_set_planes(h, _ptr(s.pln), _ptr(s.pln_material), _ptr(s.pln_object_id), len(s.pln)))
        self._check(lib.rt_scene_set_cylinders(h, _ptr(s.cyl), _ptr(s.cyl_material), _ptr(s.cyl_object_id), len(s.cyl)))
        mats = np.ascontiguousarray(s.materials)
        assert mats.dtype.itemsize == C.sizeof(RtMaterial)
        self._check(lib.rt_scene_set_materials(h, mats.ctypes.data, len(mats)))
        self._check(lib.rt_scene_set_lights(h, _ptr(s.lights), len(s.lights)))
        amb = (C.c_float * 3)(*s.ambient)
        bg = (C.c_float * 3)(*s.background)
        self._check(lib.rt_scene_set_environment(h, amb, bg))

    def commit(self, mode=COMMIT_BUILD, want_stats=True):
        """BUILD uploads + builds the LBVH (synchronous); REFIT only enqueues the refit kernels —
        pass want_stats=False to keep it asynchronous."""
        self._check(self.lib.rt_scene_commit(self.h, mode))
        return self.build_stats() if want_stats else None

    def update_vertices(self, tri_v):
        v = np.ascontiguousarray(tri_v, np.float32).reshape(-1, 9)
        self._check(self.lib.rt_scene_update_vertices(self.h, v.ctypes.data, len(v)))

    def update_vertices_device(self, dev_ptr, n_triangles):
        """Vertices (n x 9 float32) already in device memory: asynchronous D2D copy on the context's stream."""
        self._check(self.lib.rt_scene_update_vertices_device(self.h, dev_ptr, n_triangles))

    def build_stats(self):
        st = RtBuildStats()
        self._check(self.lib.rt_scene_build_stats(self.h, C.byref(st)))
        return st.as_dict()

    # ---- rendering
    @staticmethod
    def _params(max_depth, tile=(0, 0), rank=0, world=1, flags=0, steal=None):
        p = RtRenderParams()
        p.max_depth, p.tile_w, p.tile_h, p.rank, p.world_size, p.flags = max_depth, tile[0], tile[1], rank, world, flags
        if steal is not None:      # (pool_div, frame_index, cursor device pointer)
            p.steal_pool_div, p.frame_index, p.steal_cursor = steal
        return p

    def render(self, cam, max_depth, aux=False, tile=(0, 0), rank=0, world=1, flags=0, out=None):
        """Host-buffer frame (rt_render).  Returns (rgb[H,W,3] u8, prim[H,W] i32|None, t[H,W] f32|None, stats)."""
        c = camera_struct(cam)
        p = self._params(max_depth, tile, rank, world, flags)
        W, H = cam.width, cam.height
        if flags & FLAG_PACKED_TILES:
            _, owned, tb = tile_layout(W, H, tile[0], tile[1], rank, world)
            rgb = np.zeros(owned * tb, np.uint8) if out is None else out
        else:
            rgb = np.zeros((H, W, 3), np.uint8) if out is None else out
        prim = t = None
        a = None
        if aux:
            prim = np.full((H, W), -1, np.int32)
            t = np.zeros((H, W), np.float32)
            a = RtAuxOut(prim.ctypes.data, t.ctypes.data)
        st = RtFrameStats()
        self._check(self.lib.rt_render(self.h, C.byref(c), C.byref(p), rgb.ctypes.data,
                                       C.byref(a) if a is not None else None, C.byref(st)))
        return rgb, prim, t, st.as_dict()

    def render_device(self, cam, max_depth, rgb_dev_ptr, tile=(0, 0), rank=0, world=1, flags=0, want_stats=True,
                      steal=None):
        """Device-buffer frame (rt_render_device); rgb_dev_ptr is a CUDA device pointer (int).
        steal = (pool_div, frame_index, cursor_dev_ptr) enables dynamic tile stealing."""
        c = camera_struct(cam)
        p = self._params(max_depth, tile, rank, world, flags, steal)
        st = RtFrameStats()
        self._check(self.lib.rt_render_device(self.h, C.byref(c), C.byref(p), rgb_dev_ptr, None,
                                              C.byref(st) if want_stats else None))
        return st.as_dict()

    def synchronize(self):
        self._check(self.lib.rt_synchronize(self.h))

    def shared_buffer_create(self, nbytes):
        """-> (device pointer, 64-byte IPC handle) of a buffer peers on this node can map."""
        p = C.c_void_p()
        handle = C.create_string_buffer(64)
        self._check(self.lib.rt_shared_buffer_create(self.h, nbytes, C.byref(p), handle))
        return p.value, handle.raw

    def shared_buffer_open(self, handle):
        p = C.c_void_p()
        self._check(self.lib.rt_shared_buffer_open(self.h, bytes(handle), C.byref(p)))
        return p.value

    def peer_sync(self, sync_ptr, rank, world, frame_index, phase):
        self._check(self.lib.rt_peer_sync(self.h, sync_ptr, rank, world, frame_index, phase))

    def peer_barrier(self, sync_ptr, world, epoch):
        self._check(self.lib.rt_peer_barrier(self.h, sync_ptr, world, epoch))

    def render_push(self, cam_struct, params_struct, packed_ptr, frame_ptr, sync_ptr, frame_index):
        """One multi-GPU frame step (render packed -> handshake -> push -> handshake), only enqueued.  Takes
        prebuilt RtCamera / RtRenderParams so that the per-frame host cost is one ctypes call."""
        self._check(self.lib.rt_render_push(self.h, C.byref(cam_struct), C.byref(params_struct), packed_ptr, frame_ptr,
                                            sync_ptr, frame_index))

    def download(self, dev_ptr, host_array):
        self._check(self.lib.rt_download(self.h, dev_ptr, host_array.ctypes.data, host_array.nbytes))

    def assemble_tiles(self, packed_dev_ptr, src_rank, world, width, height, frame_dev_ptr, tile=(0, 0)):
        self._check(self.lib.rt_assemble_tiles(self.h, packed_dev_ptr, src_rank, world, width, height, tile[0], tile[1],
                                               frame_dev_ptr))

    # ---- per-ray queries
    def trace_rays(self, rays, flags=0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        prim = np.zeros(len(rays), np.int32)
        t = np.zeros(len(rays), np.float32)
        self._check(self.lib.rt_trace_rays(self.h, rays.ctypes.data, len(rays), flags, prim.ctypes.data, t.ctypes.data))
        return prim, t

    def shade_rays(self, rays, max_depth, flags=0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        out = np.zeros((len(rays), 3), np.float32)
        self._check(self.lib.rt_shade_rays(self.h, rays.ctypes.data, len(rays), max_depth, flags, out.ctypes.data))
        return out

    # ---- introspection
    def bvh_download(self):
        nn, nb = C.c_uint32(), C.c_uint32()
        self._check(self.lib.rt_bvh_download(self.h, None, None, None, C.byref(nn), C.byref(nb)))
        nodes = np.zeros((nn.value, 16), np.float32)
        order = np.zeros(nb.value, np.uint32)
        keys = np.zeros(nb.value, np.uint64)
        self._check(self.lib.rt_bvh_download(self.h, _ptr(nodes), _ptr(order), _ptr(keys), C.byref(nn), C.byref(nb)))
        return nodes, order, keys

    def frame_launches(self):
        """Running count of the kernels enqueued by render / render_device / render_push / peer_sync / assemble_tiles."""
        n = C.c_uint64(0)
        self._check(self.lib.rt_debug_frame_launches(self.h, C.byref(n)))
        return n.value

    def warp_times(self, max_warps=1 << 16):
        """(n, 2) uint64 {start_ns, end_ns} per warp of the last primary traversal kernel (FLAG_WARP_TIMES)."""
        out = np.zeros((max_warps, 2), np.uint64)
        n = C.c_uint32(max_warps)
        self._check(self.lib.rt_debug_warp_times(self.h, out.ctypes.data, C.byref(n)))
        return out[:n.value]

    def sort_pairs(self, keys, values):
        keys = np.ascontiguousarray(keys, np.uint64).copy()
        values = np.ascontiguousarray(values, np.uint32).copy()
        self._check(self.lib.rt_debug_sort_pairs(self.h, _ptr(keys), _ptr(values), len(keys)))
        return keys, values
