"""ctypes bindings for the C ABI (include/forkergl_b200.h) and the host facade's C entry points (host/capi.cpp).

`Fgl` wraps any shared library exporting the `fgl_*` symbols; `Host` wraps a library exporting `frh_*` (the C++
facade: scene/OBJ/TGA loaders + Render::*).  The product pair is libforkergl_b200.so + libforkerhost.so (CUDA,
fails loudly without a GPU); tests additionally load oracle/liboracle_host.so, which carries the same facade
linked against the CPU oracle.  Nothing in this module falls back from one to the other.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)

# enums (mirror include/forkergl_b200.h)
MODE_FORWARD, MODE_DEFERRED = 0, 1
PASS_FORWARD, PASS_GEOMETRY, PASS_LIGHTING, PASS_SHADOW = 0, 1, 2, 3
WRAP_NOWRAP, WRAP_REPEAT, WRAP_MIRRORED_REPEAT, WRAP_CLAMP_TO_EDGE = 0, 1, 2, 3
FILTER_NEAREST, FILTER_LINEAR = 0, 1
SHADER_DEPTH, SHADER_G, SHADER_BLINN_PHONG, SHADER_PBR = 0, 1, 2, 3
SHADOW_HARD, SHADOW_PCF, SHADOW_PCSS = 0, 1, 2
BLUR_SIMPLE_3X3, BLUR_TWO_PASS_GAUSSIAN = 0, 1
(PLANE_FRAME, PLANE_DEPTH, PLANE_SHADOW, PLANE_NORMAL, PLANE_WORLDPOS, PLANE_LIGHTNDC, PLANE_ALBEDO, PLANE_EMISSIVE,
 PLANE_PARAM, PLANE_SHADINGTYPE, PLANE_AO, PLANE_FRAME_RGB8, PLANE_SSAA_RGB8, PLANE_PRIMID_CAMERA,
 PLANE_PRIMID_LIGHT) = range(15)
PLANE_NAMES = {
    "frame": PLANE_FRAME, "depth": PLANE_DEPTH, "shadow": PLANE_SHADOW, "normal": PLANE_NORMAL,
    "worldpos": PLANE_WORLDPOS, "lightndc": PLANE_LIGHTNDC, "albedo": PLANE_ALBEDO, "emissive": PLANE_EMISSIVE,
    "param": PLANE_PARAM, "shadingtype": PLANE_SHADINGTYPE, "ao": PLANE_AO, "frame_u8": PLANE_FRAME_RGB8,
    "ssaa_u8": PLANE_SSAA_RGB8, "ids_camera": PLANE_PRIMID_CAMERA, "ids_light": PLANE_PRIMID_LIGHT,
}
SHADOW_MODES = {"hard": SHADOW_HARD, "pcf": SHADOW_PCF, "pcss": SHADOW_PCSS}


class FglMaterial(C.Structure):
    _fields_ = [("ka", C.c_float * 3), ("kd", C.c_float * 3), ("ks", C.c_float * 3), ("ke", C.c_float * 3),
                ("pbr_ke", C.c_float * 3), ("albedo", C.c_float * 3), ("roughness", C.c_float),
                ("metalness", C.c_float), ("diffuse_map", C.c_int), ("specular_map", C.c_int),
                ("normal_map", C.c_int), ("emissive_map", C.c_int), ("base_color_map", C.c_int),
                ("roughness_map", C.c_int), ("metalness_map", C.c_int), ("ao_map", C.c_int),
                ("pbr_normal_map", C.c_int), ("pbr_emissive_map", C.c_int)]

    def __init__(self, **kw):
        super().__init__()
        for f in ("diffuse_map", "specular_map", "normal_map", "emissive_map", "base_color_map", "roughness_map",
                  "metalness_map", "ao_map", "pbr_normal_map", "pbr_emissive_map"):
            setattr(self, f, -1)
        self.albedo = (C.c_float * 3)(1, 1, 1)
        for k, v in kw.items():
            if isinstance(v, (tuple, list, np.ndarray)):
                v = (C.c_float * 3)(*[float(x) for x in v])
            setattr(self, k, v)


class FglUniforms(C.Structure):
    _fields_ = [("model", C.c_float * 16), ("view", C.c_float * 16), ("projection", C.c_float * 16),
                ("normal", C.c_float * 9), ("light_space", C.c_float * 16), ("light_position", C.c_float * 3),
                ("light_color", C.c_float * 3), ("eye_position", C.c_float * 3)]

    def __init__(self, **kw):
        super().__init__()
        eye4 = np.eye(4, dtype=np.float32)
        for f in ("model", "view", "projection", "light_space"):
            setattr(self, f, (C.c_float * 16)(*eye4.ravel()))
        self.normal = (C.c_float * 9)(*np.eye(3, dtype=np.float32).ravel())
        for k, v in kw.items():
            a = np.asarray(v, dtype=np.float32).ravel()
            setattr(self, k, (C.c_float * a.size)(*a))


class FglParams(C.Structure):
    _fields_ = [("shadow_mode", C.c_int), ("pcf_filter_size", C.c_double), ("pcss_blocker_filter_size", C.c_double),
                ("area_light_size", C.c_float), ("shadow_bias_slope", C.c_float), ("shadow_bias_min", C.c_float),
                ("shadow_intensity", C.c_float), ("ssao_radius", C.c_float), ("ssao_range_check_radius", C.c_float),
                ("ssao_bias", C.c_float), ("ssao_range_check", C.c_int), ("materialize_frame_f32", C.c_int)]


class FglTiming(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("ms_total", C.c_float), ("launches", C.c_int),
                ("algorithmic_bytes", C.c_uint64)]


class FglError(RuntimeError):
    pass


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Fgl:
    """One fgl context on top of a library exporting the C ABI."""

    def __init__(self, lib, device=0, ctx=None):
        self.lib = lib if not isinstance(lib, str) else C.CDLL(lib, mode=C.RTLD_LOCAL)
        L = self.lib
        L.fgl_last_error.restype = C.c_char_p
        L.fgl_last_error.argtypes = [C.c_void_p]
        L.fgl_backend_name.restype = C.c_char_p
        for name in ("fgl_create", "fgl_upload_texture", "fgl_upload_vertices", "fgl_upload_mesh", "fgl_draw_mesh",
                     "fgl_read_plane", "fgl_write_plane", "fgl_plane_info", "fgl_set_params", "fgl_blur",
                     "fgl_draw_screen_space_pixels", "fgl_copy_plane_rows_to_device", "fgl_get_timings"):
            getattr(L, name).restype = C.c_int
        L.fgl_read_plane.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.fgl_write_plane.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.fgl_copy_plane_rows_to_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.fgl_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.fgl_host_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.fgl_host_free.argtypes = [C.c_void_p, C.c_void_p]
        self._host_bufs = {}
        L.fgl_destroy.argtypes = [C.c_void_p]
        L.fgl_destroy.restype = None
        self._own = ctx is None
        if ctx is None:
            p = C.c_void_p()
            rc = L.fgl_create(int(device), C.byref(p))
            if rc != 0:
                raise FglError("fgl_create failed: %s" % L.fgl_last_error(None).decode())
            ctx = p
        self.ctx = ctx if isinstance(ctx, C.c_void_p) else C.c_void_p(ctx)

    def host_array(self, shape, dtype, key=None):
        """numpy array on page-locked memory from fgl_host_alloc (cached per key/shape/dtype, reused between calls)."""
        k = (key, tuple(shape), np.dtype(dtype).str)
        if k not in self._host_bufs:
            n = int(np.prod(shape)) * np.dtype(dtype).itemsize
            p = C.c_void_p()
            self.call("fgl_host_alloc", n, C.byref(p))
            buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
            self._host_bufs[k] = (p, np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape))
        return self._host_bufs[k][1]

    def close(self):
        if self.ctx:
            for p, _ in self._host_bufs.values():
                self.lib.fgl_host_free(self.ctx, p)
            self._host_bufs = {}
        if self._own and self.ctx:
            self.lib.fgl_destroy(self.ctx)
        self.ctx = None

    @property
    def backend(self):
        return self.lib.fgl_backend_name().decode()

    def _ck(self, rc, what):
        if rc != 0:
            raise FglError("%s failed (%d): %s" % (what, rc, self.lib.fgl_last_error(self.ctx).decode()))

    def call(self, name, *args):
        f = getattr(self.lib, name)
        f.restype = C.c_int
        self._ck(f(self.ctx, *args), name)

    # ---- resources
    def default_params(self):
        p = FglParams()
        self.lib.fgl_default_params(C.byref(p))
        return p

    def set_params(self, p):
        self.call("fgl_set_params", C.byref(p))

    def upload_texture(self, texels, wrap=WRAP_NOWRAP, filt=FILTER_NEAREST):
        """texels: uint8 array (H, W, bpp) in TGA memory order (row 0 first; B,G,R[,A] or grey)."""
        t = np.ascontiguousarray(texels, dtype=np.uint8)
        if t.ndim == 2:
            t = t[:, :, None]
        h, w, bpp = t.shape
        out = C.c_int(-1)
        self.call("fgl_upload_texture", t.ctypes.data_as(C.c_void_p), w, h, bpp, int(wrap), int(filt), C.byref(out))
        return out.value

    def upload_vertices(self, pos, uv, nrm, tan=None):
        pos, uv, nrm = _f32(pos).reshape(-1, 3), _f32(uv).reshape(-1, 2), _f32(nrm).reshape(-1, 3)
        tanp, ntan = None, 0
        if tan is not None:
            tan = _f32(tan).reshape(-1, 3)
            tanp, ntan = tan.ctypes.data_as(C.c_void_p), len(tan)
        out = C.c_int(-1)
        self.call("fgl_upload_vertices", pos.ctypes.data_as(C.c_void_p), len(pos), uv.ctypes.data_as(C.c_void_p),
                  len(uv), nrm.ctypes.data_as(C.c_void_p), len(nrm), tanp, ntan, C.byref(out))
        return out.value

    def upload_mesh(self, vertices_id, pos_idx, uv_idx, nrm_idx, material=None, has_tangents=False,
                    support_pbr=False):
        pi, ti, ni = _i32(pos_idx).ravel(), _i32(uv_idx).ravel(), _i32(nrm_idx).ravel()
        material = material or FglMaterial()
        out = C.c_int(-1)
        self.call("fgl_upload_mesh", int(vertices_id), len(pi) // 3, pi.ctypes.data_as(C.c_void_p),
                  ti.ctypes.data_as(C.c_void_p), ni.ctypes.data_as(C.c_void_p), C.byref(material),
                  int(has_tangents), int(support_pbr), C.byref(out))
        return out.value

    # ---- state / passes (thin)
    def init_frame_buffer(self, w, h): self.call("fgl_init_frame_buffer", w, h)
    def init_depth_buffer(self, w, h): self.call("fgl_init_depth_buffer", w, h)
    def init_shadow_buffer(self, w, h): self.call("fgl_init_shadow_buffer", w, h)
    def init_geometry_buffers(self, w, h): self.call("fgl_init_geometry_buffers", w, h)
    def clear_color(self, rgb): self.call("fgl_clear_color", (C.c_float * 3)(*rgb))
    def set_viewport(self, x, y, w, h): self.call("fgl_set_viewport", x, y, w, h)
    def set_view_projection_matrix(self, m): self.call("fgl_set_view_projection_matrix", (C.c_float * 16)(*_f32(m).ravel()))
    def set_light_space_matrix(self, m): self.call("fgl_set_light_space_matrix", (C.c_float * 16)(*_f32(m).ravel()))
    def set_render_mode(self, m): self.call("fgl_set_render_mode", int(m))
    def set_pass_type(self, p): self.call("fgl_set_pass_type", int(p))
    def set_shadow_status(self, on): self.call("fgl_set_shadow_status", int(bool(on)))
    def begin_frame(self): self.call("fgl_begin_frame")
    def set_row_band(self, y0, y1): self.call("fgl_set_row_band", int(y0), int(y1))
    def set_chain_blockers_before(self, k): self.call("fgl_set_chain_blockers_before", C.c_uint64(int(k)))

    def get_chain_blockers(self):
        v = C.c_uint64(0)
        self.call("fgl_get_chain_blockers", C.byref(v))
        return v.value

    def draw_mesh(self, mesh_id, kind, uniforms): self.call("fgl_draw_mesh", int(mesh_id), int(kind), C.byref(uniforms))
    def ssao(self): self.call("fgl_ssao")
    def blur(self, plane, kind): self.call("fgl_blur", int(plane), int(kind))
    def ssaa_resolve(self, k): self.call("fgl_ssaa_resolve", int(k))
    def sync(self): self.call("fgl_sync")
    def set_stream(self, stream_ptr): self.call("fgl_set_stream", C.c_void_p(stream_ptr))

    def get_matrix(self, which):
        out = (C.c_float * 16)()
        self.call("fgl_get_%s_matrix" % which, out)
        return np.array(out, dtype=np.float32).reshape(4, 4)

    def draw_screen_space_pixels(self, eye, light_pos, light_color):
        self.call("fgl_draw_screen_space_pixels", (C.c_float * 3)(*eye), (C.c_float * 3)(*light_pos),
                  (C.c_float * 3)(*light_color))

    # ---- device-side PCSS chain hand-off (sort-first groups)
    def chain_peer_mailbox(self):
        """(device pointer, 64-byte CUDA IPC handle) of this context's mailbox."""
        ptr, h = C.c_void_p(), (C.c_uint8 * 64)()
        self.call("fgl_chain_peer_mailbox", C.byref(ptr), h, C.c_size_t(64))
        return ptr.value, bytes(h)

    def chain_peer_connect(self, next_ptr=None, next_ipc=None, wait_prev=False, enable=True):
        buf = (C.c_uint8 * 64).from_buffer_copy(next_ipc) if next_ipc else None
        self.call("fgl_chain_peer_connect", C.c_void_p(next_ptr or 0), buf, int(wait_prev), int(enable))

    def prepare_screen_space_pixels(self, eye, light_pos, light_color, ssao_follows):
        self.call("fgl_prepare_screen_space_pixels", (C.c_float * 3)(*eye), (C.c_float * 3)(*light_pos), (C.c_float * 3)(*light_color),
                  int(bool(ssao_follows)))

    # ---- buffers
    def plane_info(self, plane):
        w, h, ch, b = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self.call("fgl_plane_info", int(plane), C.byref(w), C.byref(h), C.byref(ch), C.byref(b))
        return w.value, h.value, ch.value, b.value

    def read_plane(self, plane, pinned=False):
        """Returns (H, W) or (H, W, 3) numpy array; float32, uint8 (RGB8 images) or int32 (primitive ids).
        pinned=True reads into a cached page-locked buffer (valid until the next pinned read of the same plane)."""
        if isinstance(plane, str):
            plane = PLANE_NAMES[plane]
        w, h, ch, b = self.plane_info(plane)
        dt = np.uint8 if b == 1 else (np.int32 if plane in (PLANE_PRIMID_CAMERA, PLANE_PRIMID_LIGHT) else np.float32)
        shape = (h, w, ch) if ch > 1 else (h, w)
        a = self.host_array(shape, dt, key=plane) if pinned else np.empty(shape, dtype=dt)
        self.call("fgl_read_plane", int(plane), a.ctypes.data_as(C.c_void_p), a.nbytes)
        return a

    def write_plane(self, plane, arr):
        if isinstance(plane, str):
            plane = PLANE_NAMES[plane]
        a = _f32(arr)
        self.call("fgl_write_plane", int(plane), a.ctypes.data_as(C.c_void_p), a.nbytes)

    def copy_plane_rows_to_device(self, plane, y0, y1, dst_ptr, nbytes):
        self.call("fgl_copy_plane_rows_to_device", int(plane), int(y0), int(y1), C.c_void_p(dst_ptr), nbytes)

    # ---- recorded frames (CUDA graphs)
    def frame_record_begin(self): self.call("fgl_frame_record_begin")

    def frame_record_end(self, frame_id=-1):
        v = C.c_int(int(frame_id))
        self.call("fgl_frame_record_end", C.byref(v))
        return v.value

    def frame_record_abort(self): self.call("fgl_frame_record_abort")
    def frame_replay(self, frame_id): self.call("fgl_frame_replay", int(frame_id))
    def frame_release(self, frame_id): self.call("fgl_frame_release", int(frame_id))

    def frame_info(self, frame_id):
        """(graph nodes, kernel launches) of one replay."""
        a, b = C.c_int(0), C.c_int(0)
        self.call("fgl_frame_info", int(frame_id), C.byref(a), C.byref(b))
        return a.value, b.value

    # ---- instrumentation
    def enable_timing(self, on=True): self.call("fgl_enable_timing", int(on))
    def reset_timings(self): self.call("fgl_reset_timings")

    def timings(self):
        arr = (FglTiming * 64)()
        n = C.c_int(0)
        self.call("fgl_get_timings", arr, 64, C.byref(n))
        return [dict(name=arr[i].name.decode(), ms_total=arr[i].ms_total, launches=arr[i].launches,
                     algorithmic_bytes=arr[i].algorithmic_bytes) for i in range(n.value)]

    def transfer_bytes(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self.call("fgl_transfer_bytes", C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def launch_count(self):
        v = C.c_uint64(0)
        self.call("fgl_launch_count", C.byref(v))
        return v.value


class Scene:
    def __init__(self, host, handle, info):
        self.host, self.handle = host, handle
        (self.width, self.height, self.ssaa, self.ssaa_k, self.ssao, self.deferred, self.shadow,
         self.triangles) = info
        self.buffer_width = self.width * (self.ssaa_k if self.ssaa else 1)
        self.buffer_height = self.height * (self.ssaa_k if self.ssaa else 1)

    def free(self):
        if self.handle:
            self.host.lib.frh_scene_free(self.handle)
            self.handle = None


class Host:
    """The C++ facade (ForkerGL / Scene / Render::*) behind its C entry points (host/capi.cpp)."""

    def __init__(self, host_lib_path, fgl_lib_path=None):
        # the CUDA library must be loaded first (and globally) so that the facade's NEEDED entry binds to it
        self.fgl_lib = C.CDLL(fgl_lib_path, mode=C.RTLD_GLOBAL) if fgl_lib_path else None
        self.lib = C.CDLL(host_lib_path, mode=C.RTLD_LOCAL)
        self.lib.frh_last_error.restype = C.c_char_p
        self.lib.frh_context.restype = C.c_void_p
        self.lib.frh_scene_free.argtypes = [C.c_void_p]
        self.lib.frh_scene_free.restype = None
        self.lib.frh_render.argtypes = [C.c_void_p, C.c_int, C.c_int]
        self.lib.frh_render_begin.argtypes = [C.c_void_p, C.c_int, C.c_int]
        self.lib.frh_render_finish.argtypes = [C.c_void_p]
        self.lib.frh_scene_info.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.frh_set_camera.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self._fgl = None

    def _ck(self, rc, what):
        if rc != 0:
            raise FglError("%s: %s" % (what, self.lib.frh_last_error().decode()))

    @property
    def fgl(self):
        """The facade's singleton fgl context (created on first use; raises without a usable backend)."""
        if self._fgl is None:
            ctx = self.lib.frh_context()
            if not ctx:
                raise FglError("frh_context: %s" % self.lib.frh_last_error().decode())
            self._fgl = Fgl(self.fgl_lib or self.lib, ctx=ctx)
        return self._fgl

    def load_scene(self, scene_file, assets_dir, wrap=WRAP_NOWRAP, filt=FILTER_NEAREST):
        _ = self.fgl
        h = C.c_void_p()
        self._ck(self.lib.frh_scene_load(os.fsencode(assets_dir), os.fsencode(scene_file), int(wrap), int(filt),
                                         C.byref(h)), "frh_scene_load")
        info = (C.c_int * 8)()
        self._ck(self.lib.frh_scene_info(h, info), "frh_scene_info")
        return Scene(self, h, list(info))

    def render(self, scene, shadow_mode=SHADOW_PCSS, materialize_frame_f32=True):
        if isinstance(shadow_mode, str):
            shadow_mode = SHADOW_MODES[shadow_mode]
        self._ck(self.lib.frh_render(scene.handle, int(shadow_mode), int(bool(materialize_frame_f32))), "frh_render")

    def render_replay(self, scene, shadow_mode=SHADOW_PCSS, materialize_frame_f32=True):
        """frh_render_replay: the frame as a recorded CUDA graph (first call eager, second call records + replays, then one
        graph launch per call).  Returns True if this call was a graph launch; frames that cannot be recorded are rendered
        eagerly (replay_fallback_reason() says why)."""
        if isinstance(shadow_mode, str):
            shadow_mode = SHADOW_MODES[shadow_mode]
        done = C.c_int(0)
        self.lib.frh_render_replay.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        self._ck(self.lib.frh_render_replay(scene.handle, int(shadow_mode), int(bool(materialize_frame_f32)), C.byref(done)), "frh_render_replay")
        return bool(done.value)

    def replay_fallback_reason(self):
        self.lib.frh_replay_fallback_reason.restype = C.c_char_p
        return self.lib.frh_replay_fallback_reason().decode()

    def render_begin(self, scene, shadow_mode=SHADOW_PCSS, materialize_frame_f32=True):
        if isinstance(shadow_mode, str):
            shadow_mode = SHADOW_MODES[shadow_mode]
        self._ck(self.lib.frh_render_begin(scene.handle, int(shadow_mode), int(bool(materialize_frame_f32))), "frh_render_begin")

    def render_finish(self, scene):
        self._ck(self.lib.frh_render_finish(scene.handle), "frh_render_finish")

    def set_camera(self, scene, eye, look_at):
        self._ck(self.lib.frh_set_camera(scene.handle, (C.c_float * 3)(*eye), (C.c_float * 3)(*look_at)),
                 "frh_set_camera")

    # ---- sort-first group (host/capi.cpp frh_group_*) ----
    def group_export(self, scene):
        n = self.lib.frh_group_member_bytes()
        buf = C.create_string_buffer(n)
        self.lib.frh_group_export.argtypes = [C.c_void_p, C.c_void_p]
        self._ck(self.lib.frh_group_export(scene.handle, buf), "frh_group_export")
        return buf.raw

    def group_connect(self, rank, world, members, same_process=False):
        blob = b"".join(members)
        assert len(blob) == world * self.lib.frh_group_member_bytes()
        self.lib.frh_group_connect.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int]
        self._ck(self.lib.frh_group_connect(int(rank), int(world), blob, int(bool(same_process))), "frh_group_connect")

    def group_read_frame(self, height, width, out=None):
        """Rank 0: the finished frame as a (height, width, 3) uint8 array in page-locked host memory."""
        if out is None:
            out = self.fgl.host_array((height, width, 3), np.uint8, key="group_frame")
        self.lib.frh_group_read_frame.argtypes = [C.c_void_p, C.c_size_t]
        self._ck(self.lib.frh_group_read_frame(out.ctypes.data_as(C.c_void_p), out.nbytes), "frh_group_read_frame")
        return out

    def group_disconnect(self):
        self._ck(self.lib.frh_group_disconnect(), "frh_group_disconnect")

    def test_fragments(self, scene, shadow_mode):
        if isinstance(shadow_mode, str):
            shadow_mode = SHADOW_MODES[shadow_mode]
        cap = 1 << 16
        out, n = (C.c_float * cap)(), C.c_int(0)
        self.lib.frh_test_fragments.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        self._ck(self.lib.frh_test_fragments(scene.handle, int(shadow_mode), out, cap, C.byref(n)), "frh_test_fragments")
        assert n.value <= cap
        return np.array(out[: n.value], dtype=np.float32)

    def set_per_triangle_submission(self, on):
        self.lib.frh_set_per_triangle_submission.restype = None
        self.lib.frh_set_per_triangle_submission(int(bool(on)))

    def set_point_light(self, scene, position, color):
        self.lib.frh_set_point_light.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self._ck(self.lib.frh_set_point_light(scene.handle, (C.c_float * 3)(*position), (C.c_float * 3)(*color)), "frh_set_point_light")

    def output_tga(self, directory):
        self._ck(self.lib.frh_output_tga(os.fsencode(directory)), "frh_output_tga")

    def test_matrices(self, t, rot, scale, eye, center, ratio):
        out = (C.c_float * 105)()
        self.lib.frh_test_matrices.restype = None
        self.lib.frh_test_matrices((C.c_float * 3)(*t), C.c_float(rot), C.c_float(scale), (C.c_float * 3)(*eye),
                                   (C.c_float * 3)(*center), C.c_float(ratio), out)
        return np.array(out, dtype=np.float32)


# ---- the product libraries --------------------------------------------------------------------------------
CUDA_LIB = os.path.join(HERE, "libforkergl_b200.so")
HOST_LIB = os.path.join(HERE, "libforkerhost.so")


def product_host(instance=0):
    """The product: C++ facade over the CUDA library.  No fallback: raises if either is missing or no GPU.

    instance > 0: a further, independent instance of the facade in this process (its own ForkerGL statics, scene and fgl
    context) for frames in flight — the reference's API is a set of statics, so a second frame needs a second copy of the
    library image: the shared object is loaded once more from a private copy of the file.  All instances share the one
    libforkergl_b200.so."""
    for p in (CUDA_LIB, HOST_LIB):
        if not os.path.exists(p):
            raise FglError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % p)
    if instance == 0:
        return Host(HOST_LIB)
    import shutil
    import tempfile
    C.CDLL(CUDA_LIB, mode=C.RTLD_GLOBAL)  # the copy's DT_NEEDED entry resolves to the library that is already loaded
    d = tempfile.mkdtemp(prefix="fgl_host_%d_" % instance)
    copy = os.path.join(d, "libforkerhost_%d.so" % instance)
    shutil.copy(HOST_LIB, copy)
    return Host(copy)


def product_fgl(device=0):
    if not os.path.exists(CUDA_LIB):
        raise FglError("%s is missing — run `python -c 'import __graft_entry__ as g; g.build()'`" % CUDA_LIB)
    return Fgl(CUDA_LIB, device=device)
