"""Synthetic scenes submitted straight through the C ABI (no asset files): a large unclipped ground plane plus a
procedurally generated height-field mesh with procedural textures, rendered with the pass sequence of the
reference's Render::Render (render.cpp:40-58).  Used by smoke(), by the unit tests (same submission against the CUDA
library and the CPU oracle) and by bench.py for the high-triangle-count configuration.

Matrices are built here in float32; both back ends receive the same uniform bits, so the builders do not have to
match the reference's host math (the C++ facade does that for .scene files)."""
import numpy as np

from . import binding as B


def look_at(eye, center, up=(0, 1, 0)):
    eye, center, up = (np.asarray(v, dtype=np.float32) for v in (eye, center, up))
    f = center - eye
    f = f / np.linalg.norm(f)
    s = np.cross(f, up)
    s = s / np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4, dtype=np.float32)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[0, 3], m[1, 3], m[2, 3] = -np.dot(s, eye), -np.dot(u, eye), np.dot(f, eye)
    return m.astype(np.float32)


def perspective(fov_deg, ratio, n, f):
    t = np.tan(np.radians(fov_deg) / 2)
    m = np.zeros((4, 4), dtype=np.float32)
    m[0, 0] = 1 / (ratio * t)
    m[1, 1] = 1 / t
    m[2, 2] = -(f + n) / (f - n)
    m[2, 3] = -2 * f * n / (f - n)
    m[3, 2] = -1
    return m


def orthographic(l, r, b, t, n, f):
    m = np.eye(4, dtype=np.float32)
    m[0, 0], m[1, 1], m[2, 2] = 2 / (r - l), 2 / (t - b), -2 / (f - n)
    m[0, 3], m[1, 3], m[2, 3] = -(r + l) / (r - l), -(t + b) / (t - b), -(f + n) / (f - n)
    return m


def model_matrix(translate=(0, 0, 0), rot_y_deg=0.0, scale=1.0):
    c, s = np.cos(np.radians(rot_y_deg)), np.sin(np.radians(rot_y_deg))
    r = np.array([[c, 0, s, 0], [0, 1, 0, 0], [-s, 0, c, 0], [0, 0, 0, 1]], dtype=np.float32)
    sc = np.diag([scale, scale, scale, 1]).astype(np.float32)
    t = np.eye(4, dtype=np.float32)
    t[:3, 3] = translate
    return (t @ r @ sc).astype(np.float32)


def normal_matrix(model):
    return np.linalg.inv(model[:3, :3].astype(np.float64)).T.astype(np.float32)


def height_field(n, seed=20261017, amplitude=0.15):
    """n x n quads over x,z in [-1,1]: positions, uvs (x8, exercises Repeat), analytic normals, per-vertex tangents."""
    rng = np.random.RandomState(seed)
    phi, psi = rng.uniform(0, 2 * np.pi, 4), rng.uniform(0, 2 * np.pi, 4)
    g = np.linspace(-1, 1, n + 1, dtype=np.float64)
    x, z = np.meshgrid(g, g, indexing="xy")
    y = np.zeros_like(x)
    dydx, dydz = np.zeros_like(x), np.zeros_like(x)
    for k in range(1, 5):
        a, w = amplitude * 2.0 ** -k, 2.0 ** k * np.pi
        y += a * np.sin(w * x + phi[k - 1]) * np.cos(w * z + psi[k - 1])
        dydx += a * w * np.cos(w * x + phi[k - 1]) * np.cos(w * z + psi[k - 1])
        dydz += -a * w * np.sin(w * x + phi[k - 1]) * np.sin(w * z + psi[k - 1])
    pos = np.stack([x, y, z], -1).reshape(-1, 3).astype(np.float32)
    nrm = np.stack([-dydx, np.ones_like(x), -dydz], -1).reshape(-1, 3)
    nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    tan = np.stack([np.ones_like(x), dydx, np.zeros_like(x)], -1).reshape(-1, 3)
    tan = (tan / np.linalg.norm(tan, axis=1, keepdims=True)).astype(np.float32)
    uv = (np.stack([x, z], -1).reshape(-1, 2) * 4 + 4).astype(np.float32)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="xy")
    v00 = (j * (n + 1) + i).ravel()
    v10, v01, v11 = v00 + 1, v00 + n + 1, v00 + n + 2
    idx = np.stack([v00, v01, v11, v00, v11, v10], -1).reshape(-1, 3).astype(np.int32)
    return pos, uv, nrm, tan, idx


def procedural_texture(size, seed, channels=3):
    rng = np.random.RandomState(seed)
    g = np.linspace(0, 1, size, endpoint=False)
    u, v = np.meshgrid(g, g)
    base = 0.5 + 0.25 * np.sin(2 * np.pi * 3 * u) * np.cos(2 * np.pi * 2 * v)
    img = np.stack([np.clip(base + 0.2 * rng.rand(size, size) - 0.1 + 0.1 * c, 0, 1) for c in range(channels)], -1)
    return (img * 255).astype(np.uint8)


class SyntheticScene:
    """Ground plane (2 unclipped triangles, like the reference's obj/plane) + height field; optionally PBR."""

    def __init__(self, fgl, quads=24, pbr=False, textured=True, wrap=B.WRAP_REPEAT, filt=B.FILTER_NEAREST, tex_size=64):
        self.f = fgl
        f = fgl
        # plane: large quad at y = -0.6 whose far corners project far outside the screen
        ppos = np.array([[-3, 0, -3], [3, 0, -3], [3, 0, 3], [-3, 0, 3]], dtype=np.float32)
        puv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=np.float32)
        pnrm = np.array([[0, 1, 0]], dtype=np.float32)
        pv = f.upload_vertices(ppos, puv, pnrm)
        pidx = np.array([[0, 2, 1], [0, 3, 2]], dtype=np.int32)
        pmat = B.FglMaterial(ka=(0.2, 0.2, 0.2), kd=(0.55, 0.5, 0.45), ks=(0.3, 0.3, 0.3))
        self.plane = f.upload_mesh(pv, pidx, pidx, np.zeros_like(pidx), pmat)
        self.plane_model = model_matrix((0, -0.6, -1), 0, 1)

        pos, uv, nrm, tan, idx = height_field(quads)
        hv = f.upload_vertices(pos, uv, nrm, tan)
        mat = B.FglMaterial(ka=(0.3, 0.3, 0.3), kd=(0.7, 0.4, 0.3), ks=(0.5, 0.5, 0.5), roughness=0.6, metalness=0.2,
                            albedo=(0.8, 0.6, 0.4))
        if textured:
            diffuse = f.upload_texture(procedural_texture(tex_size, 1), wrap, filt)
            spec = f.upload_texture(procedural_texture(tex_size, 2, 1)[:, :, 0], wrap, filt)
            nmap = procedural_texture(tex_size, 3)
            nmap[:, :, 0] = 200 + nmap[:, :, 0] // 5   # TGA order B,G,R: keep the normal mostly along +z (b channel = z)
            nrm_tex = f.upload_texture(nmap, wrap, filt)
            mat.diffuse_map, mat.specular_map, mat.normal_map = diffuse, spec, nrm_tex
            if pbr:
                mat.base_color_map, mat.pbr_normal_map = diffuse, nrm_tex
                mat.roughness_map = f.upload_texture(procedural_texture(tex_size, 4, 1)[:, :, 0], wrap, filt)
                mat.metalness_map = f.upload_texture(procedural_texture(tex_size, 5, 1)[:, :, 0], wrap, filt)
        self.field = f.upload_mesh(hv, idx, idx, idx, mat, has_tangents=True, support_pbr=pbr)
        self.field_model = model_matrix((0, -0.35, -1), 25, 0.8)
        self.triangles = 2 + len(idx)
        self.eye, self.center = (-1.0, 1.0, 1.0), (0.0, 0.0, -1.0)
        self.light_pos, self.light_color = (2.0, 5.0, 5.0), (2.0, 2.0, 2.0)

    def meshes(self):
        return [(self.plane, self.plane_model), (self.field, self.field_model)]

    def render(self, W, H, **kw):
        """The pass sequence of Render::Render (render.cpp:40-58) for a W x H output (raster size W*ssaa x H*ssaa)."""
        self.render_begin(W, H, **kw)
        self.render_finish()

    def render_begin(self, W, H, shadow_mode=B.SHADOW_HARD, ssao=False, ssaa=1, forward=False, shadow=True,
                     materialize_frame_f32=True, band=None):
        """Everything before the lighting loop (a sort-first driver exchanges the PCSS chain state before render_finish)."""
        f = self.f
        self._pending = (forward, ssaa)
        BW, BH = W * ssaa, H * ssaa
        p = f.default_params()
        p.shadow_mode = shadow_mode
        p.materialize_frame_f32 = int(materialize_frame_f32)
        f.set_params(p)
        f.set_shadow_status(shadow)
        f.set_render_mode(B.MODE_FORWARD if forward else B.MODE_DEFERRED)
        f.begin_frame()
        f.set_row_band(*(band if band else (0, -1)))
        f.set_viewport(0, 0, BW, BH)
        ratio = W / H
        ls = (orthographic(-3 * ratio, 3 * ratio, -3, 3, 0.1, 20) @ look_at(self.light_pos, (0, 0, 0))).astype(np.float32)
        view = look_at(self.eye, self.center)
        proj = perspective(45, ratio, 0.01, 20)
        if shadow:  # DoShadowPass, render.cpp:60-97
            f.init_shadow_buffer(BW, BH)
            f.init_depth_buffer(BW, BH)
            f.set_pass_type(B.PASS_SHADOW)
            f.set_light_space_matrix(ls)
            for mesh, model in self.meshes():
                f.draw_mesh(mesh, B.SHADER_DEPTH, B.FglUniforms(model=model, light_space=ls))
        un = lambda model: B.FglUniforms(model=model, view=view, projection=proj, normal=normal_matrix(model), light_space=ls,
                                         light_position=self.light_pos, light_color=self.light_color, eye_position=self.eye)
        if forward:  # DoForwardPass, render.cpp:99-157
            f.init_frame_buffer(BW, BH)
            f.init_depth_buffer(BW, BH)
            f.clear_color((0.12, 0.12, 0.12))
            f.set_pass_type(B.PASS_FORWARD)
            f.set_view_projection_matrix(proj @ view)
            for i, (mesh, model) in enumerate(self.meshes()):
                f.draw_mesh(mesh, B.SHADER_BLINN_PHONG, un(model))
        else:  # DoGeometryPass + DoLightingPass, render.cpp:159-212
            f.init_geometry_buffers(BW, BH)
            f.init_depth_buffer(BW, BH)
            f.set_pass_type(B.PASS_GEOMETRY)
            f.set_view_projection_matrix(proj @ view)
            for mesh, model in self.meshes():
                f.draw_mesh(mesh, B.SHADER_G, un(model))
            f.init_frame_buffer(BW, BH)
            f.set_pass_type(B.PASS_LIGHTING)
            f.clear_color((0.12, 0.12, 0.12))
            # a head start for the lighting pass: everything that does not need the previous band's chain state, and — when
            # that state is known or arrives on the device — the PCSS chain itself, overlapped with SSAO and the blur
            f.prepare_screen_space_pixels(self.eye, self.light_pos, self.light_color, ssao)
            if ssao:
                f.ssao()
                f.blur(B.PLANE_AO, B.BLUR_TWO_PASS_GAUSSIAN)

    def render_finish(self):
        f = self.f
        forward, ssaa = self._pending
        if not forward:
            f.draw_screen_space_pixels(self.eye, self.light_pos, self.light_color)
        if ssaa > 1:
            f.ssaa_resolve(ssaa)


def submit_test_frame(fgl, W, H, shadow_mode=B.SHADOW_HARD, **kw):
    scene_kw = {k: kw.pop(k) for k in ("quads", "pbr", "textured", "wrap", "filt") if k in kw}
    s = SyntheticScene(fgl, **scene_kw)
    s.render(W, H, shadow_mode=shadow_mode, **kw)
    return s
