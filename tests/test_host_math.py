"""Host-side pieces of the facade against the UNMODIFIED reference, on the CPU (no device work):

* the per-frame uniform builders (MakeModelMatrix / MakeNormalMatrix / MakeLookAtMatrix / MakePerspectiveMatrix /
  MakeOrthographicMatrix and mat x mat; reference geometry.cpp:60-68,92-179, geometry.h:784-793), bit for bit,
  on the parameter sets of parity.MATRIX_CASES (golden words produced by `ref_driver --matrices`);
* the TGA writer + Output::* (reference tgaimage.cpp:43-246, output.cpp:12-86, SURVEY.md §8 f N2): the files written by the
  facade after an oracle-backed frame are byte-identical with the reference's own files (md5 in the golden);
* a second camera / light set through frh_set_camera / frh_set_point_light instead of the .scene file."""
import hashlib
import os

import numpy as np
import pytest

import parity as P
from conftest import sha
from forkerrenderer_b200 import binding as B


def _words(host, case):
    t, rot, scale, eye, center, ratio = case[0:3], case[3], case[4], case[5:8], case[8:11], case[11]
    return host.test_matrices(t, rot, scale, eye, center, ratio).view(np.uint32)


def test_uniform_builders_match_reference_bits(oracle_host, golden):
    names = ["model"] * 16 + ["normal"] * 9 + ["lookat"] * 16 + ["persp"] * 16 + ["ortho"] * 16 + ["ortho*lookat"] * 16 + ["persp*lookat"] * 16
    assert len(golden["host_math"]) == len(P.MATRIX_CASES)
    for entry, case in zip(golden["host_math"], P.MATRIX_CASES):
        assert entry["case"] == [float(np.float32(v)) for v in case]
        want = np.array([int(w, 16) for w in entry["words"]], dtype=np.uint32)
        got = _words(oracle_host, case)
        bad = sorted({names[i] for i in np.nonzero(got != want)[0]})
        assert not bad, "case %s: %s differ from the reference" % (case, bad)


def test_uniform_builders_golden_is_the_reference_when_it_is_here(golden):
    if not P.have_ref():
        pytest.skip("oracle/_ref/ref_driver not built")
    for entry, case in zip(golden["host_math"], P.MATRIX_CASES):
        assert ["%08x" % w for w in P.run_reference_matrices(case)] == entry["words"]


def test_product_host_library_builders_match_too(golden):
    """The same check on libforkerhost.so (the product's facade) — pure host arithmetic, runs without a GPU."""
    if not (os.path.exists(B.HOST_LIB) and os.path.exists(B.CUDA_LIB)):
        pytest.skip("product libraries not built")
    host = B.Host(B.HOST_LIB)
    for entry, case in zip(golden["host_math"], P.MATRIX_CASES):
        want = np.array([int(w, 16) for w in entry["words"]], dtype=np.uint32)
        assert np.array_equal(_words(host, case), want), case


@pytest.mark.parametrize("cfg", ["c4_hard", "c2_hard"])
def test_tga_files_are_byte_identical_with_the_reference(cfg, oracle_host, golden, tmp_path):
    """Output::* after a frame (frh_output_tga): every file the reference writes for this config, md5 for md5.
    zbuffer.tga is left out: the reference converts FLT_MAX * 255 to uint8_t there (buffer.cpp:28), which is undefined
    behaviour — the float depth plane is compared instead (test_oracle_golden.py)."""
    scene, shadow, wrap, filt = P.CONFIGS[cfg]
    sc = oracle_host.load_scene(os.path.join(P.REPO, scene), P.ASSETS, wrap, filt)
    try:
        oracle_host.render(sc, shadow, True)
        oracle_host.output_tga(str(tmp_path))
    finally:
        sc.free()
    want = golden["tga_files"][cfg]
    checked = 0
    for name, md5 in want.items():
        if name == "zbuffer.tga":
            continue
        path = tmp_path / name
        assert path.exists(), name
        assert hashlib.md5(path.read_bytes()).hexdigest() == md5, "%s: %s differs from the reference's file" % (cfg, name)
        checked += 1
    assert checked >= 2


def test_camera_and_light_setters_equal_a_scene_file(oracle_host, golden):
    """frh_set_camera + frh_set_point_light on c1.scene == c1_cam2.scene (same models, another camera and light)."""
    sc = oracle_host.load_scene(os.path.join(P.REPO, "scenes/c1.scene"), P.ASSETS, 0, 0)
    try:
        oracle_host.set_camera(sc, (1.2, 0.6, 0.8), (0.1, -0.2, -1))
        oracle_host.set_point_light(sc, (-1.5, 4, 3), (1.5, 1.8, 2))
        oracle_host.render(sc, "pcss", True)
        f = oracle_host.fgl
        want = golden["c1_cam2_pcss"]["planes"]
        for name in ("depth", "shadow", "normal", "lightndc", "ids_camera", "ids_light", "frame", "frame_u8"):
            assert sha(f.read_plane(name)) == want[name]["sha256"], name
    finally:
        sc.free()


@pytest.mark.parametrize("case", sorted(P.FRAGMENT_CASES))
def test_host_shader_programs_match_the_reference(case, oracle_host, golden):
    """Shader::ProcessVertex / ProcessFragment of the four programs on the host (host/programs.cpp), with Texture::Sample /
    SampleFloat and Shadow::CalculateShadowVisibility (Hard / PCF / PCSS over the host's own mt19937 stream) behind them,
    against the reference's own programs evaluated on the same faces and barycentric points (reference shader.h:28-30,
    gshader.h, phongshader.h, pbrshader.h, texture.h:41-145, shadow.cpp:23-132): every output word, bit for bit."""
    scene, shadow, wrap, filt = P.FRAGMENT_CASES[case]
    sc = oracle_host.load_scene(os.path.join(P.REPO, scene), P.ASSETS, wrap, filt)
    try:
        got = oracle_host.test_fragments(sc, shadow).view(np.uint32)
    finally:
        sc.free()
    want = np.array([int(w, 16) for w in golden["fragments"][case]], dtype=np.uint32)
    assert got.shape == want.shape and len(want) > 100
    assert np.array_equal(got, want), "words %s differ" % np.nonzero(got != want)[0][:16]


def test_reference_mesh_draw_compiles_against_the_facade(tmp_path):
    """The reference's Mesh::Draw (mesh.cpp:10-25: shader.Use, 3 x ProcessVertex, ForkerGL::DrawTriangle) is source-compatible
    with the facade's headers.  Needs the reference sources (this container only)."""
    src = "/root/reference/src/mesh.cpp"
    if not os.path.exists(src):
        pytest.skip("reference sources are not here")
    import subprocess
    body = "".join(open(src).readlines()[9:25])
    assert "ProcessVertex" in body and "DrawTriangle" in body
    tu = tmp_path / "ref_mesh_draw.cpp"
    tu.write_text('#include "mesh.h"\n#include "forkergl.h"\n#include "shader.h"\n' + body)
    host = os.path.join(P.REPO, "forkerrenderer_b200", "host")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-fsyntax-only", "-I" + host, "-I" + os.path.join(P.REPO, "include"), str(tu)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
