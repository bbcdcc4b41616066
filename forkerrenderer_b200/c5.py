"""The synthetic high-triangle-count workload C5 as files on disk: an OBJ + MTL + TGA textures + .scene written by
tools/gen_c5.cpp (built into forkerrenderer_b200/bin/gen_c5 by __graft_entry__.build()), so that the SAME input goes through the
reference's loaders (oracle/_ref/ref_driver) and through the facade's (frh_scene_load).  The files are never committed (the full
instance is 0.96 GB of text); they are generated where they are needed, deterministically (md5 of the OBJ is pinned in
tests/golden/reference_hashes.json for the reduced instance)."""
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GEN = os.path.join(HERE, "bin", "gen_c5")
REF_ASSETS = os.path.join(REPO, "oracle", "_ref", "assets")

# name: (quads per side, width, height)
INSTANCES = {
    "c5": (2237, 7680, 4320),        # 10 008 338 triangles, 8K: the north-star scaling configuration
    "c5_small": (700, 1920, 1080),   # 980 000 triangles
    "c5_golden": (500, 1920, 1080),  # 500 000 triangles: the instance the reference itself rendered for the golden fingerprints
    "c5_cpu": (280, 960, 540),       # 156 800 triangles: bounded CPU sample of the bench's cpu_baseline leg
}


def ensure(quads, width, height, root=None, ssao=True, mode="deferred"):
    """Generates (once per root) the assets and the scene; returns (assets_dir, scene_path).  `assets_dir` also links the
    reference's obj/plane (the scene's ground plane, reference README.md:70-85)."""
    if not os.path.exists(GEN):
        raise RuntimeError("forkerrenderer_b200/bin/gen_c5 is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    root = root or os.path.join(tempfile.gettempdir(), "fgl_c5_assets")
    os.makedirs(os.path.join(root, "obj"), exist_ok=True)
    os.makedirs(os.path.join(root, "output"), exist_ok=True)
    link = os.path.join(root, "obj", "plane")
    if not os.path.exists(link):
        src = os.path.join(REF_ASSETS, "obj", "plane")
        if not os.path.isdir(src):
            raise RuntimeError("reference assets are not staged (oracle/_ref/assets/obj/plane)")
        try:
            os.symlink(src, link)
        except FileExistsError:
            pass
    scene = os.path.join(root, "c5_%d_%dx%d_%s_%s.scene" % (quads, width, height, "ssao" if ssao else "nossao", mode))
    obj = os.path.join(root, "obj", "c5_%d" % quads, "field.obj")
    done = obj + ".done"
    if not (os.path.exists(done) and os.path.exists(scene)):
        subprocess.run([GEN, str(quads), str(width), str(height), root, scene, "on" if ssao else "off", mode], check=True)
        open(done, "w").close()
    return root, scene
