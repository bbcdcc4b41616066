"""Generates tests/golden/reference_hashes.json by running the UNMODIFIED reference renderer (oracle/_ref/ref_driver,
built from /root/reference by oracle/Makefile) on every parity configuration and hashing its raw buffers.

    python tests/golden/make_golden.py            # all configs (C3 takes about a minute of CPU)

The reference has no tests or golden images of its own, so these fingerprints — sha256 of the raw
little-endian fp32 planes, 8-bit images and winner-id planes the reference produced in this container (g++ 13.3,
-O2, no FMA) — are what pins the oracle and, through it, the CUDA path.  A few sampled pixel values are stored next
to each hash to make mismatches debuggable."""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import parity as P  # noqa: E402


def fingerprint(a):
    a = np.ascontiguousarray(a)
    flat = a.reshape(-1)
    idx = np.linspace(0, flat.size - 1, 8).astype(np.int64)
    return {"sha256": hashlib.sha256(a.tobytes()).hexdigest(), "shape": list(a.shape), "dtype": str(a.dtype),
            "samples": {str(int(i)): (float(flat[i]) if a.dtype == np.float32 else int(flat[i])) for i in idx}}


def main():
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_hashes.json")
    golden = json.load(open(out_path)) if os.path.exists(out_path) else {}
    cfgs = sys.argv[1:] or (list(P.CONFIGS) + ["buffer_ops", "host_math", "tga_files", "fragments"])
    if "fragments" in cfgs:  # the reference's own Shader::ProcessVertex / ProcessFragment on selected faces (ref_driver --fragments)
        cfgs.remove("fragments")
        golden["fragments"] = {case: ["%08x" % w for w in P.run_reference_fragments(case)] for case in P.FRAGMENT_CASES}
        print("fragments ok", flush=True)
    if "host_math" in cfgs:  # the reference's host uniform builders (geometry.cpp:60-68,92-179), words as hex
        cfgs.remove("host_math")
        golden["host_math"] = [{"case": [float(np.float32(v)) for v in case], "words": ["%08x" % w for w in P.run_reference_matrices(case)]}
                               for case in P.MATRIX_CASES]
        print("host_math ok", flush=True)
    if "tga_files" in cfgs:  # md5 of the files the reference's Output::* writes (tgaimage.cpp:43-246, output.cpp:12-86)
        cfgs.remove("tga_files")
        golden["tga_files"] = {cfg: P.run_reference_tga(cfg) for cfg in ("c4_hard", "c2_hard")}
        print("tga_files ok", flush=True)
    if "buffer_ops" in cfgs:  # Buffer1f / Buffer3f SimpleBlurDenoised and TwoPassGaussianBlurDenoised (buffer.cpp:35-98, 140-203)
        cfgs.remove("buffer_ops")
        golden["buffer_ops"] = {k: {"buffer1f": fingerprint(v[0]), "buffer3f": fingerprint(v[1])} for k, v in sorted(P.run_reference_buffer_ops().items())}
        print("buffer_ops ok", flush=True)
    for cfg in cfgs:
        ref = P.run_reference(cfg)
        meta = ref.pop("meta")
        entry = {"scene": P.CONFIGS[cfg][0], "shadow": P.CONFIGS[cfg][1], "wrap": P.CONFIGS[cfg][2], "filter": P.CONFIGS[cfg][3],
                 "width": meta["width"], "height": meta["height"], "planes": {k: fingerprint(v) for k, v in sorted(ref.items())}}
        if P.CONFIGS[cfg][0].startswith("@"):
            entry["generated_obj_md5"] = P.generated_input_md5(cfg)
        golden[cfg] = entry
        print(cfg, "ok (reference frame %.2f s)" % meta["t_frame"], flush=True)
    json.dump(golden, open(out_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
