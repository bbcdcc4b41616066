"""Opcode histograms of the hot kernels of libforkergl_b200.so and the full listing of k_ssao (cuobjdump -sass; runs without a GPU).
usage: python tools/sass_report.py > profiles/r02c_sass_hot_kernels.txt"""
import collections
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "forkerrenderer_b200", "libforkergl_b200.so")
HOT = ["k_raster_blocksILi0ELi32", "k_raster_blocksILi0ELi8", "k_raster_smallILi0", "k_resolve_geometry", "k_resolve_shadow", "k_setupE", "k_ssaoILb1",
       "k_lighting_hardILi2", "k_pcss_visibilityE", "k_chunk_index", "k_pixel_masks", "k_chain_fused_r40", "k_blur_h", "k_blur_v"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    funcs, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", line):
            funcs[name].append(line)
    print("SASS of the hot kernels of forkerrenderer_b200/libforkergl_b200.so (cuobjdump -sass, sm_100a), final tree of round 2.")
    print("Per kernel: instruction count, opcode histogram, then the full listing of k_ssao.")
    allops = collections.Counter()
    for f, lines in funcs.items():
        for l in lines:
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
            if m:
                allops[m.group(1)] += 1
    tma = sum(v for k, v in allops.items() if k.startswith(("UTMA", "UBLKCP", "UTCMMA", "TCGEN")))
    print("Whole library: %d instructions; TMA / tcgen05 opcodes (UTMA*, UBLKCP, UTCMMA): %d; LDGSTS: %d; no stage of this path is a dense contraction"
          " or a tile copy the kernel waits for (DESIGN.md §9)." % (sum(allops.values()), tma, allops.get("LDGSTS", 0)))
    print()
    for key in HOT:
        for f, lines in funcs.items():
            if key in f:
                ops = collections.Counter()
                for l in lines:
                    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
                    if m:
                        ops[m.group(1)] += 1
                print("== %s: %d instructions" % (f, sum(ops.values())))
                print("   " + ", ".join("%s %d" % kv for kv in ops.most_common(18)))
    for f, lines in funcs.items():
        if "k_ssaoILb1" in f:
            print("\n==== " + f)
            for l in lines:
                print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l))


if __name__ == "__main__":
    main()
