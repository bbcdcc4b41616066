"""Where a kernel's time goes, from the SOURCE page of an .ncu-rep (needs -lineinfo / --import-source on):
the SASS lines with the most stall samples, the share of samples and of executed instructions between consecutive
block-wide barriers (the phases of a persistent kernel), and the executed-instruction mix per opcode.

usage: python tools/ncu_hotspots.py report.ncu-rep kernel_regex [top_n]"""
import collections
import csv
import subprocess
import sys


def main():
    rep, pattern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pattern], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next((r for r in rows if "Source" in r and "Instructions Executed" in r), None)
    if hdr is None:
        sys.exit("no source page for %r in %s" % (pattern, rep))
    si, sa, ie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    lines = []
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) > max(sa, ie) and r[sa].isdigit() and r[ie].isdigit():
            if lines and r[si].split()[:1] == lines[0][3].split()[:1] and len(lines) > 50 and r[si] == lines[0][3]:
                break  # the page repeats per profiled launch: keep the first
            lines.append((len(lines), int(r[sa]), int(r[ie]), r[si]))
    tot_s, tot_e = sum(l[1] for l in lines) or 1, sum(l[2] for l in lines) or 1
    print("%d SASS lines, %d stall samples, %d warp instructions executed" % (len(lines), tot_s, tot_e))
    print("\n-- top stall lines")
    for i, s, e, src in sorted(lines, key=lambda l: -l[1])[:top]:
        print("%6d  %5.1f %%  exec %-10d  #%-5d %s" % (s, 100.0 * s / tot_s, e, i, src[:90]))
    print("\n-- between block-wide barriers (BAR.SYNC): samples / executed instructions")
    prev = 0
    for b in [l[0] for l in lines if "BAR.SYNC" in l[3]] + [len(lines)]:
        s, e = sum(l[1] for l in lines[prev:b]), sum(l[2] for l in lines[prev:b])
        print("#%5d..%-5d  %5.1f %% of samples  %5.1f %% of instructions" % (prev, b, 100.0 * s / tot_s, 100.0 * e / tot_e))
        prev = b
    mix = collections.Counter()
    for _, _, e, src in lines:
        t = src.split()
        if t:
            mix[(t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]] += e
    print("\n-- executed-instruction mix")
    for op, e in mix.most_common(15):
        print("%-10s %5.1f %%" % (op, 100.0 * e / tot_e))


if __name__ == "__main__":
    main()
