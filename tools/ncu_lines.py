"""Per-source-line share of stall samples and executed instructions from an .ncu-rep captured with --import-source on
(read here, no GPU): python tools/ncu_lines.py rep [min_pct]"""
import csv, io, subprocess, sys

def main(rep, min_pct=0.7):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    fname, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]; continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}; continue
        if hdr and r[0].isdigit() and r[hdr["# Samples"]].isdigit():
            lines.append((fname, int(r[0]), r[1], int(r[hdr["# Samples"]]), int(r[hdr["Instructions Executed"]])))
    ts = sum(l[3] for l in lines) or 1; ti = sum(l[4] for l in lines) or 1
    print("total samples %d, warp instructions %d" % (ts, ti))
    for f, n, src, s, i in lines:
        if 100.0 * s / ts >= min_pct or 100.0 * i / ti >= min_pct:
            print("%-18s %4d  samples %5.1f%%  inst %5.1f%%  %s" % (f, n, 100.0 * s / ts, 100.0 * i / ti, src.strip()[:120]))

if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.7)
