"""Per-source-line instruction / sample counts of one kernel from an ncu report taken with --import-source on:
    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [min_share]
(ncu -i REPORT --page source --csv --print-source cuda,sass; lines of every file are listed, hottest first)."""
import csv, subprocess, sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur, hdr, lines = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur, hdr = r[1], None
        elif r[0] == "Line No":
            hdr = r
        elif hdr is not None and r[0].isdigit() and len(r) > 8 and r[2] == "-":        # a CUDA source line (aggregated over its SASS)
            iex, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
            lines.append((cur.split("/")[-1], int(r[0]), int(r[iex] or 0), int(r[isamp] or 0), r[1].strip()))
    tot_i, tot_s = sum(l[2] for l in lines), sum(l[3] for l in lines)
    print("kernel %s: %d warp instructions, %d samples" % (kern, tot_i, tot_s))
    for f, ln, ins, smp, src in sorted(lines, key=lambda l: -l[2]):
        if ins >= min_share * tot_i or smp >= min_share * tot_s:
            print("%5.1f%% instr %5.1f%% samples  %s:%d  %s" % (100.0 * ins / tot_i, 100.0 * smp / max(tot_s, 1), f, ln, src[:100]))


if __name__ == "__main__":
    main()
