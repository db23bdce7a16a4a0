"""Per-source-line stall samples of one kernel of an ncu report captured with --import-source on (built with -lineinfo).

    python profiles/srcview.py gpurun_out/X.ncu-rep <kernel regex> [top N]
"""
import csv
import io
import os
import subprocess
import sys


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kre], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname, lines, stall_cols, h = "?", [], [], None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = os.path.basename(r[1])
        elif r[0] == "Line No":
            h = r
            stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        elif h is not None and r[0].isdigit() and len(r) > 7 and r[4].isdigit():
            lines.append((fname, int(r[0]), r[1], int(r[4]), int(r[7]), r))
    tot = sum(l[3] for l in lines)
    tin = sum(l[4] for l in lines)
    print(f"samples {tot}  warp instructions {tin}")
    for f, ln, src, s, ins, r in sorted(lines, key=lambda l: -l[3])[:top]:
        why = sorted(((int(r[i]), c[6:]) for i, c in stall_cols if i < len(r) and r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
        print(f"{100 * s / tot:5.1f}%  instr {100 * ins / max(tin, 1):5.1f}%  {f}:{ln:<5d} {src.strip()[:100]}   [{', '.join(f'{c} {n}' for n, c in why)}]")


if __name__ == "__main__":
    main()
