"""Hot source lines of an .ncu-rep captured with --import-source on: stall samples per CUDA source line, per kernel.
usage: python tools/ncu_lines.py <file>.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout
cur, out = None, {}
for r in csv.reader(raw.splitlines()):
    if len(r) >= 2 and r[0] == "Function Name":
        cur = r[1]; out.setdefault(cur, []); continue
    if cur and len(r) > 8 and r[0].isdigit() and r[2] == "-":
        try:
            out[cur].append((int(r[4] or 0), int(r[7] or 0), int(r[0]), r[1].strip()))
        except ValueError:
            pass
for k, rows in out.items():
    tot = sum(x[0] for x in rows) or 1; ins = sum(x[1] for x in rows) or 1
    print(f"## {k[:90]}  samples {tot}  warp-instr {ins}")
    for s, n, ln, src in sorted(rows, key=lambda x: -x[0])[:top]:
        print(f"{s * 100 / tot:5.1f}% smp {n * 100 / ins:5.1f}% ins  L{ln:<4d} {src[:120]}")
