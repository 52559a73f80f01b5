"""Top warp-stall hot spots per SASS instruction of the k-th kernel in an ncu report captured with --import-source on.

    python profiles/stall_hotspots.py <x.ncu-rep> <launch-index> [top-n]

Prints total samples, the stall-reason totals and the top instructions (with the few instructions before each, for
context).  Reads the report with `ncu -i ... --page source --print-source sass --csv`; no GPU needed.
"""
import csv
import io
import subprocess
import sys

rep, k = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(k),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.splitlines()
print(lines[0])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
seen, body = set(), []
for r in rows[1:]:  # the header repeats per function, and the listing itself is printed twice
    if len(r) >= len(hdr) and r[0] != hdr[0] and r[0] not in seen:
        seen.add(r[0])
        body.append(r)
ix = {h: i for i, h in enumerate(hdr)}
samp = [int(r[ix["# Samples"]] or 0) for r in body]
total = sum(samp)
print("samples:", total)
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: sum(int(r[ix[h]] or 0) for r in body) for h in reasons}
print("stall totals:", ", ".join(f"{h[6:]} {v / max(total, 1):.1%}" for h, v in sorted(tot.items(), key=lambda x: -x[1]) if v > 0.01 * total))
order = sorted(range(len(body)), key=lambda i: -samp[i])[:top]
for i in order:
    r = body[i]
    why = sorted(((int(r[ix[h]] or 0), h[6:]) for h in reasons), reverse=True)[:2]
    print(f"{samp[i] / max(total, 1):6.1%}  {r[ix['Source']].strip():70s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
    for j in range(max(0, i - 2), i):
        print(f"          ^ {body[j][ix['Source']].strip()}")
