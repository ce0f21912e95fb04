"""Top stall-sample instructions of one kernel of an .ncu-rep (SASS view).  python tools/ncu_src_top.py rep launch_index [n]"""
import csv, subprocess, sys, io
rep, idx = sys.argv[1], int(sys.argv[2]); n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
print(lines[0][:160])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]; body = []
for r in rows[1:]:
    if len(r) < len(hdr) - 2: continue
    if r[1] == 'Source': break
    body.append(r)
iS = hdr.index("# Samples"); iSrc = hdr.index("Source"); iEx = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS] or 0) for r in body)
print("total samples", tot, "instructions", len(body), "executed warp-inst", sum(int(r[iEx] or 0) for r in body))
agg = {}
for r in body:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print({k: round(100 * v / max(tot, 1), 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(body)), key=lambda k: -int(body[k][iS] or 0))[:n]
for k in sorted(order):
    r = body[k]
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{k:5d} {100 * int(r[iS]) / tot:5.1f}%  ex={r[iEx]:>9s}  {r[iSrc].strip()[:90]:90s} {st}")
