"""Top stalled SASS instructions of one launch of an .ncu-rep:  python tools/ncu_hot.py rep.ncu-rep LAUNCH_INDEX [N]"""
import csv, io, subprocess, sys
rep, k = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", k, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(i, r[isrc].strip(), int(r[iex] or 0), int(r[ismp] or 0)) for i, r in enumerate(rows[h + 1:]) if len(r) > ismp and r[ismp].isdigit()]
tot = sum(d[3] for d in data) or 1
print("kernel:", rows[0][1][:90] if rows and len(rows[0]) > 1 else "?", " total samples", tot, " warp insts", sum(d[2] for d in data))
for i, s, e, sm in sorted(data, key=lambda d: -d[3])[:n]:
    print(f"{i:5d} {sm / tot:6.1%} exec={e:9d}  {s[:110]}")
