R=r02; O=gpurun_out
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${R}_launches_raw.csv \
    python bench.py --ncu-pass --steps 1 --warmup 2 > /dev/null 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("$O/${R}_launches_raw.csv", errors="replace")))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ix = {k: i for i, k in enumerate(rows[h])}
out = [["id", "kernel", "grid", "block", "time_us"]]
for r in rows[h + 1:]:
    if len(r) > ix["Metric Value"] and r[ix["Metric Name"]] == "gpu__time_duration.sum":
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")[:70]
        out.append([r[ix["ID"]], name, r[ix["Grid Size"]], r[ix["Block Size"]], f"{float(r[ix['Metric Value']].replace(',', '')) / 1e3:.2f}"])
csv.writer(open("$O/${R}_launches.csv", "w", newline="")).writerows(out)
print("launches:", len(out) - 1)
PY
rm -f $O/${R}_launches_raw.csv
