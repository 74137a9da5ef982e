R=r02
O=gpurun_out
python bench.py --profile-kernels > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1_kernels.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/${R}_launches_raw.csv python bench.py --ncu-pass --steps 1 --warmup 1 > /dev/null 2>&1
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
python tools/block_sweep.py --batch 8 --dtype bf16 --out $O/${R}_block_sweep_bf16_b8_full.txt > /dev/null 2>&1
for b in 1 16 64; do python tools/block_sweep.py --batch $b --dtype bf16 --quick --out $O/${R}_block_sweep_bf16_b$b.txt > /dev/null 2>&1; done
python tools/block_sweep.py --batch 8 --dtype bf16 --quick --out $O/${R}_block_sweep_bf16_b8.txt > /dev/null 2>&1
python tools/timeline.py --list --out $O/${R}_timeline.txt > /dev/null 2>&1
tail -3 $O/${R}_block_sweep_bf16_b64.txt
