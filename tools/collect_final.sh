#!/bin/bash
# Reduced end-of-round collection (one B200, ~8 GPU-minutes): bench line + kernel table, ncu launch list, ncu --set full of the
# dominant kernel and of the kernels new this session, timeline, sanitizer on the new kernels, block sweep.
#   bash tools/collect_final.sh r02
R=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py --profile-kernels > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1_kernels.txt
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${R}_launches_raw.csv \
    python bench.py --ncu-pass --steps 1 --warmup 1 > /dev/null 2>&1
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
for c in s3_mlpf s1_mlpf s3_tmc; do
  timeout 120 ncu --set full --clock-control none -k regex:"token_mixer|mlp_fused|conv_tc" -s 3 -c 1 -o /tmp/ncu_$c python tools/microbench.py --case $c --iters 2 > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/ncu_$c.ncu-rep --out /tmp/ncu_$c.csv
done
timeout 120 ncu --set full --clock-control none -k regex:"patch_embed" -s 1 -c 1 -o /tmp/ncu_pe python -m pytest tests/test_gpu_engine.py -q -m gpu -k "patch_embed and 2x4x2x512" > /dev/null 2>&1
python tools/ncu_summary.py /tmp/ncu_pe.ncu-rep --out /tmp/ncu_pe.csv
head -1 /tmp/ncu_s3_mlpf.csv > $O/${R}_ncu_dominant.csv
for c in s3_mlpf s1_mlpf s3_tmc pe; do tail -n +2 /tmp/ncu_$c.csv | sed "s/^/$c: /" >> $O/${R}_ncu_dominant.csv; done
timeout 200 python tools/microbench.py --case s1_tmf s2_tmf s3_tmc s1_mlpf s2_mlpf s3_mlpf s3_mlp1 s3_mlp2 s4_mlp1 s4_mlp2 > $O/${R}_microbench.txt 2>&1
timeout 200 python tools/block_sweep.py --batch 8 --dtype bf16 --quick --out $O/${R}_block_sweep_bf16_b8.txt > /dev/null 2>&1
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_engine.py tests/test_gpu_session.py tests/test_gpu_parity.py -m gpu -q -x \
    -k "fused_mlp or row_tap or im2col_rows or patch_embed or concurrent or radar_enhance_concat or tap_major" > $O/${R}_sanitizer_memcheck.log 2>&1
tail -5 $O/${R}_sanitizer_memcheck.log
timeout 120 python tools/timeline.py --list --out $O/${R}_timeline.txt > /dev/null 2>&1
ls -la $O | tail -12
