#!/bin/bash
# Round artefacts for profiles/ (run under gpurun on one B200; everything it writes is small text):
#   bash tools/collect_profiles.sh r02
# 1. the default bench line + per-kernel table, 2. the launch list of the same forward (ncu, durations only),
# 3. ncu --set full summaries of the dominant kernels (microbench cases), 4. block sweep, 5. compute-sanitizer on the newest kernels.
R=${1:-r02}
O=gpurun_out
mkdir -p $O
python bench.py --profile-kernels > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1_kernels.txt
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${R}_launches_raw.csv \
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
for c in s1_tmf s2_tmf s3_tmc s1_mlpf s2_mlpf s3_mlpf s3_mlp1 s3_mlp2; do
  ncu --set full --clock-control none -k regex:"token_mixer|mlp_fused|conv_tc" -s 3 -c 1 -o /tmp/ncu_$c python tools/microbench.py --case $c --iters 2 > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/ncu_$c.ncu-rep --out /tmp/ncu_$c.csv
done
head -1 /tmp/ncu_s1_tmf.csv > $O/${R}_ncu_dominant.csv
for c in s1_tmf s2_tmf s3_tmc s1_mlpf s2_mlpf s3_mlpf s3_mlp1 s3_mlp2; do tail -n +2 /tmp/ncu_$c.csv | sed "s/^/$c: /" >> $O/${R}_ncu_dominant.csv; done
python tools/microbench.py --case s1_tmf s2_tmf s3_tmc s1_mlpf s2_mlpf s3_mlpf s3_mlp1 s3_mlp2 s3_fc1v s4_mlp1 s4_mlp2 s1_core s2_core s3_core > $O/${R}_microbench.txt 2>&1
python tools/block_sweep.py --batch 8 --dtype bf16 --quick --out $O/${R}_block_sweep_bf16_b8.txt > /dev/null 2>&1
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_token_mixer.py tests/test_gpu_engine.py -m gpu -q -x \
    -k "upsample or point_reducer or head_level or class_map or fused_token_mixer or stage3 or fusion or radar_enh or shuffle or col2im or golden or mlp" > $O/${R}_sanitizer_memcheck.log 2>&1
tail -5 $O/${R}_sanitizer_memcheck.log
compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "upsample or point_reducer or class_map or col2im" > $O/${R}_sanitizer_racecheck.log 2>&1
tail -3 $O/${R}_sanitizer_racecheck.log
python tools/block_sweep.py --batch 8 --dtype bf16 --out $O/${R}_block_sweep_bf16_b8_full.txt > /dev/null 2>&1
for b in 1 16 64; do python tools/block_sweep.py --batch $b --dtype bf16 --quick --out $O/${R}_block_sweep_bf16_b$b.txt > /dev/null 2>&1; done
python tools/timeline.py --list --out $O/${R}_timeline.txt > /dev/null 2>&1
python tools/prof_train.py > $O/${R}_prof_train.txt 2>&1
