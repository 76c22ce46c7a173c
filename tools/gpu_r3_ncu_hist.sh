#!/bin/bash
# ncu --set full of asn_fast_hist (round-2 kernel and the round-1 probe) on 50 frames, i.i.d. and segmentation-like maps
mkdir -p gpurun_out
HIST_PROBE_REPS=1 timeout 600 ncu --set full --clock-control none -k regex:fast_hist_kernel -c 8 -o /tmp/hist_full python tools/hist_probe.py > gpurun_out/ncu_hist.log 2>&1
echo "exit $?"; tail -3 gpurun_out/ncu_hist.log
ncu -i /tmp/hist_full.ncu-rep --page raw --csv > gpurun_out/hist_full_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/hist_full_raw.csv")))
hdr, data = rows[0], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size"]
idx = [hdr.index(w) if w in hdr else -1 for w in want]
for d in data:
    print([d[i][:48] if i >= 0 else None for i in idx])
PY
