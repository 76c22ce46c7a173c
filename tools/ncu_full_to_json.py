#!/usr/bin/env python
"""ncu --set full raw CSV (+ the library's scope names in launch order) -> profiles/*_hot_kernels_ncu_full.json and
*_ncu_traffic.json.  usage: ncu_full_to_json.py raw.csv sequence.json out_full.json out_traffic.json
The capture covers the kernels of ONE pass of tools/hot_path_once.py --tier-b (source-image shapes); launches are matched
to scope names in order: a scope is one kernel launch except where noted in MULTI."""
import csv
import json
import re
import sys

MULTI = {"upsample_bwd": 2}   # scopes that launch more than one kernel (none on the Tier-B pass)
OURS = re.compile(r"asn::|umma::|lazy::|halo::")


def main():
    raw, seqf, out_full, out_traffic = sys.argv[1:5]
    rows = list(csv.reader(open(raw)))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    seq = json.load(open(seqf))

    def g(r, k):
        try:
            return float(r[idx[k]])
        except (KeyError, ValueError):
            return None

    launches = rows[2:]   # (the capture's -k regex already restricts it to the library's kernels)
    names = []
    for n in seq:
        names += [n] * MULTI.get(n, 1)
    aligned = len(names) == len(launches)
    kernels, traffic = [], {}
    for i, r in enumerate(launches):
        name = names[i] if aligned else None
        rec = {"#": i, "name": name, "kernel": r[idx["Kernel Name"]][:120], "grid": r[idx["launch__grid_size"]],
               "block": r[idx["launch__block_size"]], "us": g(r, "gpu__time_duration.sum"),
               "dram_rd_MB": g(r, "dram__bytes_read.sum"), "dram_wr_MB": g(r, "dram__bytes_write.sum"),
               "l2_pct": g(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
               "l2_read_MB_from_sm": (g(r, "lts__t_sectors_srcunit_tex_op_read.sum") or 0) * 32 / 1e6,
               "tensor_pct": g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
               "sm_pct": g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
               "regs": r[idx["launch__registers_per_thread"]],
               "warps_pct": g(r, "sm__warps_active.avg.pct_of_peak_sustained_active")}
        kernels.append(rec)
        if name and name not in traffic and rec["dram_rd_MB"] is not None:
            traffic[name] = {"dram_bytes_per_launch": int((rec["dram_rd_MB"] + rec["dram_wr_MB"]) * 1e6), "us_cold": rec["us"],
                             "shape": "source image (720x1280; features 90x160), first launch of the pass",
                             "ncu_kernel": rec["kernel"][:60]}
    json.dump({"what": "ncu --set full --clock-control none of every libasn_b200 kernel of one Tier-B pass of "
                       "tools/hot_path_once.py (source-image shapes; cold caches, serialised: compare shares, not absolutes)",
               "names_aligned_with_library_scopes": aligned, "kernels": kernels}, open(out_full, "w"), indent=1)
    json.dump({"source": out_full, "kernels": traffic}, open(out_traffic, "w"), indent=1)
    print("launches", len(launches), "scopes", len(names), "aligned", aligned)


if __name__ == "__main__":
    main()
