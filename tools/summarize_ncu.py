#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel:
launches, total and share of device time.  Usage: summarize_ncu.py launches.csv [out.md]"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(.*", "", name)
    m = re.search(r"umma_kernel<\(int\)(\d), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+)>", name)
    if m:
        return (f"asn::umma::umma_kernel<mode={m.group(1)},BN={m.group(2)},stages={m.group(3)},CL={m.group(4)},"
                f"MT={m.group(5)},EW={m.group(6)}>")
    name = re.sub(r"^void ", "", name)
    return name[:110]


OURS = re.compile(r"asn::|umma::|lazy::|halo::|upsample2_|per_class_iu|^(aspp_|fcd_|lazy_|ce_kernel|ce_finalize|ce_generic|softmax_kernel|softmax_generic|"
                  r"sgd_step|adam_step|upsample_|fast_hist|gan_loss|nchw_|nhwc_)")


def is_ours(short_name):
    return bool(OURS.search(short_name))


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
    agg = defaultdict(lambda: [0, 0.0])
    for name, ns in rows:
        a = agg[short(name)]
        a[0] += 1
        a[1] += ns
    total = sum(v[1] for v in agg.values())
    mine = sum(v[1] for k, v in agg.items() if is_ours(k))
    out = [f"# ncu launch list summary: {path}", "",
           f"{len(rows)} launches, {total / 1e6:.2f} ms of serialised device time; "
           f"libasn_b200 kernels {mine / 1e6:.2f} ms = {100 * mine / total:.1f} % of it", "",
           "| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
        out.append(f"| `{k}` | {n} | {ns / 1e6:.3f} | {100 * ns / total:.2f} % | {ns / n / 1e3:.1f} |")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
