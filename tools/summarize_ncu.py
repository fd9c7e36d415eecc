"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals / shares (markdown)."""
import csv
import re
import sys
from collections import defaultdict


def main(path, out):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9, "second": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
    tot = sum(ns for _, ns in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for name, ns in rows:
        short = re.sub(r"<.*", "", name.split("(")[0]).strip()
        if "gemm_bf16_tcgen05" in name:
            m = re.search(r"gemm_bf16_tcgen05_kernel<(\d+), *(true|false|\(bool\)[01]), *(true|false|\(bool\)[01])>", name)
            short = "dle::gemm_bf16_tcgen05_kernel" + (f"<BN={m.group(1)},A_MN={m.group(2)},B_MN={m.group(3)}>" if m else "")
        agg[short][0] += 1
        agg[short][1] += ns
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path})\n\n{len(rows)} launches, {tot / 1e6:.3f} ms total device time "
                "(cold-cache, serialised under ncu: compare SHARES, not absolutes)\n\n| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:110]}` | {n} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% | {ns / n / 1e3:.1f} |\n")
    print(open(out).read()[:3000])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
