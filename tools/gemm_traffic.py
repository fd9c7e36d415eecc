"""Turn `ncu -i <rep> --page raw --csv` of the first forward GEMM launches of `bench.py` (launch order per encoder layer: qkv,
attn_out+drop+res, ffn1+gelu, ffn2+drop+res) into profiles/<name>.json: per-launch DRAM read+write, algorithmic bytes, tensor-pipe
activity and TFLOP/s, plus the average DRAM traffic per launch that bench.py reports as `roofline.traffic`.
usage: python tools/gemm_traffic.py raw.csv out.json TOKENS ["source description"]"""
import csv
import json
import sys

H, I = 1024, 4096
SHAPES = [("qkv", 3 * H, H, 1), ("attn_out+drop+res", H, H, 2), ("ffn1+gelu(2 outputs)", I, H, 2), ("ffn2+drop+res", H, I, 2)]
# last field: [T,N]-sized bf16 tensors moved besides A and W (out; + residual in / second output)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}


def main(raw, out, tokens, source):
    rows = list(csv.reader(l for l in open(raw, newline="") if l.strip() and not l.startswith("==")))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        return float(r[col[name]].replace(",", "")) * UNIT.get(units[col[name]], 1.0)

    launches = []
    for li, r in enumerate(body):
        name, N, K, nout = SHAPES[li % 4]
        us = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        alg = 2.0 * (tokens * K + N * K + nout * tokens * N)
        launches.append({"launch": li, "gemm": name, "M": tokens, "N": N, "K": K, "time_us": round(us, 3),
                         "dram_read_MB": round(rd / 1e6, 1), "dram_write_MB": round(wr / 1e6, 1), "algorithmic_MB": round(alg / 1e6, 1),
                         "tensor_active_pct": round(val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), 1),
                         "tflops": round(2.0 * tokens * N * K / us / 1e6, 1)})
    avg = sum((l["dram_read_MB"] + l["dram_write_MB"]) * 1e6 for l in launches) / len(launches)
    json.dump({"source": source, "avg_dram_traffic_bytes_per_launch": int(avg),
               "avg_algorithmic_bytes_per_launch": int(sum(l["algorithmic_MB"] for l in launches) * 1e6 / len(launches)),
               "launches": launches}, open(out, "w"), indent=1)
    for l in launches:
        print(l)
    print("avg traffic/launch MB", avg / 1e6)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4] if len(sys.argv) > 4 else "ncu --set full --clock-control none")
