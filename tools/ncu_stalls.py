"""Top stall sites of one kernel from an .ncu-rep captured with --import-source on:
    python tools/ncu_stalls.py report.ncu-rep [n_top] [launch_skip]
Prints the headline metrics (duration, tensor-pipe activity, issue activity, DRAM), the stall-reason totals and the n_top SASS
instructions by sample count with their source line (needs `ncu` on PATH; run where the report is, no GPU needed)."""
import csv
import io
import subprocess
import sys


def page(rep, which, skip, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv", "--launch-skip", str(skip), "--launch-count", "1", *extra],
                         capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO("\n".join(l for l in out.splitlines() if not l.startswith("==")))))


def main(rep, n_top=40, skip=0):
    raw = page(rep, "raw", skip)
    h, vals = raw[0], raw[2]
    ci = {x: i for i, x in enumerate(h)}
    for k in ("Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"):
        if k in ci:
            print(f"{k:80s} {vals[ci[k]][:90]} {raw[1][ci[k]]}")
    src = page(rep, "source", skip, ("--print-source", "sass"))   # SASS rows; the CUDA-C view would need the sources on this box
    hdr = src[1]
    body = [r for r in src[2:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
    body = body[:len(body) // 2] if len(body) > 2 and body[0][0] == body[len(body) // 2][0] else body
    ci = {x: i for i, x in enumerate(hdr)}
    stalls = [x for x in hdr if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[ci["# Samples"]]) for r in body)
    agg = sorted(((s, sum(int(r[ci[s]]) for r in body)) for s in stalls), key=lambda kv: -kv[1])
    print(f"\n{tot} samples over {len(body)} SASS instructions; warp instructions executed {sum(int(r[ci['Instructions Executed']]) for r in body)}")
    print("stall totals:", ", ".join(f"{s[6:]} {100.0 * v / tot:.1f}%" for s, v in agg[:10]))
    top = sorted(range(len(body)), key=lambda i: -int(body[i][ci["# Samples"]]))[:n_top]
    for i in sorted(top):
        r = body[i]
        st = sorted(((s[6:], int(r[ci[s]])) for s in stalls), key=lambda kv: -kv[1])[:2]
        print(f"{i:5d} {r[ci['Source']].strip()[:64]:64s} {100.0 * int(r[ci['# Samples']]) / tot:5.1f}%  x{r[ci['Instructions Executed']]:>9s}  {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
