"""Hand-off timeline of the attention-backward kernel (measurement tooling; needs the trace build:
    python -m deeplearningexamples_b200.csrc.build --variant-trace
    B=128 python tools/attn_trace.py            # on the GPU box; sets DLE_LIB_PATH itself
Lane 0 of the MMA-issuer warp and of one compute warp per group of ONE mid-kernel CTA stamps clock64() after every barrier wait and
after every MMA issue / barrier arrive (attention_sm100.cu, DLE_ATTN_TRACE); every CTA records its SM and start/end time.  Prints
  * per-CTA duration statistics and the idle gap between consecutive CTAs on the same SM,
  * for the traced CTA: prologue / per-pair / kv-tile-end / dQ-drain durations, and for each wait site the cycles spent blocked,
  * the raw timeline of two steady-state pairs,
and writes everything to gpurun_out/attn_trace.json."""
import ctypes
import json
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("DLE_LIB_PATH", os.path.join(ROOT, "deeplearningexamples_b200", "libdle_b200_trace.so"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from deeplearningexamples_b200 import kernels as k, _lib as L  # noqa: E402

MMA = {10: "W kv_full", 11: "W dkv_read", 12: "W q_full", 13: "W s_free[1](t-1)", 14: "I S(t,0)", 15: "W dp_free[1](t-1)", 16: "I dP(t,0)",
       17: "W s_free[0](t)", 18: "I S(t,1)", 19: "W p_full(t-1)", 20: "I dV(t-1)", 21: "W dp_free[0](t)", 22: "I dP(t,1)",
       23: "W ds_full(t-1)", 24: "I dK,dQ(t-1)", 25: "W p_full(tail)", 26: "I dV(tail)", 27: "W ds_full(tail)", 28: "I dK,dQ(tail)"}
CMP = {30: "mask staged, nk known", 31: "pair begin", 32: "W s_full", 33: "S in regs, s_free", 34: "P computed", 35: "W dv_done(t-1)",
       36: "P stored, p_full", 37: "W dp_full", 38: "dP in regs, dp_free", 39: "dS computed", 40: "W pair_done(t-1)", 41: "dS stored, ds_full",
       42: "W dkv_full", 43: "dK/dV drained", 44: "dQ drained"}
NAMES = {1: "kernel entry", **MMA, **CMP}
CAP, MAXC = 4096, 8192


def main():
    B, S, A = int(os.environ.get("B", 128)), int(os.environ.get("S", 512)), int(os.environ.get("A", 16))
    p = float(os.environ.get("P", 0.1))
    H = A * 64
    qkv = torch.randn(B * S, 3 * H, device="cuda").to(torch.bfloat16)
    dctx = torch.randn(B * S, H, device="cuda").to(torch.bfloat16)
    mask = torch.zeros(B, S, device="cuda")
    ctx, lse = k.attn_fwd(qkv, mask, B, S, A, dropout_p=p, seed=1, dropout_stream=1)
    for _ in range(4):
        k.attn_bwd(qkv, mask, ctx, dctx, lse, B, S, A, dropout_p=p, seed=1, dropout_stream=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    k.attn_bwd(qkv, mask, ctx, dctx, lse, B, S, A, dropout_p=p, seed=1, dropout_stream=1)
    e1.record(); torch.cuda.synchronize()
    print(f"traced launch (B={B}, p={p}): {e0.elapsed_time(e1) * 1e3:.1f} us  (includes the delta kernel)")
    ev = (ctypes.c_ulonglong * (4 * CAP))(); cnt = (ctypes.c_int * 4)(); ctas = (ctypes.c_ulonglong * (3 * MAXC))()
    lib = ctypes.CDLL(os.environ["DLE_LIB_PATH"])
    rc = lib.dle_debug_attn_trace(ev, cnt, ctas)
    assert rc == 0, rc
    out = {"B": B, "p": p}
    # ---- per-CTA
    n_cta = min(A * B, MAXC)
    per_sm = defaultdict(list)
    for i in range(n_cta):
        sm, t0, t1 = ctas[3 * i], ctas[3 * i + 1], ctas[3 * i + 2]
        per_sm[sm].append((t0, t1, i))
    durs, gaps = [], []
    for sm, lst in per_sm.items():
        lst.sort()
        for a, b_ in zip(lst, lst[1:]):
            gaps.append(b_[0] - a[1])
        durs += [t1 - t0 for t0, t1, _ in lst]
    durs.sort(); gaps.sort()
    q = lambda v, f: v[min(len(v) - 1, int(f * len(v)))]
    t_first = min(l[0][0] for l in per_sm.values()); t_last = max(l[-1][1] for l in per_sm.values())
    out["cta"] = dict(n=n_cta, sms=len(per_sm), dur_ns_median=q(durs, .5), dur_ns_p10=q(durs, .1), dur_ns_p90=q(durs, .9),
                      gap_ns_median=q(gaps, .5), gap_ns_p90=q(gaps, .9), kernel_ns=t_last - t_first,
                      ctas_per_sm_min=min(len(l) for l in per_sm.values()), ctas_per_sm_max=max(len(l) for l in per_sm.values()))
    print("CTAs:", out["cta"])
    # ---- traced CTA
    tl = {}
    for s_, role in enumerate(("mma", "cmp_g0", "cmp_g1", "mma2")):
        rows = []
        for i in range(cnt[s_]):
            w = ev[s_ * CAP + i]
            rows.append((w >> 16, (w >> 8) & 255, w & 255))
        tl[role] = rows
    tl = {k_: v for k_, v in tl.items() if v}
    base = min(r[0][0] for r in tl.values() if r)
    out["timeline"] = {role: [(c - base, code, t) for c, code, t in rows] for role, rows in tl.items()}
    for role, rows in tl.items():
        blocked = defaultdict(int); count = defaultdict(int)
        prev = None
        for c, code, t in rows:
            if prev is not None and NAMES.get(code, "").startswith("W "):
                blocked[code] += c - prev; count[code] += 1
            prev = c
        total = rows[-1][0] - rows[0][0] if rows else 0
        print(f"\n== {role}: {len(rows)} events, {total} clk from first to last event")
        for code in sorted(blocked, key=lambda c_: -blocked[c_]):
            print(f"   blocked at {NAMES[code]:24s}: {blocked[code]:8d} clk total ({100.0 * blocked[code] / max(total, 1):5.1f} %), {blocked[code] / max(count[code], 1):7.0f} clk avg x {count[code]}")
        # segment durations between consecutive events (non-wait = work)
        work = defaultdict(int); wcount = defaultdict(int)
        prev = None
        for c, code, t in rows:
            if prev is not None and not NAMES.get(code, "").startswith("W "):
                work[code] += c - prev[0]; wcount[code] += 1
            prev = (c, code)
        for code in sorted(work, key=lambda c_: -work[c_]):
            print(f"   work  until {NAMES.get(code, code):24s}: {work[code]:8d} clk total ({100.0 * work[code] / max(total, 1):5.1f} %), {work[code] / max(wcount[code], 1):7.0f} clk avg x {wcount[code]}")
    # pair period from the MMA warp: time between successive "I S(t,0)"
    s0 = [(t, c) for c, code, t in tl["mma"] if code == 14]
    per = [b_[1] - a[1] for a, b_ in zip(s0, s0[1:])]
    if per:
        print("\npair period (clk) by t:", per)
        out["pair_period_clk"] = per
    print("\n-- merged timeline, pairs 5..6 (clk relative to CTA start)")
    merged = sorted((c - base, role, NAMES.get(code, str(code)), t) for role, rows in tl.items() for c, code, t in rows)
    for c, role, name, t in merged:
        if 5 <= t <= 6 or name in ("kernel entry", "mask staged, nk known"):
            print(f"  {c:8d}  {role:7s} t={t:2d}  {name}")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/attn_trace.json", "w"))


if __name__ == "__main__":
    main()
