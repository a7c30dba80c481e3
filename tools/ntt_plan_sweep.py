#!/usr/bin/env python3
"""Times k_ntt_pass schedules (digit plans) per size on the GPU and checks every plan against the default one.

    python tools/ntt_plan_sweep.py [--out profiles/r2_ntt_plan_sweep.json]

SWB_NTT_PLAN_<log_n>=a,b,c (read by ntt.cu at every call) selects the digits; the result of each plan must equal
the default plan's bit for bit."""
import argparse
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def plans(log_n, max_digit=11, min_digit=5, max_passes=4):
    out = []
    for m in range(1, max_passes + 1):
        for combo in itertools.product(range(min_digit, max_digit + 1), repeat=m):
            if sum(combo) == log_n and list(combo) == sorted(combo, reverse=True):
                out.append(combo)
    return out


def main():
    import torch
    from simpleworks_b200.binding import Backend
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--sizes", default="16,18,20,21,22,23,24,25,26")
    ap.add_argument("--default-only", action="store_true")
    ap.add_argument("--plan", default="", help="only this plan (digits separated by commas) beside the default, for the single size given")
    args = ap.parse_args()
    be = Backend(0)
    be.profile(True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    res = {}
    for log_n in [int(x) for x in args.sizes.split(",")]:
        n = 1 << log_n
        x = torch.randint(-2 ** 63, 2 ** 63 - 1, (n, 4), dtype=torch.int64, device="cuda:0")
        x[:, 3] &= 0x0FFFFFFFFFFFFFFF
        os.environ.pop(f"SWB_NTT_PLAN_{log_n}", None)
        want = be.ntt_(x.clone(), log_n)
        rows = []
        cands = [None] + [p for p in plans(log_n)]
        # orderings: also try the ascending order of each multiset
        cands += [tuple(reversed(p)) for p in plans(log_n) if tuple(reversed(p)) != p]
        if args.default_only:
            cands = [None]
        if args.plan:
            cands = [None, tuple(int(x) for x in args.plan.split(","))]
        for plan in cands:
            if plan is None:
                os.environ.pop(f"SWB_NTT_PLAN_{log_n}", None)
            else:
                os.environ[f"SWB_NTT_PLAN_{log_n}"] = ",".join(map(str, plan))
            try:
                y = be.ntt_(x.clone(), log_n)
            except Exception as e:
                rows.append({"plan": plan, "error": str(e)[:80]})
                continue
            ok = bool(torch.equal(y, want))
            ts = []
            for _ in range(4):
                y = x.clone()
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                be.ntt_(y, log_n)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            rows.append({"plan": plan, "ms": min(ts), "equal_default": ok, "passes": be.last_stages()})
        os.environ.pop(f"SWB_NTT_PLAN_{log_n}", None)
        rows.sort(key=lambda r: r.get("ms", 1e9))
        res[str(log_n)] = rows
        print(log_n, [(r["plan"], round(r.get("ms", -1), 3), r.get("equal_default")) for r in rows[:6]],
              "default:", [round(r["ms"], 3) for r in rows if r["plan"] is None], flush=True)
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
