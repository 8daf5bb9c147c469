#!/usr/bin/env python
"""A/B of kernel variants on ONE GPU box, with a correctness gate: every variant runs in its own process (the switches are read once
per process), is timed device-resident on the bench workloads with the L2 flushed between steps, and its output bytes are compared
with the baseline's (exact variants must reproduce them bit for bit) and with cv2.seamlessClone.

  python tools/ab_select.py [--out gpurun_out/ab] [--workloads cfg2 cfg1]        orchestrator: baseline + every variant below
  python tools/ab_select.py --worker OUT.json --workloads cfg2 ...                one variant (environment already set)

The orchestrator writes <out>/results.json, <out>/selected.json (the switches of every variant that is bit-identical to the baseline
on all workloads and at least 1.5 % faster at cfg2, combined and re-measured) and a table on stdout.  Nothing here is a bench
number: bench.py is."""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANT_LIBS = os.path.join(ROOT, "seamlesscloneoptimization_b200", "lib", "variants")

# (name, switch group, environment, variant library tag or None).  The baseline is the library's defaults.  A variant that keeps the
# arithmetic of the baseline (a different schedule of the same operations) must reproduce the baseline's bytes exactly; the ones in
# INEXACT change a summation order -- lowproj2 reorders float64 sums (the baseline's own order is not deterministic: atomics), the
# `s32` library cuts the columns of the tridiagonal solve into 32 segments instead of 16 -- and must stay within +-1 LSB of cv2 and
# within 0.01 points of the baseline's exact-byte share.  One winner per group is kept.
# (The first call of round 2 measured rhs_fold2, tri_smem, i8_p2, a leaner digitise and a digitise with the low-frequency block
# fused in: profiles/r2b_ab_variants_call1.log.  The first two became defaults, the last two gained nothing and were removed.)
VARIANTS = [
    ("tri_smem2", "tri", {"SCB_TRI_SMEM": "2"}, None),
    ("tri_s32", "tri", {}, "s32"),
    ("tri_smem2_s32", "tri", {"SCB_TRI_SMEM": "2"}, "s32"),
    ("i8_fwd_p2", "i8", {"SCB_I8_PERSISTENT": "3"}, None),
    ("i8_p2", "i8", {"SCB_I8_PERSISTENT": "2"}, None),
]


# second round, measured on top of the first round's winners (alone, the projections hide behind the column solve)
ON_TOP = [("lowproj2", {"SCB_LOWPROJ": "2"})]
INEXACT = {"lowproj2", "tri_s32", "tri_smem2_s32"}


def same_result(r, base, wls, exact):
    for wl in wls:
        if not r[wl]["host_equals_device"]:
            return False
        if r[wl]["md5"] == base[wl]["md5"]:
            continue
        if exact or r[wl].get("max_abs", 9) > 1 or abs(r[wl].get("pct_exact", 0.0) - base[wl].get("pct_exact", 100.0)) > 0.01:
            return False
    return True


def lib_of(tags):
    tags = sorted(t for t in tags if t)
    return os.path.join(VARIANT_LIBS, "libscb_" + "".join(tags) + ".so") if tags else None


def fake_worker(out_path, wls):
    """AB_FAKE=1: synthetic results, so that the orchestrator's bookkeeping can be exercised without a GPU (tests/test_tools.py)."""
    gain = {"SCB_I8_PERSISTENT": 0.003, "SCB_TRI_SMEM": 0.02, "SCB_LOWPROJ": 0.004}
    ms = 0.215 - sum(v for k, v in gain.items() if os.environ.get(k, "1") not in ("0", "1"))
    if os.environ.get("SCB_LIBRARY"):
        ms -= 0.001
    broken = os.environ.get("SCB_I8_PERSISTENT") == "2"  # pretend one variant computes something else
    res = {wl: {"ms": ms, "ms_mean": ms, "ms_min": ms, "stages_us": {}, "md5": "beef" if broken else "cafe", "host_equals_device": True, "roi": [0, 0], "engine": 4,
                "pct_exact": 99.85, "max_abs": 1} for wl in wls}
    with open(out_path, "w") as fh:
        json.dump(res, fh)


def worker(out_path, wls, steps):
    if os.environ.get("AB_FAKE"):
        return fake_worker(out_path, wls)
    import numpy as np
    import torch

    import seamlesscloneoptimization_b200 as scb
    from seamlesscloneoptimization_b200 import _capi as capi
    from seamlesscloneoptimization_b200 import workloads

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}
    for wl in wls:
        cache = f"/tmp/ab_wl_{wl}.npz"  # written once by the orchestrator: generating the 4K workload takes seconds
        if os.path.exists(cache):
            z = np.load(cache)
            src, dst, mask, p = z["src"], z["dst"], z["mask"], tuple(int(v) for v in z["p"])
        else:
            src, dst, mask, p = workloads.make_config(wl, seed=0)
        stream = torch.cuda.Stream(device=dev)
        ctx = scb.Context(0, stream=stream.cuda_stream)
        d_src, d_dst, d_mask = (torch.from_numpy(a).to(dev) for a in (src, dst, mask))
        d_blend = torch.zeros_like(d_dst)
        plan = scb.Plan(ctx, d_mask, src.shape[:2], dst.shape[:2], p, scb.MEM_DEVICE)
        g = plan.geometry
        with torch.cuda.stream(stream):
            for _ in range(5):
                plan.execute(d_src, d_dst, d_blend, scb.MEM_DEVICE)
            torch.cuda.synchronize()
            evs = []
            for _ in range(steps):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                plan.execute(d_src, d_dst, d_blend, scb.MEM_DEVICE)
                e1.record(stream)
                evs.append((e0, e1))
            torch.cuda.synchronize()
            ms = [a.elapsed_time(b) for a, b in evs]
            acc = {}
            for _ in range(5):
                flush.fill_(1)
                for k, v in plan.execute_timed(d_src, d_dst, d_blend, scb.MEM_DEVICE).items():
                    acc.setdefault(k, []).append(v)
            torch.cuda.synchronize()
        dev_blend = d_blend.cpu().numpy()
        # the HOST path (banded passes: other tile ranges of the same kernels) must give the same bytes
        host_blend = ctx.seamless_clone(src, dst, mask, p)
        interior = dev_blend[g.ry + 1 : g.ry + g.h - 1, g.rx + 1 : g.rx + g.w - 1]
        r = {
            "ms": statistics.median(ms), "ms_mean": statistics.mean(ms), "ms_min": min(ms),
            "stages_us": {k: round(1e3 * statistics.mean(v), 2) for k, v in acc.items()},
            "md5": hashlib.md5(np.ascontiguousarray(dev_blend)).hexdigest(),
            "host_equals_device": bool(np.array_equal(host_blend, dev_blend)),
            "roi": [int(g.w), int(g.h)], "engine": int(plan.engine),
        }
        ref_path = f"/tmp/ab_ref_{wl}.npy"
        if os.path.exists(ref_path):
            ref = np.load(ref_path)[g.ry + 1 : g.ry + g.h - 1, g.rx + 1 : g.rx + g.w - 1]
            d = np.abs(ref.astype(np.int16) - interior.astype(np.int16))
            r.update(pct_exact=100.0 * float((d == 0).mean()), max_abs=int(d.max()))
        res[wl] = r
        plan.close()
        ctx.close()
    with open(out_path, "w") as fh:
        json.dump(res, fh)


def run_variant(name, env_extra, wls, steps, out_dir, timeout_s):
    out = os.path.join(out_dir, f"{name}.json")
    if os.path.exists(out):
        os.remove(out)
    env = dict(os.environ)
    env.update(env_extra)
    lib = env_extra.get("SCB_LIBRARY")
    if lib and not os.path.exists(lib):
        return {"error": f"{lib} not built"}
    t0 = time.time()
    try:
        pr = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", out, "--steps", str(steps), "--workloads"] + wls, env=env, cwd=ROOT,
                            capture_output=True, text=True, timeout=timeout_s)
    except subprocess.TimeoutExpired:
        return {"error": f"timeout after {timeout_s} s"}
    if pr.returncode != 0 or not os.path.exists(out):
        return {"error": f"exit {pr.returncode}: {(pr.stderr or '')[-600:]}"}
    with open(out) as fh:
        r = json.load(fh)
    r["_wall_s"] = round(time.time() - t0, 1)
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", default=None)
    ap.add_argument("--workloads", nargs="+", default=["cfg2", "cfg1"])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "ab"))
    ap.add_argument("--timeout", type=int, default=100)
    ap.add_argument("--only", nargs="*", default=None)
    args = ap.parse_args()
    if args.worker:
        worker(args.worker, args.workloads, args.steps)
        return
    os.makedirs(args.out, exist_ok=True)
    if not os.environ.get("AB_FAKE"):
        import cv2
        import numpy as np

        from seamlesscloneoptimization_b200 import workloads

        for wl in args.workloads:  # cv2.seamlessClone once per workload
            src, dst, mask, p = workloads.make_config(wl, seed=0)
            np.savez(f"/tmp/ab_wl_{wl}.npz", src=src, dst=dst, mask=mask, p=np.array(p))
            np.save(f"/tmp/ab_ref_{wl}.npy", cv2.seamlessClone(src, dst, mask.copy(), p, cv2.NORMAL_CLONE))
    results = {"baseline": run_variant("baseline", {}, args.workloads, args.steps, args.out, args.timeout * 2)}
    base = results["baseline"]
    print("baseline", json.dumps(base), flush=True)
    if "error" in base:
        json.dump(results, open(os.path.join(args.out, "results.json"), "w"), indent=1)
        raise SystemExit("baseline failed")
    head = args.workloads[0]
    winners = {}
    for name, group, env_extra, tag in VARIANTS:
        if args.only is not None and name not in args.only:
            continue
        env_v = dict(env_extra)
        if tag:
            env_v["SCB_LIBRARY"] = lib_of([tag])
        r = run_variant(name, env_v, args.workloads, args.steps, args.out, args.timeout)
        results[name] = r
        if "error" not in r:
            r["_correct"] = same_result(r, base, args.workloads, name not in INEXACT)
            r["_speedup"] = base[head]["ms"] / r[head]["ms"]
            score = r["_speedup"] - (0.01 if tag else 0.0)  # a compile-time variant must earn its keep: +1 % over the plain switch
            if r["_correct"] and score >= 1.015 and (group not in winners or score > winners[group][3]):
                winners[group] = (name, env_extra, tag, score)
        print(name, json.dumps(r), flush=True)
        json.dump(results, open(os.path.join(args.out, "results.json"), "w"), indent=1)
    combined, tags = {}, []
    for name, env_extra, tag, sp in winners.values():
        combined.update(env_extra)
        tags.append(tag)
    if lib_of(tags):
        combined["SCB_LIBRARY"] = lib_of(tags)
    sel = {"winners": [w[0] for w in winners.values()], "env": combined}
    if combined:
        r = run_variant("combined", combined, args.workloads, args.steps, args.out, args.timeout)
        results["combined"] = r
        ok = "error" not in r and same_result(r, base, args.workloads, not any(w[0] in INEXACT for w in winners.values()))
        sel["combined_ok"] = ok
        if ok:
            sel["speedup"] = base[head]["ms"] / r[head]["ms"]
        else:  # fall back to the single best winner
            best = max(winners.values(), key=lambda w: w[3])
            sel["env"] = dict(best[1], **({"SCB_LIBRARY": lib_of([best[2]])} if best[2] else {}))
            sel["winners"] = [best[0]]
        print("combined", json.dumps(r), flush=True)
    for name, env_extra in ON_TOP:
        env_v = dict(sel["env"], **env_extra)
        r = run_variant(name + "_on_top", env_v, args.workloads, args.steps, args.out, args.timeout)
        results[name + "_on_top"] = r
        if "error" not in r:
            ref_ms = results["combined"][head]["ms"] if sel.get("combined_ok") else base[head]["ms"]
            r["_correct"] = same_result(r, base, args.workloads, False)
            r["_speedup"] = base[head]["ms"] / r[head]["ms"]
            if r["_correct"] and ref_ms / r[head]["ms"] >= 1.01:
                sel["env"] = env_v
                sel["winners"].append(name)
                sel["speedup"] = r["_speedup"]
        print(name + "_on_top", json.dumps(r), flush=True)
    json.dump(results, open(os.path.join(args.out, "results.json"), "w"), indent=1)
    json.dump(sel, open(os.path.join(args.out, "selected.json"), "w"), indent=1)
    with open(os.path.join(args.out, "selected.env"), "w") as fh:
        for k, v in sel["env"].items():
            fh.write(f"export {k}={v}\n")
        if "SCB_LIBRARY" in sel["env"]:
            fh.write(f"export SCB_TEST_LIBRARY={sel['env']['SCB_LIBRARY']}\n")
    print("\n%-16s %9s %9s %8s  %s" % ("variant", head + " ms", "speedup", "correct", "stages (us)"))
    for name, r in results.items():
        if "error" in r:
            print("%-16s ERROR %s" % (name, r["error"][:200]))
            continue
        print("%-16s %9.4f %9.3f %8s  %s" % (name, r[head]["ms"], base[head]["ms"] / r[head]["ms"], r.get("_correct", "-"), r[head]["stages_us"]))
    print("selected:", json.dumps(sel))


if __name__ == "__main__":
    main()
