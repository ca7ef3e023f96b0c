#!/usr/bin/env python
"""Times the switchable kernel variants on one B200 and checks every output against the oracle's golden.

Knobs (environment, read by libb2gpu.so per call / per handle): B2GPU_SCATTER (radix pass: 1 = round 1's k_scatter,
2 / 22 / 3 / 32 / 24 / 34 = k_scatter2 variants, zip-ada_b200/csrc/b2_scatter2.cuh), B2GPU_PM (package-merge lists: 0 = binary
searches, 1 = merge path), B2GPU_RR_GROUP, B2GPU_PIPELINE + B2GPU_BATCH_POSITIONS.  Workload: the 1 GiB text stream of
bench.py (golden markov:1073741824:5eed0001:9), input resident in HBM.

Greedy: round 1's path as the control, the package-merge and round-0 knobs one at a time, the scatter variants on top,
then the dispatch group and two batches in flight.  Every configuration is logged to
gpurun_out/variants.jsonl before ("started") and after it ran, so that a run that dies in one variant can be re-invoked
and carries on behind it; the best environment is written to gpurun_out/best_env.sh.
"""
import hashlib
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")
LOG = os.path.join(OUT, "variants.jsonl")
KNOBS = ("B2GPU_SCATTER", "B2GPU_PM", "B2GPU_R0", "B2GPU_RR_GROUP", "B2GPU_SC_PFD", "B2GPU_PIPELINE", "B2GPU_BATCH_POSITIONS")


def log(rec):
    with open(LOG, "a") as f:
        f.write(json.dumps(rec) + "\n")
        f.flush()
        os.fsync(f.fileno())


def history():
    done, started = {}, set()
    if os.path.exists(LOG):
        for line in open(LOG):
            try:
                r = json.loads(line)
            except Exception:
                continue
            key = json.dumps(r["env"], sort_keys=True)
            if r.get("status") == "started":
                started.add(key)
            else:
                done[key] = r
    return done, started


def main():
    import torch
    import corpus
    os.makedirs(OUT, exist_ok=True)
    b2 = importlib.import_module("zip-ada_b200")
    size_mb = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    n = size_mb << 20
    dev = torch.device("cuda", 0)
    gkey = "markov:%d:5eed0001:9" % n
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "stream_sha.json"))).get(gkey)
    d_in = torch.zeros(n + 256, dtype=torch.uint8, device=dev)
    d_in[:n] = corpus.workload("markov", n, 0x5EED0001, torch, dev)
    cap = int(b2.lib().b2_bound(n)) + 1024 * (n // 40000 + 16)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()

    def run(env):
        key = json.dumps(env, sort_keys=True)
        done, started = history()
        if key in done:
            return done[key]
        if key in started:
            rec = {"env": env, "status": "crashed earlier", "ok": False}
            log(rec)
            return rec
        log({"env": env, "status": "started"})
        for k in KNOBS:
            os.environ.pop(k, None)
        os.environ.update({k: str(v) for k, v in env.items()})
        rec = {"env": env, "status": "ran", "ok": False}
        try:
            with b2.Encoder(b2.block_900k, 0) as enc:
                enc.set_timing(1)
                ln = enc.encode_device_ptr(d_in.data_ptr(), n, n, d_out.data_ptr(), cap)           # warm-up: workspaces grow
                sha = hashlib.sha256(d_out[:ln].cpu().numpy().tobytes()).hexdigest()
                rec["bytes"] = int(ln)
                rec["ok"] = bool(g is not None and sha == g["sha256"] and ln == g["bytes"])
                enc.reset_stats()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(steps):
                    ln2 = enc.encode_device_ptr(d_in.data_ptr(), n, n, d_out.data_ptr(), cap)
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / steps
                st = enc.stats()
                sha2 = hashlib.sha256(d_out[:ln2].cpu().numpy().tobytes()).hexdigest()
                rec["ok"] = bool(rec["ok"] and sha2 == sha)
                rec.update({"ms_per_step": round(dt * 1e3, 2), "MBps": round(n / 1e6 / dt, 1),
                            "scatter_ms_per_step": round(st.scatter_ms / steps, 2), "sort_ms_per_step": round(st.sort_ms / steps, 2),
                            "scatter_GBps": round(24.0 * st.scatter_elems / max(1e-9, st.scatter_ms / 1e3) / 1e9, 1),
                            "scatter_launches_per_step": int(st.scatter_launches // steps)})
        except Exception as ex:
            rec["error"] = str(ex)[:300]
            rec["ok"] = False
        log(rec)
        print(json.dumps(rec), flush=True)
        if "error" in rec:
            os._exit(3)            # the CUDA context may be gone: the caller starts again and carries on behind this variant
        return rec

    def better(a, b):
        return a.get("ok") and "ms_per_step" in a and (not (b and b.get("ok")) or a["ms_per_step"] < b["ms_per_step"])

    mode = sys.argv[3] if len(sys.argv) > 3 else "full"
    best = None
    if mode == "third":
        # third run: the defaults as built (k_scatter3 variant 45, dispatch groups of 64, prefetch one SM count ahead, the
        # batch sized by the free memory) against their neighbours
        for env in ({}, {"B2GPU_SCATTER": 53}, {"B2GPU_SCATTER": 61}, {"B2GPU_SCATTER": 69}, {"B2GPU_SC_PFD": 296}, {"B2GPU_SC_PFD": 74},
                    {"B2GPU_BATCH_POSITIONS": 1610612736}):
            run(env)
        return 0
    if mode == "full":
        # round 1's path as the control, then the package-merge by merge path and the seven-pass round 0 one at a time
        control = {"B2GPU_SCATTER": 1, "B2GPU_PM": 0, "B2GPU_R0": 8}
        best = run(control)
        if not best.get("ok"):
            print("the control configuration did not produce the golden stream")
            best = None
        base = dict(control)
        for knob, v in (("B2GPU_PM", 1), ("B2GPU_R0", 7)):
            r = run(dict(control, **{knob: v}))
            if r.get("ok") and (best is None or r["ms_per_step"] < best["ms_per_step"] * 1.003):
                base[knob] = v
        scatters = (1, 2, 22, 24, 3, 32, 34, 40, 41, 43, 45, 47)
    else:
        # second run: the package-merge and round-0 knobs are settled (profiles/r02b_variants.jsonl); k_scatter2 against k_scatter3
        base = {"B2GPU_PM": 1, "B2GPU_R0": 7}
        scatters = (2, 40, 41, 42, 43, 45, 47, (41, 148), (43, 296))
    for sc in scatters:
        r = run(dict(base, B2GPU_SCATTER=sc) if isinstance(sc, int) else dict(base, B2GPU_SCATTER=sc[0], B2GPU_SC_PFD=sc[1]))
        if better(r, best):
            best = r
    if best is None:
        print("no variant produced the golden stream")
        return 1
    env = dict(best["env"])
    sc = env.get("B2GPU_SCATTER")
    extras = [{"B2GPU_RR_GROUP": 64}, {"B2GPU_RR_GROUP": 256}]
    if sc in (41, 43, 45, 47) and "B2GPU_SC_PFD" not in env:
        d0 = 592 if sc in (43, 47) else 296
        extras += [{"B2GPU_SC_PFD": d0 * 2}]
    for extra in extras:
        r = run(dict(env, **extra))
        if better(r, best):
            best = r
    # the whole 1 GiB stream as one batch (2.9 G positions, about 135 GB of workspace) instead of two
    # (for the record only: not a candidate for the default, whose workspace must leave room for the caller's own buffers)
    run(dict(best["env"], B2GPU_BATCH_POSITIONS=3221225472))
    if mode == "full":
        env = dict(best["env"])
        # two batches in flight, half the batch each (the same device memory); against the same batch size alone
        half = 805306368
        r1 = run(dict(env, B2GPU_BATCH_POSITIONS=half))
        r2 = run(dict(env, B2GPU_BATCH_POSITIONS=half, B2GPU_PIPELINE=2))
        for r in (r1, r2):
            if better(r, best):
                best = r
    with open(os.path.join(OUT, "best_env.sh"), "w") as f:
        for k, v in best["env"].items():
            f.write("export %s=%s\n" % (k, v))
    print("BEST", json.dumps(best))
    return 0


if __name__ == "__main__":
    sys.exit(main())
