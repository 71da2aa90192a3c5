"""GPU triage: run a list of (trace, schedule, implementation) cases against the oracle, each in its own
process (a CUDA fault in one case does not take the others down), and print which results differ.
    python tools/triage.py [quick|wide|all]
Development aid; the parity suite proper is tests/test_gpu_parity.py."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = {
    "quick": [
        ("gossip", 4, 2000, 50, 0), ("gossip", 16, 20000, 1000, 0), ("gossip", 64, 40000, 8192, 0),
        ("gossip", 4, 2000, 50, 1), ("gossip", 16, 20000, 1000, 1), ("gossip", 64, 40000, 8192, 1),
        ("adversarial", 16, 12000, 500, 1), ("gossip", 7, 3000, 1, 1), ("gossip", 33, 6000, 640, 1),
        ("gossip", 96, 20000, 3000, 0), ("adversarial", 128, 40000, 8192, 0), ("gossip", 256, 60000, 16384, 0),
        ("tick", 128, 12000, 2048, 0), ("gossip", 1024, 20000, 8192, 0),
    ],
}


def one(gen, M, N, K, wide):
    os.environ["SW_FORCE_WIDE"] = str(wide)
    for p in (os.path.join(ROOT, "py-swirld_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import numpy as np
    import oracle as orc
    from swirld_b200 import engine, traces
    tr = getattr(traces, gen)(M, N, 3)
    o = orc.run_oracle(tr, K)
    r = engine.run_engine(tr, K)
    out = {}
    ocs = o["oracle"].can_see()
    for k, a, b in [("can_see", ocs, r["can_see"])] + [(k, o[k], r[k]) for k in ("round", "witness", "witness_table", "famous", "consensus", "transactions")]:
        a, b = np.asarray(a), np.asarray(b)
        if a.shape != b.shape:
            out[k] = "shape %s vs %s" % (a.shape, b.shape)
        elif not np.array_equal(a, b):
            d = np.argwhere(a != b)
            out[k] = "%d diffs, first %s exp %s got %s" % (len(d), d[0].tolist(), a[tuple(d[0])].item(), b[tuple(d[0])].item())
    if o["new_c_per_call"] != r["new_c_per_call"]:
        out["new_c"] = "differs"
    st = r["stats"]
    print(json.dumps({"case": [gen, M, N, K, wide], "bad": out, "ms": {k: round(st[k], 3) for k in st if k.startswith("ms_")},
                      "launches": st["kernel_launches"]}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        a = sys.argv[2:]
        one(a[0], int(a[1]), int(a[2]), int(a[3]), int(a[4]))
        sys.exit(0)
    which = sys.argv[1] if len(sys.argv) > 1 else "quick"
    for c in CASES[which]:
        try:
            p = subprocess.run([sys.executable, __file__, "--one"] + [str(x) for x in c], capture_output=True, text=True, timeout=240)
            tail = (p.stdout.strip().splitlines() or [""])[-1]
            if p.returncode != 0:
                tail += " | rc=%d %s" % (p.returncode, p.stderr.strip().splitlines()[-1:] if p.stderr else "")
            print(tail, flush=True)
        except subprocess.TimeoutExpired:
            print(json.dumps({"case": list(c), "bad": "TIMEOUT"}), flush=True)
