"""K3 at the full config size (20 000 bags x 3025), timed like bench.py's select_20k leg (whole
cs_select_topk call, pre-allocated outputs, no host sync, L2 flushed between repetitions, CUDA
events) under each kernel variant.  The switches are read when the library loads, so every variant
runs in its own process.  Also checks that every variant produces the same selection.

    python profiles/time_select_ab.py            # all variants, one JSON line each
    python profiles/time_select_ab.py --one      # the variant of the current environment
"""
import json
import os
import subprocess
import sys

VARIANTS = {
    "cta+lookback offsets (default)": {},
    "cta+lookback, threshold always from 64 columns": {"CELLSEG_SELECT_COL32": "0"},
    "cta+recount offsets": {"CELLSEG_SELECT_OFFSETS": "recount"},
    "cta+lookback, plain launch of the clean-up pass": {"CELLSEG_SELECT_SORT_PDL": "0"},
    "cta, 12 CTAs/SM (40 registers)": {"CELLSEG_SELECT_OCC": "12"},
    "cta, 16 CTAs/SM (32 registers)": {"CELLSEG_SELECT_OCC": "16"},
    "two warps per bag, 12 CTAs/SM (78 registers)": {"CELLSEG_SELECT_CTA": "64"},
    "two warps per bag, 14 CTAs/SM (72 registers)": {"CELLSEG_SELECT_CTA": "64", "CELLSEG_SELECT_OCC64": "14"},
    "two warps per bag, 16 CTAs/SM (64 registers)": {"CELLSEG_SELECT_CTA": "64", "CELLSEG_SELECT_OCC64": "16"},
    "warp per bag": {"CELLSEG_SELECT_WARP": "1"},
    "cta+ticket (round-2 r2i build)": {"CELLSEG_SELECT_OFFSETS": "ticket"},
}


def one():
    import numpy as np
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from cellsegmentation_b200 import ops, synthetic
    dev = torch.device("cuda", 0)
    out = {}
    for nb, T in ((20000, 3025), (18000, 225), (1024, 3025)):
        g = torch.Generator(device=dev)
        g.manual_seed(7)
        p = torch.rand(nb * T, device=dev, generator=g)
        lab = torch.from_numpy(synthetic.make_labels(nb, seed=3)).to(dev)
        buf = ops.select_buffers(nb, nb * 330, dev)
        fn = lambda: ops.select_topk(p, lab, nb, T, 1, 30, sync=False, out=buf)  # noqa: E731
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        reps = 7
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for a, b in ev:
            flush.zero_()
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        kept = int(buf[2][-1].item())
        nbytes = 4.0 * p.numel() + 5.0 * kept + 4.0 * nb
        pos = torch.arange(1, kept + 1, device=dev, dtype=torch.int64)
        digest = int(((buf[0][:kept].to(torch.int64) * 2 + buf[1][:kept].to(torch.int64)) * pos % 1000003).sum().item())
        out["%dx%d" % (nb, T)] = {"ms_min": ms[0], "ms_median": ms[reps // 2], "kept": kept,
                                   "GBps_at_min": nbytes / (ms[0] * 1e-3) / 1e9, "digest": digest}
    print(json.dumps(out))


if __name__ == "__main__":
    if "--one" in sys.argv:
        one()
    else:
        digests = {}
        only = os.environ.get("AB_ONLY")                  # comma-separated substrings of variant names
        for name, env in VARIANTS.items():
            if only and not any(o in name for o in only.split(",")):
                continue
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=dict(os.environ, **env),
                               capture_output=True, text=True, timeout=600)
            line = (r.stdout.strip().splitlines() or ["{}"])[-1]
            print(name, line if r.returncode == 0 else "FAILED rc=%d %s" % (r.returncode, r.stderr[-800:]), flush=True)
            if r.returncode == 0:
                digests[name] = {k: (v["kept"], v["digest"]) for k, v in json.loads(line).items()}
        vals = list(digests.values())
        print("outputs identical across the %d variants run:" % len(vals), all(v == vals[0] for v in vals))
