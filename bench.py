#!/usr/bin/env python
"""Headline benchmark: instances/sec of (score + adaptive top-k) — BASELINE.json configs[1].

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload: ResNet-34 tile-instance inference (tile 32, interval 5 -> 3025 instances per 299x299
bag, bf16 tcgen05 path) followed by the count-driven adaptive top-k selection, on synthetic
LYSTO-shaped bags with random-init weights.  A step scores and selects one slice of
`--bags-per-step` bags of the HBM-resident bag array (the slice advances every step, so both
the u8 inputs and the activation workspace exceed L2).  One JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILE, INTERVAL, H = 32, 5, 299
T_PER_BAG = 3025
FLOP_INBOUNDS = 82.18e6      # in-bounds FLOP per 32x32 instance, ResNet-34
FLOP_NOMINAL = 149.52e6      # nominal (zero-padding taps included)
SELECT_BYTES_PER_INST = 4.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bags-per-step", type=int, default=1024, help="bags scored per step per GPU")
    ap.add_argument("--resident-bags", type=int, default=0, help="bags resident in HBM per GPU (0: auto)")
    ap.add_argument("--max-batch", type=int, default=37888, help="instances per forward batch")
    ap.add_argument("--ref-bags", type=int, default=2, help="bags per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def measured_traffic(n_inst):
    """DRAM bytes of the conv stage for n_inst instances, from the committed `ncu --set full`
    capture of one forward batch (profiles/summarize_full.py); None when no capture is there."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_n_traffic.json")))
        return {"bytes": t["dram_bytes_per_instance"] * n_inst, "per_instance": t["dram_bytes_per_instance"],
                "source": "profiles/r01_n_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum over the "
                          "%d launches of one %d-instance batch)" % (t["launches"], t["instances"])}
    except Exception:
        return None


def cpu_reference_step(n_bags, sd, labels, seed):
    """The reference's CPU path restated (oracle port): unfold -> ResNet-34 fp32 forward ->
    lexsort + adaptive top-k predicate.  Returns (instances, seconds)."""
    import numpy as np
    import torch
    from oracle import model as omodel, select as oselect, synth, tiles as otiles
    bags = synth.make_bags(n_bags, seed=seed)
    t0 = time.perf_counter()
    x = torch.from_numpy(otiles.unfold(list(bags), INTERVAL, TILE))
    probs = omodel.forward_probs(sd, x, "resnet34", batch=3025)
    tid = np.repeat(np.arange(n_bags), T_PER_BAG)
    oselect.sample_indices_loop(tid, labels[:n_bags], probs, 1, 30)
    return x.shape[0], time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch
    from oracle import model as omodel
    torch.set_num_threads(os.cpu_count())
    sd = omodel.make_state_dict("resnet34", seed=0, random_bn=False)
    labels = np.array([3, 0, 7, 1, 12, 0, 40, 5] * 8, np.int32)
    for i in range(args.warmup):
        cpu_reference_step(1 if i else 1, sd, labels, 100 + i)
    inst, secs = 0, 0.0
    for i in range(args.steps):
        n, s = cpu_reference_step(args.ref_bags, sd, labels, i)
        inst += n; secs += s
    val = inst / secs
    sample = "%d bags x 3025 instances per step (unfold + fp32 forward + lexsort/top-k loop)" % args.ref_bags
    print(json.dumps({
        "impl": "reference", "metric": "instances/sec (score+adaptive top-k)", "value": val,
        "unit": "instances/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1] ResNet-34 tile inference + adaptive top-k, tile 32 interval 5; "
                               "CPU reference path (oracle port of inference_tiles + sample) on a bounded sample"},
        "cpu_baseline": {"value": val, "unit": "instances/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": "instances/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from cellsegmentation_b200 import ops, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.bags_per_step
    resident = args.resident_bags or max(B * (args.steps + args.warmup + 1), B)
    resident = min(resident, 20000 // max(world, 1)) if args.resident_bags == 0 else resident
    resident = max(resident, B)
    bags = synthetic.make_bags_device(resident, dev, seed=rank)
    labels_h = synthetic.make_labels(resident, seed=rank)
    labels = torch.from_numpy(labels_h).to(dev)
    convs, fc_w, fc_b = synthetic.make_resnet_weights("resnet34", seed=0)
    clf = ops.TileClassifier("resnet34", convs, fc_w, fc_b, device=dev)
    n_inst = B * T_PER_BAG
    prob = torch.empty(n_inst, dtype=torch.float32, device=dev)
    cap = int(B * 330)
    launches = [0]

    def step_device(i, ev=None):
        b0 = (i * B) % (resident - B + 1)
        view = bags[b0:b0 + B]
        if ev:
            ev[0].record()
        clf.forward_tiles(view, TILE, INTERVAL, precision="bf16", max_batch=args.max_batch, prob_out=prob)
        if ev:
            ev[1].record()
        idx, pl, off = ops.select_topk(prob, labels[b0:b0 + B], B, T_PER_BAG, 1, 30, capacity=cap)
        if ev:
            ev[2].record()
        launches[0] += clf.last_launch_count + 4   # + count, scan, fast select, exact fallback
        return idx

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step_device(i)
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches[0] = 0
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin.record()
    for i in range(args.steps):
        step_device(args.warmup + i, evs[i])
    t_end.record()
    sync_all()
    ms_total = t_begin.elapsed_time(t_end)
    clocks = sampler.stop() if rank == 0 else None
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    sel_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    gpu_launches = launches[0]

    # ---- side kernels at the full config size, timed alone on this stream (rank 0 reports):
    #      K3 select over 20 000 bags x 3025 probabilities, K4b HSV refine over the resident bags
    def time_alone(fn, reps=5):
        fn(); torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for a, b in ev:
            flush.zero_()                      # > L2: evict the operands between repetitions
            a.record(); fn(); b.record()
        torch.cuda.synchronize()
        return min(a.elapsed_time(b) for a, b in ev)

    side = {}
    if rank == 0:
        nb_sel = 20000
        g = torch.Generator(device=dev); g.manual_seed(7)
        p_all = torch.rand(nb_sel * T_PER_BAG, device=dev, generator=g)
        lab_all = torch.from_numpy(synthetic.make_labels(nb_sel, seed=3)).to(dev)
        sel_out = {}

        def run_sel():
            sel_out["r"] = ops.select_topk(p_all, lab_all, nb_sel, T_PER_BAG, 1, 30, capacity=nb_sel * 330)
        ms = time_alone(run_sel)
        m_kept = int(sel_out["r"][0].numel())
        sel_bytes = 4.0 * p_all.numel() + 5.0 * m_kept + 4.0 * nb_sel
        side["select_20k"] = {"bound": "hbm", "bags": nb_sel, "instances": int(p_all.numel()), "kept": m_kept,
                              "ms": ms, "instances_per_s": p_all.numel() / (ms * 1e-3),
                              "achieved": sel_bytes / (ms * 1e-3) / 1e9, "unit": "GB/s",
                              "note": "4 launches (count, scan, fast select, exact fallback) + one .item() sync inside the timing"}
        del p_all, sel_out
        # K4b at the config-5 size class: 8000 bags (3.6 GB of traffic per launch, >> L2)
        nb_m = max(resident, 8000)
        imgs_m = bags if nb_m == resident else synthetic.make_bags_device(nb_m, dev, seed=99)
        blob = torch.rand((nb_m, H // 13 + 1, H // 13 + 1), device=dev, generator=g) < 0.45
        masks = blob.repeat_interleave(13, 1).repeat_interleave(13, 2)[:, :H, :H].contiguous().to(torch.uint8)
        out_m = torch.empty_like(masks)
        ms = time_alone(lambda: ops.hsv_refine(imgs_m, masks, 170, out=out_m))
        side["hsv_refine"] = {"bound": "hbm", "bags": nb_m, "ms": ms, "masks_per_s": nb_m / (ms * 1e-3),
                              "achieved": 5.0 * masks.numel() / (ms * 1e-3) / 1e9, "unit": "GB/s",
                              "bytes_per_mask": 5 * H * H}
        # N1: connected-component clean-up of the refined masks (2000 bags), and the whole
        # preprocess_masks chain (HSV AND + clean-up) in masks/s
        nb_c = 2000
        cc_in = out_m[:nb_c].clone()
        work = torch.empty_like(cc_in)

        def run_cc():
            work.copy_(cc_in)
            ops.remove_small_regions(work, 400, 120)
        ms_cc = time_alone(run_cc, reps=3)
        side["remove_small_regions"] = {"bags": nb_c, "ms": ms_cc, "masks_per_s": nb_c / (ms_cc * 1e-3),
                                        "note": "includes a device copy of the 2000 input masks"}
        side["preprocess_masks_chain"] = {"masks_per_s": 1.0 / (ms / nb_m * 1e-3 + ms_cc / nb_c * 1e-3)}
        del masks, out_m, cc_in, work
        # configs[3]: ResNeXt-50 32x4d backbone at a denser tile stride (interval 3 -> 8100
        # instances per bag), score + select, timed alone on a 64-bag slice
        nb_x, iv_x = 64, 3
        t_x = ((H - TILE + iv_x - 1) // iv_x + 1) ** 2
        cx, fwx, fbx = synthetic.make_resnet_weights("resnext50_32x4d", seed=0)
        clf_x = ops.TileClassifier("resnext50_32x4d", cx, fwx, fbx, device=dev)
        prob_x = torch.empty(nb_x * t_x, dtype=torch.float32, device=dev)

        def run_x():
            clf_x.forward_tiles(bags[:nb_x], TILE, iv_x, precision="bf16", max_batch=args.max_batch // 2,
                                prob_out=prob_x)
            ops.select_topk(prob_x, labels[:nb_x], nb_x, t_x, 1, 30, capacity=nb_x * 330)
        ms_x = time_alone(run_x, reps=3)
        side["resnext50_32x4d_dense_stride"] = {
            "workload": "configs[3]: ResNeXt-50 32x4d, tile 32 interval 3 (%d instances/bag), %d bags" % (t_x, nb_x),
            "ms": ms_x, "instances_per_s": nb_x * t_x / (ms_x * 1e-3), "launches": clf_x.last_launch_count + 4}
        clf_x.close()
        del prob_x

    # ---- end to end: host (pinned) bags -> H2D -> score + select -> D2H of the selection
    host = torch.empty((B, H, H, 3), dtype=torch.uint8).pin_memory()
    host.copy_(bags[:B].cpu())
    host_labels = torch.from_numpy(labels_h[:B].copy()).pin_memory()
    stage = torch.empty_like(bags[:B])
    d_lab = torch.empty(B, dtype=torch.int32, device=dev)
    out_idx = torch.empty(cap, dtype=torch.int32).pin_memory()
    out_lab = torch.empty(cap, dtype=torch.uint8).pin_memory()

    def step_e2e():
        stage.copy_(host, non_blocking=True)
        d_lab.copy_(host_labels, non_blocking=True)
        clf.forward_tiles(stage, TILE, INTERVAL, precision="bf16", max_batch=args.max_batch, prob_out=prob)
        idx, pl, off = ops.select_topk(prob, d_lab, B, T_PER_BAG, 1, 30, capacity=cap)
        m = idx.numel()
        out_idx[:m].copy_(idx, non_blocking=True)
        out_lab[:m].copy_(pl, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return m

    for _ in range(2):
        m_sel = step_e2e()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        m_sel = step_e2e()
    e1.record()
    sync_all()
    ms_e2e = e0.elapsed_time(e1)

    t = torch.tensor([ms_total, ms_e2e, fwd_ms, sel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, fwd_ms, sel_ms = t.tolist()

    if rank == 0:
        pk, pk_kind = peaks()
        value = world * n_inst * args.steps / (ms_total * 1e-3)
        e2e_v = world * n_inst * args.steps / (ms_e2e * 1e-3)
        tflops = FLOP_INBOUNDS * n_inst / (fwd_ms * 1e-3) / 1e12
        peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
        out = {
            "metric": "instances/sec (score+adaptive top-k)", "value": value, "unit": "instances/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "configs[1] ResNet-34 tile inference + adaptive top-k, tile 32 interval 5 "
                                   "(3025 instances/bag), bf16 tcgen05; step = %d-bag slice (%d instances) per GPU "
                                   "of %d HBM-resident bags, slice advances each step" % (B, n_inst, resident),
                       "bags_per_step_per_gpu": B, "instances_per_step_per_gpu": n_inst,
                       "max_batch": args.max_batch, "tiles_per_pos": 1, "topk_neg": 30,
                       "l2": "inputs (u8 slice %.0f MB + activation workspace) larger than L2; no flush" % (B * 268203 / 1e6),
                       "weights": "random-init ResNet-34, BN folded"},
            "e2e": {"value": e2e_v, "unit": "instances/s", "h2d_bytes_per_step": int(B * 268203 + 4 * B),
                    "d2h_bytes_per_step": int(m_sel * 5)},
            "gpu_launches": int(gpu_launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "conv stage: stem_win_kernel + 32 conv_halo/conv_gemm launches + head per batch",
                         "achieved": tflops, "peak": peak, "unit": "TFLOP/s", "frac": tflops / peak,
                         "peak_kind": pk_kind + " bf16_tflops_sustained",
                         "traffic": (measured_traffic(n_inst) or {}).get("bytes"),
                         "traffic_detail": measured_traffic(n_inst),
                         "flop_per_instance": FLOP_INBOUNDS,
                         "achieved_nominal": FLOP_NOMINAL * n_inst / (fwd_ms * 1e-3) / 1e12,
                         "fwd_ms_per_step": fwd_ms,
                         "select_in_step": {"bound": "hbm", "ms_per_step": sel_ms,
                                            "note": "the step's own selection over %d bags: four small launches + "
                                                    "the .item() sync, launch-latency bound; select_20k is the "
                                                    "kernel at the config's full size" % B,
                                    "achieved": SELECT_BYTES_PER_INST * n_inst / (sel_ms * 1e-3) / 1e9,
                                    "peak": pk["hbm_gbs"], "unit": "GB/s",
                                    "frac": SELECT_BYTES_PER_INST * n_inst / (sel_ms * 1e-3) / 1e9 / pk["hbm_gbs"]}},
        }
        for k, v in side.items():
            if "achieved" in v:
                v["peak"] = pk["hbm_gbs"]
                v["frac"] = v["achieved"] / pk["hbm_gbs"]
            out["roofline"][k] = v
        if not args.no_cpu_baseline:
            from oracle import model as omodel
            torch.set_num_threads(os.cpu_count())
            sd = omodel.make_state_dict("resnet34", seed=0, random_bn=False)
            lab8 = np.array([3, 0, 7, 1, 12, 0, 40, 5], np.int32)
            cpu_reference_step(1, sd, lab8, 99)
            n, s = cpu_reference_step(3, sd, lab8, 0)
            out["cpu_baseline"] = {"value": n / s, "unit": "instances/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": "3 bags x 3025 instances: unfold + fp32 forward + lexsort/top-k loop (oracle port)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
