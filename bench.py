#!/usr/bin/env python
"""bench.py -- the stereo hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): stereo pairs/s for costvol + dispreg + reproj at 544x960,
D = 192, and the fraction of the HBM roofline the dominant kernel reaches.

A *step* is one pass of the forward hot path over one batch of B = 8 synthetic
stereo pairs (BASELINE config 2, "PSMNet inference at 544x960, D=192, batch 8"):
  1. concat cost volume      [8,32,136,240] x2 -> [8,64,48,136,240]   (a1)
  2. soft-argmin             [8,192,544,960] logits -> [8,1,544,960]  (a4)
  3. patch reprojection loss (ps = 11, masked MSE) of the IR patterns warped by
     the predicted disparity, + the Fold visualisation image            (a7)
Weak scaling: every rank processes its own batch of 8 pairs; no collective on
the data path.  `value` times the step with inputs resident in
HBM; `e2e` times the same calls with every input copied from pinned host memory
and the results (disparity + loss) copied back inside the timed region.

`--impl reference` times the reference's CPU implementation of the same step
(the torch restatement in oracle/, all host threads) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "stereo pairs/s, costvol+dispreg+reproj @544x960 D=192"
UNIT = "pairs/s"
B, C, H, W, D, PS = 8, 32, 544, 960, 192, 11
HQ, WQ, DQ = H // 4, W // 4, D // 4
WORKLOAD = (f"config2: PSMNet inference hot path fwd, batch {B} x {H}x{W}, D={D}: concat volume "
            f"[{B},{2 * C},{DQ},{HQ},{WQ}] + soft-argmin [{B},{D},{H},{W}] + patch reprojection loss ps={PS} (+fold image); "
            f"logits = trilinear upsample of random [{DQ},{HQ},{WQ}] logits as psmnet.py:186-197 produces them")

# algorithmic bytes per launch (SURVEY.md §8d formulas x B pairs), fp32
ALGO_BYTES = {
    "concat_volume_fwd": 4 * (2 * C * HQ * WQ + 2 * C * DQ * HQ * WQ) * B,
    "soft_argmin_fwd": 4 * (D * H * W + H * W) * B,
    "reproj_patch_loss+fold_fwd": (4 * (2 * 1 * H * W + H * W) + H * W + 4 * H * W) * B,  # + fold image out
}


# training hot path per pair: concat fwd+bwd, 3 x soft-argmin fwd+bwd, patch reprojection fwd+bwd
TRAIN_B, TRAIN_STEPS = 2, 10
TRAIN_BYTES_PER_PAIR = (2 * 4 * (2 * C * HQ * WQ + 2 * C * DQ * HQ * WQ) + 3 * (4 * (D * H * W + H * W) + 4 * (2 * D * H * W + 2 * H * W))
                        + 4 * 3 * H * W + H * W + 4 * H * W)
BOUND = {"concat_volume_fwd": "hbm", "soft_argmin_fwd": "hbm", "reproj_patch_loss+fold_fwd": "shared-memory bandwidth (LSU wavefronts)"}


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic():
    """dram bytes per launch from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            pass
    return {}


# ----------------------------------------------------------------------------------------------
# clocks sampled DURING the timed regions
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
           "hw_power_brake_slowdown": 0x80}
    NOTE = {"sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._active = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception as e:  # pragma: no cover
            self._nv, self.error = None, repr(e)

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                    for name, bit in {**self.BAD, **self.NOTE}.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.002)

    def start(self):
        self._active.set()

    def pause(self):
        self._active.clear()

    def result(self):
        self._stop.set()
        if self._nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": int(statistics.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's torch restatement on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_reference_step(inputs, so):
    """One pass of the same forward step for `inputs` (a 1-pair sample), stock torch on CPU."""
    L, R, cost, pat_L, pat_R, mask = inputs
    vol = so.concat_volume(L, R, DQ)
    disp = so.soft_argmin(cost)
    loss, vis, _ = so.reproj_error_patch(pat_L, pat_R, disp, mask, ps=PS)
    return vol, disp, loss, vis


def make_inputs(nb, seed, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    L = torch.randn(nb, C, HQ, WQ, generator=g)
    R = torch.randn(nb, C, HQ, WQ, generator=g)
    # The logits PSMNet feeds to softmax are ALWAYS the trilinear upsample of the [1,D/4,H/4,W/4] output of the
    # 3-D aggregation (nets/psmnet/psmnet.py:186-197), so the synthetic logits are built the same way from
    # random low-resolution logits: spatially smooth, like the disparity map regressed from them.
    cost = torch.empty(nb, D, H, W)
    for b in range(nb):  # generated per pair to bound the temporary
        low = torch.randn(1, 1, DQ, HQ, WQ, generator=g) * 4.0
        cost[b] = torch.nn.functional.interpolate(low, size=(D, H, W), mode="trilinear", align_corners=False)[0, 0]
    pat_L = (torch.rand(nb, 1, H, W, generator=g) > 0.5).float()
    pat_R = (torch.rand(nb, 1, H, W, generator=g) > 0.5).float()
    mask = torch.rand(nb, 1, H, W, generator=g) > 0.2
    ts = [L, R, cost, pat_L, pat_R, mask]
    if pin:
        ts = [t.pin_memory() for t in ts]
    if device != "cpu":
        ts = [t.to(device) for t in ts]
    return ts


def time_cpu_reference(steps, warmup):
    from oracle import stereo_oracle as so

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    inputs = make_inputs(1, 1234)
    with torch.no_grad():
        for _ in range(warmup):
            cpu_reference_step(inputs, so)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_reference_step(inputs, so)
        dt = time.perf_counter() - t0
    return steps * 1 / dt, dt / steps * 1e3, cores


def run_reference(args, rank):
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)  # each step is a bounded 1-pair sample (~1 s)
    pairs_s, ms, cores = time_cpu_reference(steps, warmup)
    sample = f"1 pair of the {B}-pair batch per step, {steps} timed steps after {warmup} warm-up, torch {torch.__version__} CPU"
    line = {
        "impl": "reference", "metric": METRIC, "value": pairs_s, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": pairs_s, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": pairs_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------
# the B200 arm
# ----------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist

    from activezero_b200 import _lib, dist_util, ops
    from activezero_b200.utils import reprojection as az_rp

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist_util.init_from_env("nccl", dev)
    _lib.load()
    # weak scaling: the job is world*B pairs, rank r owns pairs [first, first+B)
    first_pair, n_pairs = dist_util.shard_pairs(world * B, rank, world)
    assert n_pairs == B

    host = make_inputs(n_pairs, 1000 + first_pair, pin=True)
    L, R, cost, pat_L, pat_R, mask = [t.to(dev, non_blocking=True) for t in host]
    torch.cuda.synchronize()
    names = ["concat_volume_fwd", "soft_argmin_fwd", "reproj_patch_loss+fold_fwd"]

    def step(Ld, Rd, costd, pLd, pRd, md, evs=None):
        if evs is not None:
            evs[0].record()
        vol = ops.build_concat_volume(Ld, Rd, DQ)
        if evs is not None:
            evs[1].record()
        disp = ops.soft_argmin(costd)
        if evs is not None:
            evs[2].record()
        loss, vis = ops.reproj_loss(pLd, pRd, disp, md, ps=PS, sign=-1.0, want_warped=True)
        if evs is not None:
            evs[3].record()
        return vol, disp, loss, vis

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    with torch.no_grad():
        for _ in range(args.warmup):
            out = step(L, R, cost, pat_L, pat_R, mask)
        del out
        # ---- device-resident timing (value) + per-kernel events for the roofline ----
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = _lib.kernel_launches
        barrier()
        clocks.start()
        start.record()
        for k in range(args.steps):
            out = step(L, R, cost, pat_L, pat_R, mask, evs[k])
        stop.record()
        barrier()
        clocks.pause()
        launches = _lib.kernel_launches - launches0
        ms_total = start.elapsed_time(stop)
        del out

        # ---- end to end: pinned host -> device, compute, results back, every step ----
        e2e_steps = max(1, min(args.steps, 10))
        res_disp = torch.empty((B, 1, H, W), dtype=torch.float32).pin_memory()
        res_loss = torch.empty((), dtype=torch.float32).pin_memory()
        h2d = sum(t.numel() * t.element_size() for t in host)
        d2h = res_disp.numel() * 4 + 4

        def e2e_step():
            dv = [t.to(dev, non_blocking=True) for t in host]
            _, disp, loss, _ = az_call(dv)
            res_disp.copy_(disp, non_blocking=True)
            res_loss.copy_(loss, non_blocking=True)

        def az_call(dv):
            # the call a user makes: the reference-named functions of the drop-in modules
            vol = ops.build_concat_volume(dv[0], dv[1], DQ)
            disp = ops.soft_argmin(dv[2])
            loss, vis, _ = az_rp.get_reproj_error_patch(dv[3], dv[4], disp, dv[5], ps=PS)
            return vol, disp, loss, vis

        for _ in range(2):
            e2e_step()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        clocks.start()
        s2.record()
        for _ in range(e2e_steps):
            e2e_step()
        e2.record()
        barrier()
        clocks.pause()
        ms_e2e = s2.elapsed_time(e2)
        loss_val = float(res_loss)

        # ---- informational variant: the same step when the soft-argmin is fed by the
        # LOW-RESOLUTION logits and fused with the trilinear upsample, i.e. the [B,192,544,960] tensor is
        # never materialised (nor shipped over PCIe).  Not part of `value` / `e2e`.
        low_host = (torch.randn(n_pairs, 1, DQ, HQ, WQ, generator=torch.Generator().manual_seed(7 + first_pair)) * 4.0).pin_memory()
        low_dev = low_host.to(dev)

        def fused_step(Ld, Rd, lowd, pLd, pRd, md):
            vol = ops.build_concat_volume(Ld, Rd, DQ)
            disp = ops.upsample_soft_argmin(lowd, (D, H, W))
            loss, vis = ops.reproj_loss(pLd, pRd, disp, md, ps=PS, sign=-1.0, want_warped=True)
            return vol, disp, loss, vis

        def fused_e2e_step():
            dv = [host[0].to(dev, non_blocking=True), host[1].to(dev, non_blocking=True),
                  low_host.to(dev, non_blocking=True)] + [t.to(dev, non_blocking=True) for t in host[3:]]
            _, disp, loss, _ = fused_step(*dv)
            res_disp.copy_(disp, non_blocking=True)
            res_loss.copy_(loss, non_blocking=True)

        for _ in range(3):
            fused_step(L, R, low_dev, pat_L, pat_R, mask)
            fused_e2e_step()
        f0, f1, f2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        barrier()
        f0.record()
        for _ in range(args.steps):
            fused_step(L, R, low_dev, pat_L, pat_R, mask)
        f1.record()
        for _ in range(e2e_steps):
            fused_e2e_step()
        f2.record()
        barrier()
        ms_fused, ms_fused_e2e = f0.elapsed_time(f1), f1.elapsed_time(f2)
        h2d_fused = h2d - host[2].numel() * 4 + low_host.numel() * 4

        # ---- informational variant: the TRAINING hot path (BASELINE target "fwd/bwd"), forward and backward of
        # concat volume + three soft-argmin heads (psmnet.py:200-217) + patch reprojection loss on the last head,
        # at the same frame size, TRAIN_B pairs per step, device-resident.  The upstream gradients of the volume
        # and of the three heads are synthetic tensors (in the network they come from dres0 and the smooth-L1 loss).
        del low_dev
        ms_train = float("nan")
        if not args.no_train_variant:
            tb = TRAIN_B
            with torch.enable_grad():
                Lt, Rt = L[:tb].clone().requires_grad_(True), R[:tb].clone().requires_grad_(True)
                heads = [cost[:tb].clone().requires_grad_(True) for _ in range(3)]
                gvol = torch.randn(tb, 2 * C, DQ, HQ, WQ, device=dev)
                gheads = [torch.randn(tb, 1, H, W, device=dev) for _ in range(3)]

                def train_step():
                    for t in [Lt, Rt] + heads:
                        t.grad = None
                    vol = ops.build_concat_volume(Lt, Rt, DQ)
                    d1, d2, d3 = (ops.soft_argmin(c_) for c_ in heads)
                    loss, _, _ = az_rp.get_reproj_error_patch(pat_L[:tb], pat_R[:tb], d3, mask[:tb], ps=PS)
                    torch.autograd.backward([vol, d1, d2, d3, loss], [gvol] + gheads + [None])

                for _ in range(3):
                    train_step()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                t0.record()
                for _ in range(TRAIN_STEPS):
                    train_step()
                t1.record()
                barrier()
                ms_train = t0.elapsed_time(t1)
                del Lt, Rt, heads, gvol, gheads

    ms_total, ms_e2e, ms_fused, ms_fused_e2e, ms_train = dist_util.max_over_ranks(
        [ms_total, ms_e2e, ms_fused, ms_fused_e2e, ms_train], dev)

    if rank == 0:
        per_kernel = {}
        for i, n in enumerate(names):
            per_kernel[n] = statistics.fmean(evs[k][i].elapsed_time(evs[k][i + 1]) for k in range(args.steps))
        peak, peak_src = _peaks()
        traffic = _traffic()
        kernels = []
        for n in names:
            gbs = ALGO_BYTES[n] / (per_kernel[n] * 1e-3) / 1e9
            kernels.append({"kernel": n, "ms": per_kernel[n], "algo_bytes": ALGO_BYTES[n], "achieved_gbs": gbs,
                            "frac": gbs / peak, "share": per_kernel[n] / sum(per_kernel.values()),
                            "traffic": traffic.get(n), "bound": BOUND[n]})
        # the roofline object is an HBM roofline: it describes the slowest of the HBM-bound kernels; the
        # patch kernel (121 taps per pixel out of shared memory, ~9 MB of compulsory traffic per pair) is
        # listed with its share in `kernels` and cannot be placed on a bandwidth roofline meaningfully
        dom = max((k for k in kernels if k["bound"] == "hbm"), key=lambda k: k["ms"])
        line = {
            "metric": METRIC, "value": world * B * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "pairs_per_gpu_per_step": B, "sharding": "batch (pairs) per rank, no data-path collective",
                       "l2": "inputs larger than L2: 6.5 GB streamed per step vs 126 MB L2"},
            "clocks": clocks.result(),
            "e2e": {"value": world * B * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                    "api": "ops.build_concat_volume + ops.soft_argmin + utils.reprojection.get_reproj_error_patch"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak,
                         "unit": "GB/s", "frac": dom["frac"], "traffic": dom["traffic"], "peak_source": peak_src,
                         "selection": "slowest HBM-bound kernel of the step (see kernels[] for all shares)",
                         "hbm_bound_share_of_step": sum(k["share"] for k in kernels if k["bound"] == "hbm"),
                         "pipeline_frac": sum(ALGO_BYTES.values()) / (ms_total / args.steps * 1e-3) / 1e9 / peak},
            "kernels": kernels,
            "variant_fused_upsample": {
                "note": "informational, SURVEY §8f-1: soft-argmin fused with the trilinear upsample reads the "
                        "[B,1,48,136,240] low-res logits; not the BASELINE config, not part of value/e2e",
                "value": world * B * args.steps / (ms_fused * 1e-3), "ms_per_step": ms_fused / args.steps,
                "e2e_value": world * B * e2e_steps / (ms_fused_e2e * 1e-3), "h2d_bytes_per_step": h2d_fused, "unit": UNIT},
            "loss_check": loss_val,
        }
        if not args.no_train_variant:
            pairs_s = world * TRAIN_B * TRAIN_STEPS / (ms_train * 1e-3)
            line["variant_train_fwd_bwd"] = {
                "note": "informational: training hot path at the same frame size -- concat volume fwd+bwd, three "
                        "soft-argmin heads fwd+bwd, patch reprojection loss (ps=11, with Fold image) fwd+bwd; "
                        "synthetic upstream gradients; not part of value/e2e",
                "value": pairs_s, "unit": UNIT, "pairs_per_gpu_per_step": TRAIN_B, "steps": TRAIN_STEPS,
                "ms_per_step": ms_train / TRAIN_STEPS, "algo_bytes_per_pair": TRAIN_BYTES_PER_PAIR,
                "hbm_frac": pairs_s / world * TRAIN_BYTES_PER_PAIR / 1e9 / peak}
        if world == 1 and not args.no_cpu_baseline:
            pairs_s, _, cores = time_cpu_reference(12, 1)  # ~10 s of CPU work on the box's host cores
            line["cpu_baseline"] = {"value": pairs_s, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "1 pair of the 8-pair batch per step, 12 timed steps after 1 warm-up"}
        emit(line)
    else:
        clocks.result()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    banner to stdout when the communicator is created), so fd 1 is pointed at stderr for the whole run
    and the JSON line is written to the saved descriptor."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-variant", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        args.warmup = max(args.warmup, 3)  # timing rule: at least 3 warm-up steps
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
