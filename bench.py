#!/usr/bin/env python
"""bench.py — QuadtreeCNN train images/s @224^2, per-GPU batch 256, bf16 tensor-core compute (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A step = forward + CrossEntropyLoss + backward + Adam over one synthetic batch (SURVEY.md §8d recipe), i.e. the
hot loop of the reference's `Quadtree_from scratch/Quadtree_train.py:60-66`.
  value : whole-job images/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e   : the same step through the public module API with HOST (pinned) inputs: H2D copies of images /
          pose vectors / labels and the D2H read of the loss are inside the timed region
  roofline : dominant kernel family = tcgen05 implicit-GEMM convolutions (tensor bound); per-launch CUDA-event
          times in a second, event-instrumented pass; achieved = algorithmic conv FLOPs / summed launch time
  cpu_baseline : the UNMODIFIED reference model (baseline/_ref, staged by __graft_entry__.build()) driven like its own
          training loop on this box's host cores (fp32, torch CPU); the oracle port only if the staged files are missing
`--impl reference` times that same CPU path as the reference arm.
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

METRIC = "QuadtreeCNN train images/sec @224^2 bs256 bf16"
UNIT = "images/s"
TRAIN_GFLOP_PER_IMG = 11.08  # BASELINE.md §3 (fwd + dgrad + wgrad)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-batch", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="quadtree_train", choices=["quadtree_train", "attention_infer", "quadtree3d_train", "cnn_lstm_train"],
                    help="quadtree_train is the BASELINE.json headline (configs[2]); the other two time configs[1] / configs[3] "
                         "for profiles/ and print an informational line with their own metric name")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured (sustained)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# --------------------------------------------------------------------------------------------------- CPU arm
REF_DIR = os.path.join(ROOT, "baseline", "_ref")  # unmodified reference model files, copied there by __graft_entry__.build()


def load_reference_module(subdir="Quadtree_from scratch"):
    """The UNMODIFIED reference `models.py` (baseline/_ref/<subdir>/models.py), imported by path. The only shim is the
    one SURVEY.md §8(c) describes: `torchvision.models.resnet18(weights=IMAGENET1K_V1)` cannot download offline, so the
    `weights=` argument is dropped (random init; throughput does not depend on the weight values)."""
    import importlib.util
    path = os.path.join(REF_DIR, subdir, "models.py")
    if not os.path.exists(path):
        return None
    import torchvision
    if not getattr(torchvision.models.resnet18, "_offline_shim", False):
        orig = torchvision.models.resnet18

        def resnet18_offline(weights=None, **kw):
            return orig(weights=None, **kw)
        resnet18_offline._offline_shim = True
        torchvision.models.resnet18 = resnet18_offline
    spec = importlib.util.spec_from_file_location("reference_models_" + subdir.replace(" ", "_").replace("+", "_"), path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_reference_rate(batch, steps, warmup):
    """The reference's own CPU path, images/s: the unmodified `QuadtreeCNN` driven exactly like the hot loop of
    `Quadtree_from scratch/Quadtree_train.py:43-66` (get_model, nn.CrossEntropyLoss, optim.Adam(lr, weight_decay), zero_grad /
    forward / loss / backward / step, loss.item()), fp32, all host threads. Falls back to the oracle port only when
    baseline/_ref is missing; returns which one ran."""
    import contextlib
    import io
    import torch
    from qtcnn_b200.data import synthetic_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    images, numerical, labels = synthetic_batch(batch, 1234)
    ref = load_reference_module()
    if ref is not None:
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            model = ref.get_model(num_classes=8, device=torch.device("cpu"), model_name="quadtree")
        criterion = torch.nn.CrossEntropyLoss()
        optimizer = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4)  # Quadtree_train.py:21-22,45
        model.train()

        def one():
            optimizer.zero_grad()
            loss = criterion(model(images, numerical), labels)
            loss.backward()
            optimizer.step()
            return loss.item()
        kind = "reference"
    else:
        from oracle import quadtree_oracle as O
        p = O.make_params("quadtree", 8, seed=0)
        box = {"state": None}

        def one():
            loss, box["state"] = O.train_step_cpu("quadtree", p, (images, numerical), labels, state=box["state"])
            return float(loss)
        kind = "port"
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, torch.get_num_threads(), kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 6)), max(1, min(args.warmup, 2))
    rate, spi, threads, kind = cpu_reference_rate(args.cpu_batch, steps, warmup)
    what = "unmodified reference QuadtreeCNN + nn.CrossEntropyLoss + optim.Adam" if kind == "reference" else "oracle port (baseline/_ref missing)"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": spi * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"QuadtreeCNN level-1 fwd+bwd+Adam 224x224 batch {args.cpu_batch} on CPU ({what})"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{steps} steps of batch {args.cpu_batch} after {warmup} warm-up"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        # a mismatched collective should end the run in minutes, not after NCCL's default 10-minute watchdog
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    os.environ.setdefault("QTCNN_QUIET_PRETRAINED", "1")  # random init on purpose (no checkpoint offline)
    from qtcnn_b200 import data as D  # synthetic-input recipe (the product arm never imports oracle/)
    from qtcnn_b200 import models as M
    from qtcnn_b200 import ops, optim, parallel

    B = args.batch
    torch.manual_seed(0)
    model = M.QuadtreeCNN(num_classes=8).to(dev).train()  # dropout 0.5 active, as in the reference script
    dp = parallel.DataParallelGrads(model) if world > 1 else None
    params = [p for p in model.parameters() if p.requires_grad]
    # optim.Adam(model.parameters(), lr, weight_decay) of Quadtree_train.py:45 as one multi-tensor launch per step
    opt = optim.Adam(params, lr=1e-4, weight_decay=1e-4)
    images_h, numerical_h, labels_h = D.synthetic_batch(B, 1234 + rank)
    images_u8_h = D.quantize_images_u8(images_h).pin_memory()  # decoded-pixel form of the same batch (e2e input path)
    images_h, numerical_h, labels_h = images_h.pin_memory(), numerical_h.pin_memory(), labels_h.pin_memory()
    images, numerical, labels = images_h.to(dev), numerical_h.to(dev), labels_h.to(dev)

    def step(x, nf, y):
        opt.zero_grad(set_to_none=True)
        # forward + nn.CrossEntropyLoss (computed inside the fused head-tail kernel) + backward + Adam
        loss, _ = model.training_loss(x, nf, y)
        loss.backward()
        if dp is not None:
            dp.finish()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(3, args.warmup)):
        step(images, numerical, labels)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = ops.launches()
    ms = timed(lambda: step(images, numerical, labels), args.steps)
    launches = ops.launches() - n0
    clocks = sampler.stop() if sampler else None
    value = B * world * args.steps / (ms * 1e-3)

    # end to end through the package API (data.BatchPrefetcher + model.training_loss + optim.Adam): every step's inputs
    # start in pinned HOST memory and are copied inside the timed region (side stream, double-buffered staging, so the
    # copy of step i+1 overlaps the kernels of step i, like a DataLoader with pin_memory + non_blocking copies); the loss
    # of every step is read back to the host (asynchronous 4-byte D2H, consumed one step later, the way a training loop
    # logs losses without stalling the GPU). Headline e2e: images travel as uint8 pixels and are normalised on the device
    # (ToTensor + Normalize fused into the stem's packing kernel); `e2e_fp32_inputs`: normalised fp32 tensors, exactly what
    # the reference's DataLoader hands to `images.to(device)` (4x the PCIe bytes).
    loss_h = [{"buf": torch.zeros((), dtype=torch.float32).pin_memory(), "ev": torch.cuda.Event()} for _ in range(2)]

    def run_e2e(host_batch, steps):
        state = {"last": None}

        def loop(nsteps):
            def batches():
                for _ in range(nsteps):
                    yield host_batch
            for i, (x, nf, y) in enumerate(D.BatchPrefetcher(batches(), dev)):
                loss = step(x, nf, y).detach()
                slot = loss_h[i & 1]
                slot["buf"].copy_(loss, non_blocking=True)
                slot["ev"].record()
                if i > 0:
                    prev = loss_h[(i + 1) & 1]
                    prev["ev"].synchronize()
                    state["last"] = float(prev["buf"])
        loop(2)
        return timed(lambda: loop(steps), 1)

    if args.no_e2e:
        ms_e2e, e2e, ms_e2e32, e2e32 = float("nan"), None, float("nan"), None
    else:
        ms_e2e = run_e2e((images_u8_h, numerical_h, labels_h), args.steps)
        e2e = B * world * args.steps / (ms_e2e * 1e-3)
        ms_e2e32 = run_e2e((images_h, numerical_h, labels_h), args.steps)
        e2e32 = B * world * args.steps / (ms_e2e32 * 1e-3)
    h2d = images_u8_h.numel() + numerical_h.numel() * 4 + labels_h.numel() * 8
    h2d32 = images_h.numel() * 4 + numerical_h.numel() * 4 + labels_h.numel() * 8

    roofline = None
    if not args.no_roofline:
        # every rank runs the instrumented steps (their gradient all-reduces must match across ranks); rank 0 reports
        ops.profile_begin()
        for _ in range(min(3, args.steps)):
            step(images, numerical, labels)
        torch.cuda.synchronize()
        prof = ops.profile_end()
    if not args.no_roofline and rank == 0:
        peaks = load_peaks()
        gemm = {k: v for k, v in prof.items() if v["flops"] > 0}
        nsteps = min(3, args.steps)
        flops = sum(v["flops"] for v in gemm.values())
        tms = sum(v["ms"] for v in gemm.values())
        nlaunch = sum(v["n"] for v in gemm.values())
        # dominant kernel = the persistent slab convolution (conv3x3_kernel: 3x3/s1 fprop + dgrad)
        dom = {k: v for k, v in gemm.items() if k.startswith("conv3x3_kernel")}
        dflops, dms, dn = (sum(v[x] for v in dom.values()) for x in ("flops", "ms", "n"))
        ach = dflops / (dms * 1e-3) / 1e12 if dms > 0 else 0.0
        ncu = {}
        ncu_path = os.path.join(ROOT, "profiles", "ncu_gemm_summary.json")
        if os.path.exists(ncu_path):
            with open(ncu_path) as f:
                ncu = json.load(f)
        roofline = {"bound": "tensor", "achieved": ach, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": ach / peaks["tflops"],
                    "traffic": ncu.get("conv3x3_kernel", {}).get("dram_bytes_per_launch"), "peak_source": peaks["src"],
                    "kernel": "conv3x3_kernel (persistent tcgen05 slab convolution, 3x3/s1 fprop + dgrad)",
                    "launches_per_step": dn / nsteps, "ms_per_launch": dms / max(dn, 1), "ms_per_step": dms / nsteps,
                    "step_share": (dms / nsteps) / (ms / args.steps),
                    "algorithmic_flops_per_launch": dflops / max(dn, 1),
                    "tensor_pipe_active_pct_ncu": ncu.get("conv3x3_kernel", {}).get("tensor_active_pct"),
                    "conv_tc_util_pct_flop_weighted_ncu": ncu.get("flop_weighted_tensor_active_pct"),
                    "all_gemm": {"achieved": flops / (tms * 1e-3) / 1e12 if tms > 0 else 0.0, "launches_per_step": nlaunch / nsteps,
                                 "ms_per_step": tms / nsteps, "step_share": (tms / nsteps) / (ms / args.steps)},
                    "per_family": {k: {"tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12, "ms_per_step": v["ms"] / nsteps,
                                       "launches_per_step": v["n"] / nsteps}
                                   for k, v in gemm.items() if v["ms"] > 0},
                    # streaming kernels: algorithmic bytes / CUDA-event time vs the measured HBM copy peak
                    "hbm_families": {k: {"GBps": v["bytes"] / (v["ms"] * 1e-3) / 1e9, "frac_of_hbm_peak": v["bytes"] / (v["ms"] * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                         "ms_per_step": v["ms"] / nsteps, "calls_per_step": v["n"] / nsteps}
                                     for k, v in prof.items() if v.get("bytes", 0) > 0 and v["ms"] > 0},
                    "hbm_peak_GBps": peaks["hbm_gbs"]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, spi, threads, kind = cpu_reference_rate(args.cpu_batch, 4, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": kind,
               "sample": f"4 steps of batch {args.cpu_batch} (fwd+bwd+Adam, fp32) after 1 warm-up, {spi:.2f} s/step"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"QuadtreeCNN level-1 (ResNet-18 + 2x2 quadtree + 47 pose features) fwd+bwd+Adam, 224x224, "
                                   f"per-GPU batch {B}, dropout 0.5", "global_batch": B * world, "parallelism": f"dp{world}",
                       "l2": "per-step activations+grads (~3 GB) exceed the 126 MB L2, no explicit flush",
                       "optimizer": "qtcnn_b200.optim.Adam (multi-tensor kernel, torch.optim.Adam semantics), lr 1e-4, wd 1e-4",
                       "loss": "nn.CrossEntropyLoss semantics inside the fused head-tail kernel (model.training_loss)"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                    "inputs": "uint8 pixels + fp32 pose vectors + int64 labels from pinned host memory, normalised on the device"},
            "e2e_fp32_inputs": {"value": e2e32, "unit": UNIT, "h2d_bytes_per_step": h2d32, "d2h_bytes_per_step": 4,
                                "ms_per_step": ms_e2e32 / args.steps},
            "gpu_launches": launches,
            "model_tflops": value * TRAIN_GFLOP_PER_IMG / 1e3 / world,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_secondary(args):
    """Informational timings of BASELINE.json configs[1] (level-1+2 inference) and configs[3] (3-D model training)."""
    import torch
    import torch.nn.functional as F
    from qtcnn_b200 import data as O
    from qtcnn_b200 import models as M
    from qtcnn_b200 import ops
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    if args.workload == "attention_infer":
        B = args.batch
        model = M.AttentionHierarchicalCNN(num_classes=8).to(dev).eval()
        images, numerical, _ = O.synthetic_batch(B, 1234)
        images, numerical = images.to(dev), numerical.to(dev)

        def step():
            with torch.no_grad():
                return model(images, numerical)
        metric, unit, per_step = "AttentionHierarchicalCNN (level-1+2) inference images/sec @224^2 bf16", "images/s", B
        flops = 3.977e9 * B
    elif args.workload == "cnn_lstm_train":
        # BASELINE.json configs[4] / SURVEY §8 a17: 16 frames x 224^2 per sample, frozen ResNet-18 (3.63 GFLOP/frame forward)
        B = 16 if args.batch == 256 else args.batch
        model = M.get_model_seq("cnn_lstm", 8, dev, seq_len=16).train()
        opt = torch.optim.Adam([q for q in model.parameters() if q.requires_grad], lr=1e-4, fused=True)
        frames, numerical, labels = O.synthetic_batch(B, 1234, seq_len=16, clip_size=224)
        frames, numerical, labels = frames.to(dev), numerical.to(dev), labels.to(dev)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = F.cross_entropy(model(frames, numerical), labels)
            loss.backward()
            opt.step()
            return loss
        metric, unit, per_step = "CnnLstm train frames/sec 16x224^2 bf16 (frozen ResNet-18 + LSTM)", "frames/s", B * 16
        flops = 3.63e9 * B * 16
    else:
        B = 32 if args.batch == 256 else args.batch
        model = M.Quadtree3DCNN(num_classes=8, sequence_length=16).to(dev).train()
        opt = torch.optim.Adam(model.parameters(), lr=5e-5, weight_decay=5e-4, fused=True)
        clips, numerical, labels = O.synthetic_batch(B, 1234, seq_len=16, clip_size=112)
        clips, numerical, labels = clips.to(dev), numerical.to(dev), labels.to(dev)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = F.cross_entropy(model(clips, numerical), labels)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)  # 3dcnn/train_3D_Quadtree_cnn_model.py:123
            opt.step()
            return loss
        metric, unit, per_step = "Quadtree3DCNN train clips/sec 16x112x112 bf16", "clips/s", B
        flops = 39.5e9 * B
    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    n0 = ops.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    value = per_step * args.steps / (ms * 1e-3)
    print(json.dumps({"metric": metric, "value": value, "unit": unit, "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
                      "ms_per_step": ms / args.steps, "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": args.workload, "batch": B}, "gpu_launches": ops.launches() - n0,
                      "model_tflops": flops * args.steps / (ms * 1e-3) / 1e12, "informational": True}), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload != "quadtree_train":
        run_secondary(a)
    else:
        run_ours(a)
