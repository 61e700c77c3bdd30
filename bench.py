#!/usr/bin/env python
"""bench.py -- headline benchmark of the purification hot path (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload purify|pgd|gender|cars]

"step" = one pass of the hot path over one batch of synthetic input:
  workload purify (default, BASELINE configs[1]): NVAE CelebA-64 "ids" + configs/ours_learned_blur_ids.yaml
      (blur on, eps 0) + VGG11 classifier, batch 512 per GPU, bf16 tensor-core path, random-init weights of the
      C32 architecture, synthetic images.  metric = purified img/s.
  workload pgd (BASELINE configs[4]): PGD-Linf (eps 8/255, step 2/255, 50 steps) through purifier + classifier.
  workload gender / cars (BASELINE configs[2] / [3]): StyleGAN-E4E @1024 + ResNet-50, batch 128 / Style-Transformer @512 +
      ResNeXt-50, batch 256 (extra measurements; the driver's headline line stays configs[1]).

value  : whole-job img/s with inputs resident in HBM, device-timed (CUDA events), max over ranks.
e2e    : the same through the reference-facing API (`NVAEDefenseModel.__call__`) from pinned HOST buffers, with the
         H2D copy of the batch and the D2H read of the logits inside the timed region.
roofline: the tcgen05 implicit-GEMM conv kernel (dominant), FLOP-weighted over all its launches in the timed
         region, timed per launch with CUDA events on the launching stream, against MEASURED_PEAKS.json.
cpu_baseline: the oracle (CPU restatement of the reference, kind "port") on the host cores, bounded sample.
--impl reference: the oracle port on the host cores for the same config, bounded sample per step.
Multi-GPU (torchrun): batch sharded data-parallel, no data-path collective; the only collective is one
all-reduce of the int64[3] accuracy counters.  scaling = weak (fixed 512 images per GPU).
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

import torch  # noqa: E402

LEARNED_BLUR_IDS = {   # /root/reference/configs/ours_learned_blur_ids.yaml, verbatim
    "interpolation_alphas": [0.000, 0.000, 0.001, 0.136, 0.131, 0.206, 0.179, 0.305, 0.347, 0.349, 0.465, 0.528, 0.551,
                             0.606, 0.681, 0.676, 0.834, 0.800, 0.938, 0.911, 1.000, 1.000, 1.000, 1.000],
    "alpha_attenuation": 0.7, "initial_noise_eps": 0.0, "gaussian_blur_input": True}
COSINE_NOISE_IDS = {   # /root/reference/configs/ours_cosine_noise_ids.yaml, verbatim
    "interpolation_alphas": [0.00, 0.02, 0.04, 0.07, 0.10, 0.15, 0.20, 0.25, 0.31, 0.37, 0.43, 0.50, 0.57, 0.63, 0.69,
                             0.75, 0.80, 0.85, 0.90, 0.93, 0.96, 0.98, 1.00, 1.00],
    "alpha_attenuation": 0.7, "initial_noise_eps": 2.0, "gaussian_blur_input": False}

# algorithmic work (unpruned, conv/linear MACs x2): purifier 15.31 + VGG11 2.49 GFLOP per image
GFLOP_PER_IMAGE_FWD = 17.8


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ----------------------------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import nvae_ref
    from gen_adversarial_b200 import synth
    from gen_adversarial_b200.nvae_spec import NvaeSpec, NVAE_C32_CONFIG, NVAE_C32_RESOLUTION
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = LEARNED_BLUR_IDS
    spec = NvaeSpec(NVAE_C32_CONFIG, NVAE_C32_RESOLUTION)
    sd = synth.make_nvae_state_dict(seed=0)
    vgg = nvae_ref.build_vgg11(synth.make_vgg11_state_dict(100, seed=1), 100)
    sample = args.ref_batch
    x, y = synth.synthetic_batch(sample, seed=42)
    alphas = [a * cfg["alpha_attenuation"] for a in cfg["interpolation_alphas"]]

    def step():
        g = torch.Generator().manual_seed(int(time.time() * 1e3) % (2 ** 31))
        noises = [torch.randn(s, generator=g) for s in spec.noise_shapes(sample)]   # the reference draws fresh noise per call
        with torch.no_grad():
            logits, _ = nvae_ref.defense_call(sd, spec, vgg, x, alphas, noises, cfg["initial_noise_eps"], cfg["gaussian_blur_input"])
        return logits

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    line = {"impl": "reference", "metric": "purified_img_per_s", "value": val, "unit": "img/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "NVAE-C32 CelebA-64 ids, ours_learned_blur_ids.yaml + VGG11, CPU oracle port of the reference path",
                       "sample_batch": sample},
            "cpu_baseline": {"value": val, "unit": "img/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} steps x batch {sample} of the same workload (oracle/nvae_ref.py, torch CPU fp32, {cores} threads)"},
            "e2e": {"value": val, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=JSON_OUT or sys.stdout, flush=True)


# ----------------------------------------------------------------------------------------------- our arm (GPU)
def cpu_baseline(sample_batch=16, runs=2):
    from oracle import nvae_ref
    from gen_adversarial_b200 import synth
    from gen_adversarial_b200.nvae_spec import NvaeSpec, NVAE_C32_CONFIG, NVAE_C32_RESOLUTION
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = LEARNED_BLUR_IDS
    spec = NvaeSpec(NVAE_C32_CONFIG, NVAE_C32_RESOLUTION)
    sd = synth.make_nvae_state_dict(seed=0)
    vgg = nvae_ref.build_vgg11(synth.make_vgg11_state_dict(100, seed=1), 100)
    x, _ = synth.synthetic_batch(sample_batch, seed=42)
    noises = synth.synthetic_noise(spec, sample_batch, seed=7)
    alphas = [a * cfg["alpha_attenuation"] for a in cfg["interpolation_alphas"]]
    times = []
    for i in range(runs + 1):
        t0 = time.perf_counter()
        with torch.no_grad():
            nvae_ref.defense_call(sd, spec, vgg, x, alphas, noises, cfg["initial_noise_eps"], cfg["gaussian_blur_input"])
        times.append(time.perf_counter() - t0)
    t = statistics.median(times[1:])
    return {"value": sample_batch / t, "unit": "img/s", "cores": cores, "kind": "port",
            "sample": f"batch {sample_batch}, 1 warm-up + median of {runs} calls of the same workload (oracle/nvae_ref.py, torch CPU fp32)"}


def run_ours(args):
    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py (impl ours) needs a GPU: the product has no CPU path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the ONE JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=dev)
    from gen_adversarial_b200 import ops, synth
    from gen_adversarial_b200.nvae_spec import NvaeSpec, NVAE_C32_CONFIG, NVAE_C32_RESOLUTION
    from gen_adversarial_b200.defenses.ours.models import NVAEDefenseModel, CelebaIdentityClassifier

    mode = args.mode
    sg = args.workload in ("gender", "cars")
    B = args.batch if args.batch is not None else {"purify": 512, "pgd": 1024, "gender": 128, "cars": 256}[args.workload]
    if sg:
        # BASELINE configs[2] / [3]: StyleGAN-E4E @1024 + ResNet-50 (ours_linear_noise_gender.yaml) and
        # Style-Transformer @512 + ResNeXt-50 (ours_cosine_blur_cars.yaml); YAML values verbatim
        from gen_adversarial_b200.defenses.ours.models import (E4EStyleGanDefenseModel, TransStyleGanDefenseModel,
                                                               CelebaGenderClassifier, CarsTypeClassifier)
        if args.workload == "gender":
            cfg = {"interpolation_alphas": [0.05, 0.11, 0.16, 0.22, 0.27, 0.33, 0.38, 0.44, 0.50, 0.55, 0.61, 0.66, 0.72, 0.77, 0.83, 0.88, 0.94, 1.00],
                   "alpha_attenuation": 1.0, "initial_noise_eps": 4.0, "gaussian_blur_input": False}
            clf = CelebaGenderClassifier(synth.make_resnet50_checkpoint(), dev, mode=mode)
            dm = E4EStyleGanDefenseModel(clf, synth.make_e4e_checkpoint(1024), cfg["interpolation_alphas"], cfg["alpha_attenuation"],
                                         cfg["initial_noise_eps"], cfg["gaussian_blur_input"], dev, mode=mode).eval()
            res, n_cls = (3, 256, 256), 2
        else:
            cfg = {"interpolation_alphas": [0.010, 0.038, 0.084, 0.146, 0.222, 0.309, 0.402, 0.500, 0.598, 0.691, 0.778, 0.854, 0.916, 0.962, 0.990, 1.000],
                   "alpha_attenuation": 0.7, "initial_noise_eps": 0.0, "gaussian_blur_input": True}
            clf = CarsTypeClassifier(synth.make_resnext50_checkpoint(), dev, mode=mode)
            dm = TransStyleGanDefenseModel(clf, synth.make_trans_checkpoint(512), cfg["interpolation_alphas"], cfg["alpha_attenuation"],
                                           cfg["initial_noise_eps"], cfg["gaussian_blur_input"], dev, mode=mode).eval()
            res, n_cls = (3, 128, 128), 4
        if args.chunk:
            dm.max_chunk = args.chunk
    else:
        cfg = COSINE_NOISE_IDS if args.workload == "pgd" else LEARNED_BLUR_IDS
        spec = NvaeSpec(NVAE_C32_CONFIG, NVAE_C32_RESOLUTION)
        nv = synth.make_nvae_checkpoint(seed=0)
        vg = {"state_dict": synth.make_vgg11_state_dict(100, seed=1, device=str(dev))}
        clf = CelebaIdentityClassifier(vg, dev, mode=mode)
        del vg
        dm = NVAEDefenseModel(clf, nv, cfg["interpolation_alphas"], cfg["alpha_attenuation"], cfg["initial_noise_eps"],
                              cfg["gaussian_blur_input"], dev, mode=mode).eval()
        res, n_cls = NVAE_C32_RESOLUTION, 100
    dm.sample_offset = rank * B                      # Philox streams keyed by the GLOBAL sample index
    use_graph = (args.cuda_graph == 1) or (args.cuda_graph < 0 and args.workload == "pgd")
    if use_graph and not sg:
        dm.enable_cuda_graph(True)                   # public API switch: whole call / PGD iteration replayed as a CUDA graph
    x_cpu, y_cpu = synth.synthetic_batch(B, res, n_cls, seed=42 + rank)
    x_host = x_cpu.pin_memory()
    x_dev = x_host.to(dev)
    y_dev = y_cpu.to(dev)
    counters = torch.zeros(3, dtype=torch.int64, device=dev)       # n_total, n_clean_correct, n_robust_correct
    logits_host = torch.empty((B, n_cls), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pgd = args.workload == "pgd"
    if pgd:
        from gen_adversarial_b200.attacks import PGDLinf
        attack = PGDLinf(8 / 255, 2 / 255, args.pgd_steps)
        with torch.no_grad():
            y_dev = dm(x_dev).argmax(dim=1)          # attack the model's own clean predictions (labels are synthetic)
        succ_host = torch.empty((B,), dtype=torch.bool).pin_memory()

    def step_resident():
        if pgd:
            succ, _, _ = attack(x_dev, y_dev, dm)
            counters[2] += (~succ).sum()
            return succ
        with torch.no_grad():
            logits = dm(x_dev)
        _, _, pred = ops.softmax_xent(logits, y_dev, want_grad=False, counter=counters[1:2].view(torch.int64))
        return logits

    def step_e2e():
        xd = x_host.to(dev, non_blocking=True)
        if pgd:
            succ, _, _ = attack(xd, y_dev, dm)
            succ_host.copy_(succ, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return succ
        with torch.no_grad():
            logits = dm(xd)
        logits_host.copy_(logits, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return logits

    # ---------------- resident timing
    for _ in range(args.warmup):
        step_resident()
    barrier()
    ops.launch_count(reset=True)
    from gen_adversarial_b200 import graphs as ga_graphs
    replayed0 = ga_graphs.REPLAYED_LAUNCHES[0]
    timer = ops.KernelTimer() if rank == 0 else None
    ops.TIMER = timer
    ops.TIME_ALL = bool(args.breakdown)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    from gen_adversarial_b200 import graphs as ga_graphs
    launches = ops.launch_count(reset=True) + ga_graphs.REPLAYED_LAUNCHES[0] - replayed0      # eager launches + graph-replayed kernel nodes
    ops.TIMER = None
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    counters[0] = B * args.steps
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)       # the path's only collective (24 bytes)
    ms = float(t_ms.item())
    value = world * B * args.steps / (ms / 1e3)

    # ---------------- end-to-end timing (host buffers, copies inside the timed region)
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    barrier()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_val = world * B * args.steps / (float(ms_e2e.item()) / 1e3)

    if rank == 0:
        pk = peaks()
        agg_all = timer.summary()
        if args.breakdown:
            rows = sorted(agg_all.items(), key=lambda kv: -kv[1]["ms"])
            print(f"{'op / shape':70s} {'n':>6s} {'ms/step':>9s} {'us/launch':>10s}", file=sys.stderr)
            for k, a in rows[:60]:
                print(f"{k[:70]:70s} {a['launches']:6d} {a['ms'] / args.steps:9.3f} {1e3 * a['ms'] / a['launches']:10.1f}", file=sys.stderr)
            print(f"sum of bracketed ops: {sum(a['ms'] for a in agg_all.values()) / args.steps:.2f} ms/step; step {ms / args.steps:.2f} ms", file=sys.stderr)
        agg = {k: a for k, a in agg_all.items() if not k.startswith(("op:", "fused:", "hbm:"))}
        fused = {k: a for k, a in agg_all.items() if k.startswith("fused:")}
        hbm = {k: a for k, a in agg_all.items() if k.startswith("hbm:")}
        tot_ms = sum(a["ms"] for a in agg.values())
        tot_fl = sum(a["flops"] for a in agg.values())
        n_l = sum(a["launches"] for a in agg.values())
        ach = tot_fl / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0
        top = sorted(agg.items(), key=lambda kv: -kv[1]["ms"])[:8]
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv, all shapes, FLOP-weighted)",
                "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({pk['src']})",
                # ncu --set full (profiles/r01_ncu_full_kernels_v7.md), per launch of the shape with the largest share of conv_tc time
                # (3x3, 32x32, C 64 -> 64, batch 512): dram__bytes_read + dram__bytes_write = 67.2 + 28.7 MB; algorithmic bytes
                # (x in + weights + y out, bf16) = 134.3 MB -- the output is still in the 126 MB L2 when the next kernel reads it
                "traffic": 95.9e6 if (not pgd and not sg and B == 512) else None,
                "traffic_note": "DRAM bytes per launch of 'k3 hw32 cin64 cout64' from ncu; algorithmic 134.3 MB",
                "launches": n_l, "share_of_step": tot_ms / ms if ms > 0 else None,
                "by_shape": [{"shape": k, "launches": a["launches"], "ms": round(a["ms"], 3),
                              "tflops": round(a["flops"] / (a["ms"] * 1e-3) / 1e12, 1) if a["ms"] > 0 else None,
                              "gbs": round(a["bytes"] / (a["ms"] * 1e-3) / 1e9, 1) if a["ms"] > 0 else None} for k, a in top]}
        # the two other kernel classes of the step, each against the roofline that bounds it
        extra = []
        if fused:
            f_ms = sum(a["ms"] for a in fused.values()); f_fl = sum(a["flops"] for a in fused.values())
            # the stage that bounds this kernel runs on the fp32 pipe: per hidden element 25 depthwise FMAs + ~5 FMA-pipe ops of the two
            # SiLUs; B200: 128 fp32 FMA lanes / clk / SM (scripts/ubench_pipes.cu: FFMA 1.0, FFMA2 0.5 warp-instructions / clk / SMSP)
            import re as _re
            simt_ops = 0.0
            for k, a in fused.items():
                m_ = _re.search(r"hw(\d+) c(\d+) hidden(\d+)", k)
                simt_ops += a["launches"] * B * int(m_.group(1)) ** 2 * int(m_.group(3)) * 30.0
            sm_hz = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
            fp32_peak = 148 * 128 * sm_hz
            extra.append({"kernel": "mbconv_fused_kernel (decoder cell: expand 1x1 -> SiLU -> depthwise 5x5 -> SiLU -> project 1x1, hidden tensor on chip)",
                          "bound": "tensor", "achieved": f_fl / (f_ms * 1e-3) / 1e12, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                          "frac": f_fl / (f_ms * 1e-3) / 1e12 / pk["tf_sustained"], "launches": sum(a["launches"] for a in fused.values()),
                          "share_of_step": f_ms / ms,
                          "fp32_pipe": {"achieved_tfma_s": simt_ops / (f_ms * 1e-3) / 1e12, "peak_tfma_s": fp32_peak / 1e12,
                                        "frac": simt_ops / (f_ms * 1e-3) / fp32_peak},
                          "note": "issue/FMA-pipe bound by the SIMT depthwise stage (ncu: fp32 FMA pipe 25%, XU 24%, issue slots 48%, tensor pipe 6%); "
                                  "HBM traffic = x + r only (ncu dram 88 MB per 32x32 launch vs 1.6 GB for the three-kernel path)",
                          "by_shape": [{"shape": k[6:], "launches": a["launches"], "us_per_launch": round(1e3 * a["ms"] / a["launches"], 1),
                                        "tflops": round(a["flops"] / (a["ms"] * 1e-3) / 1e12, 1)} for k, a in sorted(fused.items(), key=lambda kv: -kv[1]["ms"])]})
        if hbm:
            h_ms = sum(a["ms"] for a in hbm.values()); h_by = sum(a["bytes"] for a in hbm.values())
            extra.append({"kernel": "se_residual_kernel (SE gate + residual + next cell's activation copies, elementwise)", "bound": "hbm",
                          "achieved": h_by / (h_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": h_by / (h_ms * 1e-3) / 1e9 / pk["hbm_gbs"],
                          "launches": sum(a["launches"] for a in hbm.values()), "share_of_step": h_ms / ms,
                          "by_shape": [{"shape": k[4:], "launches": a["launches"], "us_per_launch": round(1e3 * a["ms"] / a["launches"], 1),
                                        "gbs": round(a["bytes"] / (a["ms"] * 1e-3) / 1e9, 1)} for k, a in sorted(hbm.items(), key=lambda kv: -kv[1]["ms"])[:4]]})
        cpu = cpu_baseline() if not args.no_cpu_baseline else None
        gflop = GFLOP_PER_IMAGE_FWD * ((2 * args.pgd_steps + 1) if pgd else 1)
        if sg:
            gflop = {"gender": 304.0, "cars": 172.0}[args.workload]      # encoder + decoder + classifier, MACs x 2
        workload = ("BASELINE configs[4]: PGD-Linf eps 8/255, step 2/255, %d steps (fwd + input-gradient each) + 1 eval forward, through "
                    "NVAE-C32 purifier (ours_cosine_noise_ids.yaml) + VGG11, random-init weights" % args.pgd_steps) if pgd else \
            ("BASELINE configs[1]: NVAE-C32 CelebA-64 ids purification (ours_learned_blur_ids.yaml: blur, eps 0, "
             "learned alphas x0.7) + VGG11 classifier, random-init weights")
        if args.workload == "gender":
            workload = ("BASELINE configs[2]: StyleGAN-E4E CelebA-HQ 256 gender (IR-SE50 + 18 map2style heads -> StyleGAN2 @1024 -> face_pool) "
                        "ours_linear_noise_gender.yaml (eps 4, linear alphas) + ResNet-50, random-init weights")
        elif args.workload == "cars":
            workload = ("BASELINE configs[3]: Style-Transformer Stanford Cars 128 (IR-SE50 + 3 transformer decoder layers -> StyleGAN2 @512) "
                        "ours_cosine_blur_cars.yaml (blur, cosine alphas x0.7) + ResNeXt-50, random-init weights")
        line = {"metric": "pgd_attacked_img_per_s" if pgd else "purified_img_per_s", "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": mode, "data": "synthetic",
                "config": {"workload": workload,
                           "nvae": None if sg else NVAE_C32_CONFIG, "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
                           "l2": "no explicit flush: each step streams > 10 GB of activations through the 126 MB L2",
                           "cuda_graph": bool(use_graph and not sg),
                           "gflop_per_image_algorithmic": gflop},
                "tflops_algorithmic": value * gflop / 1e3,
                "e2e": {"value": e2e_val, "unit": "img/s", "h2d_bytes_per_step": x_host.numel() * 4 * world,
                        "d2h_bytes_per_step": (B if pgd else logits_host.numel() * 4) * world},
                "gpu_launches": int(launches), "hbm_peak_gb": round(torch.cuda.max_memory_allocated(dev) / 1e9, 2), "clocks": clocks, "roofline": roof, "roofline_other_kernels": extra, "cpu_baseline": cpu,
                "counters": {"n_total": int(counters[0].item()), "n_clean_correct": int(counters[1].item()),
                             "n_robust_correct": int(counters[2].item())}}
        print(json.dumps(line), file=JSON_OUT or sys.stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


JSON_OUT = None          # the process's original stdout; file descriptor 1 itself is pointed at stderr (see main)


def main():
    # stdout carries exactly ONE JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its version banner
    # there even with NCCL_DEBUG_FILE set), so fd 1 is redirected to stderr for the whole run and the JSON goes to a dup of the original.
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default 512 purify / 1024 pgd: 95 GB of tape + activations)")
    ap.add_argument("--workload", default="purify", choices=["purify", "pgd", "gender", "cars"])
    ap.add_argument("--chunk", type=int, default=0, help="generator batch chunk of the StyleGAN workloads (0: automatic)")
    ap.add_argument("--pgd-steps", type=int, default=50)
    ap.add_argument("--cuda-graph", type=int, default=-1, help="1: replay the call / PGD iteration as a CUDA graph; 0: eager; default: pgd only "
                    "(the purify roofline needs per-launch events, which only exist in eager mode)")
    ap.add_argument("--ref-batch", type=int, default=16, help="bounded sample per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="CUDA-event time of EVERY op (written to stderr as a table)")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 4
        args.warmup = args.warmup if args.warmup is not None else 1
        run_reference(args)
    else:
        args.steps = args.steps if args.steps is not None else {"purify": 10, "pgd": 2}.get(args.workload, 3)
        args.warmup = args.warmup if args.warmup is not None else {"purify": 3, "pgd": 3}.get(args.workload, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
