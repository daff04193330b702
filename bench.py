"""Headline benchmark: frames/s through the BLIP ViT-L/16 @224 forward (BASELINE.json configs[1]: batch 256
synthetic 224x224 frames, bf16 operands, forward only), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's sm_100a path
    python bench.py --impl reference ...                           # the CPU port of the reference's vit.py (oracle/)
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # N ranks, each its own 256-frame batches (weak scaling)

One step = one 256-frame batch through vidil_vit_forward.  Prints ONE JSON line on rank 0 with
  value     frames/s, inputs resident in HBM, CUDA events around exactly K steps, max over ranks
  e2e       the same through the host-buffer API (VisionTransformer.encode_host_stream -> vidil_encoder_host_submit/_wait):
            pinned host frames -> H2D -> forward -> D2H of the [B,197,1024] fp32 tokens, every step, with step k+1's H2D
            and step k-1's D2H overlapping step k's forward
  roofline  the tcgen05 GEMM kernel: algorithmic FLOPs / its event-timed device time inside the timed steps
  cpu_baseline  the oracle port of models/vit.py timed on this box's host cores (rank 0, N=1 only)
  reference_ops_on_gpu  the same restatement run by PyTorch eager on the GPU (TF32 and bf16 autocast): the honest second baseline
Other workloads (not the driver's line): --workload clip | sim | text | tokenize | capfilt.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

VIT = {"large": (1024, 24, 16), "base": (768, 12, 12)}


def flops_per_frame(D, depth, tokens, patch, proj_dim=0):
    per_layer = 2 * tokens * D * 3 * D + 2 * tokens * D * D + 4 * tokens * D * 4 * D + 4 * tokens * tokens * D
    return float(depth * per_layer + 2 * (tokens - 1) * 3 * patch * patch * D + 2 * D * proj_dim)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(tflops=float(p.get("bf16_tflops", 1590.0)), tflops_sustained=float(p.get("bf16_tflops_sustained", 1400.0)),
                    hbm_gbs=float(p.get("hbm_gbs", 6650.0)), source="measured (MEASURED_PEAKS.json)")
    return dict(tflops=1590.0, tflops_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples of one GPU while the timed region runs."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                    pw.append(float(parts[2]))
                except ValueError:
                    continue
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w": round(statistics.median(pw), 1),
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus}")
        raise SystemExit(f"WORLD_SIZE={world} but --gpus {args.gpus}")
    return rank, world, local


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's vit.py on this box's host cores
# ---------------------------------------------------------------------------------------------------------------------
def cpu_vit_frames_per_s(vit, image_size, frames_per_step, steps, warmup, budget_s=None):
    import torch

    from oracle import vit_oracle, weights as W
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    D, depth, heads = VIT[vit]
    sd = W.vit_state_dict(vit, image_size, seed=0)
    x = W.frames(frames_per_step, image_size, seed=0)
    for _ in range(warmup):
        vit_oracle.vit_forward(sd, x, heads)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        vit_oracle.vit_forward(sd, x, heads)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done * frames_per_step / dt, dt / done * 1e3, cores, done


def run_reference(args):
    rank, world, _ = dist_env(args) if "RANK" in os.environ else (0, 1, 0)
    if rank != 0:
        return
    D, depth, heads = VIT[args.vit]
    tokens = (args.image_size // 16) ** 2 + 1
    fps, ms, cores, done = cpu_vit_frames_per_s(args.vit, args.image_size, args.ref_frames, args.steps, args.warmup)
    sample = (f"{args.ref_frames} frames per step x {done} steps of the {args.batch}-frame workload; oracle/vit_oracle.py "
              f"(restatement of models/vit.py:180-194, fp32, torch CPU kernels, {cores} threads)")
    line = {
        "impl": "reference", "metric": "frames/sec encoded", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, tokens),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, tokens):
    return {"workload": f"BLIP ViT-{args.vit[0].upper()}/16 @{args.image_size} forward only, batch {args.batch} synthetic "
                        f"{args.image_size}x{args.image_size}x3 frames per GPU per step (BASELINE.json configs[1])",
            "frames_per_step_per_gpu": args.batch, "tokens_per_frame": tokens, "image_size": args.image_size,
            "l2_policy": "inputs larger than L2: 154 MB of fp32 frames per step from 2 rotating buffers; every layer "
                         "streams 0.1-0.4 GB of activations",
            "parallelism": f"dp{args.gpus} (frames sharded, weights replicated, no collective in the step)"}


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def run_vit(args):
    import torch
    import torch.distributed as dist

    from vidil_b200 import _lib, distributed as vdist
    from vidil_b200.blip import create_vit

    rank, world, local = dist_env(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU path for the product (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        vdist.init_distributed_mode("nccl")

    B, K, Wm = args.batch, args.steps, args.warmup
    D, depth, heads = VIT[args.vit]
    tokens = (args.image_size // 16) ** 2 + 1
    gflop_frame = flops_per_frame(D, depth, tokens, 16) / 1e9
    peaks = measured_peaks()

    torch.manual_seed(1234 + rank)
    model, width = create_vit(args.vit, args.image_size, compute_dtype=args.dtype)
    with torch.no_grad():  # exercise bias / affine paths; values stay at init scale
        for name, p in model.named_parameters():
            if name.endswith(".bias"):
                p.normal_(0.0, 0.02)
            elif "norm" in name and name.endswith(".weight"):
                p.normal_(1.0, 0.02)
    model = model.to(dev).eval()
    frames = [torch.randn(B, 3, args.image_size, args.image_size, device=dev) for _ in range(2)]
    out = model(frames[0])  # packs weights, sizes the workspace
    assert tuple(out.shape) == (B, tokens, D) and bool(torch.isfinite(out).all())
    enc = model._ensure_packed()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ---------------------------------------------------------------------------------
    for i in range(Wm):
        model(frames[i & 1])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    enc.set_profiling(True)
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(K):
        model(frames[i & 1])
    ev1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    prof = enc.read_profile()
    enc.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None

    # same K steps without the per-kernel events, to show what the instrumentation costs
    barrier()
    ev0.record()
    for i in range(K):
        model(frames[i & 1])
    ev1.record()
    barrier()
    ms_total_plain = max_over_ranks(ev0.elapsed_time(ev1))
    ms_best = min(ms_total, ms_total_plain)
    fps = world * B * K / (ms_best / 1e3)

    # ---- end to end through the host-buffer API ------------------------------------------------------------------------
    # encode_host_stream: every step's frames start in pinned host memory and every step's [B,197,1024] fp32 tokens
    # end in pinned host memory; step k+1's H2D and step k-1's D2H overlap step k's forward (two slots).
    host_in = [torch.randn(B, 3, args.image_size, args.image_size).pin_memory() for _ in range(2)]
    host_outs = [torch.empty(B, tokens, D, dtype=torch.float32).pin_memory() for _ in range(2)]

    def stream_of(n):
        for i in range(n):
            yield host_in[i & 1]

    for o in model.encode_host_stream(stream_of(max(3, Wm // 2)), outs=host_outs):
        pass
    barrier()
    checksum = 0.0
    t0 = time.perf_counter()
    for o in model.encode_host_stream(stream_of(K), outs=host_outs):
        checksum += float(o[0, 0, 0])  # touch the delivered result on the host
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_fps = world * B * K / e2e_s
    host_out = host_outs[(K - 1) & 1]
    # the blocking single-call variant (no overlap), for reference
    t0 = time.perf_counter()
    for i in range(3):
        model.encode_host(host_in[i & 1], out=host_outs[0])
    torch.cuda.synchronize()
    e2e_blocking_fps = world * B * 3 / max_over_ranks(time.perf_counter() - t0)

    # ---- the same from DECODED frames: uint8 320x240 RGB in pinned memory -> H2D -> resize/normalise on the GPU -> ViT
    #      -> D2H of the tokens (vidil_b200.preprocess.encode_u8_stream); what a video pipeline actually feeds
    from vidil_b200 import preprocess as vpre
    u8_in = [torch.randint(0, 256, (B, 240, 320, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    for o in vpre.encode_u8_stream(model, (u8_in[i & 1] for i in range(3)), args.image_size, outs=host_outs):
        pass
    barrier()
    t0 = time.perf_counter()
    for o in vpre.encode_u8_stream(model, (u8_in[i & 1] for i in range(K)), args.image_size, outs=host_outs):
        checksum += float(o[0, 0, 0])
    torch.cuda.synchronize()
    e2e_u8_fps = world * B * K / max_over_ranks(time.perf_counter() - t0)

    # ---- the path's one collective: all-gather of per-rank result rows (JSON), rank-0 merge --------------------------
    t0 = time.perf_counter()
    rows = {f"rank{rank}_frame{i}": {"cls_l1": float(host_out[i, 0].abs().sum())} for i in range(0, B, 32)}
    merged = vdist.gather_and_write(rows, None, device=dev)
    gather_ms = (time.perf_counter() - t0) * 1e3

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    g = prof["gemm"]
    gemm_tflops = g["flops"] / (g["ms"] / 1e3) / 1e12 if g["ms"] > 0 else 0.0
    step_ms = ms_best / K
    classes = {k: {"ms_per_step": v["ms"] / K, "launches_per_step": v["launches"] / K,
                   "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["ms"] > 0 else 0.0,
                   "gbs": (v["bytes"] / (v["ms"] / 1e3) / 1e9) if v["ms"] > 0 else 0.0} for k, v in prof.items()}
    line = {
        "metric": "frames/sec encoded", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic", "config": workload_config(args, tokens),
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": B * 3 * args.image_size ** 2 * 4,
                "d2h_bytes_per_step": B * tokens * D * 4, "ms_per_step": e2e_s / K * 1e3,
                "api": "VisionTransformer.encode_host_stream -> vidil_encoder_host_submit/_wait (pinned host buffers, "
                       "copies overlapped with the neighbouring steps' forwards)",
                "blocking_call_value": e2e_blocking_fps, "checksum": checksum,
                "from_uint8_frames": {"value": e2e_u8_fps, "unit": "frames/s", "h2d_bytes_per_step": B * 240 * 320 * 3,
                                      "note": "decoded 320x240 uint8 frames in pinned memory; PIL-identical resize + "
                                              "normalise on the GPU (vidil_preprocess_frames) inside the timed region"}},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (patch-embed, qkv, proj, fc1, fc2: "
                                                  f"{g['launches'] // K} launches per step)",
                     "achieved": gemm_tflops, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": gemm_tflops / peaks["tflops_sustained"], "frac_of_burst_peak": gemm_tflops / peaks["tflops"],
                     "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the four per-layer GEMMs of one
                     # `ncu --set full` capture of this build (profiles/r01d_summary.md: fc1 471, fc2 832, qkv 366,
                     # proj 460 MB) — against 574 MB of algorithmic operand + output bytes per launch
                     "traffic": 532.3e6 if (args.vit == "large" and B == 256 and args.image_size == 224) else None,
                     "algorithmic_bytes_per_launch": g["bytes"] / max(g["launches"], 1),
                     "share_of_step": g["ms"] / ms_total},
        "whole_step": {"gflop_per_frame": gflop_frame, "tflops": fps / world * gflop_frame / 1e3,
                       "frac_of_sustained_peak": fps / world * gflop_frame / 1e3 / peaks["tflops_sustained"],
                       "frac_of_burst_peak": fps / world * gflop_frame / 1e3 / peaks["tflops"],
                       "ms_with_kernel_events": ms_total / K, "ms_without": ms_total_plain / K},
        "kernel_classes": classes,
        "clocks": clocks,
        "gather": {"ms": gather_ms, "rows": len(merged), "collective": "all_gather of length-prefixed JSON rows"},
    }
    if world == 1 and not args.no_cpu_baseline:
        fps_cpu, ms_cpu, cores, done = cpu_vit_frames_per_s(args.vit, args.image_size, 8, 6, 1, budget_s=20.0)
        line["cpu_baseline"] = {"value": fps_cpu, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": f"8-frame batches x {done} (1 warm-up) of the same workload through "
                                          f"oracle/vit_oracle.py (fp32 restatement of models/vit.py), {cores} torch threads"}
        # the reference's own op sequence (the fp32 restatement of models/vit.py: Conv2d, nn.Linear, materialised
        # softmax attention, nn.GELU, nn.LayerNorm) run by PyTorch eager on this same GPU — not a product path, a second baseline
        line["reference_ops_on_gpu"] = torch_eager_on_gpu(args.vit, args.image_size, args.batch, dev)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def torch_eager_on_gpu(vit, image_size, batch, dev, iters=5):
    """frames/s of oracle/vit_oracle.vit_forward on CUDA tensors: fp32 with TF32 matmuls (cudnn.benchmark / allow_tf32 as the
    reference sets them, run_video_CapFilt.py:234) and under bf16 autocast.  Bounded: 2 warm-up + `iters` timed batches each."""
    import torch

    from oracle import vit_oracle, weights as W
    _, _, heads = VIT[vit]
    sd = {k: v.to(dev) for k, v in W.vit_state_dict(vit, image_size, seed=0).items()}
    x = torch.randn(batch, 3, image_size, image_size, device=dev)
    out = {"what": "oracle/vit_oracle.py (same ATen ops as models/vit.py) on this GPU through PyTorch eager, batch %d" % batch}
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
    try:
        for key, ctx in (("fp32_tf32_frames_per_s", torch.autocast("cuda", enabled=False)),
                         ("bf16_autocast_frames_per_s", torch.autocast("cuda", dtype=torch.bfloat16))):
            with torch.no_grad(), ctx:
                for _ in range(2):
                    vit_oracle.vit_forward(sd, x, heads)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    vit_oracle.vit_forward(sd, x, heads)
                e1.record()
                torch.cuda.synchronize()
            out[key] = batch * iters / (e0.elapsed_time(e1) / 1e3)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    return out


def run_sim(args):
    """BASELINE configs[3] kernel in isolation: 2048 frames x 10k phrases x 768, top-5 (not the driver's line)."""
    import torch

    from vidil_b200 import _lib, ops
    dev = torch.device("cuda", 0)
    Fr, T, Dm, k = 2048, 10000, 768, 5
    img = torch.nn.functional.normalize(torch.randn(Fr, Dm, device=dev), dim=-1)
    bank = torch.nn.functional.normalize(torch.randn(T, Dm, device=dev), dim=-1)
    for _ in range(args.warmup):
        ops.sim_topk(img, bank, k)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        ops.sim_topk(img, bank, k)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    print(json.dumps({"metric": "frames/sec scored (sim + top-k)", "value": Fr / (ms / 1e3), "unit": "frames/s",
                      "ms_per_step": ms, "tflops": 2.0 * Fr * T * Dm / (ms / 1e3) / 1e12,
                      "gpu_launches": _lib.launch_count() - l0,
                      "config": {"workload": f"{Fr} frames x {T} phrases x {Dm}, top-{k}"}}), flush=True)


def run_clip(args):
    """CLIP ViT-L/14 image tower + projection throughput (not the driver's line)."""
    import torch

    from oracle import weights as W
    from vidil_b200.clip import CLIPVisionB200
    dev = torch.device("cuda", 0)
    c = W.CLIP_CONFIGS["large14"]
    m = CLIPVisionB200(**c, compute_dtype=args.dtype)
    with torch.no_grad():
        for p in m.parameters():
            p.normal_(0.0, 0.02)
        for n, p in m.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.add_(1.0)
    m = m.to(dev).eval()
    x = [torch.randn(args.batch, 3, 224, 224, device=dev) for _ in range(2)]
    for i in range(args.warmup):
        m(x[i & 1])
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        m(x[i & 1])
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    fps = args.batch / (ms / 1e3)
    gf = flops_per_frame(1024, 24, 257, 14, 768) / 1e9
    print(json.dumps({"metric": "frames/sec encoded (CLIP ViT-L/14)", "value": fps, "unit": "frames/s", "ms_per_step": ms,
                      "tflops": fps * gf / 1e3, "frac_of_sustained_peak": fps * gf / 1e3 / measured_peaks()["tflops_sustained"],
                      "config": {"workload": f"CLIP ViT-L/14 @224 image tower, batch {args.batch}"}}), flush=True)


def run_text(args):
    """CLIP text tower (phrase bank) throughput: 512-phrase batches of 77 tokens (not the driver's line)."""
    import torch

    from oracle import weights as W
    from vidil_b200.clip import CLIPTextB200
    dev = torch.device("cuda", 0)
    c = W.CLIP_TEXT_CONFIGS["large14"]
    m = CLIPTextB200(**c, compute_dtype=args.dtype)
    with torch.no_grad():
        for p in m.parameters():
            p.normal_(0.0, 0.02)
        for n, p in m.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.add_(1.0)
    m = m.to(dev).eval()
    ids = W.token_ids("large14", 512, 77, seed=0).to(dev)
    for _ in range(args.warmup):
        m(ids)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        m(ids)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.steps
    L, D, I = 77, 768, 3072
    gf = 12 * (2 * L * D * 3 * D + 2 * L * D * D + 4 * L * D * I + 4 * L * L * D) / 1e9
    print(json.dumps({"metric": "phrases/sec embedded (CLIP text tower)", "value": 512 / (ms / 1e3), "unit": "phrases/s",
                      "ms_per_step": ms, "tflops": 512 / (ms / 1e3) * gf / 1e3,
                      "config": {"workload": "CLIP ViT-L/14 text tower, 512 phrases x 77 tokens per step"}}), flush=True)


def run_tokenize(args):
    """BASELINE configs[3]/[4] shape, tokenization half: synthetic videos x 8 frames through the whole CLIP branch of
    run_visual_tokenization.py — phrase bank by the native text tower, frames through the native image tower from pinned
    host buffers, similarity + top-k per bank on the device, aggregation on the host, one all-gather of the JSON rows,
    rank 0 writes visual_tokens.json.  Videos are sharded over the ranks with the reference's slice formula."""
    import tempfile as _tf

    import torch
    import torch.distributed as dist

    from oracle import weights as W
    from vidil_b200 import distributed as vdist, visual_tokenization as vt
    from vidil_b200.clip import CLIPTextB200, CLIPVisionB200
    rank, world, local = dist_env(args)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        vdist.init_distributed_mode("nccl")
    torch.manual_seed(7)
    vision = CLIPVisionB200(**W.CLIP_CONFIGS["large14"], compute_dtype=args.dtype)
    text = CLIPTextB200(**W.CLIP_TEXT_CONFIGS["large14"], compute_dtype=args.dtype)
    with torch.no_grad():
        for m in (vision, text):
            for n, p in m.named_parameters():
                p.normal_(0.0, 0.02)
                if "norm" in n and n.endswith("weight"):
                    p.add_(1.0)
    vision, text = vision.to(dev).eval(), text.to(dev).eval()
    num_frm, k = 8, 5
    bank_sizes = {"objects": 19965, "attributes": 16693, "scenes": 365, "verbs": 7414}  # the reference's `vg` ontology
    phrases = {key: [f"{key} {i}" for i in range(n)] for key, n in bank_sizes.items()}
    videos = [f"video{i}" for i in range(args.videos)]
    start, end = vdist.shard_bounds(len(videos), world, rank)
    mine = videos[start:end]
    fb = args.batch - args.batch % num_frm                      # frames per tower call: whole videos
    host = [torch.randn(fb, 3, 224, 224).pin_memory() for _ in range(2)]
    n_calls = (len(mine) * num_frm + fb - 1) // fb

    def run_once():
        t0 = time.perf_counter()
        reps = {}
        for key, n in bank_sizes.items():                        # phrase bank, 512 phrases per call like the reference
            embs = [text(W.token_ids("large14", min(512, n - i), 77, seed=i).to(dev)) for i in range(0, n, 512)]
            reps[key] = {"text_embeds": torch.cat(embs)}
        torch.cuda.synchronize()
        t_bank = time.perf_counter() - t0
        embeds = []
        for e in vision.encode_host_stream(host[i & 1] for i in range(n_calls)):
            embeds.append(e.to(dev, non_blocking=True))
        image_embeds = torch.cat(embeds)[:len(mine) * num_frm]
        torch.cuda.synchronize()
        t_frames = time.perf_counter() - t0 - t_bank
        rows = vt.tokens_from_embeddings(image_embeds, reps, phrases, mine, [[""]] * len(mine), num_frm, k)
        t_tok = time.perf_counter() - t0 - t_bank - t_frames
        out_dir = _tf.mkdtemp() if rank == 0 else None
        merged = vdist.gather_and_write(rows, os.path.join(out_dir, "visual_tokens.json") if rank == 0 else None, device=dev)
        t_all = time.perf_counter() - t0
        return t_all, t_bank, t_frames, t_tok, (len(merged) if merged is not None else 0)

    run_once() if args.warmup else None
    if world > 1:
        dist.barrier()
    t_all, t_bank, t_frames, t_tok, n_rows = run_once()
    if world > 1:
        t = torch.tensor([t_all], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_all = float(t.item())
    if rank == 0:
        print(json.dumps({"metric": "frames/sec tokenized (CLIP towers + sim + top-k + JSON gather)",
                          "value": args.videos * num_frm / t_all, "unit": "frames/s", "n_gpus": world, "seconds": t_all,
                          "rank0_seconds": {"phrase_bank": t_bank, "frames": t_frames, "sim_topk_aggregate": t_tok},
                          "rows_merged": n_rows,
                          "config": {"workload": f"{args.videos} synthetic videos x {num_frm} frames, vg-sized phrase banks "
                                                 f"{bank_sizes}, top-{k}, tower batch {fb} frames from pinned host memory"}}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


def _randomise(module, std=0.02):
    import torch
    with torch.no_grad():
        for n, p in module.named_parameters():
            p.normal_(0.0, std)
            if ("norm" in n or "LayerNorm" in n) and n.endswith("weight"):
                p.add_(1.0)


def run_capfilt(args):
    """BASELINE.json configs[2]: `--videos` synthetic videos x 8 frames through the CapFilt models of run_video_CapFilt.py —
    captioner = BLIP ViT + med.py decoder with beam search (beams 3, max_length 20, min_length 5, :102), filterer = BLIP_ITM
    (its own ViT + the multimodal text encoder + itm_head) over every (caption, frame) pair of a video (:108-120).  One step =
    all videos once; frames start on the device.  Under torchrun the videos are sharded over the ranks with the reference's
    slice formula (run_video_CapFilt.py:239-241) and the kept captions are merged by one all-gather of JSON rows (BASELINE.json
    configs[4], CapFilt half).  Not the driver's line."""
    import torch
    import torch.distributed as dist

    from vidil_b200 import distributed as vdist
    from vidil_b200.blip import BLIP_Decoder, BLIP_ITM
    rank, world, local = dist_env(args)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        vdist.init_distributed_mode("nccl")
    torch.manual_seed(0)
    Fv = 8
    v_start, v_end = vdist.shard_bounds(args.videos, world, rank)
    V = v_end - v_start
    n_frames = V * Fv
    cap = BLIP_Decoder(image_size=args.image_size, vit=args.vit, compute_dtype=args.dtype)
    _randomise(cap)
    with torch.no_grad():
        cap.text_decoder.cls.predictions.decoder.weight.normal_(0.0, 0.1)     # spread-out logits, like a trained head
    cap = cap.to(dev).eval()
    itm = BLIP_ITM(image_size=args.image_size, vit=args.vit, compute_dtype=args.dtype, cache_identical_inputs=False)
    _randomise(itm)
    itm = itm.to(dev).eval()
    frames = torch.randn(n_frames, 3, args.image_size, args.image_size, device=dev)
    chunk = args.batch
    T_itm = 35
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    lib = __import__("vidil_b200._lib", fromlist=["_lib"])

    def step(timed):
        # captioner: ViT per `chunk` frames, then one beam search over all frames
        if timed:
            ev[0].record()
        toks = torch.cat([cap.visual_encoder(frames[i:i + chunk]) for i in range(0, n_frames, chunk)])
        if timed:
            ev[1].record()
        ids = torch.tensor([cap._prompt_ids], dtype=torch.long).repeat(n_frames, 1)
        ids[:, 0] = cap.bos_token_id
        out, scores, lens = cap.text_decoder.generate(input_ids=ids[:, :-1], max_length=20, min_length=5, num_beams=3,
                                                      eos_token_id=cap.sep_token_id, pad_token_id=cap.pad_token_id,
                                                      encoder_hidden_states=toks, return_scores=True)
        if timed:
            ev[2].record()
        # filterer: every generated caption of a video against each of its 8 frames (duplicates are not removed here: worst case)
        cap_ids = torch.zeros(n_frames, T_itm, dtype=torch.int32, device=dev)
        L = min(out.shape[1], T_itm)
        cap_ids[:, :L] = out[:, :L].int()
        cap_ids[:, 0] = 101                                     # [CLS], as the tokenizer call of blip_itm.py:46 produces
        mask = (torch.arange(T_itm, device=dev)[None] < lens[:, None].clamp(max=T_itm)).int()
        itoks = torch.cat([itm.visual_encoder(frames[i:i + chunk]) for i in range(0, n_frames, chunk)])
        if timed:
            ev[3].record()
        # frame-major pairs: frame (v, j) against the Fv captions of video v -> one cross-attention query group per frame
        pair_ids = cap_ids.view(V, 1, Fv, T_itm).expand(V, Fv, Fv, T_itm).reshape(-1, T_itm)
        pair_mask = mask.view(V, 1, Fv, T_itm).expand(V, Fv, Fv, T_itm).reshape(-1, T_itm)
        logits = []
        pc = max(Fv * Fv, args.pair_chunk - args.pair_chunk % (Fv * Fv))             # whole videos per call
        for i in range(0, pair_ids.shape[0], pc):
            j = min(i + pc, pair_ids.shape[0])
            _, _, cls = itm.text_encoder.run(pair_ids[i:j], pair_mask[i:j], itoks[i // Fv:j // Fv], want_hidden=False, want_cls=True,
                                             seqs_per_frame=Fv)
            logits.append(cls)
        prob = torch.softmax(torch.cat(logits), dim=1)[:, 1].view(V, Fv, Fv).max(dim=1).values.reshape(-1)   # max over frames, :118
        keep = (prob > 0.5).cpu()
        if timed:
            ev[4].record()
        return out.cpu(), keep

    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    before = lib.launch_count()
    acc = [0.0] * 4
    for _ in range(args.steps):
        out, keep = step(True)
        torch.cuda.synchronize()
        for i in range(4):
            acc[i] += ev[i].elapsed_time(ev[i + 1])
    launches = lib.launch_count() - before
    ms = [a / args.steps for a in acc]
    total = sum(ms)
    # per-class device time of ONE more beam search with an event pair around every kernel (not part of the timed steps): the
    # decode-step cross-attention is the dominant kernel and a pure K/V stream, so its roofline is HBM
    native = cap.text_decoder.bert._ensure_packed()
    native.set_profiling(True)
    toks = torch.cat([cap.visual_encoder(frames[i:i + chunk]) for i in range(0, n_frames, chunk)])
    ids = torch.tensor([cap._prompt_ids], dtype=torch.long).repeat(n_frames, 1)
    ids[:, 0] = cap.bos_token_id
    cap.text_decoder.generate(input_ids=ids[:, :-1], max_length=20, min_length=5, num_beams=3, eos_token_id=cap.sep_token_id,
                              pad_token_id=cap.pad_token_id, encoder_hidden_states=toks)
    prof = native.read_profile()
    native.set_profiling(False)
    del toks
    peaks = measured_peaks()
    xa = prof["attention"]
    roofline = {"bound": "hbm", "kernel": "cross_decode_mma_kernel + the prompt's attention_x_kernel (cross-attention onto the image "
                "tokens, one launch per layer and step)", "achieved": xa["bytes"] / (xa["ms"] / 1e3) / 1e9 if xa["ms"] else None,
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": xa["bytes"] / (xa["ms"] / 1e3) / 1e9 / peaks["hbm_gbs"] if xa["ms"] else None,
                "traffic": 624.4e6 if (args.vit == "large" and args.image_size == 224 and n_frames == 1024) else None,
                "algorithmic_bytes_per_launch": xa["bytes"] / max(xa["launches"], 1), "launches": xa["launches"],
                "peak_source": peaks["source"], "share_of_beam_search": xa["ms"] / max(sum(v["ms"] for v in prof.values()), 1e-9)}
    classes = {k: {"ms": v["ms"], "launches": v["launches"], "tflops": v["flops"] / (v["ms"] / 1e3) / 1e12 if v["ms"] else 0.0,
                   "gbs": v["bytes"] / (v["ms"] / 1e3) / 1e9 if v["ms"] else 0.0} for k, v in prof.items()}
    # the rows a rank contributes: video -> kept captions (token ids here: there is no vocabulary to decode with), then the one
    # collective of the path
    t0 = time.perf_counter()
    rows = {f"video{v_start + v}": [out[v * Fv + i].tolist() for i in range(Fv) if bool(keep[v * Fv + i])] for v in range(V)}
    merged = vdist.gather_and_write(rows, None, device=dev)
    gather_ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([total, gather_ms] + ms, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total, gather_ms, ms = float(t[0]), float(t[1]), [float(x) for x in t[2:]]
        cnt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(cnt)
        launches = int(cnt.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    n_frames_all = args.videos * Fv
    cpu = None
    if not args.no_cpu_baseline:
        # the reference's text side on the host cores: oracle port of med.py + the restated beam search, 2 frames of image tokens
        from oracle import med_oracle, weights as W
        name = "base_l" if args.vit == "large" else "base_b"
        torch.set_num_threads(os.cpu_count() or 1)
        sd = W.med_state_dict(name, "decoder", seed=0)
        n_tok = (args.image_size // 16) ** 2 + 1
        enc = W.image_tokens(2, n_tok, W.MED_CONFIGS[name]["encoder_width"], seed=0)
        t0 = time.perf_counter()
        med_oracle.generate(sd, enc, W.MED_SPECIAL[name]["prompt"], 12, 12, num_beams=3, max_length=20, min_length=5)
        dt = time.perf_counter() - t0
        cpu = {"value": 2 / dt, "unit": "captions/s (beam search only, image tokens given)", "cores": os.cpu_count(), "kind": "port",
               "sample": "oracle/med_oracle.generate on 2 frames of image tokens (cached decoder, beams 3, max_length 20)"}
    tokens = (args.image_size // 16) ** 2 + 1
    D, depth, _ = VIT[args.vit]
    print(json.dumps({"metric": "frames/sec through CapFilt (caption + filter)", "value": n_frames_all / (total / 1e3), "unit": "frames/s",
                      "n_gpus": world, "ms_per_step": total, "steps": args.steps, "warmup": args.warmup, "gpu_launches": launches,
                      "scaling": "strong", "gather": {"ms": gather_ms, "rows": len(merged), "collective": "all_gather of length-prefixed JSON rows"},
                      "stages_ms": {"captioner_vit": ms[0], "caption_beam_search": ms[1], "filterer_vit": ms[2], "itm_pairs": ms[3]},
                      "caption_frames_per_s": n_frames_all / ((ms[0] + ms[1]) / 1e3),
                      "captions_per_s_beam_search_only": n_frames_all / (ms[1] / 1e3), "cpu_baseline": cpu,
                      "roofline": roofline, "beam_search_kernel_classes": classes,
                      "decode_rows_per_rank": n_frames * 3, "itm_pairs_per_rank": n_frames * Fv, "dtype": args.dtype, "data": "synthetic",
                      "config": {"workload": f"{args.videos} synthetic videos x 8 frames @{args.image_size}, BLIP ViT-{args.vit[0].upper()}/16 + "
                                 f"med.py decoder (beam 3, max_length 20, min_length 5) + BLIP_ITM filter over {n_frames * Fv} "
                                 f"(caption, frame) pairs x 35 tokens per rank; {tokens} image tokens per frame; videos sharded over "
                                 f"{world} rank(s)"}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="vit", choices=["vit", "clip", "sim", "text", "tokenize", "capfilt"])
    ap.add_argument("--pair-chunk", type=int, default=2048, help="--workload capfilt: (caption, frame) pairs per ITM call")
    ap.add_argument("--videos", type=int, default=1024, help="--workload tokenize | capfilt: synthetic videos (8 frames each), all ranks")
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--vit", default="large", choices=list(VIT))
    ap.add_argument("--image-size", type=int, default=224)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--ref-frames", type=int, default=4, help="--impl reference: frames per CPU step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "sim":
        return run_sim(args)
    if args.workload == "clip":
        return run_clip(args)
    if args.workload == "text":
        return run_text(args)
    if args.workload == "tokenize":
        return run_tokenize(args)
    if args.workload == "capfilt":
        return run_capfilt(args)
    return run_vit(args)


if __name__ == "__main__":
    main()
