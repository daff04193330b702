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
  cpu_baseline  the reference's own models/vit.py (staged into oracle/_ref by oracle/stage_ref.py; the oracle port when no
            staged copy exists) timed on this box's host cores (rank 0, N=1 only)
  reference_ops_on_gpu  the same op sequence run by PyTorch eager on the GPU (TF32 and bf16 autocast): the honest second baseline
  workloads the other BASELINE.json configs, driver-timed in the same run: `clip` (CLIP ViT-L/14 tower), `sim` (configs[3]:
            2048 frames x 10 000 phrases, frames sharded over the ranks, top-5 checked against the fp32 ranking, plus the
            end-to-end index-flip report through the 16-bit towers), `capfilt` (configs[2]: 128 videos x 8 frames through
            ViT + beam search + ITM), `tokenize` (towers + text bank + sim + aggregation + JSON gather)
  pipeline  configs[4]: 50 000 synthetic frames = 6 250 videos through BOTH drivers (CapFilt on 4 frames per video,
            visual tokenization on 8), sharded with the reference's slice formula, one all-gather of JSON rows per
            driver, rank 0 writes the two result files
Single workloads on their own: --workload clip | sim | text | tokenize | capfilt | pipeline.
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


def measured_traffic(kernel: str, shape_key: str) -> dict:
    """{"traffic": bytes per launch or None, "traffic_source": ...} from profiles/traffic.json — the DRAM bytes
    (dram__bytes_read.sum + dram__bytes_write.sum) of one `ncu --set full` capture, keyed by kernel and workload shape."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            rec = json.load(f)[kernel][shape_key]
        return {"traffic": float(rec["bytes_per_launch"]), "traffic_source": rec.get("source", "profiles/traffic.json")}
    except (OSError, KeyError, ValueError, TypeError):
        return {"traffic": None, "traffic_source": "no committed ncu capture for this kernel at this shape"}


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
    """frames/s of the reference's ViT forward on this box's host cores with all the threads torch can use.  Times the
    UNMODIFIED models/vit.py when a reference tree is reachable (/root/reference in the build container, the copy staged
    under oracle/_ref by oracle/stage_ref.py on the GPU box) -> kind "reference"; the oracle's restatement otherwise ->
    kind "port".  Returns (frames_per_s, ms_per_step, cores, steps_done, kind, what)."""
    import torch

    from oracle import reference_shims, vit_oracle, weights as W
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    D, depth, heads = VIT[vit]
    sd = W.vit_state_dict(vit, image_size, seed=0)
    x = W.frames(frames_per_step, image_size, seed=0)
    if reference_shims.reference_available():
        model = reference_shims.build_reference_vit(vit, image_size, sd)
        reference_shims.uninstall_shims()
        kind = "reference"
        what = (f"the unmodified models/vit.py ({reference_shims.REFERENCE_ROOT}; timm/fairscale import shims only), fp32, "
                f"torch CPU kernels, {cores} threads")

        def fwd():
            with torch.no_grad():
                return model(x)
    else:
        kind = "port"
        what = f"oracle/vit_oracle.py (restatement of models/vit.py:180-194; no staged reference found), fp32, {cores} threads"

        def fwd():
            return vit_oracle.vit_forward(sd, x, heads)
    for _ in range(warmup):
        fwd()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        fwd()
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done * frames_per_step / dt, dt / done * 1e3, cores, done, kind, what


def run_reference(args):
    rank, world, _ = dist_env(args) if "RANK" in os.environ else (0, 1, 0)
    if rank != 0:
        return
    tokens = (args.image_size // 16) ** 2 + 1
    fps, ms, cores, done, kind, what = cpu_vit_frames_per_s(args.vit, args.image_size, args.ref_frames, args.steps, args.warmup)
    sample = (f"{args.ref_frames} frames per step x {done} steps: a bounded sample of the {args.batch}-frame-per-step workload "
              f"(the CPU forward is linear in the batch); {what}")
    cfg = workload_config(args, tokens)
    cfg["frames_per_step_per_gpu"] = args.ref_frames   # what this arm really runs per step
    cfg["native_arm_frames_per_step_per_gpu"] = args.batch
    line = {
        "impl": "reference", "metric": "frames/sec encoded", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
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

    # THE timed region of `value`: exactly K steps, no per-kernel events (the pass above exists for the roofline's per-kernel
    # times and is reported next to it as `ms_with_kernel_events`; nothing is min()-ed)
    barrier()
    sampler2 = ClockSampler(local)
    if rank == 0:
        sampler2.start()
    ev0.record()
    for i in range(K):
        model(frames[i & 1])
    ev1.record()
    barrier()
    ms_total_plain = max_over_ranks(ev0.elapsed_time(ev1))
    clocks_plain = sampler2.stop() if rank == 0 else None
    fps = world * B * K / (ms_total_plain / 1e3)

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

    # ---- the leanest host interface: uint8 frames up, 16-bit tokens down (vidil_vit_forward16): 59 + 103 MB per step instead
    #      of 154 + 207 MB — what matters when eight ranks share one host's memory system
    host_outs16 = [torch.empty(B, tokens, D, dtype=model.token_dtype16).pin_memory() for _ in range(2)]
    for o in vpre.encode_u8_stream(model, (u8_in[i & 1] for i in range(3)), args.image_size, outs=host_outs16, half_tokens=True):
        pass
    barrier()
    t0 = time.perf_counter()
    for o in vpre.encode_u8_stream(model, (u8_in[i & 1] for i in range(K)), args.image_size, outs=host_outs16, half_tokens=True):
        checksum += float(o[0, 0, 0])
    torch.cuda.synchronize()
    e2e_u8_16_fps = world * B * K / max_over_ranks(time.perf_counter() - t0)
    del host_outs16

    # ---- the path's one collective: all-gather of per-rank result rows (JSON), rank-0 merge --------------------------
    t0 = time.perf_counter()
    rows = {f"rank{rank}_frame{i}": {"cls_l1": float(host_out[i, 0].abs().sum())} for i in range(0, B, 32)}
    merged = vdist.gather_and_write(rows, None, device=dev)
    gather_ms = (time.perf_counter() - t0) * 1e3

    g = prof["gemm"]
    gemm_tflops = g["flops"] / (g["ms"] / 1e3) / 1e12 if g["ms"] > 0 else 0.0
    step_ms = ms_total_plain / K
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
                                              "normalise on the GPU (vidil_preprocess_frames) inside the timed region"},
                "from_uint8_frames_16bit_tokens": {"value": e2e_u8_16_fps, "unit": "frames/s", "h2d_bytes_per_step": B * 240 * 320 * 3,
                                                   "d2h_bytes_per_step": B * tokens * D * 2,
                                                   "note": "same, tokens delivered in the 16-bit operand type (vidil_vit_forward16)"}},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (patch-embed, qkv, proj, fc1, fc2: "
                                                  f"{g['launches'] // K} launches per step)",
                     "achieved": gemm_tflops, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                     "frac": gemm_tflops / peaks["tflops_sustained"], "frac_of_burst_peak": gemm_tflops / peaks["tflops"],
                     "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch from the latest committed `ncu --set full` capture of
                     # this kernel at this shape (profiles/traffic.json, written by tools/summarize_profiles.py); null otherwise
                     **measured_traffic("gemm_tcgen05_kernel", f"vit_{args.vit}_{args.image_size}_b{B}"),
                     "algorithmic_bytes_per_launch": g["bytes"] / max(g["launches"], 1),
                     "share_of_step": g["ms"] / ms_total},
        "whole_step": {"gflop_per_frame": gflop_frame, "tflops": fps / world * gflop_frame / 1e3,
                       "frac_of_sustained_peak": fps / world * gflop_frame / 1e3 / peaks["tflops_sustained"],
                       "frac_of_burst_peak": fps / world * gflop_frame / 1e3 / peaks["tflops"],
                       "ms_with_kernel_events": ms_total / K, "ms_without": ms_total_plain / K},
        "kernel_classes": classes,
        "clocks": clocks_plain, "clocks_during_kernel_event_pass": clocks,
        "gather": {"ms": gather_ms, "rows": len(merged) if merged is not None else 0, "collective": "all_gather of length-prefixed JSON rows"},
    }
    if world == 1 and not args.no_cpu_baseline:
        fps_cpu, ms_cpu, cores, done, kind, what = cpu_vit_frames_per_s(args.vit, args.image_size, 8, 6, 1, budget_s=20.0)
        line["cpu_baseline"] = {"value": fps_cpu, "unit": "frames/s", "cores": cores, "kind": kind,
                                "sample": f"8-frame batches x {done} (1 warm-up) of the same workload through {what}"}
        # the reference's own op sequence (the fp32 restatement of models/vit.py: Conv2d, nn.Linear, materialised
        # softmax attention, nn.GELU, nn.LayerNorm) run by PyTorch eager on this same GPU — not a product path, a second baseline
        line["reference_ops_on_gpu"] = torch_eager_on_gpu(args.vit, args.image_size, args.batch, dev)
    # the other BASELINE configs, in the same driver-timed process (every rank takes part: the records shard their work)
    del model, frames, host_in, host_outs, u8_in, out
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if not args.no_workloads:
        other_workloads(args, Ctx(rank, world, local, dev), line)
    if rank == 0:
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


class Ctx:
    """rank / world / device of this process, plus the two reductions every record needs."""

    def __init__(self, rank, world, local, dev):
        self.rank, self.world, self.local, self.dev = rank, world, local, dev

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max(self, x: float) -> float:
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x: float) -> float:
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t)
        return float(t.item())


def make_ctx(args) -> Ctx:
    import torch

    from vidil_b200 import distributed as vdist
    rank, world, local = dist_env(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200; there is no CPU path for the product (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        vdist.init_distributed_mode("nccl")
    return Ctx(rank, world, local, torch.device("cuda", local))


def _randomise(module, std=0.02):
    import torch
    with torch.no_grad():
        for n, p in module.named_parameters():
            p.normal_(0.0, std)
            if ("norm" in n or "LayerNorm" in n) and n.endswith("weight"):
                p.add_(1.0)


def build_clip_models(args, ctx, want_text=True, dtype=None):
    """CLIP ViT-L/14 image tower (+ text tower) of openai/clip-vit-large-patch14's architecture with seeded random weights."""
    import torch

    from oracle import weights as W
    from vidil_b200.clip import CLIPTextB200, CLIPVisionB200
    torch.manual_seed(7)
    vision = CLIPVisionB200(**W.CLIP_CONFIGS["large14"], compute_dtype=dtype or args.clip_dtype)
    _randomise(vision)
    vision = vision.to(ctx.dev).eval()
    text = None
    if want_text:
        text = CLIPTextB200(**W.CLIP_TEXT_CONFIGS["large14"], compute_dtype=dtype or args.clip_dtype)
        _randomise(text)
        text = text.to(ctx.dev).eval()
    return vision, text


def clip_record(args, ctx, vision, steps, warmup):
    """CLIP ViT-L/14 image tower + projection: frames/s per rank batch (weak scaling like the headline)."""
    import torch
    B = args.batch
    x = [torch.randn(B, 3, 224, 224, device=ctx.dev) for _ in range(2)]
    for i in range(warmup):
        vision(x[i & 1])
    ctx.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(steps):
        vision(x[i & 1])
    ev1.record()
    ctx.barrier()
    ms = ctx.max(ev0.elapsed_time(ev1)) / steps
    fps = ctx.world * B / (ms / 1e3)
    gf = flops_per_frame(1024, 24, 257, 14, 768) / 1e9
    peaks = measured_peaks()
    return {"metric": "frames/sec encoded (CLIP ViT-L/14 image tower + projection)", "value": fps, "unit": "frames/s",
            "n_gpus": ctx.world, "steps": steps, "warmup": warmup, "ms_per_step": ms, "gflop_per_frame": gf,
            "tflops_per_gpu": fps / ctx.world * gf / 1e3,
            "frac_of_sustained_peak": fps / ctx.world * gf / 1e3 / peaks["tflops_sustained"],
            "frac_of_burst_peak": fps / ctx.world * gf / 1e3 / peaks["tflops"],
            "dtype": vision.compute_dtype,
            "config": {"workload": f"CLIP ViT-L/14 @224 image tower, batch {B} frames per GPU per step, device-resident"}}


def sim_record(args, ctx, steps, warmup, flip_frames=0):
    """BASELINE configs[3]: 2048 frames x 10 000 phrases x 768, top-5; frames sharded over the ranks with the reference's
    slice formula, bank replicated.  Indices are checked against np.argsort(fp32 scores)[::-1][:5] (the reference's
    run_visual_tokenization.py:276,306) on up to 256 of rank 0's frames.  flip_frames > 0 adds the END-TO-END report
    BASELINE.md §5 promises: that many seeded frames through the native CLIP L/14 tower in bf16 and in fp16, ranked
    against the same bank, compared with the ranking from the fp32 tower (the oracle's restatement of transformers' CLIP
    vision path run in true fp32 on this GPU, TF32 off): flipped indices and the score gap of every flip."""
    import numpy as np
    import torch

    from oracle import tokenization_oracle
    from vidil_b200 import _lib, distributed as vdist, ops
    Fr, T, Dm, k = 2048, 10000, 768, 5
    g = torch.Generator().manual_seed(0)
    img_all = torch.nn.functional.normalize(torch.randn(Fr, Dm, generator=g), dim=-1)
    bank = torch.nn.functional.normalize(torch.randn(T, Dm, generator=torch.Generator().manual_seed(1)), dim=-1).to(ctx.dev)
    a, b = vdist.shard_bounds(Fr, ctx.world, ctx.rank)
    img = img_all[a:b].to(ctx.dev)
    for _ in range(warmup):
        ops.sim_topk(img, bank, k)
    ctx.barrier()
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        sc, idx = ops.sim_topk(img, bank, k)
    ev1.record()
    ctx.barrier()
    launches = _lib.launch_count() - l0
    ms = ctx.max(ev0.elapsed_time(ev1)) / steps
    # the same call replayed from a CUDA graph: device time without the Python / launch gaps between the three kernels
    graph_us = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ops.sim_topk(img, bank, k)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                ops.sim_topk(img, bank, k)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr.replay()
        ev0.record()
        for _ in range(steps):
            gr.replay()
        ev1.record()
        torch.cuda.synchronize()
        graph_us = ctx.max(ev0.elapsed_time(ev1)) / steps * 1e3
    except Exception as e:  # noqa: BLE001 - the eager number above stands on its own
        graph_us = f"graph capture failed: {type(e).__name__}: {e}"
    rec = {"metric": "frames/sec scored (image x phrase-bank similarity + top-5)", "value": Fr / (ms / 1e3), "unit": "frames/s",
           "n_gpus": ctx.world, "us_per_call": ms * 1e3, "us_per_call_graph_replay": graph_us, "steps": steps, "warmup": warmup,
           "tflops": 2.0 * Fr * T * Dm / (ms / 1e3) / 1e12, "gpu_launches_per_call": launches / max(steps, 1),
           "algorithmic_bytes": (b - a) * Dm * 4 + T * Dm * 2 + (b - a) * k * 8,
           "config": {"workload": f"{Fr} frames x {T} phrases x {Dm}, top-{k}; frames [{a},{b}) on rank {ctx.rank} of {ctx.world} "
                                  "(run_visual_tokenization.py:429-431 slice), bank replicated"}}
    if ctx.rank == 0:
        n = min(256, b - a)
        _, want = tokenization_oracle.sim_topk(img[:n].cpu().numpy(), bank.cpu().numpy(), k)
        got = idx[:n].cpu().numpy().astype(np.int64)
        rec["top5_indices_identical_to_fp32_argsort"] = bool(np.array_equal(got, want))
        rec["frames_checked"] = n
    if flip_frames > 0 and ctx.rank == 0:
        rec["end_to_end_flips"] = clip_flip_report(args, ctx, bank, flip_frames, k)
    return rec


def clip_flip_report(args, ctx, bank, n_frames, k):
    """See sim_record: top-k indices through the 16-bit native towers against the fp32 tower, same seeded frames and bank."""
    import numpy as np
    import torch

    from oracle import clip_oracle, weights as W
    from vidil_b200 import ops
    from vidil_b200.clip import CLIPVisionB200
    c = W.CLIP_CONFIGS["large14"]
    sd = W.clip_vision_state_dict("large14", seed=0)
    frames = W.frames(n_frames, 224, seed=5)
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        sd_dev = {kk: v.to(ctx.dev) for kk, v in sd.items()}
        ref = []
        with torch.no_grad():
            for i in range(0, n_frames, 64):
                ref.append(clip_oracle.clip_vision_forward(sd_dev, frames[i:i + 64].to(ctx.dev), c["num_attention_heads"])[0])
        ref = torch.cat(ref)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    ref_scores = (ref.double() @ bank.double().t())
    ref_top = torch.topk(ref_scores, k + 1, dim=1)
    want = ref_top.indices[:, :k].cpu().numpy()
    out = {"frames": n_frames, "bank": int(bank.shape[0]), "k": k,
           "reference": "oracle/clip_oracle.py (transformers' CLIP vision path restated) in fp32 on this GPU, TF32 off"}
    for dt in ("bf16", "fp16"):
        m = CLIPVisionB200(**c, compute_dtype=dt)
        m.load_state_dict(sd)
        m = m.to(ctx.dev).eval()
        emb = torch.cat([m(frames[i:i + 256].to(ctx.dev)) for i in range(0, n_frames, 256)])
        _, idx = ops.sim_topk(emb, bank, k)
        got = idx.cpu().numpy().astype(np.int64)
        diff = got != want
        # score gap of a flip: how far apart (in the fp32 ranking) the two swapped phrases are
        rows, cols = np.nonzero(diff)
        gaps = [abs(float(ref_scores[r, want[r, cc]] - ref_scores[r, got[r, cc]])) for r, cc in zip(rows, cols)]
        set_diff = sum(len(set(got[r]) ^ set(want[r])) > 0 for r in range(n_frames))
        out[dt] = {"flipped_positions": int(diff.sum()), "of": int(diff.size), "frames_with_any_flip": int(diff.any(axis=1).sum()),
                   "frames_whose_top_k_SET_differs": int(set_diff), "worst_score_gap": max(gaps) if gaps else 0.0,
                   "median_score_gap": float(np.median(gaps)) if gaps else 0.0,
                   "embedding_max_abs_err": float((emb - ref).abs().max()),
                   "fp32_gap_between_rank_k_and_k_plus_1_min": float((ref_top.values[:, k - 1] - ref_top.values[:, k]).min())}
        del m
    return out


def text_record(args, ctx, text, steps, warmup):
    """CLIP text tower (phrase bank) throughput: 512-phrase batches of 77 tokens."""
    import torch

    from oracle import weights as W
    ids = W.token_ids("large14", 512, 77, seed=0).to(ctx.dev)
    for _ in range(warmup):
        text(ids)
    ctx.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        text(ids)
    ev1.record()
    ctx.barrier()
    ms = ctx.max(ev0.elapsed_time(ev1)) / steps
    L, D, I = 77, 768, 3072
    gf = 12 * (2 * L * D * 3 * D + 2 * L * D * D + 4 * L * D * I + 4 * L * L * D) / 1e9
    return {"metric": "phrases/sec embedded (CLIP text tower)", "value": ctx.world * 512 / (ms / 1e3), "unit": "phrases/s",
            "n_gpus": ctx.world, "ms_per_step": ms, "tflops_per_gpu": 512 / (ms / 1e3) * gf / 1e3,
            "config": {"workload": "CLIP ViT-L/14 text tower, 512 phrases x 77 tokens per GPU per step"}}


VG_BANK_SIZES = {"objects": 19965, "attributes": 16693, "scenes": 365, "verbs": 7414}  # the reference's `vg` ontology


def tokenize_record(args, ctx, vision, text, n_videos, num_frm=8, k=5, out_dir=None, warm=True, from_uint8=True):
    """The CLIP branch of run_visual_tokenization.py over `n_videos` synthetic videos x num_frm frames (all ranks together):
    phrase bank by the native text tower (512 phrases per call like the reference), decoded uint8 frames from pinned host
    memory -> CLIP pre-processing on the GPU -> native image tower, similarity + top-k per bank on the device, aggregation on
    the host, ONE all-gather of the JSON rows, rank 0 writes visual_tokens.json.  Videos are sharded with the reference's
    slice formula (:429-431).  Wall clock between two barriers, max over ranks."""
    import tempfile as _tf

    import torch

    from oracle import weights as W
    from vidil_b200 import distributed as vdist, visual_tokenization as vt
    phrases = {key: [f"{key} {i}" for i in range(n)] for key, n in VG_BANK_SIZES.items()}
    videos = [f"video{i}" for i in range(n_videos)]
    start, end = vdist.shard_bounds(len(videos), ctx.world, ctx.rank)
    mine = videos[start:end]
    fb = args.batch - args.batch % num_frm                      # frames per tower call: whole videos
    n_calls = (len(mine) * num_frm + fb - 1) // fb
    use_u8 = from_uint8 and hasattr(vision, "encode_u8_stream")
    if use_u8:
        host = [torch.randint(0, 256, (fb, 240, 320, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    else:
        host = [torch.randn(fb, 3, 224, 224).pin_memory() for _ in range(2)]

    def run_once(write):
        t0 = time.perf_counter()
        reps = {}
        for key, n in VG_BANK_SIZES.items():
            # the bank's 512-phrase batches are split over the ranks and exchanged with one all-gather of embedding rows
            # (visual_tokenization.get_text_embeddings_clip(shard_over_ranks=True)); the reference embeds it on every rank
            starts = list(range(0, n, 512))
            per = (len(starts) + ctx.world - 1) // ctx.world
            embs = [text(W.token_ids("large14", min(512, n - i), 77, seed=i).to(ctx.dev)) for i in starts[ctx.rank * per:(ctx.rank + 1) * per]]
            mine_rows = torch.cat(embs) if embs else torch.zeros(0, 768, device=ctx.dev)
            reps[key] = {"text_embeds": vdist.all_gather_rows(mine_rows)}
        torch.cuda.synchronize()
        t_bank = time.perf_counter() - t0
        embeds = []
        stream = (host[i & 1] for i in range(n_calls))
        it = vision.encode_u8_stream(stream) if use_u8 else vision.encode_host_stream(stream)
        for e in it:
            embeds.append(e.to(ctx.dev, non_blocking=True))
        image_embeds = torch.cat(embeds)[:len(mine) * num_frm] if embeds else torch.zeros(0, 768, device=ctx.dev)
        torch.cuda.synchronize()
        t_frames = time.perf_counter() - t0 - t_bank
        rows = vt.tokens_from_embeddings(image_embeds, reps, phrases, mine, [[""]] * len(mine), num_frm, k) if mine else {}
        t_tok = time.perf_counter() - t0 - t_bank - t_frames
        path = os.path.join(out_dir or _tf.mkdtemp(), "visual_tokens.json") if (write and ctx.rank == 0) else None
        merged = vdist.gather_and_write(rows, path, device=ctx.dev)
        t_all = time.perf_counter() - t0
        return t_all, t_bank, t_frames, t_tok, (len(merged) if merged is not None else 0), path

    if warm:
        run_once(False)
    ctx.barrier()
    t_all, t_bank, t_frames, t_tok, n_rows, path = run_once(True)
    t_all = ctx.max(t_all)
    return {"metric": "frames/sec tokenized (text bank + CLIP tower + sim + top-k + aggregation + JSON gather)",
            "value": n_videos * num_frm / t_all, "unit": "frames/s", "n_gpus": ctx.world, "seconds": t_all,
            "rank0_seconds": {"phrase_bank": t_bank, "frames": t_frames, "sim_topk_aggregate": t_tok,
                              "gather_and_write": t_all - t_bank - t_frames - t_tok},
            "rows_merged": n_rows, "output_file_bytes": os.path.getsize(path) if path else None,
            "frames_input": "uint8 320x240 decoded frames, pinned host memory, CLIP pre-processing on the GPU" if use_u8
                            else "fp32 pre-processed frames, pinned host memory",
            "scaling": "strong",
            "config": {"workload": f"{n_videos} synthetic videos x {num_frm} frames, vg-sized phrase banks {VG_BANK_SIZES}, "
                                   f"top-{k}, tower batch {fb} frames"}}


class CapFiltModels:
    """Captioner (BLIP ViT + med.py decoder) and filterer (BLIP_ITM: its own ViT + multimodal encoder + itm_head), random init."""

    def __init__(self, args, ctx):
        import torch

        from vidil_b200.blip import BLIP_Decoder, BLIP_ITM
        torch.manual_seed(0)
        cap = BLIP_Decoder(image_size=args.image_size, vit=args.vit, compute_dtype=args.dtype)
        _randomise(cap)
        with torch.no_grad():
            cap.text_decoder.cls.predictions.decoder.weight.normal_(0.0, 0.1)     # spread-out logits, like a trained head
        self.cap = cap.to(ctx.dev).eval()
        itm = BLIP_ITM(image_size=args.image_size, vit=args.vit, compute_dtype=args.dtype, cache_identical_inputs=False)
        _randomise(itm)
        self.itm = itm.to(ctx.dev).eval()


def capfilt_step(args, ctx, models, frames, V, Fv, ev=None, T_itm=35):
    """One pass of the CapFilt models of run_video_CapFilt.py over V videos x Fv frames (frames on the device): captioner ViT ->
    beam search (beams 3, max_length 20, min_length 5, :102) -> filterer ViT -> ITM over every (caption, frame) pair of a
    video (:108-120), max over frames, keep > 0.5.  Returns (caption token ids on the host, keep mask)."""
    import torch
    cap, itm, dev = models.cap, models.itm, ctx.dev
    n_frames, chunk = V * Fv, args.batch
    if ev:
        ev[0].record()
    toks = torch.cat([cap.visual_encoder(frames[i:i + chunk]) for i in range(0, n_frames, chunk)])
    if ev:
        ev[1].record()
    ids = torch.tensor([cap._prompt_ids], dtype=torch.long).repeat(n_frames, 1)
    ids[:, 0] = cap.bos_token_id
    out, scores, lens = cap.text_decoder.generate(input_ids=ids[:, :-1], max_length=20, min_length=5, num_beams=3,
                                                  eos_token_id=cap.sep_token_id, pad_token_id=cap.pad_token_id,
                                                  encoder_hidden_states=toks, return_scores=True)
    del toks
    if ev:
        ev[2].record()
    # filterer: every generated caption of a video against each of its frames (duplicates are not removed here: worst case)
    cap_ids = torch.zeros(n_frames, T_itm, dtype=torch.int32, device=dev)
    L = min(out.shape[1], T_itm)
    cap_ids[:, :L] = out[:, :L].int()
    cap_ids[:, 0] = 101                                     # [CLS], as the tokenizer call of blip_itm.py:46 produces
    mask = (torch.arange(T_itm, device=dev)[None] < lens[:, None].clamp(max=T_itm)).int()
    itoks = torch.cat([itm.visual_encoder(frames[i:i + chunk]) for i in range(0, n_frames, chunk)])
    if ev:
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
    if ev:
        ev[4].record()
    return out.cpu(), keep


def capfilt_record(args, ctx, models, n_videos, steps, warmup, Fv=8, profile=True, cpu_baseline=False):
    """BASELINE.json configs[2]: `n_videos` synthetic videos x Fv frames (all ranks together) through capfilt_step; under
    torchrun the videos are sharded with the reference's slice formula (run_video_CapFilt.py:239-241) and the kept captions
    are merged by one all-gather of JSON rows."""
    import torch

    from vidil_b200 import _lib, distributed as vdist
    dev = ctx.dev
    v_start, v_end = vdist.shard_bounds(n_videos, ctx.world, ctx.rank)
    V = v_end - v_start
    n_frames = V * Fv
    frames = torch.randn(n_frames, 3, args.image_size, args.image_size, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    for _ in range(warmup):
        capfilt_step(args, ctx, models, frames, V, Fv)
    ctx.barrier()
    before = _lib.launch_count()
    acc = [0.0] * 4
    for _ in range(steps):
        out, keep = capfilt_step(args, ctx, models, frames, V, Fv, ev)
        torch.cuda.synchronize()
        for i in range(4):
            acc[i] += ev[i].elapsed_time(ev[i + 1])
    launches = _lib.launch_count() - before
    ms = [ctx.max(a / steps) for a in acc]
    total = ctx.max(sum(acc) / steps)
    rec = {"metric": "frames/sec through CapFilt (caption + filter)", "value": n_videos * Fv / (total / 1e3), "unit": "frames/s",
           "n_gpus": ctx.world, "ms_per_step": total, "steps": steps, "warmup": warmup, "gpu_launches": int(ctx.sum(launches)),
           "scaling": "strong",
           "stages_ms": {"captioner_vit": ms[0], "caption_beam_search": ms[1], "filterer_vit": ms[2], "itm_pairs": ms[3]},
           "caption_frames_per_s": n_videos * Fv / ((ms[0] + ms[1]) / 1e3),
           "captions_per_s_beam_search_only": n_videos * Fv / (ms[1] / 1e3),
           "decode_rows_per_rank": n_frames * 3, "itm_pairs_per_rank": n_frames * Fv, "dtype": args.dtype, "data": "synthetic"}
    if profile:
        # per-class device time of ONE more beam search with an event pair around every kernel (not part of the timed steps): the
        # decode-step cross-attention is the dominant kernel and a pure K/V stream, so its roofline is HBM
        cap = models.cap
        native = cap.text_decoder.bert._ensure_packed()
        native.set_profiling(True)
        toks = torch.cat([cap.visual_encoder(frames[i:i + args.batch]) for i in range(0, n_frames, args.batch)])
        ids = torch.tensor([cap._prompt_ids], dtype=torch.long).repeat(n_frames, 1)
        ids[:, 0] = cap.bos_token_id
        cap.text_decoder.generate(input_ids=ids[:, :-1], max_length=20, min_length=5, num_beams=3, eos_token_id=cap.sep_token_id,
                                  pad_token_id=cap.pad_token_id, encoder_hidden_states=toks)
        prof = native.read_profile()
        native.set_profiling(False)
        del toks
        peaks = measured_peaks()
        xa = prof["attention"]
        rec["roofline"] = {
            "bound": "hbm", "kernel": "cross_decode_mma_kernel + the prompt's attention_x_kernel (cross-attention onto the image "
            "tokens, one launch per layer and step)", "achieved": xa["bytes"] / (xa["ms"] / 1e3) / 1e9 if xa["ms"] else None,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": xa["bytes"] / (xa["ms"] / 1e3) / 1e9 / peaks["hbm_gbs"] if xa["ms"] else None,
            **measured_traffic("cross_decode_mma_kernel", f"capfilt_{args.vit}_{args.image_size}_f{n_frames}"),
            "algorithmic_bytes_per_launch": xa["bytes"] / max(xa["launches"], 1), "launches": xa["launches"],
            "peak_source": peaks["source"], "share_of_beam_search": xa["ms"] / max(sum(v["ms"] for v in prof.values()), 1e-9)}
        rec["beam_search_kernel_classes"] = {
            kk: {"ms": v["ms"], "launches": v["launches"], "tflops": v["flops"] / (v["ms"] / 1e3) / 1e12 if v["ms"] else 0.0,
                 "gbs": v["bytes"] / (v["ms"] / 1e3) / 1e9 if v["ms"] else 0.0} for kk, v in prof.items()}
    # the rows a rank contributes: video -> kept captions (token ids here: there is no vocabulary to decode with), then the one
    # collective of the path
    t0 = time.perf_counter()
    rows = {f"video{v_start + v}": [out[v * Fv + i].tolist() for i in range(Fv) if bool(keep[v * Fv + i])] for v in range(V)}
    merged = vdist.gather_and_write(rows, None, device=dev)
    rec["gather"] = {"ms": ctx.max((time.perf_counter() - t0) * 1e3), "rows": len(merged) if merged is not None else 0,
                     "collective": "all_gather of length-prefixed JSON rows"}
    if cpu_baseline and ctx.rank == 0:
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
        rec["cpu_baseline"] = {"value": 2 / dt, "unit": "captions/s (beam search only, image tokens given)", "cores": os.cpu_count(),
                               "kind": "port", "sample": "oracle/med_oracle.generate on 2 frames of image tokens (cached decoder, "
                                                          "beams 3, max_length 20)"}
    tokens = (args.image_size // 16) ** 2 + 1
    rec["config"] = {"workload": f"{n_videos} synthetic videos x {Fv} frames @{args.image_size}, BLIP ViT-{args.vit[0].upper()}/16 + "
                                 f"med.py decoder (beam 3, max_length 20, min_length 5) + BLIP_ITM filter over {n_frames * Fv} "
                                 f"(caption, frame) pairs x 35 tokens per rank; {tokens} image tokens per frame; videos sharded over "
                                 f"{ctx.world} rank(s)"}
    return rec


def pipeline_record(args, ctx, capm, vision, text, n_videos=6250, cap_frames=4, tok_frames=8, videos_per_pass=128):
    """BASELINE.json configs[4]: both drivers over 50 000 synthetic MSRVTT-shape frames = 6 250 videos (4 frames per video for
    CapFilt, 8 for visual tokenization: pipeline_config_msrvtt_test.yaml:13,35), videos sharded over the ranks with the
    reference's slice formula, one all-gather of JSON rows per driver, rank 0 writes frame_captions.json and
    visual_tokens.json.  Wall clock between barriers, max over ranks; synthetic frames are generated once per pass shape and
    re-used (the data loader is out of scope)."""
    import shutil
    import tempfile as _tf

    import torch

    from vidil_b200 import distributed as vdist
    out_dir = _tf.mkdtemp(prefix="vidil_pipeline_") if ctx.rank == 0 else None
    v_start, v_end = vdist.shard_bounds(n_videos, ctx.world, ctx.rank)
    mine = v_end - v_start
    frames = torch.randn(videos_per_pass * cap_frames, 3, args.image_size, args.image_size, device=ctx.dev)
    capfilt_step(args, ctx, capm, frames, videos_per_pass, cap_frames)      # warm-up: plans, workspaces
    ctx.barrier()
    t0 = time.perf_counter()
    rows = {}
    for v0 in range(0, mine, videos_per_pass):
        V = min(videos_per_pass, mine - v0)
        out, keep = capfilt_step(args, ctx, capm, frames[:V * cap_frames], V, cap_frames)
        for v in range(V):
            rows[f"video{v_start + v0 + v}"] = [out[v * cap_frames + i].tolist() for i in range(cap_frames) if bool(keep[v * cap_frames + i])]
    torch.cuda.synchronize()
    t_cap_compute = time.perf_counter() - t0
    merged = vdist.gather_and_write(rows, os.path.join(out_dir, "frame_captions.json") if ctx.rank == 0 else None, device=ctx.dev)
    ctx.barrier()
    t_cap = ctx.max(time.perf_counter() - t0)
    cap_rows = len(merged) if merged is not None else 0
    cap_bytes = os.path.getsize(os.path.join(out_dir, "frame_captions.json")) if ctx.rank == 0 else 0
    del frames
    tok = tokenize_record(args, ctx, vision, text, n_videos, num_frm=tok_frames, out_dir=out_dir, warm=False)
    total = t_cap + tok["seconds"]
    rec = {"metric": "frames/sec through run_video_CapFilt + run_visual_tokenization (BASELINE configs[4])",
           "value": n_videos * tok_frames / total, "unit": "frames/s", "n_gpus": ctx.world, "seconds": total, "scaling": "strong",
           "videos": n_videos, "frames": n_videos * tok_frames, "videos_per_rank": mine,
           "capfilt": {"seconds": t_cap, "rank0_compute_seconds": t_cap_compute, "frames": n_videos * cap_frames,
                       "frames_per_s": n_videos * cap_frames / t_cap, "rows_merged": cap_rows, "output_file_bytes": cap_bytes},
           "tokenize": {kk: tok[kk] for kk in ("seconds", "value", "rank0_seconds", "rows_merged", "output_file_bytes", "frames_input")},
           "collective": "one all_gather of length-prefixed JSON rows per driver; rank 0 merges in rank order and writes the files",
           "config": {"workload": f"{n_videos} synthetic videos: CapFilt on {cap_frames} frames each (ViT-{args.vit[0].upper()}/16 "
                                  f"@{args.image_size} + beam search + ITM), visual tokenization on {tok_frames} frames each (CLIP "
                                  f"ViT-L/14 + vg-sized banks, top-5); {ctx.world} rank(s)"}}
    if out_dir:
        shutil.rmtree(out_dir, ignore_errors=True)
    return rec


def other_workloads(args, ctx, line):
    """The non-headline BASELINE configs, driver-timed in the same process: failures are recorded, never fatal to the line."""
    import gc

    import torch
    out = {}

    def attempt(name, fn):
        t0 = time.perf_counter()
        try:
            out[name] = fn()
        except Exception as e:  # noqa: BLE001 - the headline line must still print
            out[name] = {"error": f"{type(e).__name__}: {e}"}
        out[name]["wall_seconds"] = time.perf_counter() - t0
        gc.collect()
        torch.cuda.empty_cache()

    vision = text = capm = None
    try:
        vision, text = build_clip_models(args, ctx)
    except Exception as e:  # noqa: BLE001
        out["clip_models"] = {"error": f"{type(e).__name__}: {e}"}
    if vision is not None:
        attempt("clip", lambda: clip_record(args, ctx, vision, max(3, args.steps // 4), 3))
        attempt("sim", lambda: sim_record(args, ctx, 20, 3, flip_frames=args.flip_frames if ctx.world == 1 else 0))
        attempt("tokenize", lambda: tokenize_record(args, ctx, vision, text, 256 * ctx.world))
    try:
        capm = CapFiltModels(args, ctx)
    except Exception as e:  # noqa: BLE001
        out["capfilt_models"] = {"error": f"{type(e).__name__}: {e}"}
    if capm is not None:
        attempt("capfilt", lambda: capfilt_record(args, ctx, capm, 128 * ctx.world, 3, 1))
    if ctx.rank == 0:
        line["workloads"] = out
    if not args.no_pipeline and vision is not None and capm is not None:
        t0 = time.perf_counter()
        try:
            rec = pipeline_record(args, ctx, capm, vision, text)
        except Exception as e:  # noqa: BLE001
            rec = {"error": f"{type(e).__name__}: {e}"}
        rec["wall_seconds"] = time.perf_counter() - t0
        if ctx.rank == 0:
            line["pipeline"] = rec


def run_single(args):
    """--workload X on its own: one JSON line with that record."""
    import torch.distributed as dist
    ctx = make_ctx(args)
    w = args.workload
    if w == "sim":
        rec = sim_record(args, ctx, args.steps, args.warmup, flip_frames=args.flip_frames)
    elif w == "clip":
        rec = clip_record(args, ctx, build_clip_models(args, ctx, want_text=False)[0], args.steps, args.warmup)
    elif w == "text":
        rec = text_record(args, ctx, build_clip_models(args, ctx)[1], args.steps, args.warmup)
    elif w == "tokenize":
        vision, text = build_clip_models(args, ctx)
        rec = tokenize_record(args, ctx, vision, text, args.videos)
    elif w == "capfilt":
        rec = capfilt_record(args, ctx, CapFiltModels(args, ctx), args.videos, args.steps, args.warmup, cpu_baseline=not args.no_cpu_baseline)
    else:
        vision, text = build_clip_models(args, ctx)
        rec = pipeline_record(args, ctx, CapFiltModels(args, ctx), vision, text, n_videos=args.videos if args.videos != 1024 else 6250)
    if ctx.rank == 0:
        print(json.dumps(rec), flush=True)
    if ctx.world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="vit", choices=["vit", "clip", "sim", "text", "tokenize", "capfilt", "pipeline"])
    ap.add_argument("--no-workloads", action="store_true", help="headline line only: skip the clip / sim / tokenize / capfilt sub-records")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the configs[4] pipeline sub-record")
    ap.add_argument("--flip-frames", type=int, default=256, help="frames of the end-to-end index-flip report in the sim record (0: off)")
    ap.add_argument("--pair-chunk", type=int, default=2048, help="--workload capfilt: (caption, frame) pairs per ITM call")
    ap.add_argument("--videos", type=int, default=1024, help="--workload tokenize | capfilt: synthetic videos (8 frames each), all ranks")
    ap.add_argument("--batch", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--vit", default="large", choices=list(VIT))
    ap.add_argument("--image-size", type=int, default=224)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--clip-dtype", default="fp16", choices=["bf16", "fp16"],
                    help="operand type of the CLIP towers (fp16: fewest end-to-end top-k flips, see the sim record)")
    ap.add_argument("--ref-frames", type=int, default=8, help="--impl reference: frames per CPU step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if os.environ.get("VIDIL_BENCH_WATCHDOG"):   # developer aid: dump every thread's Python stack and exit if the run stalls
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["VIDIL_BENCH_WATCHDOG"]), exit=True)
    if args.warmup < 3 and args.impl == "native":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference(args)
    if args.workload != "vit":
        return run_single(args)
    return run_vit(args)


if __name__ == "__main__":
    main()
