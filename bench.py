#!/usr/bin/env python
"""Benchmark of the LoongX denoising hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

  python bench.py --gpus N --steps K --warmup W            native arm (one process per GPU; N>1 via torchrun)
  python bench.py --impl reference --steps K --warmup W    reference arm: the oracle port of the reference's PyTorch
                                                            path on the host cores (the reference itself cannot be
                                                            imported here: diffusers/peft/s4torch are not installed)

A "step" is ONE EDIT of the per-GPU batch: neural conditioning (CS3, once) + 28 denoise steps of the Flux MM-DiT at
512x512 with an image condition (S = 512 + 1024 + 1024 tokens) — BASELINE.json configs[1].  metric = edits/sec.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "512x512 28-step edits/sec"
DENOISE_STEPS = 28
RES = 512
N_TXT = 512


def flops_per_forward(n_txt, n_img, n_cond, D=3072, blocks=57):
    """BASELINE.md §3 / SURVEY.md §8d algorithmic FLOPs of one DiT forward of one sample."""
    S = n_txt + n_img + n_cond
    return blocks * (24 * D * D * S + 4 * S * S * D) + 2 * 64 * D * (n_img + n_cond) + 2 * 4096 * D * n_txt + 2 * D * 64 * n_img + 10.8e9


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=d.get("bf16_tflops_sustained", 1424.9), hbm=d.get("hbm_gbs", 6445.6), which="measured (MEASURED_PEAKS.json, sustained)")
    return dict(tflops=1400.0, hbm=6650.0, which="fallback (B200_PROFILING.md: ~1.4 PF sustained, 6.65 TB/s)")


# --------------------------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: oracle port on the host cores, bounded sample
# --------------------------------------------------------------------------------------------------------------------
class CpuSample:
    """One double-stream + one single-stream block forward (fp32, full FLUX width, B=1, S=2560) of the oracle; an edit
    is 28 x (19 double + 38 single) of these (the embedders / final layer are < 0.1 % of the FLOPs)."""

    def __init__(self):
        from oracle import flux_dit as O

        torch.set_num_threads(os.cpu_count() or 1)
        self.O = O
        self.cfg = O.FluxConfig(num_layers=1, num_single_layers=1)
        self.P = O.init_params(self.cfg, seed=1234, dtype=torch.float32)
        g = torch.Generator().manual_seed(42)
        n = (RES // 16) ** 2
        D = self.cfg.inner_dim
        self.h = torch.randn(1, n, D, generator=g)
        self.e = torch.randn(1, N_TXT, D, generator=g)
        self.c = torch.randn(1, n, D, generator=g)
        self.temb = torch.randn(1, D, generator=g)
        side = RES // 16
        ids = torch.zeros(side, side, 3)
        ids[..., 1] += torch.arange(side)[:, None]
        ids[..., 2] += torch.arange(side)[None, :]
        ids = ids.reshape(-1, 3)
        cids = ids.clone()
        cids[:, 2] -= side
        self.rope = O.rope_tables(torch.cat([torch.zeros(N_TXT, 3), ids], 0))
        self.crope = O.rope_tables(cids)
        self.cores = torch.get_num_threads()
        self.sample = ("oracle port (reference block.py restated in PyTorch fp32), 1 double-stream + 1 single-stream block "
                       "forward at FLUX width, B=1, S=2560; extrapolated x28 steps x(19 double + 38 single) blocks")

    @torch.no_grad()
    def run(self):
        """-> (seconds for the double block, seconds for the single block)."""
        O, P, cfg = self.O, self.P, self.cfg
        t0 = time.perf_counter()
        e, h, c = O.block_forward(P, cfg, 0, self.h, self.e, self.c, self.temb, self.temb, self.crope, self.rope, {})
        t1 = time.perf_counter()
        x = torch.cat([e, h], 1)
        t2 = time.perf_counter()
        O.single_block_forward(P, cfg, 0, x, self.temb, self.rope, c, self.temb, self.crope, {})
        t3 = time.perf_counter()
        return t1 - t0, t3 - t2

    @staticmethod
    def edit_seconds(td, ts):
        return DENOISE_STEPS * (19 * td + 38 * ts)


def run_reference(args, rank):
    if rank != 0:
        return
    s = CpuSample()
    for _ in range(args.warmup):
        s.run()
    t0 = time.perf_counter()
    tds, tss = [], []
    for _ in range(args.steps):
        td, ts = s.run()
        tds.append(td)
        tss.append(ts)
    wall = time.perf_counter() - t0
    td, ts = sum(tds) / len(tds), sum(tss) / len(tss)
    edit_s = s.edit_seconds(td, ts)
    val = 1.0 / edit_s
    flop = 2 * 660.4e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "edits/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1, args.gpus),
        "cpu_baseline": {"value": val, "unit": "edits/s", "cores": s.cores, "kind": "port", "sample": s.sample,
                         "sample_tflops": flop / (td + ts) / 1e12, "s_per_edit_extrapolated": edit_s},
        "e2e": {"value": val, "unit": "edits/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference repo cannot be imported on this image (diffusers/peft/s4torch absent): this arm times the "
                "oracle restatement of its PyTorch path on the host cores; each step is the bounded sample described in "
                "cpu_baseline.sample",
    }
    print(json.dumps(line), flush=True)


def workload_config(batch_per_gpu, n_gpus):
    return {"workload": "BASELINE.json configs[1]: 512x512 edit, 28 denoise steps, EEG-only CS3 conditioning "
                        "(eeg_only_replace), image condition (S=512+1024+1024), guidance 3.5",
            "batch_per_gpu": batch_per_gpu, "global_batch": batch_per_gpu * n_gpus, "denoise_steps": DENOISE_STEPS,
            "parallelism": f"dp{n_gpus} (edits sharded by batch, no data-path collective)",
            "l2": "every denoise step streams ~40 GB of weights (>> 126 MB L2): inputs larger than L2, no explicit flush"}


# --------------------------------------------------------------------------------------------------------------------
# native arm
# --------------------------------------------------------------------------------------------------------------------
def run_native(args, rank, local_rank, world):
    import ctypes as C

    import torch.distributed as dist

    from loongx_b200 import _lib as L
    from src.flux.condition import Condition
    from src.flux.generate import generate
    from src.train.model import OminiModel

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    model = OminiModel("synthetic", lora_config={"r": 4, "lora_alpha": 4}, device=str(dev), model_config={
        "union_cond_attn": True, "add_cond_attn": False, "latent_lora": False})
    pipe = model.flux_pipe
    side = RES // 8  # latent 64 x 64

    def mk(seed, *shape, scale=1.0, dtype=torch.bfloat16, pin=False):
        g = torch.Generator().manual_seed(seed + 1000 * rank)
        t = (torch.randn(*shape, generator=g) * scale).to(dtype)
        return t.pin_memory() if pin else t

    host = dict(latents=mk(42, B, 16, side, side, pin=True), cond=mk(43, B, 16, side, side, pin=True),
                pe=mk(44, B, N_TXT, 4096, scale=0.1, pin=True), pooled=mk(44, B, 768, pin=True),
                eeg=mk(45, B, 4, 5000, dtype=torch.float32, pin=True))
    devt = {k: v.to(dev) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    out_host = torch.empty((B, (side // 2) ** 2, 64), dtype=torch.bfloat16).pin_memory()
    d2h = out_host.numel() * 2

    def edit(t):
        cnd = Condition("subject", condition=t["cond"], position_delta=[0, -(RES // 16)])
        return generate(model, pipe, conditions=[cnd], prompt_embeds=t["pe"], pooled_prompt_embeds=t["pooled"],
                        height=RES, width=RES, num_inference_steps=DENOISE_STEPS, latents=pipe._pack_latents(t["latents"]),
                        output_type="latent", default_lora=True, additional_condition1=t["eeg"], use_brain_condition=True,
                        fuse_flag=False, eeg_only_replace=True).images

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        out = edit(devt)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all(), "non-finite latents"

    # ---- timed region: K edits, inputs resident in HBM
    clocks = ClockSampler(local_rank)
    barrier()
    L.lib.lx_launch_count_reset()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        edit(devt)
    e1.record()
    barrier()
    clk = clocks.stop()
    launches = int(L.lib.lx_launch_count(-1))
    secs = max_over_ranks(e0.elapsed_time(e1) / 1e3)
    value = world * B * args.steps / secs

    # ---- end to end through generate() with HOST buffers: H2D of every input + D2H of the result inside the timing
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        t = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        out_host.copy_(edit(t), non_blocking=True)
        torch.cuda.synchronize()
    e2e_secs = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * B * args.steps / e2e_secs

    # ---- per-kernel-class CUDA-event timing of one more edit (events recorded on the launching stream in the C ABI)
    L.lib.lx_profile_begin()
    edit(devt)
    ms = (C.c_double * 4)()
    cnt = (C.c_int64 * 4)()
    work = (C.c_double * 4)()
    L.check(L.lib.lx_profile_end(ms, cnt, work), "lx_profile_end")
    peaks = load_peaks()
    names = ["gemm_bf16_kernel", "attention_kernel", "dit_row_kernels", "cs3_dgf_kernels"]
    tot_ms = sum(ms) or 1.0
    gemm_tf = work[0] / ms[0] / 1e9 if ms[0] else 0.0
    att_tf = work[1] / ms[1] / 1e9 if ms[1] else 0.0
    row_gbs = work[2] / ms[2] / 1e6 if ms[2] else 0.0
    traffic, traffic_src = None, None
    # DRAM bytes per launch of the dominant kernel: the newest committed ncu --set full capture (scripts/ncu_summary.py
    # gemm_traffic); re-captured whenever csrc/gemm.cu changes
    tps = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.startswith("gemm_traffic_") and f.endswith(".json"))
    tp = os.path.join(ROOT, "profiles", tps[-1]) if tps else ""
    if tp and B == 1:  # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
        tj = json.load(open(tp))
        traffic, traffic_src = tj["traffic_bytes_per_launch"], tj["source"]
    roofline = {"bound": "tensor", "kernel": "lx::gemm_bf16_kernel (tcgen05 GEMM + fused epilogues)", "achieved": gemm_tf,
                "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": gemm_tf / peaks["tflops"], "traffic": traffic,
                "traffic_unit": "bytes per launch (dram read + write), average over the 152 block GEMMs of a step",
                "traffic_source": traffic_src,
                "peak_source": peaks["which"], "launches_per_edit": int(cnt[0]), "avg_launch_us": ms[0] / max(cnt[0], 1) * 1e3,
                "share_of_edit": ms[0] / tot_ms,
                "method": "CUDA events around every launch of one extra edit, on the launching stream (lx_profile_*)"}
    kernels = {names[i]: {"ms_per_edit": ms[i], "launches": int(cnt[i]), "share": ms[i] / tot_ms} for i in range(4)}
    kernels["attention_kernel"].update({"achieved_tflops": att_tf, "frac_of_peak": att_tf / peaks["tflops"]})
    kernels["dit_row_kernels"].update({"achieved_gbs": row_gbs, "frac_of_hbm_peak": row_gbs / peaks["hbm"]})
    cs3_gbs = work[3] / ms[3] / 1e6 if ms[3] else 0.0  # algorithmic bytes of the CS3 / DGF kernels
    kernels["cs3_dgf_kernels"].update({"achieved_gbs": cs3_gbs, "frac_of_hbm_peak": cs3_gbs / peaks["hbm"],
                                       "note": "once per edit, ~30 small launches: latency-bound, 0.15 % of the edit"})
    n_img = (RES // 16) ** 2
    algo_tflops = B * DENOISE_STEPS * flops_per_forward(N_TXT, n_img, n_img) / (secs / args.steps) / 1e12

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        s = CpuSample()
        s.run()
        reps = []
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 12.0 or len(reps) < 2:
            reps.append(s.run())
        td, ts = sum(r[0] for r in reps) / len(reps), sum(r[1] for r in reps) / len(reps)
        cpu = {"value": 1.0 / s.edit_seconds(td, ts), "unit": "edits/s", "cores": s.cores, "kind": "port",
               "sample": s.sample + f" ({len(reps)} repetitions, {time.perf_counter() - t0:.1f} s)",
               "s_per_edit_extrapolated": s.edit_seconds(td, ts)}

    if rank == 0 and world == 1 and not args.no_vendor_leg:
        try:
            mix = gemm_mix_leg(dev)
            mix["in_loop_frac_of_cublas_same_shapes"] = gemm_tf / mix["cublas_tflops"]
            roofline["same_shapes_sustained"] = mix
        except Exception as e:  # noqa: BLE001  (context only: must never cost the headline line)
            roofline["same_shapes_sustained"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()

    configs = None
    if not args.no_config_legs:
        if world == 1:  # the extra legs must never cost the headline line (at N > 1 a rank-local failure tears the job down anyway)
            try:
                configs = run_config_legs(args, model, dev, rank, local_rank, world, peaks)
            except Exception as e:  # noqa: BLE001
                configs = {"error": f"{type(e).__name__}: {e}"[:400]}
        else:
            configs = run_config_legs(args, model, dev, rank, local_rank, world, peaks)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "edits/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(B, world),
            "e2e": {"value": e2e_value, "unit": "edits/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "src.flux.generate.generate(model, pipe, ...) with pinned host inputs, result copied to host"},
            "gpu_launches": launches, "clocks": clk, "roofline": roofline, "kernels": kernels,
            "algorithmic_tflops_per_gpu": algo_tflops, "frac_of_peak_end_to_end": algo_tflops / peaks["tflops"],
            "cpu_baseline": cpu, "configs": configs,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()



# --------------------------------------------------------------------------------------------------------------------
# roofline context: the GEMM mix of one denoise step (M = 2560 projections of a batch-1 edit), stand-alone and SUSTAINED
# (seconds, i.e. at the power cap like the loop itself), once with lx_gemm_bf16 and once with the vendor library
# (torch.matmul -> cuBLAS).  MEASURED_PEAKS' sustained figure was taken on an 8192^3 problem; this is what the vendor
# GEMM reaches on THESE shapes under the same power cap.  Reported next to the roofline, never part of the headline.
# --------------------------------------------------------------------------------------------------------------------
def gemm_mix_leg(dev, warm=16, passes=64):
    from loongx_b200 import ops

    M, ncopy = 2560, 4
    shapes = {"qkv": (9216, 3072), "attn_out": (3072, 3072), "ff_up": (12288, 3072), "ff_down": (3072, 12288),
              "single_qkv_mlp": (21504, 3072), "single_out": (3072, 15360)}
    g = torch.Generator(device=dev).manual_seed(5)
    W = {k: [torch.randn(n, kk, generator=g, device=dev, dtype=torch.bfloat16) * 0.02 for _ in range(ncopy)]
         for k, (n, kk) in shapes.items()}  # 4 x 452 MB: every launch streams its weights from HBM, as in the loop
    A = {kk: torch.randn(M, kk, generator=g, device=dev, dtype=torch.bfloat16) for kk in {v[1] for v in shapes.values()}}
    O = {n: torch.empty(M, n, device=dev, dtype=torch.bfloat16) for n in {v[0] for v in shapes.values()}}
    seq = ["qkv", "attn_out", "ff_up", "ff_down"] * 19 + ["single_qkv_mlp", "single_out"] * 38
    flop = sum(2.0 * M * shapes[k][0] * shapes[k][1] for k in seq)
    out = {"what": "GEMM mix of one denoise step (19 x [qkv, attn_out, ff_up, ff_down] + 38 x [single qkv|mlp, single out], "
                   f"M = 2560, rotating weight copies > L2), {warm} warm-up + {passes} timed passes back to back per library "
                   "(sustained: at the power cap), CUDA events", "tflop_per_pass": flop / 1e12}
    for which in ("lx", "cublas"):
        def one(it):
            for j, k in enumerate(seq):
                n, kk = shapes[k]
                w = W[k][(it + j) % ncopy]
                if which == "lx":
                    ops.gemm(A[kk], w, None, O[n])
                else:
                    torch.matmul(A[kk], w.t(), out=O[n])
        for it in range(warm):
            one(it)
        clocks = ClockSampler(dev.index or 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clocks.start()
        e0.record()
        for it in range(passes):
            one(it)
        e1.record()
        torch.cuda.synchronize()
        clk = clocks.stop()
        out[which + "_tflops"] = passes * flop / e0.elapsed_time(e1) / 1e9
        out[which + "_sm_mhz"] = clk.get("sm_mhz")
    return out


# --------------------------------------------------------------------------------------------------------------------
# config legs: BASELINE.json configs[2], [3], [4] measured after the headline region of the same run (extra keys of the
# same JSON line; the headline value / roofline above are untouched).  One warm-up pass + one timed pass each (a pass is
# 28-50 DiT forwards), CUDA events, max over ranks, clocks sampled during the timed pass.
# --------------------------------------------------------------------------------------------------------------------
def run_config_legs(args, model, dev, rank, local_rank, world, peaks):
    import gc

    import torch.distributed as dist

    from loongx_b200 import train as T
    from src.flux.condition import Condition
    from src.flux.generate import generate

    pipe = model.flux_pipe

    def mk(seed, *shape, scale=1.0, dtype=torch.bfloat16):
        g = torch.Generator().manual_seed(seed + 1000 * rank)
        return (torch.randn(*shape, generator=g) * scale).to(dtype).to(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn, warm=1, reps=1):
        for _ in range(warm):
            fn()
        clocks = ClockSampler(local_rank)
        barrier()
        clocks.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        clk = clocks.stop()
        secs = e0.elapsed_time(e1) / 1e3 / reps
        if world > 1:
            t = torch.tensor([secs], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
        return secs, clk

    def free():
        model.transformer._plans.clear()
        gc.collect()
        torch.cuda.empty_cache()

    def edit_leg(B, res, steps, label):
        side = res // 8
        n_img = (res // 16) ** 2
        t = dict(latents=mk(42, B, 16, side, side), cond=mk(43, B, 16, side, side), pe=mk(44, B, N_TXT, 4096, scale=0.1),
                 pooled=mk(44, B, 768), eeg=mk(45, B, 4, 5000, dtype=torch.float32), fnirs=mk(46, B, 6, 600, dtype=torch.float32),
                 ppg=mk(47, B, 4, 256, dtype=torch.float32), motion=mk(48, B, 6, 100, dtype=torch.float32))

        def edit():
            cnd = Condition("subject", condition=t["cond"], position_delta=[0, -(res // 16)])
            return generate(model, pipe, conditions=[cnd], prompt_embeds=t["pe"], pooled_prompt_embeds=t["pooled"], height=res,
                            width=res, num_inference_steps=steps, latents=pipe._pack_latents(t["latents"]), output_type="latent",
                            default_lora=True, additional_condition1=t["eeg"], additional_condition2=t["fnirs"],
                            additional_condition3=t["ppg"], additional_condition4=t["motion"], use_brain_condition=True,
                            fuse_flag=True).images

        free()
        secs, clk = timed(edit)
        tf = B * steps * flops_per_forward(N_TXT, n_img, n_img) / secs / 1e12
        free()
        return {"workload": label, "batch_per_gpu": B, "global_batch": B * world, "resolution": res, "denoise_steps": steps,
                "tokens": N_TXT + 2 * n_img, "ms_per_edit_batch": secs * 1e3, "edits_per_s": world * B / secs,
                "algorithmic_tflops_per_gpu": tf, "frac_of_peak_end_to_end": tf / peaks["tflops"], "clocks": clk,
                "timing": "1 warm-up pass + 1 timed pass, CUDA events, max over ranks"}

    out = {}
    out["c2_dgf_b4"] = edit_leg(4, 512, DENOISE_STEPS, "BASELINE.json configs[2]: 512x512, 28 steps, EEG+PPG -> fuse_eeg (DGF) and "
                                "fNIRS+Motion -> fuse_fnirs, DUAN fuse with the text embeddings (fuse_flag), per-GPU batch 4 "
                                "(batch 32 on 8 GPUs)")
    out["c3_1024_50"] = edit_leg(1, 1024, 50, "BASELINE.json configs[3]: 1024x1024, 50 steps, neural (all four signals, fuse_flag) + "
                                 "text (speech -> prompt embeddings) conditioning, per-GPU batch 1 (batch 8 on 8 GPUs)")
    # configs[4]: training step, per-GPU batch 8 (global batch 64 on 8 GPUs), one mean all-reduce of the flat gradient bucket
    B, side = 8, RES // 8
    g = torch.Generator().manual_seed(7 + 1000 * rank)
    r = lambda *s, scale=1.0, dt=torch.bfloat16: (torch.randn(*s, generator=g) * scale).to(dt).to(dev)  # noqa: E731
    batch = dict(image=r(B, 16, side, side), condition=r(B, 16, side, side), prompt_embeds=r(B, N_TXT, 4096, scale=0.1),
                 pooled_prompt_embeds=r(B, 768), position_delta=[[0, -(RES // 16)]], condition_type=["subject"] * B,
                 eeg=r(B, 4, 5000, dt=torch.float32), fnirs=r(B, 6, 600, dt=torch.float32),
                 ppg=r(B, 4, 256, dt=torch.float32), motion=r(B, 6, 100, dt=torch.float32))
    free()
    model.use_brain_condition, model.fuse_flag = True, True
    opt = torch.optim.AdamW(model.trainable_parameters() if hasattr(model, "trainable_parameters") else model.lora_layers, lr=1e-4)

    def one_step():
        opt.zero_grad(set_to_none=True)
        loss = model.step(batch)
        loss.backward()
        opt.step()
        return loss

    T.ALLREDUCE_TIMING = True
    one_step()  # builds the trainer (transposed panels, activation slabs)
    T.ALLREDUCE_LOG.clear()
    secs, clk = timed(one_step, warm=1, reps=2)
    ar = T.ALLREDUCE_LOG[-2:]
    ar_ms = [e[0].elapsed_time(e[1]) for e in ar]
    skew_ms = [e[3].elapsed_time(e[0]) for e in ar]  # waiting for the slowest rank's backward before the exchange
    T.ALLREDUCE_TIMING = False
    n_img = (RES // 16) ** 2
    S, D, nb = N_TXT + 2 * n_img, 3072, 57
    F = nb * (24 * D * D * S + 4 * S * S * D)
    tr = model._trainer_obj
    flop = (2 if tr.recompute else 1) * F + nb * 24 * D * D * S + 2.5 * nb * 4 * S * S * D  # executed work per sample
    tf = B * flop / secs / 1e12
    out["c4_train"] = {
        "workload": "BASELINE.json configs[4]: OminiModel.step forward + backward + AdamW, Flux-DiT LoRA r=4 + CS3/DGF "
                    "conditioning (step fuse order), 512x512 + image condition, per-GPU batch 8 (global 64 on 8 GPUs)",
        "batch_per_gpu": B, "global_batch": B * world, "micro_batch": tr.B, "recompute": tr.recompute,
        "ms_per_step": secs * 1e3, "samples_per_s": world * B / secs, "algorithmic_tflops_per_gpu": tf,
        "frac_of_peak_end_to_end": tf / peaks["tflops"],
        "allreduce": {"ms": (sum(ar_ms) / len(ar_ms)) if ar_ms else 0.0, "bytes": (ar[-1][2] if ar else tr.grad_flat.numel() * 4),
                      "calls_per_step": 1, "rank_skew_ms": (sum(skew_ms) / len(skew_ms)) if skew_ms else 0.0,
                      "what": "NCCL all-reduce (sum) + 1/world scale of the flat fp32 gradient bucket, timed after a one-element "
                              "all-reduce has lined the ranks up (rank_skew_ms = the wait for the slowest rank's backward, not "
                              "part of the exchange); 0 ms at n_gpus = 1 (no collective is issued)"},
        "trainable_elements": int(tr.grad_flat.numel()), "clocks": clk,
        "timing": "1 build step + 1 warm-up step + 2 timed steps, CUDA events, max over ranks"}
    return out

# --------------------------------------------------------------------------------------------------------------------
# training workload (BASELINE.json configs[4]; not the headline metric): OminiModel.step forward + backward + all-reduce
# --------------------------------------------------------------------------------------------------------------------
def run_train(args, rank, local_rank, world):
    import torch.distributed as dist

    from loongx_b200 import _lib as L
    from src.train.model import OminiModel

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    model = OminiModel("synthetic", lora_config={"r": 4, "lora_alpha": 4}, device=str(dev), model_config={
        "union_cond_attn": True, "add_cond_attn": False, "latent_lora": False}, use_brain_condition=True, fuse_flag=True)
    side = RES // 8
    g = torch.Generator().manual_seed(7 + 1000 * rank)
    r = lambda *s, scale=1.0, dt=torch.bfloat16: (torch.randn(*s, generator=g) * scale).to(dt).to(dev)  # noqa: E731
    batch = dict(image=r(B, 16, side, side), condition=r(B, 16, side, side), prompt_embeds=r(B, N_TXT, 4096, scale=0.1),
                 pooled_prompt_embeds=r(B, 768), position_delta=[[0, -(RES // 16)]], condition_type=["subject"] * B,
                 eeg=r(B, 4, 5000, dt=torch.float32), fnirs=r(B, 6, 600, dt=torch.float32),
                 ppg=r(B, 4, 256, dt=torch.float32), motion=r(B, 6, 100, dt=torch.float32))
    model.step(batch)  # fixes the geometry, builds the transposed panels
    opt = torch.optim.AdamW(model.lora_layers, lr=1e-4)

    def one_step():
        opt.zero_grad(set_to_none=True)
        loss = model.step(batch)  # CS3/DGF conditioning (step fuse order) + native DiT forward
        loss.backward()           # native backward + ONE NCCL all-reduce of the flat LoRA-gradient bucket
        opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        loss = one_step()
    clocks = ClockSampler(local_rank)
    barrier()
    L.lib.lx_launch_count_reset()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = one_step()
    e1.record()
    barrier()
    clk = clocks.stop()
    secs = e0.elapsed_time(e1) / 1e3
    if world > 1:
        t = torch.tensor([secs], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    launches = int(L.lib.lx_launch_count(-1))
    n_img = (RES // 16) ** 2
    S, D, nb = N_TXT + 2 * n_img, 3072, 57
    F = nb * (24 * D * D * S + 4 * S * S * D)
    tr = model._trainer_obj
    # SURVEY.md §8d per sample: forward F (+ F again only when the blocks are recomputed in the backward) + backward
    # [dX of every Linear + 2.5x the attention]
    flop = (2 if tr.recompute else 1) * F + nb * 24 * D * D * S + 2.5 * nb * 4 * S * S * D
    peaks = load_peaks()
    if rank == 0:
        tf = B * args.steps * flop / secs / 1e12
        print(json.dumps({
            "metric": "512x512 LoRA train samples/sec", "value": world * B * args.steps / secs, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[4]: train step, Flux-DiT LoRA r=4 + CS3/DGF conditioning (EEG+PPG, "
                                   "fNIRS+Motion, fuse_flag), 512x512 + image condition, LoRA + CS3/DGF encoder gradients, AdamW on the "
                                   "LoRA factors (model.py:533-558), mean all-reduce of the one flat gradient bucket",
                       "batch_per_gpu": B, "global_batch": B * world, "micro_batch": tr.B, "recompute": tr.recompute,
                       "parallelism": f"dp{world} (NCCL all-reduce of {tr.grad_flat.numel() / 1e6:.1f} M fp32)"},
            "gpu_launches": launches, "clocks": clk, "loss": float(loss.detach()),
            "algorithmic_tflops_per_gpu": tf, "frac_of_peak_end_to_end": tf / peaks["tflops"]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------------------------
# VAE workload (extra, not the headline metric): latents -> pixels either side of the loop
# --------------------------------------------------------------------------------------------------------------------
def run_vae(args, rank, local_rank, world):
    import ctypes as C

    from loongx_b200 import _lib as L
    from loongx_b200.vae import NativeVae, VaeConfig, VaeWeights, synthetic_params

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    B, side = args.batch, RES // 8
    cfg = VaeConfig()
    P = synthetic_params(cfg, 1234)
    vae = NativeVae(VaeWeights(cfg, P, dev))
    g = torch.Generator().manual_seed(11 + rank)
    z_host = torch.randn(B, 16, side, side, generator=g).pin_memory()
    img_host = torch.empty(B, 3, RES, RES, dtype=torch.float32).pin_memory()
    z_dev = z_host.to(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        barrier()
        secs = e0.elapsed_time(e1) / 1e3
        if world > 1:
            t = torch.tensor([secs], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
        return secs

    def e2e_step():
        img_host.copy_(vae.decode(z_host.to(dev, non_blocking=True), return_dict=False)[0], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    clocks = ClockSampler(local_rank)
    clocks.start()
    L.lib.lx_launch_count_reset()
    secs = timed(lambda: vae.decode(z_dev, return_dict=False))
    launches = int(L.lib.lx_launch_count(-1)) * args.steps // (args.steps + args.warmup)
    clk = clocks.stop()
    secs_e2e = timed(e2e_step)
    L.lib.lx_profile_begin()
    vae.decode(z_dev, return_dict=False)
    ms, n, w = (C.c_double * 4)(), (C.c_int64 * 4)(), (C.c_double * 4)()
    L.lib.lx_profile_end(ms, n, w)
    peaks = load_peaks()
    if rank == 0:
        out = {
            "metric": "512x512 VAE decodes/sec", "value": world * B * args.steps / secs, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "FLUX.1-dev AutoencoderKL decoder (generate.py:375-380), latents [B,16,64,64] -> image "
                                   "[B,3,512,512], seeded synthetic weights", "batch_per_gpu": B, "global_batch": B * world,
                       "l2": "each decode writes and re-reads 10.5 GB of im2col panels per sample (>> 126 MB L2)"},
            "e2e": {"value": world * B * args.steps / secs_e2e, "unit": "images/s", "h2d_bytes_per_step": z_host.numel() * 4,
                    "d2h_bytes_per_step": img_host.numel() * 4, "api": "pipeline.vae.decode(z) with pinned host latents, image copied to host"},
            "gpu_launches": launches, "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": "lx::im2col_scatter_kernel + GroupNorm statistics (VAE row kernels)",
                         "achieved": w[2] / max(ms[2], 1e-9) / 1e6, "peak": peaks["hbm"], "unit": "GB/s",
                         "frac": w[2] / max(ms[2], 1e-9) / 1e6 / peaks["hbm"], "traffic": None, "peak_source": peaks["which"],
                         "share_of_step": ms[2] / max(ms[0] + ms[2], 1e-9),
                         "method": "CUDA events around every launch of one extra decode (lx_profile_*)"},
            "kernels": {"gemm_bf16_kernel": {"ms": ms[0], "launches": int(n[0]), "achieved_tflops": w[0] / max(ms[0], 1e-9) / 1e9},
                        "vae_row_kernels": {"ms": ms[2], "launches": int(n[2])}},
        }
        if not args.no_cpu_baseline:
            from oracle import vae as OV

            torch.set_num_threads(os.cpu_count() or 1)
            zs = z_host[:1, :, :side // 2, :side // 2].clone()
            OV.decode_raw(P, zs, OV.VaeConfig())
            t0 = time.perf_counter()
            OV.decode_raw(P, zs, OV.VaeConfig())
            dt = time.perf_counter() - t0
            out["cpu_baseline"] = {"value": 1.0 / (4 * dt), "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": f"oracle/vae.py decode of one 256x256 image in fp32 ({dt:.2f} s), x4 pixels extrapolated"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="edits per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vendor-leg", action="store_true",
                    help="skip the sustained lx / cuBLAS GEMM-mix context measurement (roofline.same_shapes_sustained)")
    ap.add_argument("--no-config-legs", action="store_true",
                    help="skip the BASELINE.json configs[2..4] legs that follow the headline measurement")
    ap.add_argument("--workload", default="edit", choices=["edit", "train", "vae"],
                    help="edit = the headline metric (default); train = BASELINE.json configs[4]; vae = the decoder either side "
                         "of the loop (extras, not the headline)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N bench.py --gpus N ...")
    if args.workload == "train":
        run_train(args, rank, local_rank, world)
        return
    if args.workload == "vae":
        run_vae(args, rank, local_rank, world)
        return
    run_native(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
