#!/usr/bin/env python
"""bench.py — denoise-steps/s of the FLUX.1-dev DiT hot path on N B200s of one node.

One "step" = one denoise step of one candidate: DiT forward (19 double + 38 single blocks,
1024x1024 -> 4608 tokens) + flow-match Euler update.  At N GPUs every rank runs its own candidate
(the outer search loop's embarrassingly parallel axis: weak scaling) and the ranks exchange their
verifier scores with one NCCL all-gather at the end of the timed region, as a search round does.

  python bench.py [--gpus N] [--steps K] [--warmup W]          product arm (CUDA kernels)
  python bench.py --impl reference ...                          reference arm: the reference's CPU
                                                                path (oracle port) on host cores
Prints ONE JSON line (see the task contract): value = whole-job steps/s with inputs resident in
HBM; e2e = same through the public Python API with pinned-host inputs and a host read-back;
roofline = dominant kernel vs the measured bf16 tensor peak; cpu_baseline = oracle on host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STEPS_PER_IMAGE = 28
H = W = 1024
N_TXT = 512
N_IMG = (H // 16) * (W // 16)
D_MODEL = 3072


def algorithmic_tflop(n_tok: int, layers: int = 57, d: int = D_MODEL) -> float:
    """57 * 24 D^2 N (GEMM) + 57 * 4 N^2 D (attention), 2 FLOP/MAC."""
    return (layers * 24 * d * d * n_tok + layers * 4 * n_tok * n_tok * d) / 1e12


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"tflops_burst": j["bf16_tflops"], "tflops_sustained": j["bf16_tflops_sustained"],
                "hbm_gbs": j["hbm_gbs"], "source": "measured"}
    return {"tflops_burst": 1590.0, "tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.t = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(s for s, p in zip(sm, pw) if p > 300) or sorted(sm)
        return {"sm_mhz": busy[len(busy) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw)}


def cpu_reference_leg(steps: int, warmup: int):
    """The reference's CPU implementation of the path (oracle port of block.py / transformer.py over
    restated diffusers leaves) on the host cores.  A full FLUX.1-dev step is 74.4 TFLOP and the
    model is 23.8 GB, so each timed "step" here is a BOUNDED SAMPLE: one double-stream and one
    single-stream block at the real geometry (N = 4608, D = 3072); the per-step time is
    19 x t_double + 38 x t_single (the embedders and the final layer are < 0.1 % of the FLOPs)."""
    import torch
    from oracle import flux_oracle as fo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = fo.FluxConfig.tiny_depth(1, 1)
    model = fo.FluxTransformer2DModel(cfg)
    fo.init_weights_(model, seed=0)
    model.eval()
    g = torch.Generator().manual_seed(0)
    bf = torch.bfloat16
    img = torch.randn(1, N_IMG, D_MODEL, generator=g).to(bf)
    txt = torch.randn(1, N_TXT, D_MODEL, generator=g).to(bf)
    temb = torch.randn(1, D_MODEL, generator=g).to(bf)
    ids = torch.cat([torch.zeros(N_TXT, 3, dtype=bf), fo.prepare_latent_image_ids(H // 16, W // 16)])
    rope = model.pos_embed(ids)
    lora = fo.LoraSet()
    times = []
    with torch.no_grad():
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            t_, i_, _ = fo.double_block(model.transformer_blocks[0], "transformer_blocks.0.", lora,
                                        img, txt, None, temb, None, rope, None, {})
            t1 = time.perf_counter()
            x = torch.cat([t_, i_], dim=1)
            t2 = time.perf_counter()
            fo.single_block(model.single_transformer_blocks[0], "single_transformer_blocks.0.", lora,
                            x, temb, rope, None, None, None, {})
            t3 = time.perf_counter()
            if it >= warmup:
                times.append((t1 - t0, t3 - t2))
    td = sum(t[0] for t in times) / len(times)
    ts = sum(t[1] for t in times) / len(times)
    step_s = 19 * td + 38 * ts
    return {"value": 1.0 / step_s, "unit": "denoise-steps/s", "cores": cores, "kind": "port",
            "sample": (f"1 double + 1 single FLUX.1-dev block at N=4608, bf16, torch CPU "
                       f"({td:.2f} s, {ts:.2f} s); step = 19 x double + 38 x single = {step_s:.1f} s "
                       f"(extrapolated, {len(times)} timed samples)"),
            "ms_per_step": step_s * 1e3}


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL's version banner, ...) write to fd 1; the contract is ONE JSON line on stdout.
    Route fd 1 to stderr for the whole run and keep the real stdout for the final line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def _emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=STEPS_PER_IMAGE)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-text", action="store_true", help="skip the T5/CLIP encode timing")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layers", default="19,38", help="double,single layer counts (debug only)")
    args = ap.parse_args()
    K, Wm = args.steps, max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "configs[1]: FLUX.1-dev DiT 1024x1024 (4608 tokens), 28-step schedule, "
                          "1 candidate per GPU, no condition stream, guidance 3.5, batch 1",
              "parallelism": f"candidate-sharded x{args.gpus} (replicated weights)",
              "l2": "inputs larger than L2: every step streams 23.8 GB of weights (L2 = 126 MB)",
              "weights": "random-init (seeded), FLUX.1-dev architecture, 11.9 B params bf16"}

    if args.impl == "reference":
        if rank != 0:
            return
        leg = cpu_reference_leg(max(1, min(K, 2)), 1 if Wm else 0)
        line = {"impl": "reference", "metric": "denoise-steps/sec", "value": leg["value"],
                "unit": "denoise-steps/s", "n_gpus": args.gpus, "steps": K, "warmup": Wm,
                "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
                "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": leg["value"], "unit": "denoise-steps/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        _emit(line)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        _emit({"error": "no CUDA device: the product arm has no CPU fallback"})
        sys.exit(2)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from reflectionflow_b200 import _lib as L
    from reflectionflow_b200.config import FluxDiTConfig
    from reflectionflow_b200.pipeline import B200FluxPipeline, flow_match_schedule
    from reflectionflow_b200.tts.search import gather_scores, stub_verifier_score

    nl, ns = (int(x) for x in args.layers.split(","))
    cfg = FluxDiTConfig(num_layers=nl, num_single_layers=ns)
    pipe = B200FluxPipeline.from_synthetic(cfg, seed=0, device=dev, with_vae=True)
    model = pipe.transformer
    lib = L.load()

    # synthetic inputs of the named shape (seeded per rank = per candidate)
    g = torch.Generator().manual_seed(1000 + rank)
    lat_host = torch.randn(1, N_IMG, 64, generator=g).to(torch.bfloat16).pin_memory()
    txt_host = torch.randn(1, N_TXT, cfg.joint_attention_dim, generator=g).to(torch.bfloat16).pin_memory()
    pool_host = torch.randn(1, cfg.pooled_projection_dim, generator=g).to(torch.bfloat16).pin_memory()
    lat, txt, pool = lat_host.to(dev), txt_host.to(dev), pool_host.to(dev)
    img_ids = pipe._prepare_latent_image_ids(1, H // 16, W // 16, dev, torch.bfloat16)
    txt_ids = torch.zeros(N_TXT, 3, dtype=torch.bfloat16, device=dev)

    def sched(n):
        ts, sig = flow_match_schedule(max(n, 1), N_IMG)  # the mu-shifted n-step schedule
        return (ts.to(torch.bfloat16) / 1000), sig

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also captures the step graph)
    if Wm:
        t_w, s_w = sched(Wm)
        model.denoise(lat, txt, pool, t_w, s_w, 3.5, img_ids, txt_ids)
    barrier()

    # ---- timed region: K steps, inputs resident in HBM
    t_k, s_k = sched(K)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = lib.rf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    final = model.denoise(lat, txt, pool, t_k, s_k, 3.5, img_ids, txt_ids)
    score = stub_verifier_score(final)
    if world > 1:
        gather_scores(score, rank, world, dev)  # the round's one exchange step (NCCL all-gather)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.rf_launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: the public API call with HOST inputs and a host read-back
    def e2e_call(n):
        # everything the device-timed region above does (denoise, score, the round's exchange step)
        # plus the host<->device copies of the call a user makes
        out = pipe(prompt_embeds=txt_host, pooled_prompt_embeds=pool_host, latents=lat_host,
                   num_inference_steps=n, guidance_scale=3.5, height=H, width=W,
                   output_type="latent")
        sc = stub_verifier_score(out.images)
        if world > 1:
            gather_scores(sc, rank, world, dev)
        sc.cpu()
        return out.images.cpu()
    e2e_call(1)
    barrier()
    e0.record()
    res = e2e_call(K)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = t.item()
    h2d = (lat_host.numel() + txt_host.numel() + pool_host.numel()) * 2
    d2h = res.numel() * 2 + 8

    # ---- e2e at the reference's PER-STEP call granularity (tranformer_forward, transformer.py:47):
    # every step copies its latents host->device from pinned memory and reads the prediction back
    from reflectionflow_b200.transformer import tranformer_forward
    t_k16 = t_k.to(torch.bfloat16)
    gd = torch.tensor([3.5], device=dev)
    txt_d, pool_d = txt_host.to(dev), pool_host.to(dev)
    pred_host = torch.empty(1, N_IMG, 64, dtype=torch.bfloat16).pin_memory()

    def step_call(i):
        x = lat_host.to(dev, non_blocking=True)
        out = tranformer_forward(model, None, None, None, {}, 0, hidden_states=x, encoder_hidden_states=txt_d,
                                 pooled_projections=pool_d, timestep=t_k16[i:i + 1].to(dev), img_ids=img_ids,
                                 txt_ids=txt_ids, guidance=gd, return_dict=False)[0]
        pred_host.copy_(out, non_blocking=True)
    step_call(0)
    barrier()
    e0.record()
    for i in range(K):
        step_call(i)
    e1.record()
    barrier()
    ms_fwd = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_fwd], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_fwd = t.item()

    # ---- VAE decode of the final latent (the per-image tail: generate.py:302-307), device-timed
    for _ in range(2):
        pipe.vae.decode_packed(final, H, W, "u8")
    barrier()
    e0.record()
    for _ in range(3):
        img = pipe.vae.decode_packed(final, H, W, "u8")
    e1.record()
    barrier()
    ms_vae = e0.elapsed_time(e1) / 3

    # ---- per-kernel breakdown of one eager forward (CUDA events around every launch)
    prof = None
    # (taken hot: a 6-step denoise runs immediately before, so the kernels are timed in the same
    # power-capped state as the timed region above and compare against the SUSTAINED peak)
    if rank == 0:
        t_h, s_h = sched(6)
        model.denoise(lat, txt, pool, t_h, s_h, 3.5, img_ids, txt_ids)
        L.profile_start()
        model(hidden_states=lat, encoder_hidden_states=txt, pooled_projections=pool,
              timestep=torch.tensor([1.0], dtype=torch.bfloat16), img_ids=img_ids, txt_ids=txt_ids,
              guidance=torch.tensor([3.5]), return_dict=False)
        prof = L.profile_stop()
    # ---- text encoders of one candidate prompt (T5-XXL 512 tokens + CLIP-L 77 tokens: generate.py:148-161)
    ms_text = None
    if rank == 0 and not args.no_text:
        from reflectionflow_b200.text import B200TextEncoders
        enc = B200TextEncoders(device=dev).init_synthetic_weights(seed=3)
        gi = torch.Generator().manual_seed(5)
        ids_t5 = torch.randint(0, 32000, (1, N_TXT), generator=gi).pin_memory()
        ids_clip = torch.randint(3, 49000, (1, 77), generator=gi)
        ids_clip[:, -1] = 49407
        ids_clip = ids_clip.pin_memory()
        for _ in range(2):
            enc.t5_encode(ids_t5); enc.clip_encode(ids_clip)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            enc.t5_encode(ids_t5); enc.clip_encode(ids_clip)
        e1.record()
        torch.cuda.synchronize()
        ms_text = e0.elapsed_time(e1) / 3
        enc.close()
        del enc

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    pk = peaks()
    n_tok = N_TXT + N_IMG
    step_tflop = algorithmic_tflop(n_tok, nl + ns)
    steps_per_s = args.gpus * K / (ms / 1e3)
    tot_ms = sum(v["ms"] for v in prof.values())
    dom_name, dom = max(prof.items(), key=lambda kv: kv[1]["ms"])
    # tensor-bound kernels: FLOPs / time vs the SUSTAINED measured peak (kernel timed inside a long step)
    ach = dom["flops"] / (dom["ms"] / 1e3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(dom_name)
    roofline = {"kernel": dom_name, "bound": "tensor", "achieved": ach,
                "peak": pk["tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tflops_sustained"],
                "peak_source": pk["source"] + " (sustained cuBLAS bf16)", "traffic": traffic,
                "traffic_note": "DRAM bytes per launch (launch-weighted over the kernel's shapes) from the "
                                "committed ncu --set full capture (profiles/ncu_traffic.json, "
                                "profiles/r01_ncu_v5_summary.md); algorithmic bytes per launch = "
                                + str(round(dom["bytes"] / dom["launches"]))
                                + "; the K=12288/15360 launches re-read W once per wave (operands exceed L2)",
                "launches_per_step": dom["launches"], "avg_launch_ms": dom["ms"] / dom["launches"],
                "share_of_step": dom["ms"] / tot_ms,
                "how": "CUDA events around every launch of one eager forward inside this run"}
    kernels = {k: {"launches": v["launches"], "ms": round(v["ms"], 4), "share": round(v["ms"] / tot_ms, 4),
                   "tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1) if v["flops"] else None,
                   "gbs": round(v["bytes"] / (v["ms"] / 1e3) / 1e9, 1)} for k, v in prof.items()}
    line = {"metric": "denoise-steps/sec", "value": steps_per_s, "unit": "denoise-steps/s",
            "n_gpus": args.gpus, "steps": K, "warmup": Wm, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": config,
            "images_per_sec_dit_only": steps_per_s / STEPS_PER_IMAGE,
            "vae_decode_ms": ms_vae,
            "text_encode_ms": ms_text,
            "images_per_sec": args.gpus / ((STEPS_PER_IMAGE * ms / K + ms_vae + (ms_text or 0.0)) / 1e3),
            "images_note": "one T5-XXL + CLIP-L prompt encode (native, ids from pinned host) + 28 denoise steps at "
                           "the measured step time + one native VAE decode to uint8",
            "step_tflop": step_tflop,
            "step_tflops_achieved": step_tflop / (ms / K / 1e3),
            "step_frac_of_tensor_peak": step_tflop / (ms / K / 1e3) / pk["tflops_sustained"],
            "e2e": {"value": args.gpus * K / (ms_e2e / 1e3), "unit": "denoise-steps/s",
                    "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
                    "api": "B200FluxPipeline.__call__(prompt_embeds=<pinned host>, latents=<pinned host>, "
                           "output_type='latent') -> score -> .cpu()",
                    "per_step_forward_api": {"value": args.gpus * K / (ms_fwd / 1e3), "unit": "denoise-steps/s",
                                             "h2d_bytes_per_step": lat_host.numel() * 2 + 2,
                                             "d2h_bytes_per_step": lat_host.numel() * 2,
                                             "api": "tranformer_forward(hidden_states=<pinned host>.to(dev)) -> "
                                                    "pinned host, once per step (no CUDA graph)"},
                    "note": "same work as `value` plus the copies; the copies are < 1 MB per K-step call, so "
                            "e2e ~= value within run-to-run clock noise (the GPU is power-capped)"},
            "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels, "clocks": clocks}
    if not args.no_cpu_baseline and args.gpus == 1:  # the CPU baseline is timed at N=1 only
        leg = cpu_reference_leg(1, 1)
        line["cpu_baseline"] = {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")}
    _emit(line)


if __name__ == "__main__":
    main()
