#!/usr/bin/env python
"""bench.py -- audio-seconds per second (1/RTF) of the SenseVoiceSmall-shaped hot path.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
  python bench.py --impl reference ...                     (the CPU arm: oracle port on host cores)

One step = one pass of the hot path (front-end -> CMVN -> 70-layer int8 encoder -> CTC head ->
greedy ids) over one batch of 64 synthetic 16 kHz x 16 s clips per GPU (BASELINE.json configs[1];
weak scaling: 64 clips per rank, configs[3] at N=8).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CLIPS_PER_GPU = 64
N_SAMPLES = 256000
AUDIO_S_PER_CLIP = N_SAMPLES / 16000.0
METRIC = "audio-sec/sec (1/RTF) SenseVoiceSmall 16kHz"
UNIT = "audio-s/s"


def linear_flops_per_clip(cfg, T):
    d, din, ffn, v = cfg.d_model, cfg.d_in, cfg.ffn, cfg.vocab
    per = lambda k, n: 2.0 * T * k * n
    tot = per(din, 3 * d) + (cfg.n_layers - 1) * per(d, 3 * d) + cfg.n_layers * (per(d, d) + per(d, ffn) + per(ffn, d)) + per(d, v)
    return tot


def cpu_baseline_run(blob, n_threads, first_clip):
    """Times the CPU restatement of lele's path (oracle port): one 16 s clip per thread, each
    thread single-threaded like lele itself (Par::Seq, src/kernels/gemm.rs:196)."""
    from lele_b200.sensevoice_weights import synth_pcm
    from oracle.binding import SenseVoiceRef
    ref = SenseVoiceRef(blob)
    clips = [synth_pcm(first_clip + i, N_SAMPLES) for i in range(n_threads)]
    ref.pcm_to_ids(clips[0][:16000])  # touch code / page in the blob
    out = [None] * n_threads

    def work(i):
        out[i] = ref.pcm_to_ids(clips[i])

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    dt = time.perf_counter() - t0
    return n_threads * AUDIO_S_PER_CLIP / dt, dt


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.gpu = gpu_index
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(pw)), "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args):
    """--impl reference: lele's own CPU implementation of the path.  The reference is Rust and
    cannot be built here (no cargo/rustc; nightly + un-vendored crates), so this arm times the
    oracle port of its algorithm on all host cores (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from lele_b200.sensevoice_weights import SenseVoiceConfig, build_blob
    cfg = SenseVoiceConfig()
    blob = build_blob(cfg, seed=1234)
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_baseline_run(blob, cores, 0)
    vals, total = [], 0.0
    for s in range(args.steps):
        v, dt = cpu_baseline_run(blob, cores, s * cores)
        vals.append(v); total += dt
    value = args.steps * cores * AUDIO_S_PER_CLIP / total
    sample = f"{cores} clips x 16 s per step (one clip per host thread, full 70-layer network + front-end)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * total / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 x u8 -> i32 (f32 epilogue / attention)", "data": "synthetic",
            "config": {"workload": "SenseVoiceSmall-shaped ASR, synthetic 16 kHz x 16 s clips, random-init int8 weights", "clips_per_step": cores},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from lele_b200 import Context, SenseVoice
    from lele_b200.distributed import broadcast_blob, gather_ids, max_over_ranks, shard_range
    from lele_b200.sensevoice_weights import SenseVoiceConfig, blob_nbytes, build_blob, synth_batch

    cfg = SenseVoiceConfig()
    B = CLIPS_PER_GPU
    n_total = B * world
    # a dedicated (non-default) stream: the library launches on exactly this stream and every CUDA event
    # below is recorded on it (a NULL handle would make the library create its own private stream)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx = Context(local_rank, stream.cuda_stream)

    # ---- weights: built on rank 0, one NCCL broadcast to the other ranks ----
    nbytes = blob_nbytes(cfg)
    if rank == 0:
        blob_host = build_blob(cfg, seed=1234)
        blob_dev = torch.from_numpy(blob_host).to(dev)
    else:
        blob_host = None
        blob_dev = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    broadcast_blob(blob_dev, 0)
    hdr_len = 256 + 16 * (10 + cfg.n_layers * 21)
    header = blob_dev[:hdr_len].cpu().numpy()
    full_hdr = np.zeros(nbytes, np.uint8) if False else None  # (header only is needed on the host)
    blob_for_ctor = np.zeros(hdr_len, np.uint8); blob_for_ctor[:] = header
    model = SenseVoice.__new__(SenseVoice)
    # construct over the already-resident device blob (no second copy)
    _init_model_from_device(model, ctx, blob_for_ctor, nbytes, blob_dev.data_ptr(), B, N_SAMPLES)
    T = model.rows(N_SAMPLES)

    # ---- inputs: this rank's shard of the global clip index ----
    s0, s1 = shard_range(n_total, rank, world)
    pcm_np = synth_batch(s0, s1 - s0, N_SAMPLES)
    pcm_pinned = torch.from_numpy(pcm_np).pin_memory()
    pcm_dev = pcm_pinned.to(dev, non_blocking=True)
    ids_dev = torch.empty((B, T), dtype=torch.int32, device=dev)
    ids_pinned = torch.empty((B, T), dtype=torch.int32).pin_memory()
    torch.cuda.synchronize(dev)

    def step_device():
        model.forward_pcm_dev(pcm_dev.data_ptr(), B, N_SAMPLES, ids_dev.data_ptr())

    def step_e2e():
        model.transcribe_host_ptr(pcm_pinned.data_ptr(), B, N_SAMPLES, ids_pinned.data_ptr())
        if world > 1:
            allids = gather_ids(ids_pinned.to(dev, non_blocking=True), n_total, 0)
            if allids is not None:
                allids.cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    th0 = time.perf_counter()
    for _ in range(args.steps):
        step_device()
    host_enqueue_ms = (time.perf_counter() - th0) * 1000.0 / args.steps   # CPU time to enqueue one step (no sync inside)
    e1.record(stream)
    barrier()
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    ms_total = max_over_ranks(e0.elapsed_time(e1), dev)
    ms_per_step = ms_total / args.steps
    value = n_total * AUDIO_S_PER_CLIP / (ms_per_step / 1000.0)

    # ---- e2e: host PCM -> host ids through the C-ABI host entry (H2D + compute + D2H timed) ----
    for _ in range(2):
        step_e2e()
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_e2e()
    e1.record(stream)
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    e2e_value = n_total * AUDIO_S_PER_CLIP / (e2e_ms / 1000.0)

    # ---- per-kernel-class device times of one extra (untimed) profiled pass -> roofline ----
    model.set_profiling(True)
    step_device()
    prof = model.last_profile()
    model.set_profiling(False)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16_sus = peaks.get("bf16_tflops_sustained")
        hbm = peaks.get("hbm_gbs", 6650.0)
        peak_src = "2 x MEASURED_PEAKS.bf16_tflops_sustained (int8 tensor rate is nominally 2x bf16; no int8 peak is measured)" if bf16_sus else \
                   "2 x fallback 1.4 PFLOP/s sustained bf16 (B200_PROFILING.md)"
        peak = 2.0 * (bf16_sus if bf16_sus else 1400.0)
        g = prof.get("gemm_i8_tcgen05", {"ms": 0.0, "calls": 0})
        flops_step = B * linear_flops_per_clip(cfg, T)
        achieved = flops_step / (g["ms"] / 1000.0) / 1e12 if g["ms"] > 0 else None
        roofline = {"bound": "tensor", "kernel": "gemm_i8_tc_kernel (tcgen05 kind::i8, fused dequant epilogue)", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": None, "peak_source": peak_src,
                    "launches_per_step": g["calls"], "avg_launch_ms": (g["ms"] / g["calls"]) if g["calls"] else None,
                    "algorithmic_flops_per_step": flops_step, "share_of_step": (g["ms"] / sum(v["ms"] for k, v in prof.items() if not k.startswith("gemm_i8:"))) if prof else None}
        try:   # DRAM bytes per launch of the same kernel from the committed ncu capture (tools/summarize_ncu.py traffic)
            tr = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json")))
            roofline["traffic"] = tr["dram_bytes_per_launch"]
            roofline["traffic_unit"] = "bytes/launch (DRAM read + write)"
            roofline["traffic_source"] = tr["source"]
        except Exception:
            pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            if blob_host is None:
                blob_host = build_blob(cfg, seed=1234)
            v, dt = cpu_baseline_run(blob_host, cores, 0)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{cores} clips x 16 s, one per host thread, full network ({dt:.1f} s wall); scalar C restatement of lele's x86 path, not lele's AVX2 kernels (lele publishes 39.1 audio-s/s on one Apple-Silicon core, README.md:19)"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 x u8 -> i32 (f32 epilogue / attention)", "data": "synthetic",
                "config": {"workload": "SenseVoiceSmall-shaped ASR (70 SANM layers d512 h4 ffn2048, CTC 25055), synthetic 16 kHz x 16 s clips, random-init int8 weights",
                           "clips_per_gpu": B, "global_clips": n_total, "rows_per_clip": T, "parallelism": f"clip-sharded x{world}",
                           "l2": "inputs + weights per step (65.5 MB PCM + 240 MB blob) exceed the 126 MB L2 and every step streams >50 GB of activations"},
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(B * N_SAMPLES * 4), "d2h_bytes_per_step": int(B * T * 4)},
                "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "kernel_breakdown_ms": {k: round(v["ms"], 3) for k, v in prof.items()}, "hbm_peak_gbs": hbm}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _init_model_from_device(model, ctx, header_np, nbytes, dev_ptr, max_clips, max_samples):
    """SenseVoice over a blob that is already resident in HBM (it arrived by NCCL broadcast)."""
    import ctypes as C
    from lele_b200._lib import call, i32, sz, vp
    hdr = header_np[:256].view(np.int32)
    model.ctx = ctx
    model.header = np.ascontiguousarray(header_np)
    model.n_layers, model.d_model, model.d_in, model.vocab = int(hdr[2]), int(hdr[3]), int(hdr[4]), int(hdr[8])
    model._own = None
    h = vp()
    call("lele_b200_sensevoice_create", ctx.h, vp(dev_ptr), sz(nbytes), model.header.ctypes.data_as(vp), sz(model.header.size), i32(max_clips), i32(max_samples), C.byref(h))
    model.h = h
    model.max_clips, model.max_samples = max_clips, max_samples


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
