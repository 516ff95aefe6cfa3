#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: vocoder samples/s (x real-time) + encoder frames/s.

    python bench.py --gpus N --steps K --warmup W                 # our arm (N>1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path, timed beside it

Headline workload (N = 1): BASELINE.json configs[2] -- ``Vocoder.generate`` batch 1, 1 s @ 16 kHz (Tc = 50 code
frames -> 16 000 autoregressive steps) from random codes + speaker id, random-init weights (seed 13), synthetic
inputs.  A "step" is one full generate call (conditioning + one persistent sample-loop launch).  With N > 1
every rank generates its own utterance (utterances shard, weak scaling) and rank 0 gathers the waveforms (NCCL).
``value`` is measured with inputs resident in HBM; ``e2e`` goes through the public Python API from pinned host
buffers and reads the waveform back.  ``extra`` carries the other BASELINE configs measured in the same run
(VQ lookup 1 M frames, Encoder.encode one 2 s utterance, batched encode) -- parity for all of them lives in tests/.
Prints exactly ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "vocoder samples/sec (x real-time) + encoder frames/sec"
SR = 16000


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]),
                    bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_generate_sample(n_steps: int, warm: int = 50):
    """The restated reference generate loop (oracle/vocoder.py, unhoisted as the reference writes it) on the
    host cores: returns (samples/s, seconds).  Bounded sample of the B=1 workload; the loop is stationary."""
    import torch
    from oracle import fixtures
    from oracle import vocoder as ovoc
    sd = ovoc.init_state_dict(seed=13)
    codes, spk, u = fixtures.vocoder_inputs(1, 50, seed=0)
    with torch.no_grad():
        ovoc.generate(sd, codes, spk, u, n_steps=warm)
        t0 = time.perf_counter()
        ovoc.generate(sd, codes, spk, u, n_steps=n_steps)
        dt = time.perf_counter() - t0
    return n_steps / dt, dt


def reference_arm(args):
    """The reference's own CPU path for the headline workload: the restated rnnms generate loop (oracle/vocoder.py) in PyTorch
    on the host cores.  torchrun exports OMP_NUM_THREADS=1, so the thread count is set EXPLICITLY: the loop is timed with
    1, 16 and all host threads (small GEMVs do not scale with threads) and the best is reported with its count."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n_ar = 800
    ncpu = os.cpu_count() or 1
    counts = sorted({1, min(16, ncpu), ncpu})
    per_threads = {}
    for c in counts:
        torch.set_num_threads(c)
        cpu_generate_sample(200, warm=20)
        sps, _ = cpu_generate_sample(n_ar, warm=20)
        per_threads[c] = sps
    best = max(per_threads, key=per_threads.get)
    torch.set_num_threads(best)
    for _ in range(max(args.warmup, 1)):
        cpu_generate_sample(200, warm=20)
    t = []
    for _ in range(args.steps):
        sps, dt = cpu_generate_sample(n_ar, warm=20)
        t.append(dt)
    total = sum(t)
    value = n_ar * args.steps / total
    sample = (f"{n_ar} AR steps of the B=1 generate loop per step (the 100-step prenet is paid per {n_ar} steps here instead of "
              "per 16000: a small bias AGAINST the reference), restated rnnms loop in PyTorch CPU")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "vocoder_generate_b1_1s (BASELINE configs[2])", "sample": sample},
        "x_realtime": value / SR,
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": best, "kind": "port", "sample": sample,
                         "samples_per_s_by_threads": {str(k): v for k, v in per_threads.items()}, "host_cpus": ncpu,
                         "note": "rank 0 only: one host process, not scaled by n_gpus"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0



def run_other_workload(args, world, rank, local, dev, timed, peaks, lib, workload=None, steps=None, warmup=None):
    """BASELINE configs[3] (batched encode, 4096 x 3 s, STRONG scaling over the ranks) and configs[4] (end-to-end
    convert: encode + generate, 64 x 3 s per GPU, weak scaling).  One JSON line, same contract as the headline."""
    import torch
    import torch.distributed as dist
    from oracle import encoder as oenc
    from oracle import fixtures
    from oracle import vocoder as ovoc
    from vectorquantizedcpc_b200 import ConfEncoder, Encoder, Vocoder
    from vectorquantizedcpc_b200 import dist as vdist

    sd = fixtures.encoder_init_state(768, seed=13)
    enc = Encoder(ConfEncoder(channels=768))
    enc.load_state_dict(sd)
    enc = enc.to(dev).eval()
    import types
    args = types.SimpleNamespace(**vars(args))
    if workload is not None:
        args.workload = workload
    if steps is not None:
        args.steps = steps
    if warmup is not None:
        args.warmup = warmup
    sampler = ClockSampler(local)
    n0 = lib.vqcpc_launch_count()
    if args.workload == "encode_4096":
        n_total, T = 4096, 300
        lo, hi = vdist.shard_range(n_total, rank, world)
        mel_h = fixtures.synthetic_mel(n_total, T, seed=0)[lo:hi].contiguous().pin_memory()
        mel_d = mel_h.to(dev)
        idx_h = torch.empty(hi - lo, 150, dtype=torch.int64).pin_memory()

        def gather_idx(idx):
            if world > 1:
                vdist.gather_utterances(idx, n_total, dst=0)

        def step_resident():
            with torch.no_grad():
                _, _, idx = enc.encode(mel_d)
            gather_idx(idx)

        def step_e2e():
            with torch.no_grad():
                # chunks of ~512 utterances (whole GEMM waves): each chunk's GEMM stack runs under the next chunk's host->device copy
                _, _, idx = enc.encode_from_host(mel_h)
                gather_idx(idx)
                idx_h.copy_(idx, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        if rank == 0:
            sampler.start()
        t_ms = timed(step_resident, args.steps, args.warmup)
        launches = (lib.vqcpc_launch_count() - n0) * args.steps // (args.steps + args.warmup)
        t_e2e = timed(step_e2e, args.steps, args.warmup)
        t_c = None
        if world > 1:
            # the same pass with the context vectors c gathered as well (629 MB into rank 0 -- the one collective
            # that could limit this path)
            def step_gather_c():
                with torch.no_grad():
                    _, c, idx = enc.encode(mel_d)
                gather_idx(idx)
                vdist.gather_utterances(c, n_total, dst=0)
            t_c = timed(step_gather_c, args.steps, 1)
        clocks = sampler.stop() if rank == 0 else {}
        frames = n_total * 150
        value, e2e_v = frames * args.steps / (t_ms * 1e-3), frames * args.steps / (t_e2e * 1e-3)
        flops = frames * 6.029e6 * 3           # bf16x3: three MMA terms per product
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16x3 (bf16 hi/lo split, fp32 accumulate)", "data": "synthetic",
                "config": {"workload": "encode_4096x3s (BASELINE configs[3])", "utterances": n_total, "mel_frames": T,
                           "channels": 768, "parallelism": f"utterances sharded over {world} GPU(s), NCCL gather of indices",
                           "timing": "CUDA events per step, L2 flushed between steps"},
                "e2e": {"value": e2e_v, "unit": "frames/s", "h2d_bytes_per_step": int(mel_h.numel() * 4 * world),
                        "d2h_bytes_per_step": int(idx_h.numel() * 8 * world), "ms_per_step": t_e2e / args.steps},
                "gpu_launches": int(launches), "clocks": clocks,
                "with_gather_of_c": None if t_c is None else {"ms_per_step": t_c / args.steps, "value": frames * args.steps / (t_c * 1e-3),
                                                               "unit": "frames/s", "gathered_bytes": int(n_total * 150 * 256 * 4)},
                "roofline": {"kernel": "gemm_tc_ln_kernel", "bound": "tensor", "unit": "TFLOP/s",
                             "achieved": flops * args.steps / (t_ms * 1e-3) / 1e12 / world, "peak": peaks["bf16_tflops_sustained"],
                             "frac": flops * args.steps / (t_ms * 1e-3) / 1e12 / world / peaks["bf16_tflops_sustained"],
                             "traffic": None, "note": "per GPU, whole encode step (GEMMs + LN + VQ + LSTM) against the sustained "
                             "cuBLAS bf16 peak; the GEMM kernel alone: profiles/r01_ncu_gemm_tc_summary.txt (0.84)"}}
        if rank == 0 and world == 1 and not args.no_cpu:
            t0 = time.perf_counter()
            with torch.no_grad():
                oenc.encode(sd, fixtures.synthetic_mel(64, T, seed=0))
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 64 * 150 / dt, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "one chunk of 64 utterances x 3 s through oracle/encoder.py (fp32)"}
        return line

    # ---- convert_b64: encode + generate, 64 utterances x 3 s per GPU
    B, T = 64, 300
    voc = Vocoder()
    vsd = ovoc.init_state_dict(seed=13)
    voc.load_state_dict(vsd)
    voc = voc.to(dev).eval()
    mel_h = fixtures.synthetic_mel(B, T, seed=rank).pin_memory()
    mel_d = mel_h.to(dev)
    g = torch.Generator().manual_seed(rank)
    spk_h = torch.randint(0, 102, (B,), generator=g).pin_memory()
    spk_d = spk_h.to(dev)
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    L = 320 * 150
    wav_h = torch.empty(B, L).pin_memory()

    def convert(mel, spk):
        with torch.no_grad():
            _, _, idx = enc.encode(mel)
            wav = voc.generate(idx, spk, generator=gen)
        if world > 1:
            vdist.gather_utterances(wav, B * world, dst=0)
        return wav

    def step_e2e():
        wav = convert(mel_h.to(dev, non_blocking=True), spk_h.to(dev, non_blocking=True))
        wav_h.copy_(wav, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    steps, warm = min(args.steps, 3), 1
    if rank == 0:
        sampler.start()
    t_ms = timed(lambda: convert(mel_d, spk_d), steps, warm)
    launches = (lib.vqcpc_launch_count() - n0) * steps // (steps + warm)
    t_e2e = timed(step_e2e, steps, warm)
    clocks = sampler.stop() if rank == 0 else {}
    samples = B * L * world
    value, e2e_v = samples * steps / (t_ms * 1e-3), samples * steps / (t_e2e * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": t_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "x_realtime": value / SR,
            "config": {"workload": "convert_b64x3s (BASELINE configs[4]): Encoder.encode + Vocoder.generate", "batch_per_gpu": B,
                       "mel_frames": T, "samples_per_utterance": L,
                       "parallelism": f"utterances sharded over {world} GPU(s), NCCL gather of waveforms"},
            "e2e": {"value": e2e_v, "unit": "samples/s", "x_realtime": e2e_v / SR, "h2d_bytes_per_step": int(mel_h.numel() * 4 + B * 8),
                    "d2h_bytes_per_step": int(B * L * 4), "ms_per_step": t_e2e / steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"kernel": "ar_batch_kernel", "bound": "latency", "achieved": 1e3 * t_ms / steps / L, "unit": "us/step",
                         "peak": None, "frac": None, "traffic": None,
                         "note": "batched sample loop (64 utterances advance per step), latency-bound; see the headline workload"}}
    if rank == 0 and world == 1 and not args.no_cpu:
        # reference CPU: oracle encode of the 64 utterances + restated generate at B=64 truncated to 200 steps, extrapolated
        codes, spk, u = fixtures.vocoder_inputs(B, 150, seed=0, n_steps=200)
        t0 = time.perf_counter()
        with torch.no_grad():
            oenc.encode(sd, mel_h)
            ovoc.generate(vsd, codes, spk, u, n_steps=200)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": B * 200 / dt, "unit": "samples/s", "x_realtime": B * 200 / dt / SR,
                                "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "oracle encode of 64 x 3 s + restated generate at B=64 truncated to 200 of 48000 steps"}
    return line


# ------------------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="generate_b1", choices=["generate_b1", "encode_4096", "convert_b64"],
                    help="generate_b1 = BASELINE configs[2] (headline); encode_4096 = configs[3]; convert_b64 = configs[4]")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configs (VQ / encoder)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from oracle import fixtures
    from oracle import vocoder as ovoc
    from vectorquantizedcpc_b200 import ConfEncoder, Encoder, Vocoder, VQEmbeddingEMA, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    _lib.check(lib.vqcpc_device_check(local), "device check")
    peaks = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def flush_l2():
        flush_buf.fill_(1)

    def timed(fn, steps, warmup, flush=True):
        """per-step CUDA events on the current stream, L2 flushed (untimed) between steps -> total ms, max over ranks"""
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        for _ in range(steps):
            if flush:
                flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        barrier()
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))

    if args.workload != "generate_b1":
        line = run_other_workload(args, world, rank, local, dev, timed, peaks, lib)
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- model + inputs (random-init weights, seed 13; synthetic inputs per SURVEY.md 8d)
    vsd = ovoc.init_state_dict(seed=13)
    voc = Vocoder()
    voc.load_state_dict(vsd)
    voc = voc.to(dev).eval()
    if args.workload == "generate_b1":
        B, Tc, wname = 1, 50, "vocoder_generate_b1_1s (BASELINE configs[2])"
    else:
        B, Tc, wname = 64, 150, "convert_generate_b64_3s (BASELINE configs[4], generate part)"
    L = 320 * Tc
    codes, spk, _ = fixtures.vocoder_inputs(B, Tc, seed=rank)
    codes_d, spk_d = codes.to(dev), spk.to(dev)
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    codes_h, spk_h = codes.pin_memory(), spk.pin_memory()
    wav_h = torch.empty(B, L).pin_memory()
    gathered = [torch.empty(B, L, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None

    def gather(wav):
        if world > 1:
            dist.gather(wav, gathered, dst=0)

    def step_resident():
        with torch.no_grad():
            wav = voc.generate(codes_d, spk_d, generator=gen)
        gather(wav)

    def step_e2e():
        with torch.no_grad():
            c = codes_h.to(dev, non_blocking=True)
            s = spk_h.to(dev, non_blocking=True)
            wav = voc.generate(c, s, generator=gen)
            gather(wav)
            wav_h.copy_(wav, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    with torch.no_grad():
        voc.pack_weights()
    sampler = ClockSampler(local)
    n0 = lib.vqcpc_launch_count()
    if rank == 0:
        sampler.start()
    t_ms = timed(step_resident, args.steps, args.warmup)
    launches = (lib.vqcpc_launch_count() - n0) * args.steps // (args.steps + args.warmup)
    t_e2e_ms = timed(step_e2e, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else {}
    samples_per_step = B * L * world
    value = samples_per_step * args.steps / (t_ms * 1e-3)
    e2e_value = samples_per_step * args.steps / (t_e2e_ms * 1e-3)

    # ---- dominant kernel alone (ar_cluster_kernel): CUDA events around the C-ABI generate call, conditioning excluded
    import ctypes as C
    w, _keep = voc.pack_weights()
    with torch.no_grad():
        G = voc.condition(codes_d, spk_d)
    uni = torch.rand(B, L, device=dev, generator=gen)
    wav = torch.empty(B, L, device=dev)
    ws_bytes = lib.vqcpc_vocoder_workspace_bytes(1, 1)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)

    def ar_only():
        _lib.check(lib.vqcpc_vocoder_generate(C.byref(w), _lib.ptr(G), _lib.ptr(uni), B, 2 * Tc, L, _lib.ptr(ws), ws_bytes,
                                              _lib.ptr(wav), None, None, _lib.current_stream_ptr()), "generate")

    t_ar_ms = timed(ar_only, args.steps, 3) / args.steps / B          # per launch (one utterance per launch)
    # The sample loop is latency-bound (its algorithmic HBM traffic is 14.8 MB per 16 000 steps).  SURVEY 8(d) names its
    # bound: per-step floor = t_smem + ONE empty grid-scope exchange, where t_smem = resident weight bytes touched per step
    # (10.81 MB fp32) / aggregate on-chip bandwidth (148 SMs x 128 B/clk x SM clock) and the exchange is measured live here
    # (vqcpc_debug_exchange_floor: 128 CTAs publish and poll, no compute).
    fl_ws = torch.empty(1 << 17, dtype=torch.uint8, device=dev)
    mean_cyc = C.c_double(0.0)
    _lib.check(lib.vqcpc_debug_exchange_floor(_lib.ptr(fl_ws), fl_ws.numel(), 2000, C.byref(mean_cyc),
                                              _lib.current_stream_ptr()), "exchange floor")
    sm_mhz = float(clocks.get("sm_mhz") or 1965.0)      # median SM clock sampled during the timed region
    us_per_step = 1e3 * t_ar_ms / L
    t_smem_us = 4.0 * (2688 * 896 + 256 * 896 + 256 * 256) / (148 * 128 * sm_mhz)      # bytes / (bytes per us)
    t_exch_us = mean_cyc.value / sm_mhz
    floor_us = t_smem_us + t_exch_us
    ar_bytes = 4 * (2688 * 896 + 256 * 896 + 256 * 256 + 2688 + 512 + 256 * 2688 + 2 * Tc * 2688 + 2 * L + 256)
    roofline = {"kernel": "ar_cluster_kernel" if lib.vqcpc_ar_cluster_active() else "ar_kernel", "bound": "latency", "unit": "us/step", "achieved": us_per_step, "peak": floor_us,
                "frac": floor_us / us_per_step,
                "floor": {"t_smem_us": t_smem_us, "one_empty_grid_exchange_us": t_exch_us, "exchange_cycles": mean_cyc.value,
                          "sm_mhz_used": sm_mhz},
                "traffic": None, "algorithmic_hbm_bytes_per_launch": ar_bytes,
                "hbm_gbs_if_it_were_bandwidth": ar_bytes / (t_ar_ms * 1e-3) / 1e9, "hbm_peak_gbs": peaks["hbm_gbs"],
                "peak_source": peaks["source"],
                "note": "latency-bound persistent kernel (lower is better): achieved = measured us per autoregressive step of the "
                        "dominant kernel alone, peak = SURVEY 8(d)'s floor (t_smem + one measured empty grid exchange), "
                        "frac = floor / achieved.  One grid-scope exchange per step is what the cluster kernel has "
                        "(DESIGN.md 4.1); the remaining gap is two DSMEM hops + the serial gate/softmax/sample chain."}
    latency = {"us_per_step": us_per_step, "x_realtime_kernel_only": (L / (t_ar_ms * 1e-3)) / SR, "target_us_per_step": 1.25}

    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wname, "batch_per_gpu": B, "code_frames": Tc, "samples_per_utterance": L,
                   "weights": "random init seed 13", "timing": "CUDA events per step, L2 flushed (256 MiB write) between steps",
                   "parallelism": f"utterances sharded over {world} GPU(s), NCCL gather of waveforms"},
        "x_realtime": value / SR, "x_realtime_per_utterance": value / SR / (B * world),
        "e2e": {"value": e2e_value, "unit": "samples/s", "x_realtime": e2e_value / SR,
                "h2d_bytes_per_step": int(codes.numel() * 8 + spk.numel() * 8), "d2h_bytes_per_step": int(B * L * 4),
                "ms_per_step": t_e2e_ms / args.steps},
        "gpu_launches": int(launches), "roofline": roofline, "latency": latency, "clocks": clocks,
    }

    # ---- N > 1: BASELINE configs[3] (encode 4096 x 3 s, STRONG scaling, indices gathered; and again with c gathered) and
    # configs[4] (convert 64 x 3 s per GPU, weak scaling) ride along, so that the driver's scaling run records them
    if not args.no_extra and world > 1:
        extra = {}
        for wl, st, wu in (("encode_4096", 5, 3), ("convert_b64", 2, 1)):
            sub = run_other_workload(args, world, rank, local, dev, timed, peaks, lib, workload=wl, steps=st, warmup=wu)
            extra[wl] = {k: sub[k] for k in ("value", "unit", "ms_per_step", "scaling", "steps", "e2e", "gpu_launches", "config",
                                              "with_gather_of_c", "x_realtime") if k in sub}
        line["extra"] = extra

    # ---- the other BASELINE configs, measured in the same run on rank 0's GPU (N = 1 only)
    if not args.no_extra and world == 1:
        extra = {}
        # sample-loop throughput beyond B = 1: B = 4 (one interleaved launch of the latency kernel), B = 64 (the batched
        # tensor-core kernel, one group of 64 utterances) and B = 128 (its two-group variant), 1 s each
        for Bb in (4, 64, 128):
            cb, sb, _ = fixtures.vocoder_inputs(Bb, 50, seed=100 + Bb)
            cbd, sbd = cb.to(dev), sb.to(dev)
            with torch.no_grad():
                ms = timed(lambda: voc.generate(cbd, sbd, generator=gen), 3, 3) / 3
            extra[f"generate_b{Bb}_1s"] = {"ms": ms, "samples_per_s": Bb * 16000 / (ms * 1e-3),
                                          "x_realtime_aggregate": Bb / (ms * 1e-3), "us_per_step": ms * 1e3 / 16000}
        # configs[1]: VQ lookup microbench, 512x64 codebook, 1 M frames (HBM-bound: 520 B/frame algorithmic)
        for kind in ("init", "trained"):
            x, cb = fixtures.vq_inputs(1_000_000, kind=kind, seed=1234)
            vq = VQEmbeddingEMA(512, 64)
            vq.embedding.copy_(cb)
            vq = vq.to(dev)
            xd = x.to(dev)
            ms = timed(lambda: vq.encode(xd), 5, 3) / 5                    # through the Python API (incl. the status sync)
            # the kernel's own launch duration: 20 back-to-back C-ABI calls between two CUDA events (codebook preparation
            # launch included), status checked after the timed region
            xf, cbf = xd.reshape(-1, 64).contiguous(), vq.embedding.contiguous()
            qo, io = torch.empty_like(xf), torch.empty(xf.shape[0], dtype=torch.int64, device=dev)

            vws_bytes = lib.vqcpc_vq_workspace_bytes()
            vws = torch.empty(vws_bytes, dtype=torch.uint8, device=dev)

            def vq_launches():
                for _ in range(20):
                    _lib.check(lib.vqcpc_vq_lookup(_lib.ptr(xf), _lib.ptr(cbf), xf.shape[0], 512, 64, _lib.ptr(qo), _lib.ptr(io),
                                                   _lib.ptr(vws), vws_bytes, _lib.current_stream_ptr()), "vq_lookup")
            kms = timed(vq_launches, 1, 1) / 20
            _lib.check(lib.vqcpc_check_status(_lib.ptr(vws), _lib.current_stream_ptr()), "vq_lookup")
            extra[f"vq_lookup_1M_{kind}"] = {"frames_per_s": 1e6 / (ms * 1e-3), "ms": ms, "kernel_ms": kms,
                                            "roofline": {"kernel": "vq_tc_kernel", "bound": "hbm",
                                                         "achieved": 520e6 / (kms * 1e-3) / 1e9,
                                                         "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                                         "frac": 520e6 / (kms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                                         "traffic": 470.2e6,   # ncu, profiles/r01_ncu_vq_tc_summary.txt
                                                         "achieved_through_api": 520e6 / (ms * 1e-3) / 1e9,
                                                         "note": "achieved = 520 B/frame / launch duration (20 back-to-back C-ABI launches "
                                                                 "between CUDA events); achieved_through_api = VQEmbeddingEMA.encode incl. its "
                                                                 "per-call status sync"}}
            del xf, qo, io
            del xd
        # wave -> log-mel front-end (preprocess.py:53-75), 64 utterances x 3 s
        from vectorquantizedcpc_b200 import LogMel
        fe = LogMel().to(dev)
        wv = (torch.rand(64, 48000, generator=torch.Generator().manual_seed(3)) * 2 - 1).to(dev)
        with torch.no_grad():
            ms = timed(lambda: fe(wv), 5, 3) / 5
        extra["logmel_64x3s"] = {"ms": ms, "audio_seconds_per_s": 64 * 3.0 / (ms * 1e-3), "frames_per_s": 64 * 301 / (ms * 1e-3)}
        del wv
        # configs[0]: Encoder.encode on one 2 s utterance (latency), C = 768 and 512
        for Cc in (768, 512):
            sd = fixtures.encoder_init_state(Cc, seed=13)
            enc = Encoder(ConfEncoder(channels=Cc))
            enc.load_state_dict(sd)
            enc = enc.to(dev).eval()
            mel = fixtures.synthetic_mel(1, 200, seed=0).to(dev)
            with torch.no_grad():
                ms = timed(lambda: enc.encode(mel), 20, 3) / 20
            extra[f"encode_1utt_2s_c{Cc}"] = {"ms": ms, "frames_per_s": 100 / (ms * 1e-3), "gemm_mode": "fp32"}
            if Cc == 768:
                # configs[3]: batched encode, 512 utterances x 3 s (the per-GPU share at 8 GPUs) and all 4096 on one GPU;
                # tcgen05 bf16 hi/lo-split GEMMs ("bf16x3", fp32-grade) and the fp32 CUDA-core parity path
                for Bb in (512, 4096):
                    melb = fixtures.synthetic_mel(Bb, 300, seed=0).to(dev)
                    for gm in ("bf16x3", "fp32"):
                        if gm == "fp32" and Bb == 4096:
                            continue
                        enc.gemm_mode = gm
                        with torch.no_grad():
                            ms = timed(lambda: enc.encode(melb), 3, 3) / 3
                        frames = Bb * 150
                        extra[f"encode_batch_{Bb}x3s_c768_{gm}"] = {
                            "ms": ms, "frames_per_s": frames / (ms * 1e-3),
                            "tflops_fp32_equivalent": frames * 6.029e6 / (ms * 1e-3) / 1e12,
                            "roofline": {"bound": "tensor", "unit": "TFLOP/s",
                                         "achieved": frames * 6.029e6 * (3 if gm == "bf16x3" else 1) / (ms * 1e-3) / 1e12,
                                         "peak": peaks["bf16_tflops_sustained"] if gm == "bf16x3" else None,
                                         "note": "whole encode (GEMMs + LN + VQ + LSTM); bf16x3 issues 3 MMA terms per product"}}
                    del melb
                enc.gemm_mode = "auto"
        line["extra"] = extra

    # ---- CPU baseline (rank 0, N = 1): the restated reference loop on the host cores, bounded sample
    if rank == 0 and world == 1 and not args.no_cpu:
        n_ar = 1500
        sps, dt = cpu_generate_sample(n_ar, warm=100)
        line["cpu_baseline"] = {"value": sps, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                                "x_realtime": sps / SR,
                                "sample": f"{n_ar} AR steps (after 100 warm-up) of the B=1 generate loop, oracle/vocoder.py"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
