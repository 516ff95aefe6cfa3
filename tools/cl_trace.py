#!/usr/bin/env python
"""Phase-level trace of the cluster sample loop (csrc/vocoder_cluster.cu; see vqcpc_debug_set_ar_cluster).
usage: cl_trace.py TAG [first_poll_delay] [poll_mode] [enable]
Writes gpurun_out/cl_trace_<tag>.json: cycles per phase of the chain warp and of M warp 0 on two CTAs, the oracle check
of the first 400 samples, and the measured us/step of whole generate calls."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fixtures, vocoder as ovoc  # noqa: E402
from vectorquantizedcpc_b200 import Vocoder, _lib  # noqa: E402

# chain-warp stamps (each taken when the named value is available): slot -> meaning
C_SLOTS = {0: "step start (x known)", 1: "h published", 2: "own fc1 rows done", 3: "fc2 partials sent", 4: "RS arrived",
           5: "RS summed (AG sent next)", 6: "hb read, gh ready", 7: "AG arrived", 12: "sampled"}
C_ORDER = [0, 1, 2, 3, 4, 5, 7, 12, 6]
M_ORDER = [(8, "bar5 released"), (9, "poll done"), (10, "W_hh dots done"), (11, "W_hh reduced, hb stored")]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "run"
    delay = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    mode = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    enable = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    dev = torch.device("cuda:0")
    voc = Vocoder()
    sd = ovoc.init_state_dict(seed=13)
    voc.load_state_dict(sd)
    voc = voc.to(dev).eval()
    Tc = 50
    codes, spk, u = fixtures.vocoder_inputs(1, Tc, seed=0)
    lib = _lib.lib()
    lib.vqcpc_debug_set_ar_cluster(enable, delay, mode)
    n, t0 = 256, 4000
    out = {"first_poll_delay": delay, "poll_mode": mode, "cluster_kernel": enable}
    cd, sdv, ud = codes.to(dev), spk.to(dev), u.to(dev)
    with torch.no_grad():
        wav, mu = voc.generate(cd, sdv, uniforms=ud, return_mulaw=True)   # warm
        # oracle check of a prefix (free-running, same uniforms)
        ns = 400
        _, ocodes, _ = ovoc.generate(sd, codes, spk, u[:, :ns], n_steps=ns, return_all=True)
        match = (mu[0, :ns].cpu() == ocodes[0]).float().mean().item()
        out["sample_match_first_400"] = match
        if enable:
            for cta in (0, 77):
                buf = torch.zeros(n + 1, 32, dtype=torch.int64, device=dev)
                lib.vqcpc_debug_set_ar_trace(buf.data_ptr(), cta, t0, n + 1)
                voc.generate(cd, sdv, uniforms=ud)
                lib.vqcpc_debug_set_ar_trace(None, 0, 0, 0)
                ts = buf.cpu().double()
                q = lambda v: [float(x) for x in torch.quantile(v.double(), torch.tensor([0.0, 0.5, 0.9, 1.0], dtype=torch.float64))]
                d = {}
                for a_, b_ in zip(C_ORDER[:-1], C_ORDER[1:]):
                    d[f"{C_SLOTS[a_]} -> {C_SLOTS[b_]}"] = q(ts[:n, b_] - ts[:n, a_])
                d["hb read, gh ready -> next step start"] = q(ts[1:n + 1, 0] - ts[:n, 6])
                m = {"h published (chain) -> bar5 released (M warp 0)": q(ts[:n, 8] - ts[:n, 1])}
                for (a_, an), (b_, bn) in zip(M_ORDER[:-1], M_ORDER[1:]):
                    m[f"{an} -> {bn}"] = q(ts[:n, b_] - ts[:n, a_])
                m["h published (chain) -> F warp 4 fc1 rows done"] = q(ts[:n, 13] - ts[:n, 1])
                m["F warp 4: fc1 rows done -> fc2 partials sent"] = q(ts[:n, 14] - ts[:n, 13])
                m["W warp 0 hb stored -> chain sampled (positive = slack)"] = q(ts[:n, 12] - ts[:n, 11])
                m["poll done after publish, per M warp (median)"] = [float(torch.median(ts[:n, 16 + w] - ts[:n, 1])) for w in range(7)]
                pd = ts[:n, 16:23] - ts[:n, 1:2]
                m["slowest M warp poll done after publish [min,med,p90,max]"] = q(pd.max(dim=1).values)
                m["poll rounds per M warp (mean)"] = [float(ts[:n, 24 + w].mean()) for w in range(7)]
                out[f"cta{cta}"] = {"cycles_per_step[min,med,p90,max]": q(ts[1:n + 1, 0] - ts[:n, 0]), "chain": d, "mwarp0": m}
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        a.record()
        for _ in range(3):
            voc.generate(cd, sdv, uniforms=ud)
        b.record()
    torch.cuda.synchronize()
    out["ms_per_generate"] = a.elapsed_time(b) / 3
    out["us_per_step"] = out["ms_per_generate"] * 1e3 / (320 * Tc)
    out["x_realtime"] = 320 * Tc / 16000 / (out["ms_per_generate"] * 1e-3)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"cl_trace_{tag}.json"), "w"), indent=1)
    print(f"[{tag}] cluster={enable} delay {delay} mode {mode}: {out['us_per_step']:.3f} us/step, {out['x_realtime']:.1f}x RT, "
          f"match400 {match:.4f}")
    if enable:
        c = out["cta77"]
        print("   step", [round(v) for v in c["cycles_per_step[min,med,p90,max]"]])
        for k, v in c["chain"].items():
            print(f"   chain  {k:58s} min {round(v[0]):5d} med {round(v[1]):5d} p90 {round(v[2]):5d}")
        for k, v in c["mwarp0"].items():
            if len(v) == 4:
                print(f"   mwarp0 {k:58s} min {round(v[0]):5d} med {round(v[1]):5d} p90 {round(v[2]):5d}")
            else:
                print(f"   mwarp0 {k}: {[round(x, 2) for x in v]}")


if __name__ == "__main__":
    main()
