#!/usr/bin/env python
"""Phase-level latency trace of the persistent sample loop (diagnostic; see vqcpc_debug_set_ar_trace).
usage: ar_trace.py TAG [first_poll_delay] [backoff] [B]
Writes gpurun_out/ar_trace_<tag>.json: cycles per phase of utterance 0's chain on two CTAs, and the measured
us/step for the whole generate call."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fixtures, vocoder as ovoc  # noqa: E402
from vectorquantizedcpc_b200 import Vocoder, _lib  # noqa: E402

PHASES = ["gates+publish_h+flag", "poll_h", "sts+fc1+publish_r", "flag+poll_r", "fc2+publish_o", "flag+poll_o",
          "softmax+sample"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "run"
    delay = int(sys.argv[2]) if len(sys.argv) > 2 else 600
    backoff = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    B = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    dev = torch.device("cuda:0")
    voc = Vocoder()
    voc.load_state_dict(ovoc.init_state_dict(seed=13))
    voc = voc.to(dev).eval()
    Tc = 50
    codes, spk, u = fixtures.vocoder_inputs(B, Tc, seed=0)
    lib = _lib.lib()
    lib.vqcpc_debug_set_ar_poll_gap(delay | (backoff << 12))
    n, t0 = 256, 4000
    out = {"first_poll_delay": delay, "backoff": backoff, "B": B}
    cd, sd, ud = codes.to(dev), spk.to(dev), u.to(dev)
    with torch.no_grad():
        voc.generate(cd, sd, uniforms=ud)   # warm
        for cta in (0, 77):
            buf = torch.zeros(n + 1, 8, dtype=torch.int64, device=dev)
            lib.vqcpc_debug_set_ar_trace(buf.data_ptr(), cta, t0, n + 1)
            voc.generate(cd, sd, uniforms=ud)
            lib.vqcpc_debug_set_ar_trace(None, 0, 0, 0)
            ts = buf.cpu().double()
            d = torch.empty(n, 7)
            d[:, 0] = ts[1:n + 1, 1] - ts[0:n, 7]
            for k in range(1, 7):
                d[:, k] = ts[1:n + 1, k + 1] - ts[1:n + 1, k]
            step = ts[1:n + 1, 7] - ts[0:n, 7]
            q = lambda v: [float(x) for x in torch.quantile(v.double(), torch.tensor([0.0, 0.5, 0.9, 1.0], dtype=torch.float64))]
            out[f"cta{cta}"] = {"cycles_per_step[min,med,p90,max]": q(step),
                                "phases[min,med,p90,max]": {p: q(d[:, i]) for i, p in enumerate(PHASES)}}
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        a.record()
        for _ in range(3):
            voc.generate(cd, sd, uniforms=ud)
        b.record()
    torch.cuda.synchronize()
    out["ms_per_generate"] = a.elapsed_time(b) / 3
    out["us_per_step"] = out["ms_per_generate"] * 1e3 / (320 * Tc)
    out["x_realtime_aggregate"] = B * 320 * Tc / 16000 / (out["ms_per_generate"] * 1e-3)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"ar_trace_{tag}.json"), "w"), indent=1)
    print(f"[{tag}] delay {delay} backoff {backoff} B {B}: {out['us_per_step']:.3f} us/step, "
          f"{out['x_realtime_aggregate']:.1f}x RT aggregate")
    for cta in ("cta77",):
        print("   step", [round(v) for v in out[cta]["cycles_per_step[min,med,p90,max]"]])
        print("   med ", {k: round(v[1]) for k, v in out[cta]["phases[min,med,p90,max]"].items()})
        print("   min ", {k: round(v[0]) for k, v in out[cta]["phases[min,med,p90,max]"].items()})


if __name__ == "__main__":
    main()
