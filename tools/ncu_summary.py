#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full --import-source on) into the text summary committed under profiles/:
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [n_hot_instructions] > profiles/rNN_ncu_x_summary.txt
Raw page: duration, clocks, DRAM bytes, pipe utilisation; source page: stall-reason totals, opcode histogram of the
executed SASS and the hottest instructions by stall samples."""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return out


def main():
    rep = sys.argv[1]
    n_hot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(io.StringIO(page(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    print(f"# ncu summary of {rep.split('/')[-1]}  (ncu --set full --clock-control none --import-source on; times under ncu are")
    print("# cold-cache and at whatever clock the replay ran -- compare shares and counters, not absolutes)")
    for k, row in enumerate(rows[2:]):
        name = row[hdr.index("Kernel Name")]
        print(f"\n== launch {k}: {name}")
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print(f"  {key:84s} {row[i]:>16s} {units[i]}")
    src = page(rep, "source")
    # the source page repeats a 2-line header per kernel launch; summarise the first launch
    lines = list(csv.reader(io.StringIO(src)))
    blocks, cur = [], None
    for ln in lines:
        if ln and ln[0] == "Kernel Name":
            cur = {"name": ln[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = ln
        elif cur is not None and ln:
            cur["rows"].append(ln)
    if not blocks:
        return
    b = blocks[0]
    h = b["hdr"]
    i_src, i_all, i_exec = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = collections.Counter()
    ops = collections.Counter()
    n_samples = 0
    for r in b["rows"]:
        try:
            ns, ne = int(r[i_all]), int(r[i_exec])
        except ValueError:
            continue
        n_samples += ns
        op = r[i_src].split()
        op = [o for o in op if not o.startswith("@")]
        if op:
            ops[op[0].split(".")[0]] += ne
        for i, c in stall_cols:
            try:
                tot[c] += int(r[i])
            except ValueError:
                pass
    print(f"\n== source page, launch 0 ({b['name']}): {n_samples} warp samples")
    print("  stall reasons (share of samples): " + ", ".join(f"{c[6:]} {100.0 * v / max(n_samples, 1):.1f}%" for c, v in tot.most_common(9)))
    tot_ops = sum(ops.values())
    print("  executed warp-instructions by opcode (top 24): " + ", ".join(f"{o} {100.0 * v / tot_ops:.1f}%" for o, v in ops.most_common(24)))
    print(f"  total executed warp-instructions: {tot_ops}")
    hot = sorted(b["rows"], key=lambda r: -int(r[i_all]) if r[i_all].isdigit() else 0)[:n_hot]
    print(f"  hottest {n_hot} instructions by samples:")
    for r in hot:
        st = sorted(((int(r[i]), c[6:]) for i, c in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
        print(f"    {int(r[i_all]):7d} ({100.0 * int(r[i_all]) / max(n_samples, 1):4.1f}%)  {r[i_src].strip():60s} " + ", ".join(f"{c} {v}" for v, c in st))


if __name__ == "__main__":
    main()
