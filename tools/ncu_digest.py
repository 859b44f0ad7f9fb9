#!/usr/bin/env python
"""Digest of `ncu --page raw --csv` (+ optional `--page source --csv`): key metrics per launch, top stall sites.

    python tools/ncu_digest.py raw.csv [src.csv [kernel_index]]
"""
import csv, sys, re
KEYS = [("gpu__time_duration.sum", "us"), ("sm__inst_executed_pipe_tensor_op_umma.avg.pct_of_peak_sustained_active", "umma_inst%"),
        ("sm__pipe_tensor_subpipe_tf32_cycles_active.avg.pct_of_peak_sustained_active", "tensor_tf32%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lsu_smem%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("lts__t_bytes.sum", "l2_bytes"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_thr%"),
        ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs")]
def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(hdr)}
    tensor_cols = [n for n in hdr if "tensor" in n and "pct" in n]
    out = []
    for r in rows[2:]:
        d = {"name": re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void mdgan::", ""), "grid": r[idx.get("Grid Size", 0)]}
        for k, short in KEYS:
            if k in idx: d[short] = r[idx[k]] + (" " + units[idx[k]] if short in ("us", "dram_rd", "dram_wr", "l2_bytes") else "")
        d["_tensor_cols"] = {n: r[idx[n]] for n in tensor_cols}
        out.append(d)
    return out
def src(path, which):
    rows = list(csv.reader(open(path)))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name": cur = {"name": r[1], "hdr": None, "rows": []}; secs.append(cur)
        elif r and r[0] == "Address": cur["hdr"] = r
        elif cur and cur["hdr"] and len(r) >= len(cur["hdr"]) - 2: cur["rows"].append(r)
    sec = secs[which]
    h = sec["hdr"]; idx = {n: i for i, n in enumerate(h)}
    tot = sum(int(r[idx["# Samples"]] or 0) for r in sec["rows"]) or 1
    stall = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    print(f"-- source hot spots of launch {which}: {sec['name'][:90]}  ({tot} samples, {len(secs)} launches in file)")
    agg = {c: sum(int(r[idx[c]] or 0) for r in sec["rows"]) for c in stall}
    print("   stall totals:", ", ".join(f"{c[6:]} {100*v/tot:.0f}%" for c, v in sorted(agg.items(), key=lambda t: -t[1])[:8]))
    for r in sorted(sec["rows"], key=lambda r: -int(r[idx["# Samples"]] or 0))[:18]:
        s = int(r[idx["# Samples"]] or 0)
        st = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall), reverse=True)[:2]
        print(f"   {100*s/tot:5.1f}%  {r[idx['Source']][:84]:84s} {st}")
if __name__ == "__main__":
    for i, d in enumerate(raw(sys.argv[1])):
        t = d.pop("_tensor_cols")
        print(i, {k: v for k, v in d.items()})
        if i == 0: print("   tensor metric names:", list(t)[:8])
    if len(sys.argv) > 2:
        src(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
