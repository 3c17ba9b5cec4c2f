#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, on the CPU box) into JSON + markdown.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep --name ensemble_decomp --spectra 296 \
        --scale-spectra 12500 --out profiles/r01_ensemble_decomp

Writes <out>.json / <out>.md and updates profiles/ncu_summary.json (read by bench.py for
roofline.traffic).  Also aggregates the per-instruction stall samples of the source page by SASS
region so the hot loop / barrier / serial-phase split is visible without the GUI.
"""
import argparse
import csv
import io
import json
import os
import subprocess

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active": "dmma_pipe_pct",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefront_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "launch__registers_per_thread": "registers",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__cycles_elapsed.avg.per_second": "sm_ghz",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
}
STALLS = ["barrier", "math_pipe_throttle", "wait", "short_scoreboard", "long_scoreboard", "not_selected",
          "selected", "mio_throttle", "dispatch_stall", "branch_resolving", "no_instruction"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9,
        "Kbyte/block": 1e3, "byte/block": 1.0}


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--name", required=True)
    ap.add_argument("--spectra", type=int, default=0)
    ap.add_argument("--scale-spectra", type=int, default=0)
    ap.add_argument("--out", required=True)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    rows = ncu_csv(a.rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {"kernel": vals[hdr.index("Kernel Name")], "report": os.path.basename(a.rep), "note": a.note}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            x = float(v.replace(",", ""))
            if KEYS[h] in ("duration", "dram_read", "dram_write", "smem_dynamic"):
                x *= UNIT.get(u, 1.0)
            m[KEYS[h]] = x
        for s in STALLS:
            if h == f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio":
                m.setdefault("stall_per_issue", {})[s] = float(v)
    m["dram_bytes_per_launch"] = m.get("dram_read", 0) + m.get("dram_write", 0)
    if a.spectra and a.scale_spectra:
        m["spectra_in_capture"] = a.spectra
        m["dram_bytes_per_launch_at_bench_size"] = m["dram_bytes_per_launch"] / a.spectra * a.scale_spectra
    # ---- source page: stall samples by SASS region ------------------------------------------------
    src = ncu_csv(a.rep, "source")
    sh, data = src[1], src[2:]
    isrc, isamp, iex = sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
    cols = {k: sh.index(k) for k in ("stall_barrier", "stall_math", "stall_wait", "stall_short_sb", "stall_not_selected")}
    tot = sum(int(r[isamp]) for r in data) or 1
    regions, cur = [], None
    for r in data:
        ex = int(r[iex])
        if cur is None or abs(ex - cur["ex"]) > 0.02 * max(ex, cur["ex"], 1):
            cur = {"ex": ex, "n": 0, "samples": 0, "dmma": 0, **{k: 0 for k in cols}}
            regions.append(cur)
        cur["n"] += 1
        cur["samples"] += int(r[isamp])
        cur["dmma"] += "DMMA" in r[isrc]
        for k, i in cols.items():
            cur[k] += int(r[i])
    m["regions"] = [dict(instrs=g["n"], executed_per_instr=g["ex"], dmma_instrs=g["dmma"],
                         pct_of_samples=round(100 * g["samples"] / tot, 2),
                         **{k: round(100 * g[k] / tot, 2) for k in cols})
                    for g in regions if g["samples"] > 0.005 * tot]
    # ---- executed warp instructions by class (issue-slot budget) ------------------------------------
    def cls(text):
        t = text.split()
        if not t:
            return "other"
        op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
        if op == "DMMA":
            return "dmma"
        if op in ("DFMA", "DMUL", "DADD", "DSETP"):
            return "fp64"
        if op in ("FFMA", "FMUL", "FADD", "FSETP", "FSEL", "FMNMX", "HMMA"):
            return "fp32"
        if op == "MUFU":
            return "mufu"
        if op in ("I2F", "F2I", "F2F", "I2FP", "F2FP"):
            return "convert"
        if op in ("LDS", "STS", "LDSM", "ATOMS"):
            return "shared_mem"
        if op in ("LDG", "STG", "LD", "ST", "LDL", "STL", "ATOMG", "RED", "LDC", "LDCU"):
            return "other_mem"
        if op in ("BAR", "BSSY", "BSYNC", "BRA", "EXIT", "CALL", "RET", "WARPSYNC", "NANOSLEEP"):
            return "control"
        if op in ("SHFL", "VOTE", "MATCH"):
            return "shuffle"
        return "integer"
    mix, tot_ex = {}, 0
    for r in data:
        ex = int(r[iex])
        tot_ex += ex
        k = cls(r[isrc])
        mix[k] = mix.get(k, 0) + ex
    m["warp_instructions"] = tot_ex
    m["instruction_mix_pct"] = {k: round(100.0 * v / max(tot_ex, 1), 2) for k, v in sorted(mix.items(), key=lambda kv: -kv[1])}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(m, open(a.out + ".json", "w"), indent=1)
    with open(a.out + ".md", "w") as f:
        f.write(f"# ncu summary — {a.name}\n\n`{m['kernel']}`  (report {m['report']}; {a.note})\n\n")
        f.write("| metric | value |\n|---|---|\n")
        for k, v in m.items():
            if k not in ("regions", "stall_per_issue", "kernel", "report", "note", "instruction_mix_pct"):
                f.write(f"| {k} | {v:.6g} |\n" if isinstance(v, float) else f"| {k} | {v} |\n")
        f.write("\n## warp stall reasons (warps stalled per issue-active cycle)\n\n| reason | ratio |\n|---|---|\n")
        for k, v in sorted(m.get("stall_per_issue", {}).items(), key=lambda kv: -kv[1]):
            f.write(f"| {k} | {v:.3f} |\n")
        f.write("\n## executed warp instructions by class (% of all)\n\n| class | % |\n|---|---|\n")
        for k, v in m["instruction_mix_pct"].items():
            f.write(f"| {k} | {v} |\n")
        f.write("\n## SASS regions (consecutive instructions with equal execution count), % of all warp samples\n\n")
        f.write("| instrs | exec/instr | DMMA instrs | samples % | barrier % | math-pipe % | wait % | short-sb % |\n|---|---|---|---|---|---|---|---|\n")
        for g in m["regions"]:
            f.write(f"| {g['instrs']} | {g['executed_per_instr']} | {g['dmma_instrs']} | {g['pct_of_samples']} | "
                    f"{g['stall_barrier']} | {g['stall_math']} | {g['stall_wait']} | {g['stall_short_sb']} |\n")
    summ_path = os.path.join(os.path.dirname(a.out), "ncu_summary.json")
    summ = json.load(open(summ_path)) if os.path.exists(summ_path) else {}
    keep = {k: summ.get(a.name, {}).get(k) for k in ("dram_bytes_per_launch_at_bench_size", "dram_bytes_source", "kernel_sources_sha")}
    summ[a.name] = {k: v for k, v in m.items() if k != "regions"}
    for k, v in keep.items():          # bench-size DRAM figure and the source hash it belongs to (bench.py quotes it only while they match)
        if v is not None and (k not in summ[a.name] or keep.get("dram_bytes_source")):   # a measured bench-size figure wins over the scaled capture
            summ[a.name][k] = v
    json.dump(summ, open(summ_path, "w"), indent=1)
    print(json.dumps({k: v for k, v in m.items() if k != "regions"}, indent=1))


if __name__ == "__main__":
    main()
