#!/usr/bin/env python
"""Turns `ncu -i <rep> --page raw --csv` into the summaries committed under profiles/:
    python profiles/ncu_summary.py <raw.csv> <tag>      -> profiles/<tag>_kernels.md, issue.json, traffic.json
One row per kernel launch of the capture (profiles/prof_step.py: one launch sequence of 256 pairs at 720x576, then bag of
words, text and MIP): duration, warp instructions, issue-slot and pipe utilisation, DRAM bytes, achieved occupancy.
issue.json / traffic.json hold, per kernel FAMILY as bench.py names them, the warp instructions and DRAM bytes of one launch
sequence - the figures bench.py cannot measure live and reports in `roofline_issue` / `roofline.traffic`."""
import csv
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
FAMILY = {"k_front": "front", "k_absdiff": "absdiff", "k_twothird": "pyramid", "k_half": "pyramid", "k_scores": "scores",
          "k_row_counts": "scan", "k_scan_blocks": "scan", "k_scan_partials": "scan", "k_scan_add": "scan", "k_fill_candidates": "fill_candidates",
          "k_integral_left": "integral", "k_integral_strips": "integral", "k_pa_test": "phase_a", "k_pa_above": "phase_a", "k_pa_below": "phase_a",
          "k_pa_emit": "phase_a", "k_phase_b": "phase_b", "k_keep_flags": "compact", "k_scatter_rows": "compact", "k_advance": "compact",
          "k_describe": "describe", "k_bow_assign": "bow_assign", "k_text_lengths": "text", "k_text_write": "text", "k_mip_only": "mip_only",
          "k_set_pairs": "set_pairs"}
COLS = [("us", "gpu__time_duration.sum"), ("warp_inst_M", "smsp__inst_executed.sum"), ("issue_pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        ("alu_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), ("fma_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        ("lsu_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"), ("xu_pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        ("shared_pipe_pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"), ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread")]


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def scale(value, unit, want):
    if value is None:
        return None
    f = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "inst": 1e-6}
    return value * f[unit] if want in ("us", "MB", "M") and unit in f else value


def main():
    raw, tag = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {k: (hdr.index(m) if m in hdr else None) for k, m in COLS}
    ik = hdr.index("Kernel Name")
    out, fam_inst, fam_dram, fam_us, seen_seq = [], {}, {}, {}, set()
    for r in data:
        km = re.search(r"(k_[a-z0-9_]+)", r[ik])
        name = km.group(1) if km else r[ik].split("(")[0]
        rec = {"kernel": name}
        for k, _ in COLS:
            i = idx[k]
            v = num(r[i]) if i is not None else None
            if k == "us":
                v = scale(v, units[i], "us")
            elif k == "warp_inst_M":
                v = scale(v, units[i], "M")
            elif k in ("dram_rd_MB", "dram_wr_MB"):
                v = scale(v, units[i], "MB")
            rec[k] = v
        out.append(rec)
        fam = FAMILY.get(name, name)
        # the capture holds ONE launch sequence; text/bow/mip follow it (their scans belong to `text`)
        if fam == "scan" and any(o["kernel"] == "k_bow_assign" for o in out):
            fam = "text"
        fam_inst[fam] = fam_inst.get(fam, 0.0) + (rec["warp_inst_M"] or 0.0) * 1e6
        fam_dram[fam] = fam_dram.get(fam, 0.0) + ((rec["dram_rd_MB"] or 0.0) + (rec["dram_wr_MB"] or 0.0)) * 1e6
        fam_us[fam] = fam_us.get(fam, 0.0) + (rec["us"] or 0.0)
    with open(os.path.join(HERE, tag + "_kernels.md"), "w") as f:
        f.write("# ncu --set full, one launch sequence of 256 pairs at 720x576 + BoW + text + MIP (%s)\n\n" % tag)
        f.write("| kernel | " + " | ".join(k for k, _ in COLS) + " |\n|---|" + "---|" * len(COLS) + "\n")
        for rec in out:
            f.write("| %s | " % rec["kernel"] + " | ".join("-" if rec[k] is None else ("%.1f" % rec[k]) for k, _ in COLS) + " |\n")
        seq = [fam for fam in fam_us if fam not in ("bow_assign", "text", "mip_only")]
        f.write("\nLaunch sequence total: %.0f us (serialised, cold caches), %.0f M warp instructions, %.0f MB DRAM traffic.\n" %
                (sum(fam_us[k] for k in seq), sum(fam_inst[k] for k in seq) / 1e6, sum(fam_dram[k] for k in seq) / 1e6))
    issue = {"capture": tag}
    for fam in fam_inst:
        pipes = {}
        for rec in out:
            if FAMILY.get(rec["kernel"], rec["kernel"]) == fam and (not pipes or rec["us"] > pipes.get("_us", 0)):
                pipes = {"_us": rec["us"], "kernel": rec["kernel"], "issue_pct": rec["issue_pct"], "alu_pct": rec["alu_pct"], "fma_pct": rec["fma_pct"],
                         "lsu_pct": rec["lsu_pct"], "shared_pipe_pct": rec["shared_pipe_pct"], "dram_pct": rec["dram_pct"]}
        pipes.pop("_us", None)
        issue[fam] = {"warp_instructions_per_launch": fam_inst[fam], "ncu_us": fam_us[fam], "pipes": pipes}
    json.dump(issue, open(os.path.join(HERE, "issue.json"), "w"), indent=1, sort_keys=True)
    traffic = {k: int(v) for k, v in fam_dram.items()}
    traffic["_source"] = "%s: dram__bytes_read.sum + dram__bytes_write.sum per launch sequence of 256 pairs (ncu --set full)" % tag
    json.dump(traffic, open(os.path.join(HERE, "traffic.json"), "w"), indent=1, sort_keys=True)
    print(open(os.path.join(HERE, tag + "_kernels.md")).read())


if __name__ == "__main__":
    main()
