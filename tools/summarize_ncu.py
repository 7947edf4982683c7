"""Turns the raw ncu CSV exports brought back from the GPU box (gpurun_out/r2_*_raw.csv, *_source.csv.gz) into the small
summaries committed under profiles/ (and profiles/cvf_profile_facts.json, which bench.py reads).  Usage:
python tools/summarize_ncu.py"""
import collections
import csv
import gzip
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw_rows(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1]
    out = []
    for r in rows[2:]:
        if len(r) == len(hdr):
            out.append({h: (v, u) for h, v, u in zip(hdr, r, units)})
    return out


def summarize(row):
    s = {"kernel": row["Kernel Name"][0]}
    for k in KEYS:
        if k in row:
            s[k] = f"{row[k][0]} {row[k][1]}".strip()
    for k, (v, u) in row.items():
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
            try:
                if float(v) >= 0.02:
                    s[k] = v
            except ValueError:
                pass
    return s


def to_bytes(v, u):
    f = float(v.replace(",", ""))
    return f * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)


def source_mix(path):
    rows = list(csv.reader(gzip.open(path, "rt")))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    mix = collections.Counter()
    stalls = collections.Counter()
    total = 0
    for r in rows[2:]:
        try:
            n = int(r[ix["Instructions Executed"]])
        except (ValueError, IndexError):
            continue
        t = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip())
        op = t.split()[0].split(".")[0] if t else "?"
        mix[op] += n
        total += n
        for k in hdr:
            if k.startswith("stall_") and "(" not in k:
                stalls[k] += int(r[ix[k]] or 0)
    return total, mix, stalls


def main():
    facts = {}
    for wl, mode, name in (("C4", 0, "exact"), ("C4", 1, "mixed"), ("C3", 0, "exact")):
        raw = os.path.join(OUT, f"r2_cvf_{wl}_mode{mode}_raw.csv")
        if not os.path.exists(raw):
            continue
        row = raw_rows(raw)[-1]
        s = summarize(row)
        src = os.path.join(OUT, f"r2_cvf_{wl}_mode{mode}_source.csv.gz")
        if os.path.exists(src):
            total, mix, stalls = source_mix(src)
            tot_s = sum(stalls.values()) or 1
            s["executed_warp_instructions"] = total
            s["instruction_mix_top"] = {k: v for k, v in mix.most_common(16)}
            s["warp_state_samples_pct"] = {k[6:]: round(100.0 * v / tot_s, 1) for k, v in stalls.most_common(10)}
        json.dump(s, open(os.path.join(PROF, f"r2_cvf_{wl}_mode{mode}_summary.json"), "w"), indent=1)
        dram = to_bytes(*row["dram__bytes_read.sum"]) + to_bytes(*row["dram__bytes_write.sum"])
        ws = s.get("warp_state_samples_pct", {})
        facts[f"{wl}:{name}"] = {
            "dram_bytes_per_launch": int(dram),
            "limiter": ("instruction issue / latency at 12 warps per SM (registers): issue-active "
                        f"{row['smsp__issue_active.avg.pct_of_peak_sustained_active'][0][:4]} %, no pipe above "
                        f"{max(float(row[k][0]) for k in KEYS if 'pipe_' in k and k in row):.0f} % busy "
                        f"(xu {float(row['sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'][0]):.0f} / fp64 "
                        f"{float(row['sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'][0]):.0f} / lsu "
                        f"{float(row['sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'][0]):.0f} %), DRAM "
                        f"{float(row['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'][0]):.0f} % of peak; warp-state samples: {ws}"),
            "issue_active_pct": float(row["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
            "executed_warp_instructions": int(float(row["smsp__inst_executed.sum"][0])),
            "warps_per_sm": round(float(row["sm__warps_active.avg.pct_of_peak_sustained_active"][0]) * 64 / 100, 1),
            "source": f"profiles/r2_cvf_{wl}_mode{mode}_summary.json (ncu --set full, one launch, gpurun)",
        }
    json.dump(facts, open(os.path.join(PROF, "cvf_profile_facts.json"), "w"), indent=1)
    other = os.path.join(OUT, "r2_other_kernels_raw.csv")
    if os.path.exists(other):
        json.dump([summarize(r) for r in raw_rows(other)], open(os.path.join(PROF, "r2_other_kernels_summary.json"), "w"), indent=1)
    # the other kernels of the frame: CVC builds (with their source pages) and the small per-frame kernels
    for tag in ("cvc_v0", "cvc_v1", "cvc_fast"):
        raw = os.path.join(OUT, f"r2_{tag}_raw.csv")
        if not os.path.exists(raw):
            continue
        s = summarize(raw_rows(raw)[-1])
        src = os.path.join(OUT, f"r2_{tag}_source.csv.gz")
        if os.path.exists(src):
            total, mix, stalls = source_mix(src)
            tot_s = sum(stalls.values()) or 1
            s["executed_warp_instructions"] = total
            s["instruction_mix_top"] = {k: v for k, v in mix.most_common(14)}
            s["warp_state_samples_pct"] = {k[6:]: round(100.0 * v / tot_s, 1) for k, v in stalls.most_common(10)}
        json.dump(s, open(os.path.join(PROF, f"r2_{tag}_summary.json"), "w"), indent=1)
    small = os.path.join(OUT, "r2_small_kernels_raw.csv")
    if os.path.exists(small):
        json.dump([summarize(r) for r in raw_rows(small)], open(os.path.join(PROF, "r2_small_kernels_summary.json"), "w"), indent=1)
    print(json.dumps(facts, indent=1)[:1500])


if __name__ == "__main__":
    main()
