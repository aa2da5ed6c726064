#!/usr/bin/env python3
"""tools/ncu_summary.py <tag>  -- turn the captures of tools/profile_all.sh (gpurun_out/<tag>_launches.csv,
<tag>_full_raw.csv, <tag>_meta.json) into the committed evidence:
  profiles/<tag>_kernel_table.txt     per-kernel shares of one proof + the key ncu metrics of every kernel's first launch
  profiles/<tag>_launches.csv, profiles/<tag>_ncu_full_raw.csv   (copies)
  profiles/accum_kernel_summary.json  DRAM traffic of the G1 accumulation stage of one full-density MSM, with the kernel_rev
                                      and configuration of the capture (bench.py reports it as roofline.traffic only when they
                                      match the running code)
"""
import collections
import csv
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    name = re.sub(r"g16::", "", name).split("(")[0].replace("void ", "")
    name = re.sub(r"Fp<(\w+?)_(F[qr])P>", r"\2", name)
    return re.sub(r"Fp2<\w+, \(int\)\d>", "Fq2", name)


def rows(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return lines


def main():
    tag = sys.argv[1]
    O = os.path.join(ROOT, "gpurun_out")
    out = []
    # ---- launch list ----
    r = csv.DictReader(rows(os.path.join(O, f"{tag}_launches.csv")))
    seq, agg, tot = [], collections.OrderedDict(), 0.0
    for row in r:
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        us = v / 1e3 if row["Metric Unit"] == "ns" else (v if row["Metric Unit"] in ("us", "usecond") else v * 1e3)
        n = short(row["Kernel Name"])
        seq.append((n, us, row["Grid Size"]))
        agg.setdefault(n, [0, 0.0])
        agg[n][0] += 1
        agg[n][1] += us
        tot += us
    out.append(f"# one proof, MSMs serialised (ncu --metrics gpu__time_duration.sum --clock-control none): {len(seq)} launches, {tot / 1e3:.2f} ms of kernel time")
    out.append("# (cold-cache, serialised: compare SHARES, not absolutes)\n")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{us / 1e3:8.3f} ms {100 * us / tot:5.1f} %  x{n:3d}  {k}")
    # ---- full metrics ----
    lines = rows(os.path.join(O, f"{tag}_full_raw.csv"))
    rd = csv.reader(lines)
    hdr, units = next(rd), next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [("gpu__time_duration.sum", "ms"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
            ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
            ("dram__bytes.sum.per_second", "dramTB/s"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
            ("lts__t_sector_hit_rate.pct", "L2hit%"),
            ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
            ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
            ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math")]
    cols = [(c, l) for c, l in cols if c in ix]
    out.append("\n# ncu --set full, first launch of every kernel (and the G1 accumulation stage of the first full-density MSM)\n")
    out.append(f"{'kernel':44s} " + " ".join(f"{l:>9s}" for _, l in cols))
    seen = collections.Counter()
    stage = {"dram": 0.0, "ms": 0.0, "on": False, "done": False}

    def to_bytes(v, u):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    def dur_ms(row):
        return float(row[ix["gpu__time_duration.sum"]].replace(",", "")) * {"us": 1e-3, "usecond": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[ix["gpu__time_duration.sum"]], 1)

    def dram_bytes(row):
        """dram__bytes_read.sum + dram__bytes_write.sum; the section-based capture (no --set full) only carries the rate:
        dram__bytes.sum.per_second x gpu__time_duration.sum"""
        if "dram__bytes_read.sum" in ix:
            return to_bytes(row[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + to_bytes(row[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        k = "dram__bytes.sum.per_second"
        rate = float(row[ix[k]].replace(",", "")) * {"byte/s": 1, "Kbyte/s": 1e3, "Mbyte/s": 1e6, "Gbyte/s": 1e9, "Tbyte/s": 1e12}[units[ix[k]]]
        return rate * dur_ms(row) * 1e-3

    for row in rd:
        n = short(row[ix["Kernel Name"]])
        seen[n] += 1
        # stage of the FIRST G1 MSM whose digits kernel is followed by ba_forward<Fq>: from ba_forward (or accum_l0) to accum_l0
        if not stage["done"]:
            if n in ("ba_forward_kernel<Fq>", "ba_forward_kernel<Fq, 0>") and not stage["on"] and seen[n] == 1:
                stage["on"] = True
            if stage["on"] and (n.startswith("ba_") or n.startswith("msm_accum_l0")) and "Fq2" not in n:
                stage["dram"] += dram_bytes(row)
                stage["ms"] += float(row[ix["gpu__time_duration.sum"]].replace(",", "")) * (1e-3 if units[ix["gpu__time_duration.sum"]] in ("us", "usecond") else 1)
                if n.startswith("msm_accum_l0"):
                    stage["done"] = True
        if seen[n] > 1:
            continue
        out.append(f"{n[:44]:44s} " + " ".join(f"{row[ix[c]][:9]:>9s}" for c, _ in cols))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"{tag}_kernel_table.txt"), "w") as f:
        f.write("\n".join(out) + "\n")
    shutil.copy(os.path.join(O, f"{tag}_launches.csv"), os.path.join(ROOT, "profiles", f"{tag}_launches.csv"))
    shutil.copy(os.path.join(O, f"{tag}_full_raw.csv"), os.path.join(ROOT, "profiles", f"{tag}_ncu_full_raw.csv"))
    meta_p = os.path.join(O, f"{tag}_meta.json")
    meta = json.load(open(meta_p)) if os.path.exists(meta_p) else {}
    summ = {"source": f"profiles/{tag}_ncu_full_raw.csv (ncu --set full --clock-control none, one proof, MSMs serialised)",
            "kernel_rev": meta.get("kernel_rev"), "config": meta.get("config"), "curve": meta.get("curve"), "log_n": meta.get("log_n"),
            "workload": "synthetic",
            "g1_dram_bytes_per_launch": stage["dram"] if stage["done"] else None, "g1_stage_ms_under_ncu": stage["ms"],
            "note": "DRAM bytes (read + write) summed over the kernels of the G1 accumulation stage (batched-affine rounds + "
                    "msm_accum_l0) of the first full-density G1 MSM of the proof"}
    with open(os.path.join(ROOT, "profiles", "accum_kernel_summary.json"), "w") as f:
        json.dump(summ, f, indent=1)
        f.write("\n")
    print("\n".join(out[:30]))
    print(json.dumps(summ, indent=1))


if __name__ == "__main__":
    main()
