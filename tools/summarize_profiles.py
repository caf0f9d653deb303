#!/usr/bin/env python
"""Turns ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_r01_v8.csv profiles/r01_v8_launches.md
    python tools/summarize_profiles.py full gpurun_out/prof_v8_wc2_c64.ncu-rep profiles/r01_v8_window_c2_c64.md
"""
import collections
import csv
import re
import subprocess
import sys


def short(name):
    m = re.search(r"(\w+(<[^>]*>)?)\(", name.replace("conan::", "").replace("<unnamed>::", "").replace("unnamed>::", ""))
    return (m.group(1) if m else name)[:70]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    byid = collections.OrderedDict()
    for r in rows:
        d = byid.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        v, u = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if u == "ns" else (v if u == "us" else v * 1e3)
        else:
            d[r["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    agg, tot = collections.OrderedDict(), 0.0
    for d in byid.values():
        a = agg.setdefault(short(d["name"]), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d["us"]
        a[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        tot += d["us"]
    with open(dst, "w") as f:
        f.write(f"# ncu launch list of ONE warmed-up chunk step (S = 1024 streams), source `{src}`\n\n")
        f.write("`ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                "--clock-control none python bench.py --ncu-step` (per-launch times are serialised / cold-cache: compare shares).\n\n")
        f.write(f"Total {tot / 1e3:.2f} ms over {len(byid)} launches.\n\n| kernel | launches | time (us) | share | DRAM bytes (GB) | DRAM GB/s |\n|---|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% | {v[2] / 1e9:.2f} | {v[2] / v[1] / 1e3 if v[1] else 0:.0f} |\n")
        f.write("\n## every launch, in order\n\n| # | kernel | grid | block | us | DRAM MB |\n|---|---|---|---|---|---|\n")
        for i, d in enumerate(byid.values()):
            b = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
            f.write(f"| {i} | `{short(d['name'])}` | {d['grid']} | {d['block']} | {d['us']:.1f} | {b / 1e6:.0f} |\n")
    print("wrote", dst)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__cycles_active.avg"]


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    srcp = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    srows = list(csv.reader(srcp.splitlines()))
    starts = [i for i, r in enumerate(srows) if r and r[0] == "Kernel Name"]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary, source `{src}`\n\n`ncu --set full --clock-control none --import-source on` on one launch inside a warmed-up "
                "chunk step (S = 1024 streams).\n\n")
        for kidx, r in enumerate(rows[2:]):
            f.write(f"## {short(r[idx['Kernel Name']])}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n\n| metric | value |\n|---|---|\n")
            for w in WANT:
                if w in idx:
                    f.write(f"| {w} | {r[idx[w]]} {units[idx[w]]} |\n")
            if kidx < len(starts):
                ks = starts[kidx]
                h = srows[ks + 1]
                hi = {c: i for i, c in enumerate(h)}
                body = []
                for rr in srows[ks + 2:]:
                    if rr and rr[0] == "Kernel Name":
                        break
                    if len(rr) == len(h):
                        body.append(rr)
                stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
                agg = {c: sum(int(x[hi[c]]) for x in body) for c in stall_cols}
                tot = sum(int(x[hi["# Samples"]]) for x in body) or 1
                f.write(f"\nSASS instructions {len(body)}, warp-state samples {tot}. Stall reasons (share of samples): ")
                f.write(", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]) + "\n\n")
                f.write("| samples | executed | SASS | top stalls |\n|---|---|---|---|\n")
                for x in sorted(body, key=lambda q: -int(q[hi["# Samples"]]))[:14]:
                    st = sorted(((c[6:], int(x[hi[c]])) for c in stall_cols if int(x[hi[c]]) > 0), key=lambda kv: -kv[1])[:2]
                    f.write(f"| {x[hi['# Samples']]} | {x[hi['Instructions Executed']]} | `{x[hi['Source']].strip()[:70]}` | {st} |\n")
            f.write("\n")
    print("wrote", dst)


def traffic(src, dst):
    """ncu launch list (gpu__time_duration + dram__bytes_read/write per launch) -> JSON {kernel family: measured DRAM bytes per
    launch}, read by bench.py for the `roofline.traffic` field."""
    import json
    lines = [l for l in open(src) if not l.startswith("==")]
    byid = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = byid.setdefault(r["ID"], {"name": r["Kernel Name"]})
        v, u = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
        if r["Metric Name"] != "gpu__time_duration.sum":
            d[r["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    fam = collections.OrderedDict()
    for d in byid.values():
        n = d["name"]
        key = ("resblock_fused_kernel" if "resblock_fused" in n else "ffn_fused_kernel" if "ffn_fused" in n else
               "block_fused_kernel" if "block_fused" in n else
               "conv_window_tc_kernel" if "conv_window_tc" in n else
               # fp16-operand vocoder layers: the CTA-pair kernel (conv_gemm_tc2_kernel) and the single-CTA ring variants whose last
               # template argument (SP, four-tile split stages) is 0; SP = 1 marks a split-fp16 Emformer / Conan GEMM
               ("conv_gemm_tc_kernel (" if ("conv_gemm_tc2" in n or re.search(r"0\)?>", n.split("(")[0] if "<" in n.split("(")[0] else n)) else "conv_gemm_tc_kernel, split")
               if "conv_gemm_tc" in n else
               "conv_gemm_ffma_kernel" if "conv_gemm_ffma" in n else None)
        if key is None:
            continue
        a = fam.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    out = {k: {"launches": v[0], "dram_bytes_per_launch": v[1] / v[0]} for k, v in fam.items()}
    out["_source"] = src
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
