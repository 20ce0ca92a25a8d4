"""Summarise ncu captures for profiles/.

    python profiles/summarize_ncu.py step  gpurun_out/r2_step_metrics.csv            > profiles/r2_step_dram.md
    python profiles/summarize_ncu.py full  NAME gpurun_out/r2_X_raw.csv [...]         -> profiles/r2_NAME_ncu_full.md + profiles/ncu_traffic.json

`step`: a launch list with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch (bench.py --no-graph, PDL off):
per-kernel totals of ONE training step.  `full`: the raw page of an `ncu --set full` capture: the metrics the roofline discussion needs.
"""
import collections
import csv
import hashlib
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows_of(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.DictReader(lines))


def num(v):
    try:
        return float(str(v).replace(",", ""))
    except ValueError:
        return float("nan")


def short(name):
    name = re.sub(r"\(.*", "", name)
    return re.sub(r"^void ", "", name).replace("hca::<unnamed>::", "")[:100]


def step(path):
    per = collections.OrderedDict()                  # launch id -> dict
    for r in rows_of(path):
        d = per.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        v = num(r["Metric Value"])
        unit = r["Metric Unit"]
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["us"] = v / 1000 if unit.startswith("n") else (v * 1000 if unit.startswith("m") else v)
        elif m.startswith("dram__bytes"):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            d["rd" if "read" in m else "wr"] = v * scale
    launches = list(per.values())
    emb = [i for i, l in enumerate(launches) if "embedding_fwd" in l["name"]]
    s, e = emb[-3], emb[-2]
    st = launches[s:e]
    agg = collections.OrderedDict()
    for l in st:
        a = agg.setdefault(short(l["name"]), [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += l.get("us", 0); a[2] += l.get("rd", 0); a[3] += l.get("wr", 0)
    tot = [sum(a[i] for a in agg.values()) for i in (1, 2, 3)]
    print(f"# One training step under ncu ({len(st)} launches): device time and DRAM traffic per kernel\n")
    print("`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none` around "
          "`bench.py --steps 2 --warmup 3 --no-graph` with programmatic dependent launch off; the step between two consecutive "
          "`embedding_fwd` launches.  Times under ncu are cold-cache and serialised: compare SHARES.\n")
    print(f"Step totals: {tot[0]:.0f} us, {tot[1] / 1e6:.0f} MB read, {tot[2] / 1e6:.0f} MB written "
          f"(algorithmic minimum of the step: ~250 MB, SURVEY section 8d).\n")
    print("| us | share | launches | DRAM read MB | DRAM write MB | GB/s | kernel |\n|---:|---:|---:|---:|---:|---:|---|")
    for k, (n, us, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = (rd + wr) / (us * 1e-6) / 1e9 if us > 0 else 0
        print(f"| {us:.1f} | {100 * us / tot[0]:.1f}% | {n} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbs:.0f} | `{k}` |")


WANT = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__cluster_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]


def full(name, paths):
    out = [f"# ncu --set full: {name}\n", "Captured with `ncu --set full --clock-control none --import-source on` under gpurun (one B200); raw pages read with "
           "`ncu -i X.ncu-rep --page raw --csv`.  One column per captured launch.\n"]
    traffic_file = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic = json.load(open(traffic_file)) if os.path.exists(traffic_file) else {"kernels": {}}
    for path in paths:
        rows = rows_of(path)
        if not rows:
            continue
        cols = rows[0].keys()
        units = rows[0] if all(not str(v).replace(".", "").isdigit() for v in list(rows[0].values())[:3]) else None
        launches = rows[1:] if rows and rows[0].get("ID", "") == "" else rows
        out.append(f"## {os.path.basename(path)}\n")
        out.append("| metric | " + " | ".join(short(l.get("Kernel Name", "?"))[:60] for l in launches) + " |")
        out.append("|---|" + "---:|" * len(launches))
        for m in WANT:
            if m in cols:
                u = rows[0].get(m, "") if launches is not rows else ""
                out.append(f"| `{m}` {('[' + u + ']') if u else ''} | " + " | ".join(str(l.get(m, "")) for l in launches) + " |")
        for l in launches:
            kn = l.get("Kernel Name", "")
            key = None
            if "lstm_rec_kernel<1>" in kn or "lstm_rec_kernel<true>" in kn or "lstm_rec_kernel<(bool)1>" in kn:
                key = "lstm_rec_bwd"
            elif "lstm_rec_kernel" in kn:
                key = "lstm_rec_fwd"
            elif "gemm_tc_kernel" in kn and name.startswith("wgrad"):
                key = "wgrad_dWv"
            elif "gemm_tc_kernel" in kn and name.startswith("pv"):
                key = "pv_proj"
            elif "hv_kernel" in kn:
                key = "hv_bwd" if ("<1>" in kn or "true" in kn or "(bool)1" in kn) else "hv_fwd"
            if key and "dram__bytes_read.sum" in l:
                def to_bytes(metric):
                    u = rows[0].get(metric, "byte") if launches is not rows else "byte"
                    return num(l[metric]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                traffic["kernels"][key] = {"dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
                                           "source": f"profiles/r2_{name}_ncu_full.md ({os.path.basename(path)})"}
        out.append("")
    open(os.path.join(ROOT, "profiles", f"r2_{name}_ncu_full.md"), "w").write("\n".join(out) + "\n")
    json.dump(traffic, open(traffic_file, "w"), indent=1, sort_keys=True)
    print("wrote", f"profiles/r2_{name}_ncu_full.md", "and profiles/ncu_traffic.json")


if __name__ == "__main__":
    if sys.argv[1] == "step":
        step(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3:])
