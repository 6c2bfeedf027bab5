#!/usr/bin/env python
"""Turns the raw captures in gpurun_out/ into the committed summaries under profiles/.

    python tools/make_profiles.py <kernels.ncu-rep | raw-page.csv> <launches.csv> [round-prefix]

  profiles/<r>_kernels_ncu.md          one row per profiled kernel launch (--set full) + its top warp-stall reasons
  profiles/<r>_roofline_traffic.json   DRAM bytes of the stage-1 window-attention launch (bench.py's roofline.traffic)
  profiles/<r>_bench_launches.csv      the launch list as captured
  profiles/<r>_bench_launches_summary.md  aggregated per kernel
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, launches = sys.argv[1], sys.argv[2]
rnd = sys.argv[3] if len(sys.argv) > 3 else "r1"
out_dir = os.path.join(ROOT, "profiles")

if rep.endswith(".csv"):      # already exported on the GPU box: ncu -i <rep> --page raw --csv
    raw = open(rep).read()
else:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
cols = [
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_sectors_op_read.sum", "L2 read sectors"), ("lts__t_sectors_op_write.sum", "L2 write sectors"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu (mufu) %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
]
cols = [(c, n) for c, n in cols if c in col]
stall_cols = [h for h in hdr if "warps_issue_stalled" in h and "not_issued" not in h and h.endswith(".sum")] or \
             [h for h in hdr if "warps_issue_stalled" in h and "not_issued" not in h]


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def to_bytes(v, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return num(v) * scale.get(unit, 1.0)


lines = [f"# ncu --set full, one launch per kernel ({os.path.basename(rep)}; B = 32, 1024^2 geometry, bf16; cold caches, serialised)", "",
         "Launch order follows tools/prof_kernels.py.  `stalls` = the five largest warp-stall reasons (share of stall samples).", "",
         "| kernel | " + " | ".join(f"{n} [{units[col[c]]}]" if units[col[c]] else n for c, n in cols) + " | stalls |",
         "|---|" + "---|" * (len(cols) + 1)]
traffic = None
tensor_pct = {}
traffic_by_kernel = {}
for r in data:
    name = r[col["Kernel Name"]].replace("void ", "").replace("sodt::<unnamed>::", "").split("(")[0]
    cells = []
    for c, _ in cols:
        v = num(r[col[c]])
        cells.append(r[col[c]] if v is None else f"{v:.4g}")
    st = sorted(((num(r[col[h]]) or 0.0, h.split("issue_stalled_")[-1].split(".")[0].replace("_per_warp_active", "")) for h in stall_cols),
                reverse=True)
    tot = sum(v for v, _ in st) or 1.0
    stalls = ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in st[:5])
    lines.append(f"| `{name[:60]}` | " + " | ".join(cells) + f" | {stalls} |")
    tp = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    if any(kname in name for kname in ("window_attn_win8_kernel<16>", "attn_block_kernel", "mlp_tc_kernel")) and name.split("::")[-1] not in traffic_by_kernel:
        traffic_by_kernel[name.split("::")[-1]] = (to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) +
                                                   to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]]))
    if tp in col and ("window_attn" in name or "cattn" in name or "attn_block" in name) and "prep" not in name:
        tensor_pct.setdefault(name.split("::")[-1], num(r[col[tp]]))
    if traffic is None and "window_attn_win8_kernel<16>" in name:
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        traffic = {"kernel": name, "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes": rd + wr,
                   "time_ms_under_ncu": num(r[col["gpu__time_duration.sum"]]),
                   "source": f"ncu --set full --clock-control none, {os.path.basename(rep)}, first stage-1 launch (B=32, 256x256 tokens, C=192)"}
open(os.path.join(out_dir, f"{rnd}_kernels_ncu.md"), "w").write("\n".join(lines) + "\n")
if traffic:
    import hashlib
    traffic["traffic_by_kernel"] = traffic_by_kernel        # DRAM bytes of one launch of the kernels bench.py reports a roofline on
    traffic["attention_tensor_pipe_pct"] = tensor_pct       # BASELINE.json's "attn tensor-pipe %" (sm__pipe_tensor_cycles_active)
    # stamp: bench.py reports these numbers only while the kernel sources are the ones that were profiled
    csrc = os.path.join(ROOT, "small-object-detection-transformers_b200", "csrc")
    traffic["source_sha256"] = {f: hashlib.sha256(open(os.path.join(csrc, f), "rb").read()).hexdigest()
                                for f in ("window_attn_win8.cu", "window_attn_flash.cu", "cattn.cu", "attn_block.cu", "mlp_tc.cu", "tc05.cuh", "tma.cuh")}
    json.dump(traffic, open(os.path.join(out_dir, f"{rnd}_roofline_traffic.json"), "w"), indent=1)
# trim the launch list to whole steps: from the first front-end launch to the last one (exclusive)
lrows = open(launches).read().splitlines()
hi = next(i for i, l in enumerate(lrows) if l.startswith('"ID"'))
body = lrows[hi + 1:]
fe = [i for i, l in enumerate(body) if "frontend" in l and "kernel" in l]
if len(fe) >= 2:
    body = body[fe[0]:fe[-1]]
launches_trimmed = os.path.join(out_dir, f"{rnd}_bench_launches.csv")
open(launches_trimmed, "w").write("\n".join(lrows[:hi + 1] + body) + "\n")
n_steps = max(1, len(fe) - 1)
launches = launches_trimmed
summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), launches], capture_output=True, text=True).stdout
open(os.path.join(out_dir, f"{rnd}_bench_launches_summary.md"), "w").write(
    f"# Kernel launches of {n_steps} eager bench step(s), trimmed to whole steps (`bench.py --steps 2 --warmup 3 --no-graph` under "
    "`ncu --metrics gpu__time_duration.sum --clock-control none`)\n\nPer-launch times are cold-cache and serialised: compare shares.\n\n" + summ)
print("\n".join(lines[:3]))
print(traffic)
