#!/usr/bin/env python
"""Turn an `ncu --set full` capture of one resident bench step into the tracked summaries under profiles/:

    python profiles/summarize_ncu.py gpurun_out/<name>.ncu-rep profiles/<out>.md profiles/r01_traffic.json

Reads the report with `ncu -i ... --page raw --csv` (ncu is in the image; no GPU needed) and writes a per-launch table
(time, DRAM bytes, tensor-pipe / tensor-core shared-memory pipe / DRAM utilisation, issue slots) plus the per-step DRAM
traffic that bench.py's `roofline.traffic` fields quote."""
import csv, io, json, subprocess, sys

rep, out_md, out_json = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, name, scale=1.0):
    try:
        return float(r[ix[name]].replace(",", "")) * scale
    except Exception:
        return float("nan")


def unit_scale(name, want):
    u = units[ix[name]].lower()
    table = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    return table.get(u, 1.0) / (1e6 if want == "MB" else 1.0)


lines, conv_bytes, s1_bytes, tot_ms, conv_ms = [], 0.0, 0.0, 0.0, 0.0
for k, r in enumerate(data):
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    ms = f(r, "gpu__time_duration.sum", unit_scale("gpu__time_duration.sum", "ms"))
    rd = f(r, "dram__bytes_read.sum", unit_scale("dram__bytes_read.sum", "MB"))
    wr = f(r, "dram__bytes_write.sum", unit_scale("dram__bytes_write.sum", "MB"))
    tensor = f(r, "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active") if "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active" in ix else float("nan")
    if tensor != tensor:
        for cand in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"):
            if cand in ix:
                tensor = f(r, cand)
                break
    tcsm = f(r, "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")
    dram = float("nan")
    for cand in ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"):
        if cand in ix and f(r, cand) == f(r, cand):
            dram = f(r, cand)
            break
    issue = f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    regs = r[ix["launch__registers_per_thread"]]
    lines.append(f"| {k} | `{name}` | {ms:.3f} | {rd:.1f} | {wr:.1f} | {tensor:.1f} | {tcsm:.1f} | {dram:.1f} | {issue:.1f} | {regs} |")
    tot_ms += ms
    if "conv" in name:
        conv_bytes += (rd + wr) * 1e6
        conv_ms += ms
    elif "avgpool" not in name:
        s1_bytes += (rd + wr) * 1e6
with open(out_md, "w") as fh:
    fh.write("| # | kernel | time (ms) | DRAM read (MB) | DRAM write (MB) | tensor pipe active % | tensor-core smem pipe % | DRAM throughput % | issue slots % | regs |\n")
    fh.write("|---|---|---|---|---|---|---|---|---|---|\n")
    fh.write("\n".join(lines) + "\n\n")
    fh.write(f"Captured time {tot_ms:.2f} ms: conv kernels {100 * conv_ms / tot_ms:.0f} %, other kernels {100 * (1 - conv_ms / tot_ms):.0f} %.  "
             f"DRAM traffic per step: conv stack {conv_bytes / 1e9:.2f} GB, stage 1 {s1_bytes / 1e9:.2f} GB.\n")
json.dump({"source": rep.split("/")[-1], "conv_dram_bytes_per_step": conv_bytes, "stage1_dram_bytes_per_step": s1_bytes,
           "conv_launches": sum(1 for r in data if "conv" in r[ix["Kernel Name"]])}, open(out_json, "w"), indent=1)
print(open(out_md).read())
