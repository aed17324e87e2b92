"""Summarises an .ncu-rep (read here, no GPU needed) into a small CSV for profiles/:
key raw metrics of the first captured launch + warp-stall samples split at the DMMA span."""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
lines = ["metric,unit,value"]
for h, u, v in zip(rr[0], rr[1], rr[2]):
    if h in KEEP or h in ("Kernel Name", "Block Size", "Grid Size") or "issue_stalled" in h and "_per_issue_active" in h:
        lines.append("%s,%s,%s" % (h, u, v.replace(",", ";")))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
if len(rows) > 3:
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    # a report with several kernels repeats the header block per kernel: keep the first kernel only
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr) or r == hdr:
            break
        data.append(r)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    dm = [i for i, r in enumerate(data) if "DMMA" in r[ix["Source"]]]
    spans = {"all": (0, len(data))}
    if dm:
        spans.update({"before_first_dmma": (0, dm[0]), "dmma_span": (dm[0], dm[-1] + 1),
                      "after_last_dmma": (dm[-1] + 1, len(data))})
    for name, (a, b) in spans.items():
        lines.append("samples[%s],count,%d" % (name, sum(int(r[ix["# Samples"]]) for r in data[a:b])))
        d = {s: sum(int(r[ix[s]]) for r in data[a:b]) for s in stalls}
        for s, v in sorted(d.items(), key=lambda x: -x[1])[:6]:
            lines.append("samples[%s].%s,count,%d" % (name, s, v))
    ops = {}
    for r in data:
        parts = r[ix["Source"]].split()
        op = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "")
        op = op.split(".")[0] + ("." + ".".join(op.split(".")[1:3]) if op.startswith(("DMMA", "UBLKCP", "UTMALDG", "LDS", "STG", "SYNCS")) else "")
        ops[op] = ops.get(op, 0) + int(r[ix["Instructions Executed"]] or 0)
    for op, v in sorted(ops.items(), key=lambda x: -x[1])[:14]:
        lines.append("inst_executed[%s],warp-inst,%d" % (op, v))
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out, len(lines), "lines")
