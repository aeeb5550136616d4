"""Summarise an .ncu-rep (read on the build box): key raw metrics + top stall sources by SASS/source line."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__inst_executed_pipe_uniform.sum"]
d = dict(zip(hdr, vals))
print("kernel:", d.get("Kernel Name", "?")[:80])
for w in want:
    if w in d:
        print(f"  {w} = {d[w]}")
for k, v in d.items():
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
        try:
            if float(v) > 0.3:
                print(f"  {k.split('issue_stalled_')[1][:-len('_per_issue_active.ratio')]:24s} {float(v):.2f}")
        except ValueError:
            pass
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[0]
    idx = {n: i for i, n in enumerate(h)}
    key = [n for n in h if "Warp Stall Sampling (All" in n or n == "# Samples" or "Sampling" in n][:1]
    print("columns:", h[:12])
    samp = None
    for n in h:
        if "Stall Sampling (All" in n:
            samp = idx[n]
            break
    if samp is not None:
        body = [r for r in rows[1:] if len(r) > samp and r[samp].replace('.', '').isdigit()]
        body.sort(key=lambda r: -float(r[samp]))
        tot = sum(float(r[samp]) for r in body)
        for r in body[:int(sys.argv[2])]:
            print(f"  {float(r[samp]) / tot * 100:5.1f}%  {r[idx.get('Source', 1)][:110]}")
