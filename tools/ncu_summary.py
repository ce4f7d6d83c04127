"""Summarise an .ncu-rep (raw page + source page) into the handful of numbers used in profiles/*.md.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("----")
    d = dict(zip(hdr, r))
    for w in WANT + [h for h in hdr if "pipe_tensor" in h and h.endswith("pct_of_peak_sustained_active")]:
        if w in d:
            print(f"{w} = {d[w]} {units[hdr.index(w)]}")
    try:
        wf = float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"].replace(",", ""))
        cyc = float(d["sm__cycles_elapsed.max"].replace(",", ""))
        print(f"shared wavefronts per SM-cycle = {wf / (148 * cyc):.3f}")
    except Exception:
        pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if hi:
    sec = rows[hi[0] + 1: (hi[1] - 1 if len(hi) > 1 else None)]
    hdr = rows[hi[0]]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = 0
    by = collections.Counter(); wf = collections.Counter(); smp = collections.Counter(); stall = collections.Counter()
    for r in sec:
        if len(r) < len(hdr):
            continue
        s = r[ix["Source"]].split()
        op = s[1] if s[0].startswith("@") else s[0]
        key = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG")) else op.split(".")[0]
        n = int(r[ix["Instructions Executed"]])
        by[key] += n; tot += n
        wf[key] += int(r[ix["L1 Wavefronts Shared"]]) if "L1 Wavefronts Shared" in ix else 0   # column absent when a kernel uses no shared memory
        smp[key] += int(r[ix["# Samples"]])
        for k in ix:
            if k.startswith("stall_") and "Not Issued" not in k:
                stall[k] += int(r[ix[k]])
    print("---- per-opcode (first kernel): instructions, share, shared wavefronts, stall samples")
    for k, n in by.most_common(12):
        print(f"{k:12s} {n:>13d} {100 * n / tot:5.1f}%  wf {wf[k]:>13d}  samples {smp[k]}")
    st = sum(stall.values())
    print("stalls:", ", ".join(f"{k[6:]} {100 * v / st:.1f}%" for k, v in stall.most_common(8)))
