#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, without a GPU): headline metrics, stall mix and the SASS
regions that collect the stall samples.   python profiles/ncu_summary.py gpurun_out/X.ncu-rep [kernel-regex]"""
import csv, io, subprocess, sys

def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))

def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
            "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
            "sm__cycles_elapsed.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
            "sm__inst_executed_pipe_lsu.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum"]
    for r in data:
        print("== kernel:", r[col["Kernel Name"]][:90])
        for w in want:
            if w in col:
                print(f"  {w:70s} {r[col[w]]:>16s} {units[col[w]]}")
        print("  -- stalls per issue-active (ratio)")
        for h, i in col.items():
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                v = float(r[i] or 0)
                if v > 0.05:
                    print(f"     {h.split('stalled_')[1].replace('_per_issue_active.ratio', ''):28s} {v:6.2f}")
    rows = page(rep, "source")
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not hi:
        return
    hdr = rows[hi[0]]; data = rows[hi[0] + 1:]
    end = next((i for i, r in enumerate(data) if r and r[0] == "Kernel Name"), len(data))
    data = data[:end]
    iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    names = ["stall_long_sb", "stall_barrier", "stall_short_sb", "stall_mio", "stall_wait", "stall_math", "stall_lg", "stall_not_selected", "stall_membar", "stall_branch_resolving"]
    idx = {n: hdr.index(n) for n in names if n in hdr}
    tot = sum(int(r[iS]) for r in data) or 1
    toti = sum(int(r[iI]) for r in data) or 1
    print(f"  -- SASS: {len(data)} instructions, {tot} samples, {toti/1e6:.2f}M warp-instr executed")
    step = max(60, len(data) // 24)
    for b in range(0, len(data), step):
        blk = data[b:b + step]
        s = sum(int(r[iS]) for r in blk); ie = sum(int(r[iI]) for r in blk)
        parts = " ".join(f"{n[6:]}={100*sum(int(r[i]) for r in blk)/tot:4.1f}" for n, i in idx.items() if sum(int(r[i]) for r in blk) / tot > 0.004)
        ops = {}
        for r in blk:
            t = r[iSrc].split()
            op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
            ops[op] = ops.get(op, 0) + 1
        top = ",".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:4])
        print(f"   [{b:5d}] samples {100*s/tot:5.1f}%  inst {100*ie/toti:5.1f}%  {parts}  | {top}")

if __name__ == "__main__":
    main()
