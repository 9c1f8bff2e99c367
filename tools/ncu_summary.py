"""Summarise an .ncu-rep (one kernel launch, --set full) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.txt [profiles/name.json tiles_in_launch "capture note" [cells_in_launch]]
The optional JSON carries the raw counters bench.py turns into pipe-utilisation figures (roofline.executed)."""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, val = rows[0], rows[1], rows[2]
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed.sum.per_cycle_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
lines = []
for h, u, v in zip(hdr, units, val):
    if h in WANT or h.startswith("smsp__average_warps_issue_stalled") and "not_issued" not in h:
        lines.append(f"{h:80s} {v:>22s} {u}")
rows = list(csv.reader(src.splitlines()))
shdr, data = rows[1], rows[2:]
I, T, S = shdr.index("Instructions Executed"), shdr.index("Thread Instructions Executed"), shdr.index("# Samples")
tot = sum(int(r[I]) for r in data)
tott = sum(int(r[T]) for r in data)
tots = max(1, sum(int(r[S]) for r in data))
lines.append("")
lines.append(f"SASS instructions executed: {tot}   average active threads per instruction: {tott / tot:.2f}")
ops = {}
for r in data:
    op = r[1].strip().split()
    op = [x for x in op if not x.startswith("@")]
    name = op[0].rstrip(";") if op else "?"
    ops[name] = ops.get(name, 0) + int(r[I])
lines.append("executed instruction mix (top 24):")
for name, c in sorted(ops.items(), key=lambda kv: -kv[1])[:24]:
    lines.append(f"  {name:28s} {c:14d}  {100.0 * c / tot:6.2f} %")
open(out, "w").write("\n".join(lines) + "\n")
print(open(out).read())
if len(sys.argv) > 4:
    JW = WANT + ["sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
                 "sm__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_active.avg", "gpc__cycles_elapsed.max"]
    rec = {"tiles": int(sys.argv[4]), "capture": sys.argv[5] if len(sys.argv) > 5 else rep, "unit_of": {}}
    if len(sys.argv) > 6:
        rec["cells"] = int(sys.argv[6])
    for h, u, v in zip(hdr, units, val):
        if h in JW:
            try:
                rec[h] = float(v.replace(",", ""))
                # ncu prints byte counters in scaled units
                scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u)
                if scale:
                    rec[h] *= scale
                rec["unit_of"][h] = u
            except ValueError:
                rec[h] = v
    # pipe attribution of the executed SASS mix (warp instructions): ALU pipe = logic / min-max / permute / compare / add,
    # FMA pipe = IMAD family
    ALU = ("LOP3", "VIADDMNMX", "VIMNMX3", "VIMNMX", "PRMT", "IADD3", "SEL", "VIADD", "ISETP", "LEA", "SHF", "PLOP3", "HSET2", "IABS", "IADD", "LOP", "BREV", "FLO", "POPC")
    rec["alu_pipe_warp_inst"] = sum(c for k, c in ops.items() if k.split(".")[0] in ALU)
    rec["fma_pipe_warp_inst"] = sum(c for k, c in ops.items() if k.split(".")[0] == "IMAD")
    rec["sass_instructions_executed"] = tot
    rec["instruction_mix"] = {k: c for k, c in sorted(ops.items(), key=lambda kv: -kv[1])[:24]}
    json.dump(rec, open(sys.argv[3], "w"), indent=1)
