"""Summarise an .ncu-rep (raw page): the metrics DESIGN.md / profiles/ quote.  usage: ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__average_warps_issue_stalled_"
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    for k in KEYS:
        if k in d:
            print(f"{k:75s} {d[k]:>22s} {u.get(k,'')}")
    st = sorted(((float(v), k[len(STALL):].replace('_per_issue_active.ratio', '')) for k, v in d.items() if k.startswith(STALL) and v not in ("", "n/a")), reverse=True)
    print("stall reasons (warps per issue-active cycle):", ", ".join(f"{n}={v:.2f}" for v, n in st[:9]))
    print()

# ---- dynamic instruction mix (source page, needs --import-source on / SASS): executed warp instructions per opcode, and the
# share of the ALU-pipe opcodes that is VABSDIFF4 (what separates the ALU pipe's utilisation from the roofline fraction)
ALU_PIPE = {"VABSDIFF4", "SHF", "PRMT", "LOP3", "IADD3", "VIADD", "ISETP", "VIMNMX", "VIMNMX3", "LEA", "SEL", "IABS", "IMNMX", "FMNMX",
            "PLOP3", "FSETP", "SGXT", "BMSK", "FLO", "POPC", "MOV", "CS2R", "IADD", "FSEL", "VOTE", "P2R", "R2P"}
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
hist, name = {}, None
def flush():
    if not hist:
        return
    tot = sum(hist.values())
    alu = sum(v for k, v in hist.items() if k in ALU_PIPE)
    print(f"dynamic instruction mix of {name}: {tot} warp instructions")
    print("  " + ", ".join(f"{k} {v / tot * 100:.1f}%" for k, v in sorted(hist.items(), key=lambda kv: -kv[1])[:14]))
    if hist.get("VABSDIFF4"):
        print(f"  ALU-pipe opcodes {alu / tot * 100:.1f}% of all; VABSDIFF4 {hist['VABSDIFF4'] / alu * 100:.1f}% of the ALU-pipe opcodes")
    print()
col = None
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "Kernel Name":
        flush(); hist, name, col = {}, r[1], None
    elif len(r) > 5 and r[0] == "Address":
        col = (r.index("Source"), r.index("Instructions Executed"))
    elif col and len(r) > max(col) and r[0].startswith("0x"):
        toks = r[col[0]].split()
        if toks and toks[0].startswith("@"):
            toks = toks[1:]
        if toks:
            try:
                hist[toks[0].split(".")[0].rstrip(";")] = hist.get(toks[0].split(".")[0].rstrip(";"), 0) + int(r[col[1]])
            except ValueError:
                pass
flush()
