"""Summarise an .ncu-rep (read here, no GPU): key raw metrics, stall mix, opcode mix, hot source
lines. Usage: ncu_summary.py report.ncu-rep [n_lines]"""
import collections, csv, io, re, subprocess, sys

def run(args):
    return subprocess.run(["ncu", "-i", sys.argv[1]] + args, capture_output=True, text=True).stdout

def main():
    nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
            "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "smsp__inst_executed_op_shared_ld.sum",
            "smsp__inst_executed_op_shared_st.sum", "sm__inst_executed_pipe_uniform.sum",
            "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_alu.sum",
            "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_lsu.sum",
            "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_cbu.sum",
            "smsp__inst_executed_pipe_adu.sum", "smsp__inst_executed_pipe_uniform.sum"]
    print("== raw metrics ==")
    for i, h in enumerate(hdr):
        if h in want:
            print("  %-62s %-14s %s" % (h, units[i], vals[i]))
    rows = list(csv.reader(io.StringIO(run(["--page", "source", "--print-source", "cuda,sass", "--csv"]))))
    hdr = None
    lines, sass = [], []
    cur = None
    for r in rows:
        if len(r) > 10 and r[0] == "Line No":
            hdr = r
            ix = {h: i for i, h in enumerate(hdr)}
            continue
        if hdr and len(r) == len(hdr):
            def f(k):
                try:
                    return float(r[ix[k]])
                except ValueError:
                    return 0.0
            if r[0] != "":
                cur = (r[0], r[1].strip()[:70])
                lines.append((f("Instructions Executed"), f("# Samples"), f("Thread Instructions Executed"), cur))
            else:
                st = {h: f(h) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
                sass.append((r[3].strip(), f("Instructions Executed"), f("# Samples"), st, cur))
    tot = sum(x[1] for x in sass) or 1
    ts = sum(x[2] for x in sass) or 1
    print("== totals: %.3e warp-instructions, %d samples ==" % (tot, ts))
    stall = collections.Counter()
    for x in sass:
        for k, v in x[3].items():
            stall[k] += v
    print("== stall mix ==")
    for k, v in stall.most_common(9):
        print("  %-26s %5.1f%%" % (k, 100 * v / ts))
    ops = collections.Counter()
    osm = collections.Counter()
    for x in sass:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", x[0])
        k = m.group(2) if m else x[0][:8]
        ops[k] += x[1]
        osm[k] += x[2]
    print("== opcode mix ==")
    for k, v in ops.most_common(22):
        print("  %-10s %5.1f%% inst %5.1f%% samples" % (k, 100 * v / tot, 100 * osm[k] / ts))
    # device functions follow the kernel body in address order (SASS-only page); each ends with a RET
    srows = list(csv.reader(io.StringIO(run(["--page", "source", "--print-source", "sass", "--csv"]))))
    shdr = next((r for r in srows if len(r) > 10 and r[0] == "Address"), None)
    if shdr:
        six = {h: i for i, h in enumerate(shdr)}
        body = [r for r in srows if len(r) == len(shdr) and r[0] != "Address"]
        def num(r, k):
            try:
                return float(r[six[k]])
            except ValueError:
                return 0.0
        ti = sum(num(r, "Instructions Executed") for r in body) or 1
        tsm = sum(num(r, "# Samples") for r in body) or 1
        print("== code segments in address order (split after RET.REL.NODEC; the first is the kernel body) ==")
        start = 0
        for i, r in enumerate(body):
            if "RET.REL.NODEC" in r[six["Source"]] or i == len(body) - 1:
                blk = body[start:i + 1]
                ie = sum(num(b, "Instructions Executed") for b in blk)
                sm = sum(num(b, "# Samples") for b in blk)
                if ie / ti > 0.002 or sm / tsm > 0.002:
                    print("  instr %5d-%5d  %5.1f%% inst %5.1f%% smp | first: %s" % (
                        start, i, 100 * ie / ti, 100 * sm / tsm, blk[0][six["Source"]].strip()[:48]))
                start = i + 1
    print("== hot source lines ==")
    for a in sorted(lines, key=lambda x: -x[1])[:nlines]:
        print("  %5.2f%% inst %5.2f%% smp thr %4.1f | %s | %s" % (
            100 * a[0] / tot, 100 * a[1] / ts, a[2] / max(a[0], 1), a[3][0], a[3][1]))

main()
