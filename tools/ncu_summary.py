#!/usr/bin/env python
"""Summarise an .ncu-rep (one profiled launch) into a small markdown file for profiles/: headline metrics from the raw
page and the hottest source lines (instructions executed, stall samples) from the cuda,sass source page."""
import csv, io, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum"]


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    srcdir = sys.argv[4] if len(sys.argv) > 4 else None          # the csrc directory the profiled library was built from (default: the working tree)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu summary: %s\n\n%s\n\n" % (rep.split("/")[-1], note))
        for vals in rows[2:]:
            f.write("## launch: %s\n\n| metric | value | unit |\n|---|---|---|\n" % vals[hdr.index("Kernel Name")][:80])
            for m in WANT:
                if m in hdr:
                    i = hdr.index(m); f.write("| %s | %s | %s |\n" % (m, vals[i], units[i]))
            f.write("\n")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
        cur = "?"; h = None; lines = []
        for r in csv.reader(io.StringIO(src)):
            if len(r) >= 2 and "File" in r[0]:
                cur = r[1]; continue
            if len(r) > 5 and r[0] == "Line No":
                h = r; continue
            if h is None or len(r) < len(h) or not r[0].isdigit():
                continue
            ie = r[h.index("Instructions Executed")]; sm = r[h.index("# Samples")]; te = r[h.index("Thread Instructions Executed")]; ls = r[h.index("stall_long_sb")]
            if ie.isdigit():
                lines.append((int(ie), int(sm) if sm.isdigit() else 0, int(te) if te.isdigit() else 0, int(ls) if ls.isdigit() else 0, cur.split("/")[-1], int(r[0]), r[1].strip()[:100]))
        tot = sum(x[0] for x in lines) or 1; ts = sum(x[1] for x in lines) or 1
        # time by function: map kj_core*.h lines to the enclosing `static KJ_DEV ... kj_xxx(` definition
        import os, re
        fmap = {}
        for fn in ("kj_core.h", "kj_core_greedy.h"):
            path = os.path.join(srcdir or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "kaiju_b200", "csrc"), fn)
            cur = "(top)"; m = {}
            try:
                for i, l in enumerate(open(path).read().split("\n"), 1):
                    mm = re.match(r"^(?:template <[^>]*>\s*)?(?:static )?(?:KJ_DEV|KJ_NOINLINE) .*?\b(kj_\w+)\s*\(", l)
                    if mm:
                        cur = mm.group(1)
                    m[i] = cur
            except OSError:
                pass
            fmap[fn] = m
        agg = {}
        for x in lines:
            key = fmap.get(x[4], {}).get(x[5], x[4]) if x[4] in fmap else x[4]
            a = agg.setdefault(key, [0, 0]); a[0] += x[0]; a[1] += x[1]
        f.write("## time by function (stall-sample share = share of warp time; valid when the report was taken from the committed source)\n\n| function / file | warp instructions % | warp time % |\n|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
            f.write("| %s | %.1f | %.1f |\n" % (k, 100.0 * a[0] / tot, 100.0 * a[1] / ts))
        f.write("\n")
        f.write("## hottest source lines (share of warp instructions executed / of stall samples; active threads per instruction; long-scoreboard samples)\n\n")
        f.write("| inst %% | samples %% | thr/inst | long_sb | location | source |\n|---|---|---|---|---|---|\n")
        for x in sorted(lines, reverse=True)[:40]:
            f.write("| %.1f | %.1f | %.1f | %d | %s:%d | `%s` |\n" % (100.0 * x[0] / tot, 100.0 * x[1] / ts, x[2] / max(x[0], 1), x[3], x[4], x[5], x[6].replace("|", "\\|")))


def footprint(rep, out):
    """hot code footprint: how many 128-byte instruction lines hold 90 / 95 / 99 % of the executed warp instructions"""
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rows = list(csv.reader(io.StringIO(sass)))
    if len(rows) < 3:
        return
    h = rows[1]; ia, ie = h.index("Address"), h.index("Instructions Executed"); ins = []
    for x in rows[2:]:
        try:
            ins.append((int(x[ia], 16), int(x[ie])))
        except Exception:
            pass
    if not ins:
        return
    base = ins[0][0]; tot = sum(e for _, e in ins) or 1; acc = 0; lines = set(); marks = [0.9, 0.95, 0.99]; res = []
    for a, e in sorted(ins, key=lambda t: -t[1]):
        acc += e; lines.add((a - base) // 128)
        while marks and acc >= marks[0] * tot:
            res.append("%.0f %% of the executed instructions come from %.1f KB" % (marks[0] * 100, len(lines) * 128 / 1024)); marks.pop(0)
    with open(out, "a") as f:
        f.write("\n## instruction footprint (kernel: %d instructions = %.1f KB of SASS)\n\n%s\n" % (len(ins), len(ins) * 16 / 1024, "; ".join(res)))


if __name__ == "__main__":
    main()
    footprint(sys.argv[1], sys.argv[2])
