"""Prints the handful of ncu counters the design notes quote, from a .ncu-rep (read on the CPU box).
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [title]"""
import csv
import io
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'dram__bytes.sum.per_second', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
    'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum',
    'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum',
]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    print(f"# {title}\n\n| metric | unit | value |\n|---|---|---|")
    for k in KEYS:
        if k in d:
            print(f"| {k} | {d[k][0]} | {d[k][1]} |")


def opcode_mix(rep, top=16):
    """Appendix: executed warp instructions and predicated-on thread instructions per SASS opcode (source page)."""
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ia, ie, it = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Predicated-On Thread Instructions Executed')
    warp, thread = {}, {}
    for r in rows[2:]:
        t = r[ia].split()
        if not t:
            continue
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        warp[op] = warp.get(op, 0) + int(r[ie])
        thread[op] = thread.get(op, 0) + int(r[it])
    tot = sum(warp.values())
    print(f"\n## SASS opcode mix (source page; {tot} warp instructions)\n\n| opcode | warp instructions | share | predicated-on thread instructions |\n|---|---|---|---|")
    for op, n in sorted(warp.items(), key=lambda kv: -kv[1])[:top]:
        print(f"| {op} | {n} | {100.0 * n / tot:.1f} % | {thread[op]} |")


if __name__ == '__main__':
    main()
    opcode_mix(sys.argv[1])
