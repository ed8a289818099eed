"""Extract the judged metrics from an .ncu-rep (run where ncu is installed; no GPU needed):
    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rN_name.md"""
import collections
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum',
        'l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum', 'l1tex__m_l1tex2xbar_write_bytes.sum',
        'launch__occupancy_limit_warps', 'sm__maximum_warps_per_active_cycle_pct']


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print('# ncu summary of `%s`\n' % path)
    for r in rows[2:]:
        print('## %s (launch id %s)\n' % (r[hdr.index('Kernel Name')], r[0]))
        print('| metric | value | unit |\n|---|---|---|')
        for w in WANT:
            if w in hdr:
                print('| %s | %s | %s |' % (w, r[hdr.index(w)], units[hdr.index(w)]))
        stalls = []
        for i, h in enumerate(hdr):
            if 'pcsamp_warps_issue_stalled' in h and not h.endswith('_not_issued'):
                try:
                    stalls.append((float(r[i]), h.split('issue_stalled_')[1]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        print('\nwarp stall samples: ' + ', '.join('%s %.1f%%' % (h, 100 * v / tot) for v, h in sorted(stalls, reverse=True)[:8]))
        print()
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    if heads:
        h = rows[heads[0]]
        ci, cs = h.index('Instructions Executed'), h.index('Source')
        end = heads[1] - 1 if len(heads) > 1 else len(rows)
        ops = collections.Counter()
        static = 0
        for r in rows[heads[0] + 1:end]:
            parts = r[cs].split() if len(r) > ci else []
            if not parts:
                continue
            op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
            try:
                ops[op] += int(r[ci])
                static += 1
            except ValueError:
                pass
        tot = sum(ops.values())
        print('## dynamic SASS mix of the first launch (%d static instructions, %d warp instructions)\n' % (static, tot))
        print('| opcode | share |\n|---|---|')
        for op, n in ops.most_common(16):
            print('| %s | %.1f%% |' % (op, 100.0 * n / tot))


if __name__ == '__main__':
    main(sys.argv[1])
