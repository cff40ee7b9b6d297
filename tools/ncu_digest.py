"""Digest of an `ncu --page raw --csv` + `--page source --csv` export: key metrics, stall reasons, per-opcode
mix and the most stalled SASS instructions.  Usage: python tools/ncu_digest.py raw.csv source.csv [top]"""
import collections
import csv
import re
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.per_cycle_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__icc_request_hit_rate.pct',
        'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__inst_executed.sum', 'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_shared_st.sum',
        'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum',
        'sm__ops_path_tensor_src_fp64.avg.peak_sustained']


def main():
    raw, src = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    rows = list(csv.reader(open(raw)))
    hdr, units, r = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, r)}
    print('| metric | value | unit |\n|---|---|---|')
    for k in KEYS:
        if k in d:
            print(f'| {k} | {d[k][0]} | {d[k][1]} |')
    print('\n| stall reason | warps per issue |\n|---|---|')
    for h in hdr:
        m = re.match(r'smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio', h)
        if m and float(d[h][0]) > 0.01:
            print(f'| {m.group(1)} | {float(d[h][0]):.3f} |')
    rows = list(csv.reader(open(src)))
    hdr, data = rows[1], rows[2:]
    ia, ie, it = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
    print(f'\nSASS instructions: {len(data)}; samples: {sum(int(x[ia]) for x in data)}')
    op, ops = collections.Counter(), collections.Counter()
    for x in data:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', x[1])
        o = m.group(2).split('.')[0] if m else '?'
        op[o] += int(x[ie])
        ops[o] += int(x[ia])
    print('\n| opcode | warp instructions executed | stall samples |\n|---|---|---|')
    for o, c in op.most_common(14):
        print(f'| {o} | {c} | {ops[o]} |')
    sb = hdr.index('stall_barrier')
    names = hdr[sb:sb + 17]
    print('\n| stall | samples |\n|---|---|')
    for j, nme in enumerate(names):
        s = sum(int(x[sb + j] or 0) for x in data)
        if s:
            print(f'| {nme} | {s} |')
    print('\nmost stalled instructions:')
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][ia]))[:top]
    for i in idx:
        x = data[i]
        st = {names[j][6:]: int(x[sb + j] or 0) for j in range(17) if int(x[sb + j] or 0) > 0}
        t3 = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f'  [{i}] {x[1].strip()[:58]:58s} samples {x[ia]:>6s} exec {x[ie]:>8s} {t3}')


if __name__ == '__main__':
    main()
