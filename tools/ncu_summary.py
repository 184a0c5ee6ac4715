"""Per-launch summary of an `ncu --page raw --csv` export.  usage: python tools/ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
W = [('gpu__time_duration.sum', 'us'), ('sm__cycles_elapsed.max', 'cyc'), ('dram__bytes_read.sum', 'rd'),
     ('dram__bytes_write.sum', 'wr'), ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'),
     ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tc%'),
     ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
     ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lsuwf%'),
     ('l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'tcwf%'),
     ('launch__registers_per_thread', 'regs'), ('smsp__inst_executed.sum', 'inst'),
     ('l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'ldreq'),
     ('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'ldsec'),
     ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'bankconf'),
     ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smemwf')]
for k, r in enumerate(rows[2:]):
    name = r[hdr.index('Kernel Name')]
    name = name.replace('void arl::tc::tc_kernel<arl::', '').replace('void tc_kernel<', '')[:60]
    out = []
    for m, short in W:
        if m in hdr:
            v = r[hdr.index(m)]
            try:
                f = float(v)
                v = ('%.3g' % f) if f < 1e6 else ('%.3e' % f)
            except ValueError:
                pass
            out.append('%s=%s%s' % (short, v, units[hdr.index(m)] if short in ('rd', 'wr') else ''))
    print('%2d %-45s %s' % (k, name, ' '.join(out)))
