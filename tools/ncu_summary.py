"""Summarise an .ncu-rep (ncu --set full) into the few metrics the roofline discussion uses.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xyz.txt"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
keys = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'l1tex__t_sector_hit_rate.pct']
for n, r in enumerate(rows[2:]):
    print(f"## launch {n}")
    vals = {}
    for k in keys:
        for i, x in enumerate(h):
            if x == k:
                print(f"{k:70s} {r[i]} {units[i]}")
                vals[k] = (r[i], units[i])
    import re
    extra = re.compile(r"^(lts__throughput|l1tex__throughput|lts__t_sectors\.sum$|lts__t_sectors_op_read\.sum$|l1tex__m_xbar2l1tex_read_bytes\.sum"
                       r"|l1tex__m_l1tex2xbar_write_bytes\.sum|sm__cycles_elapsed\.max|smsp__inst_executed\.sum$|lts__t_sectors_srcunit_tex\.sum$"
                       r"|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$|sm__inst_executed_pipe_uniform|smsp__cycles_active\.avg$)")
    for i, x in enumerate(h):
        if extra.match(x) and x not in keys:
            print(f"{x:70s} {r[i]} {units[i]}")
    try:
        t = float(vals['gpu__time_duration.sum'][0].replace(',', ''))
        tu = vals['gpu__time_duration.sum'][1]
        t_us = t * {'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 'ns': 1e-3, 'nsecond': 1e-3}.get(tu, 1)
        def mb(k):
            v, u = vals[k]
            return float(v.replace(',', '')) * {'Mbyte': 1, 'Kbyte': 1e-3, 'Gbyte': 1e3, 'byte': 1e-6}.get(u, 1)
        d = mb('dram__bytes_read.sum') + mb('dram__bytes_write.sum')
        print(f"{'DRAM bytes / duration':70s} {d / t_us * 1e3:.0f} GB/s   ({d:.1f} MB in {t_us:.1f} us)")
    except Exception as ex:
        print("#", ex)
    print()
