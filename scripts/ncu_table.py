"""Condense an `ncu --page raw --csv` export into the table the profile summaries quote.
Usage: python scripts/ncu_table.py raw.csv"""
import csv, re, sys
rows = list(csv.DictReader(open(sys.argv[1])))
units = rows[0]                                   # the raw page's second line holds the unit of every column
rows = [r for r in rows[1:] if r.get('Kernel Name')]
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3,
         'usecond': 1e-3, 'nsecond': 1e-6, 'msecond': 1.0, 'second': 1e3}
nv = 'thread Domain:Push/Pop_Range:PL_Type:PL_Value:CLR_Type:Color:Msg_Type:Msg'
def num(r, k):
    try:
        return float(str(r.get(k, '')).replace(',', '')) * SCALE.get(units.get(k, ''), 1.0)
    except ValueError:
        return float('nan')
print('| range | kernel | ms | DRAM rd MB | DRAM wr MB | DRAM % | tensor % (elapsed) | L2->SM GB | L2 hit % | regs | smem KB |')
print('|---|---|---|---|---|---|---|---|---|---|---|')
for r in rows:
    rng = re.sub(r'^.*?:', '', r.get(nv, ''), count=0)
    m = re.search(r'((?:fwd|dgrad|wgrad):[\w\.]+)', r.get(nv, ''))
    rng = m.group(1) if m else ''
    name = re.sub(r'\(.*$', '', r['Kernel Name']).replace('void ', '').replace('fp::', '')
    ms = num(r, 'gpu__time_duration.sum')
    print(f"| {rng} | `{name[:44]}` | {ms:.3f} | {num(r,'dram__bytes_read.sum')/1e6:.0f} | {num(r,'dram__bytes_write.sum')/1e6:.0f} | "
          f"{num(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
          f"{num(r,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.0f} | "
          f"{num(r,'l1tex__m_xbar2l1tex_read_bytes.sum')/1e9:.2f} | {num(r,'lts__t_sector_hit_rate.pct'):.0f} | "
          f"{num(r,'launch__registers_per_thread'):.0f} | {num(r,'launch__shared_mem_per_block_dynamic')/1024:.0f} |")
