import csv, sys, re
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[0]; units=rows[1]
idx={h:i for i,h in enumerate(hdr)}
def g(r,h):
    try: return float(r[idx[h]].replace(',',''))
    except: return float('nan')
seen=set()
for r in rows[2:]:
    name=re.sub(r'\(.*','',r[idx['Kernel Name']]).replace('<unnamed>::','')
    key=(name, r[idx['launch__grid_size']])
    if key in seen: continue
    seen.add(key)
    dur=g(r,'gpu__time_duration.sum'); u=units[idx['gpu__time_duration.sum']]
    if u=='ms': dur*=1000
    elif u=='ns': dur/=1000
    elif u in ('s','second'): dur*=1e6
    rd=g(r,'dram__bytes_read.sum'); ru=units[idx['dram__bytes_read.sum']]; wr=g(r,'dram__bytes_write.sum'); wu=units[idx['dram__bytes_write.sum']]
    f={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}
    rd*=f.get(ru,1); wr*=f.get(wu,1)
    stalls={h.replace('smsp__pcsamp_warps_issue_stalled_',''):g(r,h) for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')}
    tot=sum(v for v in stalls.values() if v==v) or 1
    top=sorted(stalls.items(), key=lambda kv:-(kv[1] if kv[1]==kv[1] else 0))[:4]
    print(f"{name[:58]:58s} grid={r[idx['launch__grid_size']]:>6s} dur={dur:8.1f}us dram_rd={rd/1e6:7.1f}MB wr={wr/1e6:7.1f}MB ({(rd+wr)/dur/1e3:6.0f} GB/s) occ={g(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f}% regs={r[idx['launch__registers_per_thread']]} L1hit={g(r,'l1tex__t_sector_hit_rate.pct'):4.0f}% L2hit={g(r,'lts__t_sector_hit_rate.pct'):4.0f}% inst={g(r,'smsp__inst_executed.sum')/1e6:6.1f}M")
    print("      stalls: "+", ".join(f"{k}={v/tot:.0%}" for k,v in top))
