"""Per-source-line stall samples of one profiled launch.  usage: tools_ncu_lines.py report.ncu-rep <launch index> [min share]"""
import csv, subprocess, sys, collections
rep, skip = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.015
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass,cuda','--launch-skip',skip,'--launch-count','1'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
agg=collections.OrderedDict(); fname=None; hdr=None; name=None
stall_cols=None
for r in rows:
    if len(r)==2 and r[0]=='File Path': fname=r[1].split('/')[-1]; continue
    if len(r)==2 and r[0]=='Function Name': name=r[1][:60]; continue
    if len(r)>4 and r[0]=='Line No': hdr=r; ix={}; 
    if len(r)>4 and r[0]=='Line No':
        for i,h in enumerate(hdr):
            ix.setdefault(h,i)
        stall_cols=[(h,i) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        continue
    if hdr and len(r)>=len(hdr)-2 and r[2]=='-' and r[0].isdigit():
        # a source line row: Line No, Source text, no address
        try: s=int(r[ix['# Samples']] or 0)
        except: continue
        if s==0: continue
        key=(fname,r[0])
        st={h:int(r[i] or 0) for h,i in stall_cols}
        if key in agg:
            agg[key][0]+=s
            for h in st: agg[key][2][h]+=st[h]
        else: agg[key]=[s,r[1].strip()[:100],collections.Counter(st)]
tot=sum(v[0] for v in agg.values()) or 1
print(name,'samples',tot)
for (f,l),(s,src,st) in agg.items():
    if s/tot>=thr:
        top=", ".join(f"{h[6:]}={c/max(1,sum(st.values())):.0%}" for h,c in st.most_common(2))
        print(f"{f}:{l:>4s} {s/tot:6.1%}  {src:100s} [{top}]")
