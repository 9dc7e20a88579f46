"""Sample distribution of one profiled launch by source line (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys, collections
rep, skip = sys.argv[1], sys.argv[2]
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass','--launch-skip',skip,'--launch-count','1'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
for i,r in enumerate(rows):
    if len(r)>3 and r[0]=='Address': hdr=r; start=i+1; break
ix={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[start:] if len(r)>=len(hdr) and r[0]!='Address']
half=len(data)//2
if half and [r[1] for r in data[:half]]==[r[1] for r in data[half:]]: data=data[:half]
samp=[int(r[ix['# Samples']] or 0) for r in data]
tot=sum(samp) or 1
print('instr',len(data),'samples',tot, 'cols', [h for h in hdr if 'ource' in h or 'ile' in h][:4])
B=int(sys.argv[3]) if len(sys.argv)>3 else 40
for b in range(0,len(data),B):
    s=sum(samp[b:b+B])
    if s/tot<0.015: continue
    ops=collections.Counter(r[1].split()[0] if not r[1].startswith('@') else r[1].split()[1] for r in data[b:b+B])
    print(f"{b:5d} {s/tot:6.1%} "+", ".join(f"{k}:{v}" for k,v in ops.most_common(7)))
