import csv, sys, subprocess
rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
sel = ['--launch-skip',kern,'--launch-count','1'] if kern.isdigit() else ['--kernel-name',f'regex:{kern}']
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass']+sel,capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=None
for i,r in enumerate(rows):
    if len(r)>3 and r[0]=='Address': hdr=r; start=i+1; break
ix={h:i for i,h in enumerate(hdr)}
data=[]
for r in rows[start:]:
    if len(r)<len(hdr): break
    data.append(r)
samp=[int(r[ix['# Samples']] or 0) for r in data]
tot=sum(samp) or 1
print('kernel',kern,'instructions',len(data),'samples',tot)
order=sorted(range(len(data)), key=lambda i:-samp[i])[:topn]
for i in sorted(order):
    r=data[i]
    print(f"{i:5d} {samp[i]:6d} {samp[i]/tot:5.1%} exec={r[ix['Instructions Executed']]:>9s}  {r[1].strip()[:100]}")
