import csv, collections, subprocess, sys
rep=sys.argv[1]; lo=int(sys.argv[2]); hi=int(sys.argv[3])
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[2:] if len(r)>=len(hdr)]
by=collections.Counter()
for r in data:
    by[int(r[ix['Instructions Executed']] or 0)]+=1
tot=sum(k*v for k,v in by.items())
print("total",tot)
for k,v in sorted(by.items(), key=lambda kv:-kv[0]*kv[1])[:10]:
    print("exec count %d : %d instructions -> %.1f%%" % (k, v, 100.0*k*v/tot))
c=collections.Counter(); n=0
for r in data:
    ex=int(r[ix['Instructions Executed']] or 0)
    if lo <= ex <= hi:
        t=r[ix['Source']].split()
        op=(t[1] if t[0].startswith('@') else t[0])
        op='.'.join(op.split('.')[:2]) if op.startswith(('VI','LDS','STG','LDG','IMAD')) else op.split('.')[0]
        c[op]+=1; n+=1
print(n, "instr in range")
print(sorted(c.items(), key=lambda kv:-kv[1]))
