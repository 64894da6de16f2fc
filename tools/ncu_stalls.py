import csv,re,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
# split per kernel: sections start with "Kernel Name" line
blocks=out.split('"Kernel Name"')
for blk in blocks[1:]:
    lines=('"Kernel Name"'+blk).splitlines()
    name=lines[0].split('","')[1][:60]
    rows=list(csv.reader(lines[1:]))
    hdr=rows[0]; data=[r for r in rows[1:] if len(r)==len(hdr)]
    ie=hdr.index("Instructions Executed"); src=hdr.index("Source"); smp=hdr.index("# Samples")
    stall=[i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tsm=sum(float(r[smp] or 0) for r in data) or 1; tot=sum(float(r[ie] or 0) for r in data) or 1
    agg={}
    for r in data:
        for i in stall: agg[hdr[i]]=agg.get(hdr[i],0)+float(r[i] or 0)
    print("\n==",name,"inst",int(tot),"samples",int(tsm))
    print({k[6:]:round(100*v/tsm,1) for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:6]})
    top=sorted(range(len(data)), key=lambda i:-float(data[i][smp] or 0))[:8]
    for i in sorted(top):
        r=data[i]
        st={hdr[k][6:]:int(float(r[k] or 0)) for k in stall if float(r[k] or 0)>0.25*float(r[smp] or 1)}
        print(" ",i, f"{100*float(r[smp])/tsm:5.2f}% inst {100*float(r[ie] or 0)/tot:4.2f}%", r[src][:70], st)
