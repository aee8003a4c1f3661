import csv,sys,subprocess,re
f=sys.argv[1]
out=subprocess.run(['ncu','-i',f,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[1]
isrc=hdr.index('Source'); isamp=hdr.index('Warp Stall Sampling (All Samples)'); iex=hdr.index('Instructions Executed')
data=[(int(r[isamp] or 0), r[isrc].strip(), int(r[iex] or 0)) for r in rows[2:] if len(r)>isamp]
tot=sum(d[0] for d in data); print('total samples',tot)
acc=0; prev=0
for i,(s,src,ex) in enumerate(data):
    acc+=s
    if any(k in src for k in ('SYNCS.PHASECHK','SYNCS.ARRIVE','LDTM','STTM','UTCBAR','EXIT','UTMALDG','BAR.SYNC','UTCHMMA')):
        print(f'#{i:4d} cum={acc:5d} (+{acc-prev:4d} {100*(acc-prev)/tot:4.1f}%) ex={ex:8d} {src[:76]}')
        prev=acc
