"""Per-device-function breakdown of an ncu source page: warp-instructions, stall samples and the share of the
instruction-fetch / long-scoreboard / fixed-latency stalls, attributed through the SASS addresses of the noinline
device functions inside correct_kernel.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv
       cuobjdump -xelf all talc_b200/libtalc_b200.so   (the build that was profiled)
       python tools/ncu_functions.py src.csv talc_b200.sm_100a.cubin"""
import csv,sys,subprocess,re,collections
rep_csv=sys.argv[1]; cubin=sys.argv[2]; kname=sys.argv[3] if len(sys.argv)>3 else '_Z14correct_kernelILb1EEv11CorrectArgs'
# function table
out=subprocess.run(['readelf','-sW',cubin],capture_output=True,text=True).stdout
funcs=[]
for l in out.splitlines():
    p=l.split()
    if len(p)>=8 and p[3]=='FUNC' and ('$'+kname+'$' in p[7] or p[7]==kname or p[7].startswith('$__internal')):
        sz=int(p[2],0); funcs.append((int(p[1],16),sz,p[7]))
funcs.sort()
rows=[]
hdr=None
for r in csv.reader(open(rep_csv)):
    if not r: continue
    if r[0]=='Line No': hdr=r; continue
    if hdr and r[0]=='' and len(r)>10 and r[2].startswith('0x'):
        rows.append((int(r[2],16),r[3],int(r[6] or 0),int(r[7] or 0), r))
base=min(a for a,_,_,_,_ in rows)
ist=hdr.index('stall_no_inst'); ilsb=hdr.index('stall_long_sb'); iw=hdr.index('stall_wait')
agg=collections.defaultdict(lambda:[0,0,0,0,0,0])
def fn(off):
    best=kname
    for o,s,n in funcs:
        if n==kname: continue
        if o<=off<o+s: 
            return n
    return best
for a,sass,smp,inst,r in rows:
    f=fn(a-base)
    g=agg[f]; g[0]+=smp; g[1]+=inst; g[2]+=int(r[ist] or 0); g[3]+=int(r[ilsb] or 0); g[4]+=int(r[iw] or 0); g[5]+=1
ts=sum(v[0] for v in agg.values()); ti=sum(v[1] for v in agg.values())
print('total samples',ts,'inst',ti)
print('%-40s %7s %7s %7s %7s %7s %6s'%('function','smp%','inst%','noinst%','longsb%','wait%','nSASS'))
for f,v in sorted(agg.items(),key=lambda x:-x[1][0]):
    name=re.sub(r'.*\$_ZNK?\d*','',f)[:40]
    print('%-40s %7.2f %7.2f %7.2f %7.2f %7.2f %6d'%(name,100*v[0]/ts,100*v[1]/ti,100*v[2]/max(1,v[0]),100*v[3]/max(1,v[0]),100*v[4]/max(1,v[0]),v[5]))
