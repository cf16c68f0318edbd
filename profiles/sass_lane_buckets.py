import csv,sys,subprocess
rep,kern=sys.argv[1],sys.argv[2]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--kernel-name',kern],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hi=[i for i,r in enumerate(rows) if 'Instructions Executed' in r][0]
hdr=rows[hi]; ia=hdr.index("Instructions Executed"); it=hdr.index("Thread Instructions Executed"); isrc=hdr.index("Source")
def f(x):
    try: return float(x.replace(',',''))
    except: return None
data=[]
for r in rows[hi+1:]:
    if len(r)==len(hdr) and f(r[ia]) is not None: data.append(r)
    elif r and r[0]=="Kernel Name": break   # first instance only
tot=sum(f(r[ia]) for r in data); tott=sum(f(r[it]) for r in data)
print(len(data),"total inst %.3g thread inst %.3g avg %.2f"%(tot,tott,tott/tot))
for lo,hi_ in [(0,4),(4,8),(8,12),(12,16),(16,24),(24,33)]:
    s=sum(f(r[ia]) for r in data if f(r[ia])>0 and lo<=f(r[it])/f(r[ia])<hi_)
    print("avg threads [%d,%d): %.1f%% of warp-instr"%(lo,hi_,100*s/tot))
w=int(sys.argv[3]) if len(sys.argv)>3 else 48
for k in range(0,len(data),w):
    seg=data[k:k+w]; a=sum(f(r[ia]) for r in seg); t=sum(f(r[it]) for r in seg)
    if a/tot>0.02:
        ops={}
        for r in seg:
            op=r[isrc].split()[0] if not r[isrc].lstrip().startswith('@') else r[isrc].split()[1]
            ops[op.split('.')[0]]=ops.get(op.split('.')[0],0)+f(r[ia])
        top=sorted(ops.items(),key=lambda x:-x[1])[:5]
        print("sass[%4d:%4d] share %4.1f%% avg thr %4.1f  %s"%(k,k+w,100*a/tot,t/a if a else 0," ".join("%s:%.0f%%"%(o,100*v/a) for o,v in top)))
