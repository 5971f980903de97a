import sys, subprocess, bisect, collections, re
path, exe = sys.argv[1], sys.argv[2]
maps=[]; samples=[]
for l in open(path):
    if l.startswith("MAP "):
        f=l[4:].split()
        if len(f)>=6 and f[5]==exe:
            a,b=f[0].split('-'); maps.append((int(a,16),int(b,16),int(f[2],16)))
    elif l.startswith("S"):
        samples.append([int(x,16) for x in l.split()[1:]])
base=min(m[0]-m[2] for m in maps)
lo=min(m[0] for m in maps); hi=max(m[1] for m in maps)
syms=[]
for l in subprocess.run(["nm","-C","--defined-only",exe],capture_output=True,text=True).stdout.splitlines():
    p=l.split(' ',2)
    if len(p)==3 and p[1] in "tTwW":
        syms.append((int(p[0],16),p[2]))
syms.sort(); addrs=[s[0] for s in syms]
def name(a):
    if not (lo<=a<hi): return None
    i=bisect.bisect_right(addrs,a-base)-1
    return syms[i][1] if i>=0 else None
def short(n): 
    n=re.sub(r"\(.*","",n); return n
self_c=collections.Counter(); incl=collections.Counter()
for s in samples:
    names=[short(n) for n in (name(a) for a in s) if n]
    if not names: self_c["<outside exe (BLAS/libc)>"]+=1; continue
    self_c[names[0] if name(s[0]) else "<lib> under "+names[0]]+=1
    for n in set(names): incl[n]+=1
N=len(samples)
print("samples",N)
print("--- inclusive top")
for n,c in incl.most_common(int(sys.argv[3]) if len(sys.argv)>3 else 45): print("%6.2f%%  %s"%(100*c/N,n))
print("--- self top")
for n,c in self_c.most_common(25): print("%6.2f%%  %s"%(100*c/N,n))
