import re,sys,subprocess,bisect,collections
cubin=sys.argv[1]; kern=sys.argv[2]; srcf=sys.argv[3]
out=subprocess.run(["nvdisasm","-g","-c",cubin],capture_output=True,text=True).stdout
starts=[]
for i,t in enumerate(open(srcf).read().split("\n"),1):
    m=re.search(r"__device__\s+(?:static\s+)?(?:__forceinline__\s+|__noinline__\s+)?(?:constexpr\s+)?[\w:<>\*&]+\s+(\w+)\s*\(",t)
    if m: starts.append((i,m.group(1)))
    elif re.match(r"\s*qp_(solve|factor)_kernel\(",t): starts.append((i,t.strip().split("(")[0]))
sl=[s[0] for s in starts]
cnt=collections.Counter(); infn=False; cur=None; curfile=None
for line in out.split("\n"):
    if line.startswith(".text."): infn = kern in line
    if line.startswith("$__internal"): curfile="<internal>"; cur=0
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)',line)
    if m: curfile=m.group(1); cur=int(m.group(2)); continue
    if infn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+[A-Z@]",line):
        if curfile=="<internal>": cnt["<libdevice slow paths>"]+=1
        elif curfile and curfile.endswith("qp_kernel.cuh"):
            i=bisect.bisect_right(sl,cur)-1; cnt[starts[i][1] if i>=0 else "?"]+=1
        else: cnt["<"+(curfile or "?").split("/")[-1]+">"]+=1
tot=sum(cnt.values()); print("total %.1f KB"%(tot*16/1024))
for k,v in cnt.most_common(30): print("%6.1f KB %s"%(v*16/1024,k))
