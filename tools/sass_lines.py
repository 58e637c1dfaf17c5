import collections, glob, os, re, subprocess, tempfile, sys
ROOT='/root/repo'
lib = os.path.abspath(sys.argv[1]) if len(sys.argv)>1 else os.path.join(ROOT,"vboc_b200","libvboc_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump","-xelf","all",lib],cwd=tmp,stdout=subprocess.DEVNULL,stderr=subprocess.DEVNULL)
cubin = glob.glob(os.path.join(tmp,"*.cubin"))[0]
dis = subprocess.run(["nvdisasm","-g","-c",cubin],capture_output=True,text=True).stdout
cnt=collections.Counter(); cur_fn=None; cur=None; tot=0
ops=collections.Counter()
for l in dis.split("\n"):
    m=re.match(r"\s*\.section\s+\.text\.(\S+)",l)
    if m: cur_fn=m.group(1); continue
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split("/")[-1],int(m.group(2))); continue
    if cur_fn and "solve_kernelILi3ELi0ELi5ELb0" in cur_fn and re.match(r"\s+/\*[0-9a-f]+\*/",l):
        cnt[cur]+=1; tot+=1
        mm=re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)",l)
        if mm: ops[mm.group(1).split('.')[0]]+=1
print("total",tot, tot*16/1024,"KB")
src={}
for (f,ln),v in cnt.most_common(45):
    if f not in src:
        try: src[f]=open(os.path.join(ROOT,"vboc_b200","csrc",f)).read().split("\n")
        except Exception: src[f]=None
    t=src[f][ln-1].strip()[:80] if src[f] and ln-1<len(src[f]) else ""
    print(v,f,ln,t)
print(ops.most_common(25))
src_w=open(os.path.join(ROOT,"vboc_b200","csrc","ocp_warp.h")).read().split("\n")
meth={}; name="?"
for i,l in enumerate(src_w,1):
    m=re.match(r"\s+VB_DEV\s+[\w:<>,\s\*&]+?\s+(\w+)\(",l)
    if m: name=m.group(1)
    meth[i]=name
byfn=collections.Counter()
for (f,ln),v in cnt.items():
    byfn[meth.get(ln,"?") if f=="ocp_warp.h" else f]+=v
print(sorted(byfn.items(), key=lambda kv:-kv[1]))
