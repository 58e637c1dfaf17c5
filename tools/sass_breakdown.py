"""Histogram of SASS instructions of solve_kernel<3,VBOC> by source function (needs -lineinfo)."""
import collections, glob, os, re, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "vboc_b200", "libvboc_b200.so")], cwd=tmp,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
src = open(os.path.join(ROOT, "vboc_b200", "csrc", "ocp_warp.h")).read().split("\n")
meth, name = {}, "?"
for i, l in enumerate(src, 1):
    m = re.match(r"\s+VB_DEV\s+[\w:<>,\s\*&]+?\s+(\w+)\(", l)
    if m:
        name = m.group(1)
    meth[i] = name
cnt, cur_fn, cur_line = collections.Counter(), None, None
for l in dis.split("\n"):
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", l)
    if m:
        cur_fn = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if cur_fn and "solve_kernelILi3ELi0" in cur_fn and re.match(r"\s+/\*[0-9a-f]+\*/", l):
        f, ln = cur_line if cur_line else ("?", 0)
        cnt[meth.get(ln, "?") if f == "ocp_warp.h" else f] += 1
print("total", sum(cnt.values()))
for k, v in cnt.most_common(18):
    print(v, k)
