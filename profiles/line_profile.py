"""Per-source-line share of executed warp instructions of one kernel: joins the SASS page of an ncu
report (--set full --import-source on) with nvdisasm's line table of the object that was profiled.
usage: python profiles/line_profile.py <report.ncu-rep> <object.o> <mangled-kernel-substring> [top]"""
import collections, csv, os, re, subprocess, sys, tempfile

def main(rep, obj, kern, top=50):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]; ai, si, ii = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
    data = [(int(r[ai], 16), r[si].strip(), int(r[ii])) for r in rows[2:] if len(r) > ii]
    base = data[0][0]
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", os.path.join(td, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(dis) if ".section" in l and ".text." in l and kern in l)
    cur, off2line = None, {}
    for l in dis[start + 1:]:
        if ".section" in l and ".text." in l:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(\S.*?);", l)
        if m:
            off2line[int(m.group(1), 16)] = cur
    agg, tot = collections.Counter(), 0
    for a, s, n in data:
        agg[off2line.get(a - base)] += n; tot += n
    srcdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dust_b200", "csrc")
    cache = {}
    print(f"total warp instructions {tot}")
    for k, n in agg.most_common(top):
        txt = ""
        if k and os.path.isfile(os.path.join(srcdir, k[0])):
            lines = cache.setdefault(k[0], open(os.path.join(srcdir, k[0])).read().split("\n"))
            txt = lines[k[1] - 1].strip()[:105] if k[1] - 1 < len(lines) else ""
        print(f"{n / tot * 100:5.2f}% {k[0] if k else '?'}:{k[1] if k else 0:<4} {txt}")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 50)
