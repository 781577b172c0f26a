"""Record the DRAM traffic of a kernel from an `ncu --set full` report into profiles/traffic.json, together with a
hash of the sources the capture was taken on.  bench.py emits `roofline.traffic` only while that hash still matches
the tree (a stale figure is dropped, not repeated).

usage: python profiles/update_traffic.py [--sum] <report.ncu-rep> <kernel key> <source files ...>
   --sum: the kernel runs as several launches per call (phi_tc_kernel: whole row tiles, then the rest): add them up
   e.g. python profiles/update_traffic.py gpurun_out/fused_r2c.ncu-rep svmpc_instance_kernel=svmpc_warp_kernel \
            dust_b200/csrc/rollout.cu dust_b200/csrc/models.cuh dust_b200/csrc/common.cuh"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench_common import source_sha  # noqa: E402


def main():
    argv = [a for a in sys.argv[1:] if a != "--sum"]
    total = "--sum" in sys.argv[1:]
    rep, key, files = argv[0], argv[1], argv[2:]
    key, _, match = key.partition("=")      # "<json key>=<substring of the kernel name>" when the two differ
    match = match or key
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cand = [r for r in rows[2:] if match in r[hdr.index("Kernel Name")]]
    assert cand, f"no launch of {key} in {rep}"
    r = cand[-1]

    def get(name):
        i = hdr.index(name)
        u = units[i].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        return sum(float(q[i].replace(",", "")) * scale for q in (cand if total else [r]))

    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    path = os.path.join(ROOT, "profiles", "traffic.json")
    db = json.load(open(path)) if os.path.isfile(path) else {}
    db[key] = {"kernel": r[hdr.index("Kernel Name")], "report": os.path.basename(rep), "dram_bytes_read": rd, "dram_bytes_write": wr,
               "dram_bytes_per_launch": rd + wr, "source_files": files, "source_sha": source_sha(files),
               "launches": len(cand) if total else 1,
               "note": "ncu --set full, one call of the bench workload, --clock-control none"}
    json.dump(db, open(path, "w"), indent=1)
    print(key, db[key])


if __name__ == "__main__":
    main()
