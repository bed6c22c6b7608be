#!/usr/bin/env python
"""Copies the outputs of tools/gpu_refresh.sh from gpurun_out/<tag>/ into profiles/ under a round prefix and summarises the
ncu captures (read here, no GPU needed):   python tools/collect_profiles.py <tag> r02"""
import glob
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def main():
    tag, prefix = sys.argv[1], sys.argv[2]
    src, dst = os.path.join(ROOT, "gpurun_out", tag), os.path.join(ROOT, "profiles")
    for f in sorted(glob.glob(os.path.join(src, "bench_*.json")) + glob.glob(os.path.join(src, "sweep_*.txt")) + glob.glob(os.path.join(src, "sweep_*.json")) + glob.glob(os.path.join(src, "pytest_gpu.txt"))
                    + glob.glob(os.path.join(src, "launches_*.csv")) + glob.glob(os.path.join(src, "pcie_ceiling.txt"))):
        shutil.copy(f, os.path.join(dst, f"{prefix}_{os.path.basename(f)}"))
    traffic_path = os.path.join(dst, "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for rep in sorted(glob.glob(os.path.join(src, "prof_*.ncu-rep"))):
        name = os.path.basename(rep)[len("prof_"):-len(".ncu-rep")]
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
        open(os.path.join(dst, f"{prefix}_ncu_{name}.txt"), "w").write(out)
        rd = [float(l.split("=")[1].split()[0]) * (1e9 if "Gbyte" in l else 1e6 if "Mbyte" in l else 1.0) for l in out.splitlines() if "dram__bytes_read.sum" in l]
        wr = [float(l.split("=")[1].split()[0]) * (1e9 if "Gbyte" in l else 1e6 if "Mbyte" in l else 1.0) for l in out.splitlines() if "dram__bytes_write.sum" in l]
        kern = [l.split(": ", 1)[1].split("(")[0].replace("void ", "").strip() for l in out.splitlines() if l.startswith("== ")]
        if rd and wr and kern:
            # bench.py matches `kernel` as a substring of fft_b200_last_kernel(): keep the template name up to the first two arguments
            k = kern[0].replace("(int)", "").replace(" ", "")
            k = k.split(",")[0] if "<" in k else k
            traffic[name] = {"bytes": int(rd[0] + wr[0]), "kernel": k.replace("cfb::", ""), "source": f"profiles/{prefix}_ncu_{name}.txt"}
    json.dump(traffic, open(traffic_path, "w"), indent=1)
    print("profiles updated:", prefix)


if __name__ == "__main__":
    main()
