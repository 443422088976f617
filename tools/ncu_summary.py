"""Turn ncu outputs brought back from the GPU box into the small text summaries committed under profiles/.

  python tools/ncu_summary.py rep  gpurun_out/x.ncu-rep           # key metrics per profiled launch (--set full capture)
  python tools/ncu_summary.py list gpurun_out/launches.csv        # per-kernel share of a launch list (gpu__time_duration)
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"== {r[ki]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:70s} {r[i]:>18s} {units[i]}")


def launch_list(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui], 1.0)
        name = r[ki].split("(")[0][-70:]
        tot[name] += v; cnt[name] += 1
    total = sum(tot.values())
    print(f"{'share':>7s} {'total ms':>10s} {'launches':>8s}  kernel   (device time, serialised cold-cache replay: compare SHARES)")
    for k, v in tot.most_common(25):
        print(f"{100 * v / total:6.1f}% {v / 1e6:10.3f} {cnt[k]:8d}  {k}")
    print(f"total {total / 1e6:.3f} ms over {sum(cnt.values())} launches")


def traffic(path, out_json="profiles/traffic.json"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the three MLP kernels (mean over the captured launches: one
    step = the coarse and the fine launch of each) -> profiles/traffic.json, which bench.py quotes as roofline.traffic together
    with the name of the capture it came from."""
    import json
    import os
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    acc = collections.defaultdict(list)
    for r in rows[2:]:
        name = r[ki].split("(")[0].split("<")[0].replace("void ", "").replace("spn::", "").replace("_kernel", "")
        acc[name].append(float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]])
    res = {k: sum(v) / len(v) for k, v in acc.items()}
    res["_per_launch"] = {k: v for k, v in acc.items()}
    res["_source"] = os.path.basename(path)
    res["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none, mean of the step's "
                    "coarse (64 samples/ray) and fine (128 samples/ray) launch of each kernel; static: measured once per kernel change "
                    "under ncu (tools/gpu_round_checks.sh), not in the timed bench run")
    json.dump(res, open(out_json, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"rep": rep, "list": launch_list, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
