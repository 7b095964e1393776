"""Summarise ncu output for profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches.csv            -> per-kernel launch count / time / share
  python scripts/ncu_summary.py full gpurun_out/top_kernels.ncu-rep         -> per-launch DRAM traffic etc. of a --set full capture

The launch list comes from `ncu --metrics gpu__time_duration.sum --clock-control none --csv` (cold-cache, serialised
launches: only the SHARES are comparable with bench.py's CUDA-event numbers)."""
import csv
import re
import subprocess
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("b200::", "")
    return name.strip()


def launches(path: str) -> None:
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    head = rows[0]
    iN, iM, iV = head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[iM] != "gpu__time_duration.sum":
            continue
        k = short(r[iN])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iV].replace(",", "")) / 1e3
    tot = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total us | us/launch | share |\n|---|---|---|---|---|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {us:.1f} | {us / n:.1f} | {us / tot:.3f} |")
    print(f"| all | {sum(a[0] for a in agg.values())} | {tot:.1f} | | 1.000 |")


def full(path: str) -> None:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(head)}
    want = [
        ("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "dram read"),
        ("dram__bytes_write.sum", "dram write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "regs/thread"),
        ("launch__occupancy_limit_registers", "occ. limit regs (CTAs/SM)"),
        ("launch__occupancy_limit_shared_mem", "occ. limit smem (CTAs/SM)"),
        ("smsp__inst_executed.sum", "warp instructions"),
    ]
    want = [(m, t) for m, t in want if m in idx]
    print("| kernel | " + " | ".join(f"{t} [{units[idx[m]]}]" for m, t in want) + " |")
    print("|---|" + "---|" * len(want))
    for r in rows[2:]:
        print(f"| `{short(r[idx['Kernel Name']])}` | " + " | ".join(r[idx[m]] for m, _ in want) + " |")


def traffic(path: str) -> None:
    """Mean DRAM bytes (read + write) per launch of every kernel class of a --set full capture, as JSON on stdout
    (bench.py reports it as roofline.traffic)."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(head)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    cls = {"k_sweep<0": "sweep_fwd", "k_sweep<1": "sweep_bwd", "k_amul": "amul", "k_bicg": "vector"}
    agg = {}
    for r in rows[2:]:
        name = short(r[idx["Kernel Name"]])
        key = next((v for k, v in cls.items() if name.startswith(k)), name)
        b = sum(float(r[idx[m]].replace(",", "")) * scale[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        agg.setdefault(key, []).append(b)
    print(json.dumps({"source": path, "dram_bytes_per_launch": {k: sum(v) / len(v) for k, v in agg.items()},
                      "launches_captured": {k: len(v) for k, v in agg.items()}}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
