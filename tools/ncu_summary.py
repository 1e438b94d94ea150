"""Summarise ncu outputs into small text files for profiles/ (run here, on the CPU box).

    python tools/ncu_summary.py launches gpurun_out/launches_r01_fp32.csv  > profiles/launches_r01_fp32.txt
    python tools/ncu_summary.py full     gpurun_out/prof_fp32_r01.ncu-rep  > profiles/ncu_fp32_r01.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__cycles_active.avg",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "").replace("pvb::", "")
        ns = float(r[-1].replace(",", ""))
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += ns
        total += ns
    print(f"# {path}: {len(rows)} launches, {total / 1e6:.3f} ms of kernel time (ncu serialised, cold cache)")
    print(f"{'kernel':<52}{'launches':>9}{'total ms':>10}{'avg us':>10}{'share':>8}")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:<52}{n:>9}{ns / 1e6:>10.3f}{ns / n / 1e3:>10.1f}{ns / total:>8.1%}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: ncu --set full, per launch")
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:100])
        for k in KEYS:
            if k in hdr and r[hdr.index(k)] not in ("", "n/a"):
                print(f"   {k:<66}{r[hdr.index(k)]:>16} {units[hdr.index(k)]}")
        st = [(h, float(r[i].replace(",", ""))) for i, h in enumerate(hdr)
              if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued") and r[i] not in ("", "n/a")]
        tot = sum(v for _, v in st) or 1.0
        top = sorted(st, key=lambda x: -x[1])[:5]
        print("   stall samples: " + ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * v / tot:.0f}%" for h, v in top))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
