"""profiles/traffic_r02.json from ncu --set full captures of tools/prof_kernels.py (conv1 shape, batch 32): DRAM bytes per
launch next to the algorithmic bytes, the commit the capture was taken at and the sha1 of each kernel's source file (bench.py
flags kernels whose source changed since).   python tools/make_traffic.py <tc32 f16 .ncu-rep> [<bf16 .ncu-rep>]"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = 32
x = B * 32 * 17 * 62 * 62 * 4          # conv1 input, fp32
y = B * 32 * 15 * 60 * 60 * 4          # conv1 output
gzp = B * 32 * 19 * 64 * 64 * 4        # its gradient, zero-padded by 2 (the tensor the backward kernels read)
CSRC = "predict_pv_yield_b200/csrc/"
# ncu kernel-name fragment, launch index among the matches -> bench kernel class, source, algorithmic bytes, description
FP32 = [
    ("conv3d_igemm_tf32x3_pair_kernel", 0, "conv3d_fwd_f16x2", CSRC + "conv3d_igemm_tf32x3.cu", x + y,
     "conv1 forward, two-way fp16 split, CTA-pair kernel (x read, blocked y written)"),
    ("conv3d_igemm_tf32x3_pair_kernel", 1, "conv3d_dgrad_f16x2", CSRC + "conv3d_igemm_tf32x3.cu", gzp + x + x,
     "conv1 data gradient (padded gz read, ReLU-mask source read, padded gx written)"),
    ("conv3d_wgrad_bf16x3_kernel", 0, "conv3d_wgrad_f16x2", CSRC + "conv3d_wgrad_bf16x3.cu", x + gzp,
     "conv1 weight gradient, two-way fp16 split (x halo rows re-read per block of 6 output rows: 8/6)"),
]
BF16 = [
    ("conv3d_wgrad_bf16_rows_kernel", 0, "conv3d_wgrad_bf16", CSRC + "conv3d_wgrad_bf16_rows.cu", (x + gzp) // 2,
     "conv1 weight gradient, blocked bf16, row-step kernel (tensor-map loads)"),
]


def launches(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    res = []
    for r in rows[2:]:
        g = lambda k: float(r[hdr.index(k)].replace(",", ""))  # noqa: E731
        unit = lambda k: rows[1][hdr.index(k)]  # noqa: E731
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tr = sum(g(k) * scale[unit(k)] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        res.append((r[hdr.index("Kernel Name")], tr, g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                    g("gpu__time_duration.sum")))
    return res


def main():
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    table = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full (tools/prof_kernels.py: conv1 layer "
                         "shape 32->32, in 17x62x62, batch 32); algorithmic = bytes of each tensor touched once; made by tools/make_traffic.py",
             "_commit": commit}
    for rep, spec in zip(sys.argv[1:], (FP32, BF16)):
        ls = launches(rep)
        for frag, idx, cls, src, alg, what in spec:
            m = [l for l in ls if frag in l[0]]
            if len(m) <= idx:
                continue
            table[cls] = {"traffic": int(m[idx][1]), "algorithmic": alg, "launch": what, "tensor_pipe_active_pct": m[idx][2],
                          "duration_us": m[idx][3], "source_file": src,
                          "source_sha1": hashlib.sha1(open(os.path.join(ROOT, src), "rb").read()).hexdigest(), "capture": os.path.basename(rep)}
    json.dump(table, open(os.path.join(ROOT, "profiles", "traffic_r02.json"), "w"), indent=1)
    for k, v in table.items():
        if not k.startswith("_"):
            print(f"{k:24s} traffic {v['traffic'] / 1e6:8.1f} MB  algorithmic {v['algorithmic'] / 1e6:8.1f} MB  tensor pipe {v['tensor_pipe_active_pct']:.0f} %  {v['duration_us']:.0f} us")


if __name__ == "__main__":
    main()
