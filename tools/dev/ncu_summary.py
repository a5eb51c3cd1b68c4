"""Turn the raw ncu CSV pages of the evidence job (tools/dev/jobs/evidence.sh) into the summaries bench.py and DESIGN.md cite.

    python tools/dev/ncu_summary.py conv  <pose_raw.csv> <flow_raw.csv> <out.json>   # per-launch DRAM bytes, tensor-pipe %, time
    python tools/dev/ncu_summary.py corr  <corr_raw.csv> <out.json>

The JSON records the SHA-256 of the profiled kernel's own translation unit + csrc headers (flowtrack/pytorch_b200/build.py:
kernel_hash); bench.py refuses a capture whose kernel hash differs from
the sources the library it runs was built from.  Run it on the tree the capture was taken with."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from flowtrack.pytorch_b200 import build as ft_build  # noqa: E402

UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def rows(path):
    r = list(csv.reader(open(path)))
    hdr, units = r[0], r[1]
    for row in r[2:]:
        if len(row) != len(hdr):
            continue
        d = {}
        for h, u, v in zip(hdr, units, row):
            try:
                d[h] = float(v.replace(",", "")) * UNIT.get(u, 1.0)
            except ValueError:
                d[h] = v
        yield d


def conv(pose_csv, flow_csv, out):
    n = 0
    dram = t = tw = 0.0
    per_net = {}
    for name, path in (("pose", pose_csv), ("flow", flow_csv)):
        k = [d for d in rows(path) if "conv_tc_kernel" in str(d.get("Kernel Name", ""))]
        tt = sum(d["gpu__time_duration.sum"] for d in k)
        per_net[name] = {"launches": len(k), "time_us_under_ncu": tt * 1e6,
                         "tensor_pipe_active_pct_time_weighted": sum(d["gpu__time_duration.sum"] * d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"] for d in k) / tt,
                         "dram_bytes": sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in k),
                         "best_layer_tensor_pipe_pct": max(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"] for d in k)}
        n += len(k)
        t += tt
        dram += per_net[name]["dram_bytes"]
        tw += sum(d["gpu__time_duration.sum"] * d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"] for d in k)
    rec = {"source": "ncu --set full --clock-control none over every conv_tc_kernel launch of one PoseResNet-50 B=64 forward and one FlowNet2S B=16 forward "
                     "(tools/dev/jobs/evidence.sh; raw pages: %s, %s)" % (os.path.basename(pose_csv), os.path.basename(flow_csv)),
           "kernel_sources": list(ft_build.CONV_SOURCES), "kernel_sources_sha256": ft_build.kernel_hash(ft_build.CONV_SOURCES), "launches": n, "dram_bytes_per_launch": dram / n, "time_us_per_launch_under_ncu": t / n * 1e6,
           "tensor_pipe_active_pct_time_weighted": tw / t, "per_net": per_net}
    json.dump(rec, open(out, "w"), indent=1)
    print(json.dumps(rec, indent=1))


def corr(corr_csv, out):
    k = [d for d in rows(corr_csv) if "corr_tc_kernel" in str(d.get("Kernel Name", "")) or "corr_pair_kernel" in str(d.get("Kernel Name", ""))]
    recs = []
    for d in k:
        recs.append({"kernel": d["Kernel Name"], "time_us_under_ncu": d["gpu__time_duration.sum"] * 1e6,
                     "dram_bytes_read": d["dram__bytes_read.sum"], "dram_bytes_write": d["dram__bytes_write.sum"],
                     "tensor_pipe_active_pct": d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                     "dram_throughput_pct": d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                     "l1tex_xbar_read_bytes": d.get("l1tex__m_xbar2l1tex_read_bytes.sum"), "l1tex2xbar_write_bytes": d.get("l1tex__m_l1tex2xbar_write_bytes.sum")})
    rec = {"source": "ncu --set full --clock-control none, the correlation kernel at [32,256,48,64] (tools/dev/jobs/evidence.sh; raw page %s)" % os.path.basename(corr_csv),
           "kernel_sources": list(ft_build.CORR_SOURCES), "kernel_sources_sha256": ft_build.kernel_hash(ft_build.CORR_SOURCES), "launches": recs}
    json.dump(rec, open(out, "w"), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    {"conv": conv, "corr": corr}[sys.argv[1]](*sys.argv[2:])
