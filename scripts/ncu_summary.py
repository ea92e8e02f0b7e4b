"""Turn ncu outputs brought back in gpurun_out/ into the small, tracked summaries under profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches_X.csv profiles/launches_X.md   [skip_first_n_launches]
      per-kernel totals / shares of a `--metrics gpu__time_duration.sum` launch list
  python scripts/ncu_summary.py full gpurun_out/prof_X.ncu-rep profiles/ncu_full_X.md
      one row per captured launch of an `ncu --set full` report: duration, tensor-pipe %, DRAM bytes, registers ...
"""
from __future__ import annotations

import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("lx::", "")
    return name[:70]


def launches(src: str, dst: str, skip: int = 0) -> None:
    text = [l for l in open(src, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(text))))
    rows = [r for r in rows if r["Metric Name"] == "gpu__time_duration.sum"]
    rows = rows[skip:]
    agg: "OrderedDict[str, list]" = OrderedDict()
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary: `{src}` (first {skip} launches skipped)\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none`: per-launch times are cold-cache and serialised; "
                "use the SHARES, not the absolutes.\n\n")
        f.write(f"{len(rows)} launches, {total/1e3:.2f} ms total\n\n| kernel | launches | total ms | share | avg us | grid (last) | block |\n|---|---|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1]/1e3:.3f} | {a[1]/total*100:.1f}% | {a[1]/a[0]:.1f} | {a[2]} | {a[3]} |\n")
    print(open(dst).read())


FULL_METRICS = [
    ("gpu__time_duration.sum", "dur"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor inst % (hmma)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "bf16 MMA ops % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed_pipe_xu.sum", "XU (MUFU) inst"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("smsp__cycles_active.avg", "smsp active cycles"),
]


def full(src: str, dst: str) -> None:
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary: `{src}`\n\nCaptured with `ncu --set full --clock-control none --import-source on`; "
                "durations under ncu are serialised / cold-cache replays and are never bench values.\n\n")
        for r in rows[2:]:
            name = short(r[hdr.index("Kernel Name")])
            f.write(f"## launch {r[0]}: `{name}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for m, label in FULL_METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"| {label} (`{m}`) | {r[i]} | {units[i]} |\n")
            f.write("\n")
    print(open(dst).read()[:6000])


def gemm_traffic(src: str, dst: str) -> None:
    """profiles/gemm_traffic_*.json (bench.py's roofline.traffic) from an `ncu --set full` capture of ONE denoise step of
    a 1 + 1-block model at full width (LX_LAYERS=1,1 scripts/probe_dit.py 1 512 2, -k regex:gemm_bf16 -s 44 -c 8): launches
    are [x_embedder, double qkv / out / ff_up / ff_down, single qkv_mlp / proj_out, final proj_out]."""
    import json

    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    gem = [r for r in rows[2:] if "gemm_bf16_kernel" in r[hdr.index("Kernel Name")]]
    assert len(gem) >= 7, f"{len(gem)} GEMM launches captured, expected one step (8)"

    def mb(r, m):
        i = hdr.index(m)
        v = float(r[i].replace(",", ""))
        return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[units[i]]

    D, M = 3072, 2560
    shapes = [("double.qkv  M2560 N9216  K3072  (3 panels)", 19, 9216, 3072, 3), ("double.out  M2560 N3072  K3072  (3 panels)", 19, 3072, 3072, 3),
              ("double.ff_up M2560 N12288 K3072 (3 panels)", 19, 12288, 3072, 3), ("double.ff_down M2560 N3072 K12288 (3 panels)", 19, 3072, 12288, 3),
              ("single.qkv_mlp M2560 N21504 K3072 (2 panels)", 38, 21504, 3072, 2), ("single.proj_out M2560 N3072 K15360 (2 panels)", 38, 3072, 15360, 2)]
    per, tot_t, tot_a, n = {}, 0.0, 0.0, 0
    for (name, mult, N, K, panels), r in zip(shapes, gem[1:7]):
        rd, wr = mb(r, "dram__bytes_read.sum"), mb(r, "dram__bytes_write.sum")
        alg = (panels * N * K * 2 + M * K * 2 + M * N * 2 + (M * D * 2 if "out" in name or "down" in name else 0)) / 1e6
        per[name] = {"launches_per_step": mult, "dram_read": round(rd, 2), "dram_write": round(wr, 2), "algorithmic": round(alg, 1),
                     "duration_us": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")),
                     "tensor_pipe_active_pct": float(r[hdr.index("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active")].replace(",", ""))
                     if "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active" in hdr else None}
        tot_t += mult * (rd + wr)
        tot_a += mult * alg
        n += mult
    json.dump({"source": f"ncu --set full --clock-control none ({src}; dram__bytes_read.sum + dram__bytes_write.sum per launch; cold L2 "
                         "under ncu), B=1, 512x512 + image condition", "kernel": "lx::gemm_bf16_kernel", "per_shape_MB": per,
               "launches_per_step": n, "traffic_bytes_per_launch": tot_t / n * 1e6, "algorithmic_bytes_per_launch": tot_a / n * 1e6},
              open(dst, "w"), indent=1)
    print(open(dst).read())


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "gemm_traffic":
        gemm_traffic(sys.argv[2], sys.argv[3])
    elif mode == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
    else:
        full(sys.argv[2], sys.argv[3])
