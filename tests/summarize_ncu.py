"""Turn the CSV pages of the round-end ncu captures (tests/final_evidence.sh) into the summaries kept under profiles/.

    python tests/summarize_ncu.py gpurun_out profiles r02
"""
import csv
import json
import sys
from collections import OrderedDict

src, dst, tag = sys.argv[1], sys.argv[2], sys.argv[3]
COLS = [("gpu__time_duration.sum", "us"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("dram__bytes_read.sum", "DRAM read MB"),
        ("dram__bytes_write.sum", "DRAM write MB"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs")]


def full_summary(path, out, title):
    rows = list(csv.reader(open(path)))
    h = rows[0]
    idx = {n: i for i, n in enumerate(h)}
    units = rows[1]
    lines = [title, "kernel | " + " | ".join(c[1] for c in COLS)]
    for r in rows[2:]:
        vals = []
        for c, _ in COLS:
            v = r[idx[c]] if c in idx else ""
            if c == "gpu__time_duration.sum" and c in idx:
                f = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(units[idx[c]], 1.0)
                v = f"{float(v) * f:.1f}"
            elif c.startswith("dram__bytes") and c in idx:
                u = units[idx[c]]
                f = {"Mbyte": 1.0, "Gbyte": 1e3, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0)
                v = f"{float(v) * f:.1f}"
            else:
                try:
                    v = f"{float(v):.1f}"
                except ValueError:
                    pass
            vals.append(v)
        lines.append(r[idx["Kernel Name"]][:72] + " | " + " | ".join(vals))
    open(out, "w").write("\n".join(lines) + "\n")
    return rows


def launch_summary(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        name, val = r[4].split("(")[0].strip(), float(r[-1]) / 1e6  # ns -> ms
        a = agg.setdefault(name, [0.0, 0])
        a[0] += val
        a[1] += 1
        total += val
    lines = [title, f"# total {total:.3f} ms", ""]
    for name, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        lines.append(f"{ms:8.3f} ms  {100 * ms / total:5.1f} %  x{n:3d}  {name}")
    open(out, "w").write("\n".join(lines) + "\n")


full_summary(f"{src}/{tag}_step_full2_raw.csv", f"{dst}/{tag}_ncu_full_summary.txt",
             "# ncu --set full --clock-control none, 17 consecutive launches of the 2nd step of config 2 (mel, stem x8 for two chunk groups, then LN, qkv, attention, out_proj, LN, fc1, fc2 of layer 0, LN of layer 1)")
full_summary(f"{src}/{tag}_prefill_full2_raw.csv", f"{dst}/{tag}_ncu_prefill_summary.txt",
             "# ncu --set full --clock-control none, the 19 launches of one decoder-prefill call (tests/ncu_prefill.py: 2 layers, 1.7B widths, 64 x 407 prompt rows)\n"
             "# gemm_bf16_sm100<256,6,0,EPI,2,0>: EPI 0 = q|k|v (bf16 store), 2 = o_proj / down (fp32 TMA reduce-add), 10 = gate|up SwiGLU, 3 = lm_head (fp32 TMA store)")
launch_summary(f"{src}/{tag}_launches_final.csv", f"{dst}/{tag}_ncu_launches_summary.txt",
               "# ncu launch list of ONE step (config 2: 64 x 30 s, eager launches, 180 kernels) -- gpu__time_duration.sum, --clock-control none\n"
               "# cold-cache, serialised: compare SHARES with bench.py's event-timed `kernels`, not absolutes")
print(open(f"{dst}/{tag}_ncu_launches_summary.txt").read())

# DRAM bytes (read + write) per launch for bench.py's roofline.traffic, from the 17-launch full capture (launch order is fixed)
# (fused waveform -> embeddings path: ONE mel launch, conv1 applies the clamp / rescale; 180 launches per step)
ORDER = ["mel_logmel", "conv1", "conv2_igemm", "conv3_igemm", "conv_out_gemm", "conv1", "conv2_igemm", "conv3_igemm",
         "conv_out_gemm", "layernorm", "gemm_qkv", "window_attention", "gemm_out_proj", "layernorm", "gemm_fc1", "gemm_fc2", "layernorm"]
_names = [r[[i for i, n in enumerate(list(csv.reader(open(f"{src}/{tag}_step_full2_raw.csv")))[0]) if n == "Kernel Name"][0]]
          for r in list(csv.reader(open(f"{src}/{tag}_step_full2_raw.csv")))[2:]]
_key = {"mel_logmel": "mel_logmel", "conv1": "conv1_gelu", "layernorm": "layernorm", "window_attention": "window_attention"}
for _o, _n in zip(ORDER, _names):  # the fixed order must match what was captured
    assert (_key[_o] in _n) if _o in _key else ("gemm_bf16_sm100" in _n), (_o, _n)
rows = list(csv.reader(open(f"{src}/{tag}_step_full2_raw.csv")))
h, units = rows[0], rows[1]
idx = {n: i for i, n in enumerate(h)}
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
acc = {}
for name, r in zip(ORDER, rows[2:]):
    b = sum(float(r[idx[c]]) * scale.get(units[idx[c]], 1.0) for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    acc.setdefault(name, []).append(b)
json.dump({"source": f"profiles/{tag}_ncu_full_17kernels_raw.csv: ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per launch, "
                     "config 2 (64 x 30 s); conv kernels: per launch of a 1024/896-chunk group (mean of the two)",
           "bytes_per_launch": {k: sum(v) / len(v) for k, v in acc.items()}}, open(f"{dst}/ncu_traffic.json", "w"), indent=1)
