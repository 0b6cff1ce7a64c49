"""Turns the raw captures brought back in gpurun_out/ by scripts/gpu_validate.sh <tag> into the tracked summaries under profiles/
(run here, in the build container; no GPU needed):  python scripts/make_profile_summary.py r2"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
out = [f"# Profile summary ({tag}): tcgen05 value net + fp64 CFR wave kernels + device-side self-play on B200\n"]


def num(x):
    return float(x.replace(",", "")) if x not in ("", "-") else 0.0


# ---- launch list (ncu --metrics gpu__time_duration.sum)
lc = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(lc):
    rows = list(csv.reader(open(lc)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]; data = rows[hdr + 1:]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) > vi:
            agg.setdefault(r[ki], []).append(num(r[vi]))
    tot = sum(sum(v) for v in agg.values())
    out.append("## Launch list\n`ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 300 --csv python bench.py --steps 1 --warmup 1 --iters 64 "
               "--no-cpu-baseline` (the default data-generation workload with 64-iteration waves; raw: `" + tag + "_launches.csv`; cold-cache, "
               "serialised: compare SHARES).\n")
    out.append("| kernel | launches | avg us | share |\n|---|---|---|---|")
    for k, v in agg.items():
        out.append(f"| `{k[:78]}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |")
    shutil.copy(lc, os.path.join(P, f"{tag}_launches.csv"))

# ---- full-set capture of the two hot kernels inside a self-play wave
rep = os.path.join(G, f"{tag}_prof.ncu-rep")
traffic_path = os.path.join(P, "traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__warps_eligible.avg.per_cycle_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct"]
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    out.append("\n## `ncu --set full --clock-control none --import-source on` of one launch of each hot kernel\n"
               "(inside a self-play wave of 8192 1x6f games; driver: `scripts/datagen_probe.py --iters 64`; report not tracked, extract below and in `"
               + tag + "_ncu_full_extract.csv`).\n")
    with open(os.path.join(P, f"{tag}_ncu_full_extract.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [r[hdr.index("Kernel Name")][:60] for r in rows[2:]])
        for m in keep[1:] + stall:
            if m in hdr:
                i = hdr.index(m)
                w.writerow([m, units[i]] + [r[i] for r in rows[2:]])
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
    for r in rows[2:]:
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        name = r[hdr.index("Kernel Name")]
        b = num(r[rd]) * scale[units[rd]] + num(r[wr]) * scale[units[wr]]
        key = "value_net_dram_bytes_per_launch" if "leaf_mlp" in name else "cfr_dram_bytes_per_launch"
        traffic.setdefault("datagen_1x6_8192", {})[key] = b
        traffic["datagen_1x6_8192"]["source"] = f"profiles/{tag}_ncu_full_extract.csv (ncu --set full, one launch inside a self-play wave of 8192 games)"
        out.append(f"### `{name[:90]}`\n")
        out.append("| metric | value |\n|---|---|")
        for m in keep[1:]:
            if m in hdr:
                i = hdr.index(m)
                out.append(f"| {m} | {r[i]} {units[i]} |")
        st = sorted(((num(r[hdr.index(h)]), h) for h in stall), reverse=True)[:6]
        out.append("| top stalls (warps per issue) | " + ", ".join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in st) + " |")
        out.append("")
    json.dump(traffic, open(traffic_path, "w"), indent=1)

# ---- bench lines
for suffix, title in (("bench", "default workload: self-play data generation"), ("bench_solve", "--workload solve: homogeneous root-subgame waves"),
                      ("bench_reference", "--impl reference: the reference's loop on the host cores")):
    bj = os.path.join(G, f"{tag}_{suffix}.json")
    if not os.path.exists(bj) or not open(bj).read().strip():
        continue
    b = json.loads(open(bj).read().strip().splitlines()[-1])
    open(os.path.join(P, f"{tag}_{suffix}.json"), "w").write(json.dumps(b, indent=1))
    out.append(f"\n## bench.py, {title} (no profiler attached), same build\n")
    out.append(f"* value = {b['value']:.4e} subgame-iters/s ({b['ms_per_step']:.1f} ms per step), e2e = {b['e2e']['value']:.4e}; clocks {b.get('clocks')}")
    r = b.get("roofline")
    if r:
        out.append(f"* value-net kernel: {r['avg_launch_ms'] * 1e3:.1f} us/launch for {r.get('rows_per_launch', 0):.0f} rows ({r.get('launch_timing', '')}) = {100 * r['share_of_step']:.0f}% of the step; "
                   f"{r['achieved']:.0f} TFLOP/s algorithmic = {100 * r['frac']:.1f}% of the measured sustained bf16 peak ({r['peak']} TFLOP/s); DRAM traffic per launch {r.get('traffic')}")
        c = r.get("cfr_kernel")
        if c:
            out.append(f"* CFR kernel: {c['avg_launch_ms'] * 1e3:.1f} us/launch, {c['algorithmic_bytes_per_launch'] / 1e6:.1f} MB algorithmic (fp32-equivalent) per launch = "
                       f"{c['achieved']:.0f} GB/s = {100 * c['frac']:.1f}% of the measured HBM copy bandwidth ({c['peak']} GB/s); {100 * c['share_of_step']:.0f}% of the step")
    if "cpu_baseline" in b:
        out.append(f"* CPU baseline in the same run: {b['cpu_baseline']}")
for name in ("parity_notes.log", "pytest_gpu.log", "tc_trace.log", "datagen_probe.log", "cfr_probe.log", "smoke.log", "gelu_table.log"):
    src = os.path.join(G, f"{tag}_{name}")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, f"{tag}_{name}"))
        out.append(f"\n`{tag}_{name}`: " + {"parity_notes.log": "measured parity numbers written by the -m gpu tests", "pytest_gpu.log": "tail of pytest -m gpu",
                                           "tc_trace.log": "clock64 timeline of one CTA of the value-net kernel", "datagen_probe.log": "per-kernel time inside self-play waves (with the net / zero net)",
                                           "cfr_probe.log": "CFR kernel alone (zero net), root and self-play waves", "smoke.log": "__graft_entry__.smoke()",
                                           "gelu_table.log": "error of tanh.approx.f16x2 and of both fast-GELU evaluations on every fp16 input"}[name])
open(os.path.join(P, f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:4000])
