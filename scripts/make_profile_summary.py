"""Turns the raw captures brought back in gpurun_out/ into the tracked summaries under profiles/ (run here, no GPU)."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
out = [f"# Profile summary ({tag}): tcgen05 value net + fp64 CFR wave kernels on B200\n"]

# ---- launch list (ncu --metrics gpu__time_duration.sum)
lc = os.path.join(G, f"launches_{tag}.csv")
if not os.path.exists(lc): lc = os.path.join(G, "launches_tc.csv")
if os.path.exists(lc):
    rows = list(csv.reader(open(lc)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]; data = rows[hdr + 1:]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) > vi:
            agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out.append("## Launch list\n`ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv python bench.py --steps 1 --warmup 1 "
               "--iters 100 --no-cpu-baseline` (raw: `" + tag + "_launches.csv`; cold-cache, serialised: compare SHARES).\n")
    out.append("| kernel | launches | avg us | share |\n|---|---|---|---|")
    for k, v in agg.items():
        out.append(f"| `{k[:78]}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |")
    with open(os.path.join(P, f"{tag}_launches.csv"), "w") as f:
        f.write(open(lc).read())

# ---- full-set capture of the two hot kernels
rep = os.path.join(G, f"prof_{tag}.ncu-rep")
if not os.path.exists(rep): rep = os.path.join(G, "prof_r1b.ncu-rep")
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
               "(1x6f, K = 8192 root subgames, 540 672 value-net rows; driver: `scripts/ncu_target.py`; report not tracked, extract below and in `"
               + tag + "_ncu_full_extract.csv`).\n")
    with open(os.path.join(P, f"{tag}_ncu_full_extract.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [r[hdr.index("Kernel Name")][:60] for r in rows[2:]])
        for m in keep[1:] + stall:
            if m in hdr:
                i = hdr.index(m)
                w.writerow([m, units[i]] + [r[i] for r in rows[2:]])
    traffic = {}
    for r in rows[2:]:
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
        traffic["cfr_iter_kernel" if "cfr_iter_kernel" in r[hdr.index("Kernel Name")] else "leaf_mlp_tc_kernel" if "leaf_mlp_tc" in r[hdr.index("Kernel Name")] else r[hdr.index("Kernel Name")][:40]] = {
            "dram_bytes_per_launch": float(r[rd].replace(",", "")) * scale[units[rd]] + float(r[wr].replace(",", "")) * scale[units[wr]],
            "workload": "1x6f K=8192 root subgames", "source": f"profiles/{tag}_ncu_full_extract.csv (ncu --set full, one launch)"}
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    for r in rows[2:]:
        out.append(f"### `{r[hdr.index('Kernel Name')][:90]}`\n")
        out.append("| metric | value |\n|---|---|")
        for m in keep[1:]:
            if m in hdr:
                i = hdr.index(m)
                out.append(f"| {m} | {r[i]} {units[i]} |")
        st = sorted(((float(r[hdr.index(h)].replace(',', '') or 0), h) for h in stall), reverse=True)[:6]
        out.append("| top stalls (warps per issue) | " + ", ".join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in st) + " |")
        out.append("")

# ---- bench line
bj = os.path.join(G, f"bench_{tag}.json")
if not os.path.exists(bj): bj = os.path.join(G, "bench_r1_tc.json")
if os.path.exists(bj):
    b = json.loads(open(bj).read().strip().splitlines()[-1])
    open(os.path.join(P, f"{tag}_bench.json"), "w").write(json.dumps(b, indent=1))
    r = b["roofline"]
    out.append("## bench.py (no profiler attached), same build\n")
    out.append(f"* value = {b['value']:.4e} subgame-iters/s ({b['ms_per_step']:.1f} ms per 8192x1024 wave), e2e = {b['e2e']['value']:.4e}; "
               f"clocks {b['clocks']}")
    out.append(f"* value-net kernel: {r['avg_launch_ms'] * 1e3:.0f} us/launch (CUDA events, {r.get('launch_timing', '')}) = {100 * r['share_of_step']:.0f}% of the step; "
               f"{r['achieved']:.0f} TFLOP/s algorithmic = {100 * r['frac']:.1f}% of the measured sustained bf16 peak ({r['peak']} TFLOP/s); DRAM traffic per launch {r.get('traffic')}")
    if "cpu_baseline" in b:
        out.append(f"* CPU baseline in the same run: {b['cpu_baseline']}")
pn = os.path.join(G, "parity_notes.log")
if os.path.exists(pn):
    open(os.path.join(P, f"{tag}_parity_notes.log"), "w").write(open(pn).read())
    out.append(f"\nMeasured parity numbers of the same build: `{tag}_parity_notes.log`.")
open(os.path.join(P, f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
