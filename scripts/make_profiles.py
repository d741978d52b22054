"""Condense the round's gpurun_out/ artefacts (bench line, ncu launch list, ncu --set full raw pages) into the tracked
summaries under profiles/.   usage: python scripts/make_profiles.py r01"""
import csv, json, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
out = []

bench = json.load(open(os.path.join(G, f"bench_{tag}_final.json")))
json.dump(bench, open(os.path.join(P, f"bench_{tag}.json"), "w"), indent=1)
out.append(f"# profiles {tag}\n")
out.append(f"bench.py (N=1 B200, {bench['steps']} steps, {bench['warmup']} warm-up, clocks {bench['clocks']}):")
out.append(f"  value {bench['value']:.1f} {bench['unit']} (device-resident), e2e {bench['e2e']['value']:.1f}, {bench['ms_per_step']:.3f} ms/step, "
           f"{bench['gpu_launches']} launches; cpu_baseline {bench.get('cpu_baseline', {}).get('value')} on {bench.get('cpu_baseline', {}).get('cores')} cores")
r = bench["roofline"]
out.append(f"  roofline: {r['kernel']}: {r['achieved']:.0f} GB/s algorithmic = {r['frac']:.3f} of measured {r['peak']} GB/s; family share of step {r.get('family_share_of_step', r.get('share_of_step', 0)):.3f}\n")
steps = bench["steps"]
out.append(f"  roofline fp32 fraction {r.get('frac_fp32')}; whole step {r.get('whole_step')}")
out.append("## kernel families inside a bench step (separate CUDA-event pass; us per step, share, HBM and fp32 roofline fractions)")
for k, v in bench.get("kernel_families", {}).items():
    out.append(f"  {k:14s} {v['us_per_step']:8.1f} us  {100 * v['share_of_step']:5.1f} %  launches {v['launches_per_step']:5.1f}  frac_hbm {v['frac_hbm']:.3f}  frac_fp32 {v['frac_fp32']:.3f}")
out.append("## per (operator, kernel) launch shape")
for v in bench.get("kernels", []):
    out.append(f"  {v['op']:5s} {v['kernel']:16s} x{v['launches_per_call']:4.1f}  {v['avg_launch_us']:8.1f} us/launch  frac_hbm {v['frac_hbm']}  frac_fp32 {v['frac_fp32']}")
if bench.get("extra"):
    out.append("## extra (BASELINE configs 3-5 in the same run)")
    for k, v in bench["extra"].items():
        out.append(f"  {k}: {json.dumps(v)}")
out.append("")

# ncu launch list (cold-cache, serialised): shares
rows = list(csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))))
i0 = [i for i, r_ in enumerate(rows) if r_ and r_[0] == "ID"][0]
hdr = rows[i0]
agg = collections.OrderedDict()
for r_ in rows[i0 + 1:]:
    if len(r_) < len(hdr): continue
    rec = dict(zip(hdr, r_))
    name = rec["Kernel Name"].replace("void xl_kernel<", "").replace(">(T1::Params)", "").replace(">(Params)", "")
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(rec["Metric Value"]) / 1e3
tot = sum(v[1] for v in agg.values())
out.append(f"## ncu launch list of `python bench.py --steps 2 --warmup 1` (gpu__time_duration.sum; {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms total; cold cache + serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"  {k:34s} n={v[0]:3d}  total {v[1]:9.1f} us  avg {v[1] / v[0]:7.1f} us  {100 * v[1] / tot:5.1f} %")
out.append("")
import shutil
shutil.copy(os.path.join(G, f"launches_{tag}.csv"), os.path.join(P, f"launches_{tag}.csv"))

want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
traffic = {}
OPS = {"rsgrad": "rs", "vrsgrad": "vrs", "cztgrad": "czt", "vcztgrad": "vczt", "highna": "highna"}
for name in ("rsgrad", "vrsgrad", "cztgrad", "vcztgrad", "highna"):
    fn = os.path.join(G, f"ncu_{tag}_{name}_raw.csv")
    if not os.path.exists(fn): continue
    rows = list(csv.reader(open(fn)))
    hdr, units = rows[0], rows[1]
    mode = {"rsgrad": "grad"}.get(name, name)
    out.append(f"## ncu --set full --clock-control none, scripts/prof_rs.py 2048 {mode} (2nd iteration; one row per launch)")
    short = [w.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "").replace(".avg.pct_of_peak_sustained_active", "%").replace(".avg.pct_of_peak_sustained_elapsed", "%el") for w in want]
    for r_ in rows[2:]:
        kn = r_[hdr.index("Kernel Name")].replace("void xl_kernel<", "").replace(">(Params)", "")
        out.append(f"### {kn}")
        vals = {}
        for w, sh in zip(want, short):
            if w in hdr:
                v = r_[hdr.index(w)]; u = units[hdr.index(w)]
                vals[w] = (v, u)
                out.append(f"  {sh:52s} {v} {u}")
        try:
            rd, wr = vals["dram__bytes_read.sum"], vals["dram__bytes_write.sum"]
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            tb = float(rd[0]) * scale[rd[1]] + float(wr[0]) * scale[wr[1]]
            out.append(f"  {'traffic = dram read + write':52s} {tb / 1e6:.1f} MB")
            traffic.setdefault(kn + ":" + OPS[name], tb)
            if OPS[name] in ("rs", "czt"): traffic.setdefault(kn, tb)
        except Exception as e:
            pass
    out.append("")
json.dump(traffic, open(os.path.join(P, f"ncu_traffic_{tag}.json"), "w"), indent=1)
open(os.path.join(P, f"summary_{tag}.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:60]))
