#!/usr/bin/env python
"""Turns the raw ncu outputs in gpurun_out/ into the tracked summaries under profiles/ (run here, no GPU needed)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2a"
launches = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    agg[r[ik].split("(")[0]].append(v / 1000 if r[iu] == "ns" else v)
tot = sum(sum(v) for v in agg.values())
lines = [f"# {tag} ncu launch list summary (gpu__time_duration.sum, --clock-control none; cold-cache serialised launches: compare SHARES)",
         "# command: python bench.py --steps 1 --warmup 3 --ticks-per-step 40 --fused-chunk 40 --no-cpu-baseline --no-e2e --no-stagger --no-config4",
         "# (tower_kernel<2, 0, true> = the fused tick kernel, 40 ticks per launch; tower_kernel<2, 0, false> + advance_kernel = the 64 instrumented separate ticks)",
         "kernel,launches,mean_us,share"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    lines.append(f"{k},{len(v)},{sum(v) / len(v):.2f},{sum(v) / tot:.4f}")
open(os.path.join(out_dir, f"{tag}_launches_summary.csv"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
env_rep = os.path.join(ROOT, "gpurun_out", f"prof_env_{tag}.ncu-rep")
if os.path.exists(env_rep):   # the env kernel is captured from scripts/dbg_env_ncu.py (16 Mi boards, operands from DRAM)
    raw_env = subprocess.run(["ncu", "-i", env_rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(out_dir, f"{tag}_ncu_env_raw.csv"), "w").write(raw_env)
else:
    raw_env = None
open(os.path.join(out_dir, f"{tag}_ncu_full_raw.csv"), "w").write(raw)
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "gpc__cycles_elapsed.avg.per_second", "launch__cluster_dim_x", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
idx = [(w, h.index(w)) for w in want if w in h]
units = rr[1]
summ = []
for r in rr[2:]:
    summ.append({w: (r[i] + (" " + units[i] if units[i] else "")) for w, i in idx})
fused_rep = os.path.join(ROOT, "gpurun_out", f"prof_fused_{tag}.ncu-rep")
FUSED_TICKS = 40   # scripts/capture_profiles.sh: --ticks-per-step 40 --fused-chunk 40
fused_traffic = None
if os.path.exists(fused_rep):   # the fused tick kernel: one launch = 40 ticks
    raw_f = subprocess.run(["ncu", "-i", fused_rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(out_dir, f"{tag}_ncu_fused_raw.csv"), "w").write(raw_f)
    rf = list(csv.reader(raw_f.splitlines()))
    hf, uf = rf[0], rf[1]
    for r in rf[2:]:
        d = {w: (r[hf.index(w)] + (" " + uf[hf.index(w)] if uf[hf.index(w)] else "")) for w in want if w in hf}
        d["ticks_per_launch"] = FUSED_TICKS
        summ.append(d)

    def _bytes(s_):
        v, u_ = s_.split()[0], (s_.split() + [""])[1]
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u_, 1)
    tr_f = [(_bytes(x["dram__bytes_read.sum"]) + _bytes(x["dram__bytes_write.sum"])) / FUSED_TICKS for x in summ if x.get("ticks_per_launch")]
    if tr_f:
        fused_traffic = sum(tr_f) / len(tr_f)
if raw_env:   # same metrics for the env kernel (its raw page has its own column set)
    re_ = list(csv.reader(raw_env.splitlines()))
    he, ue = re_[0], re_[1]
    for r in re_[2:]:
        summ.append({w: (r[he.index(w)] + (" " + ue[he.index(w)] if ue[he.index(w)] else "")) for w in want if w in he})
# the SGD step on the device: launch list of scripts/dbg_train_time.py (all launches of the process: 3 warm + 2 timed steps) and the full
# capture of the two tcgen05 GEMM kernels
tl = os.path.join(ROOT, "gpurun_out", f"train_launches_{tag}.csv")
if os.path.exists(tl):
    rows_t = [r for r in csv.reader(open(tl)) if len(r) > 5]
    ht = rows_t[0]
    jk, jv, ju = ht.index("Kernel Name"), ht.index("Metric Value"), ht.index("Metric Unit")
    agg_t = collections.defaultdict(list)
    for r in rows_t[1:]:
        try:
            v = float(r[jv].replace(",", ""))
        except ValueError:
            continue
        agg_t[r[jk].split("(")[0]].append(v / 1000 if r[ju] == "ns" else v)
    tot_t = sum(sum(v) for v in agg_t.values())
    lt = [f"# {tag} ncu launch list of the device SGD step (python scripts/dbg_train_time.py 20 128 2: ResidualTower-20, batch 128, 5 steps incl. warm-up;",
          "# gpu__time_duration.sum, --clock-control none; cold-cache serialised launches: compare SHARES)", "kernel,launches,mean_us,share"]
    for k, v in sorted(agg_t.items(), key=lambda kv: -sum(kv[1])):
        lt.append(f"{k},{len(v)},{sum(v) / len(v):.2f},{sum(v) / tot_t:.4f}")
    open(os.path.join(out_dir, f"{tag}_train_launches_summary.csv"), "w").write("\n".join(lt) + "\n")
    print("\n".join(lt))
train_rep = os.path.join(ROOT, "gpurun_out", f"prof_train_{tag}.ncu-rep")
if os.path.exists(train_rep):
    raw_t = subprocess.run(["ncu", "-i", train_rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(out_dir, f"{tag}_ncu_train_raw.csv"), "w").write(raw_t)
    rt = list(csv.reader(raw_t.splitlines()))
    hq, uq = rt[0], rt[1]
    for r in rt[2:]:
        summ.append({w: (r[hq.index(w)] + (" " + uq[hq.index(w)] if uq[hq.index(w)] else "")) for w in want if w in hq})
json.dump(summ, open(os.path.join(out_dir, f"{tag}_ncu_kernels.json"), "w"), indent=1)
tw = [x for x in summ if "tower_kernel" in x["Kernel Name"] and not x.get("ticks_per_launch")]
tpath = os.path.join(out_dir, "tower_traffic.json")
traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
if tw:
    def to_bytes(s):
        v, u = s.split()[0], (s.split() + [""])[1]
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    tr = [to_bytes(x["dram__bytes_read.sum"]) + to_bytes(x["dram__bytes_write.sum"]) for x in tw]
    traffic["dram_bytes_per_launch"] = sum(tr) / len(tr)
    traffic["source"] = f"profiles/{tag}_ncu_full_raw.csv (ncu --set full, tower_kernel<2>, {len(tr)} launches)"
if fused_traffic is not None:
    traffic["dram_bytes_per_tick_fused"] = fused_traffic
    traffic["source_fused"] = (f"profiles/{tag}_ncu_fused_raw.csv (ncu --set full, fused tick kernel, bytes per launch / {FUSED_TICKS} ticks: the weights are "
                               "read from DRAM once per launch and stay in L2, so are the search trees)")
json.dump(traffic, open(tpath, "w"))
for x in summ:
    print({k: v for k, v in x.items()})
