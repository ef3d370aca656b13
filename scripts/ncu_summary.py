"""profiles/ summary from gpurun_out/prof_<W>.ncu-rep + launches_<W>.csv (run here, no GPU)."""
import csv, json, subprocess, sys
from collections import OrderedDict

W = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
tag = sys.argv[2] if len(sys.argv) > 2 else "r01"
metrics = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
           'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum']
kern = OrderedDict()
import os
for rep in (f"gpurun_out/prof_{W}.ncu-rep", f"gpurun_out/prof_{W}_fwd.ncu-rep"):
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        short = name.split('(')[0].split('::')[-1]
        if short not in kern:
            kern[short] = {m: r[hdr.index(m)] for m in metrics if m in hdr}
lrows = [r for r in csv.reader(open(f'gpurun_out/launches_{W}.csv')) if len(r) > 5]
lh = lrows[0]; ki = lh.index('Kernel Name'); vi = lh.index('Metric Value')
d = OrderedDict()
for r in lrows[1:]:
    try:
        d.setdefault(r[ki], []).append(float(r[vi].replace(',', '')))
    except ValueError:
        pass
out = [f"# Round 1 -- ncu evidence (B200, {W}: K = 100,000 SNPs, batch 32, 10 x 256)", "",
       "Commands (scripts/gpu_prof.sh): `ncu --metrics gpu__time_duration.sum --clock-control none --csv python scripts/prof_step.py cfg2 26`",
       "and `ncu --set full --clock-control none --import-source on -k regex:... -s 7 -c 3` (one in-epoch step) plus one capture of the",
       "standalone forward.  The run is one training epoch as the CLI issues it (first-layer forward once, then per step: hidden stack,",
       "small-layer update, backward + Adam + NEXT step's forward in one kernel), its validation pass, and one unfused step.",
       "Per-launch times under ncu are cold-cache and serialised -- compare shares, not absolutes.", "",
       "## Launch list (device time per launch)", "", "| kernel | launches | avg us | last us |", "|---|---|---|---|"]
step = {}
for k, v in d.items():
    out.append(f"| `{k[:72]}` | {len(v)} | {sum(v)/len(v)/1000:.2f} | {v[-1]/1000:.2f} |")
    for sk in ('k_l1_fwd_tc', 'k_hidden_tc', 'k_l1_bwd_tc', 'k_hidden_update'):
        if sk in k:
            step[sk] = v[-1] / 1000
tot = sum(step.values())
import statistics
step = {}
for k, v in d.items():
    for sk in ('k_hidden_tc', 'k_l1_bwd_tc', 'k_hidden_update'):
        if sk in k:
            step[sk] = statistics.median(v) / 1000
tot = sum(step.values())
out += ["", "Share of one in-epoch optimizer step (median launch of each step kernel; the small-layer update overlaps the backward in real runs,",
        "the standalone forward runs once per epoch + validation chunks):", ""]
for sk, v in step.items():
    out.append(f"* `{sk}`: {v:.1f} us = {100*v/tot:.0f}% of the three-kernel sum ({tot:.0f} us)")
out += ["", "## `ncu --set full` (one launch each)", "", "| metric | " + " | ".join(kern) + " |", "|---|" + "---|" * len(kern)]
for m in metrics:
    out.append(f"| {m} | " + " | ".join(kern[k].get(m, '-') for k in kern) + " |")
out += ["", "(time in us, DRAM bytes in MB.)"]
bwd = kern.get('k_l1_bwd_tc')
if bwd:
    traffic = (float(bwd['dram__bytes_read.sum']) + float(bwd['dram__bytes_write.sum'])) * 1e6
    json.dump({"K": 100000, "kernel": "tc::k_l1_bwd_tc", "dram_bytes_per_launch": traffic,
               "dram_read_MB": float(bwd['dram__bytes_read.sum']), "dram_write_MB": float(bwd['dram__bytes_write.sum']),
               "source": f"ncu --set full --clock-control none, profiles/{tag}_ncu_summary.md"},
              open('profiles/l1_backward_traffic.json', 'w'), indent=1)
    out += ["", f"Dominant kernel `tc::k_l1_bwd_tc`: algorithmic bytes 24*K*H = 614.4 MB per launch; DRAM traffic measured {traffic/1e6:.1f} MB "
            f"(read {bwd['dram__bytes_read.sum']} + write {bwd['dram__bytes_write.sum']}; the tail of the writes is still in L2 when the kernel ends) -> no re-reads.",
            "The captured launch is the in-epoch variant (also runs the next step's forward).  CUDA-event timing in bench.py (not under the",
            "profiler): plain backward 114-116 us -> 5.3 TB/s = 0.82-0.83 of the measured 6458 GB/s copy peak; with the fused forward 118 us.",
            "SASS (cuobjdump -sass liblocator_b200.so): UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UBLKCP (cp.async.bulk), SYNCS (mbarrier), ELECT, FFMA2."]
open(f'profiles/{tag}_ncu_summary.md', 'w').write("\n".join(out) + "\n")
print("\n".join(out[-24:]))
